// Functional probe: can the smem (B) operand of tcgen05.mma start at a row that is NOT a multiple of 8 rows (the swizzle
// repeat) when the descriptor's base_offset field carries the row phase?  Layout under test (what a 2-D halo TMA box
// (C, 16 w, H h) leaves in shared memory): row index = h * 16 + w, rows of RB bytes (128 / 64 / 32) with the matching
// swizzle; the operand of filter tap kw is rows {h * 16 + kw + tw : tw < 8, h < 16} -> start = kw rows, SBO = 16 rows.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/shift_probe tools/shift_probe.cu && tools/shift_probe
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

constexpr int M = 128, N = 128, ROWS = 16 * 18;

// K = RB / 2 elements per row (one swizzle row), K16 steps = RB / 32
__global__ void __launch_bounds__(128) probe(const __nv_bfloat16* A, const __nv_bfloat16* Bm, float* D, int RB, int shift, int use_base_offset, int pitch) {
    extern __shared__ uint8_t raw[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int K = RB / 2, chunks = RB / 16;
    // write ROWS rows the way TMA would: row i at i * RB, 16-byte chunk c at position c ^ ((i * RB / 128) % (RB / 16))  [address bits 4-6 ^= bits 7-9]
    for (int i = threadIdx.x; i < ROWS * chunks; i += blockDim.x) {
        const int r = i / chunks, c = i % chunks;
        const uint32_t off = (uint32_t)r * RB + (uint32_t)c * 16;
        const uint32_t sw = off ^ (((off >> 7) & 7u) << 4);
        const uint32_t swm = RB == 128 ? sw : RB == 64 ? (off ^ (((off >> 7) & 3u) << 4)) : (off ^ (((off >> 7) & 1u) << 4));
        *reinterpret_cast<uint4*>(smem + swm) = *reinterpret_cast<const uint4*>(Bm + (size_t)r * K + c * 8);
    }
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(256) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    const uint32_t a_tm = tmem + 128;
    {
        const int m = warp * 32 + lane;
        for (int k0 = 0; k0 < K; k0 += 16) {
            const uint4 v0 = *reinterpret_cast<const uint4*>(A + (size_t)m * K + k0);
            const uint4 v1 = *reinterpret_cast<const uint4*>(A + (size_t)m * K + k0 + 8);
            const uint32_t addr = a_tm + ((uint32_t)(warp * 32) << 16) + (uint32_t)(k0 / 2);
            asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                         ::"r"(addr), "r"(v0.x), "r"(v0.y), "r"(v0.z), "r"(v0.w), "r"(v1.x), "r"(v1.y), "r"(v1.z), "r"(v1.w) : "memory");
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (threadIdx.x == 0) {
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        const uint32_t swz = RB == 128 ? 2u : RB == 64 ? 4u : 6u;
        const uint32_t sbo = (uint32_t)pitch * RB;           // `pitch` rows between 8-row groups
        const uint32_t row_phase = ((uint32_t)shift * RB >> 7) & 7u;
        const uint64_t hi = (uint64_t)(((sbo >> 4) & 0x3FFFu) | (1u << 14) | ((use_base_offset ? row_phase : 0u) << 17) | (swz << 29)) << 32;
        const uint32_t b_lo = (((smem_u32(smem) + (uint32_t)shift * RB) >> 4) & 0x3FFFu) | (1u << 16);
        for (int i = 0; i < K / 16; ++i) {
            const uint32_t acc = i > 0;
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                         "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                         ::"r"(tmem), "r"(a_tm + (uint32_t)i * 8u), "l"(hi | (b_lo + 2u * i)), "r"(idesc), "r"(acc) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    {
        uint32_t ok = 0;
        while (!ok)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int m = warp * 32 + lane;
    for (int n0 = 0; n0 < N; n0 += 16) {
        uint32_t v[16];
        const uint32_t addr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)n0;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                       "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) : "r"(addr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 16; ++j) D[(size_t)m * N + n0 + j] = __uint_as_float(v[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256) : "memory");
}

int main() {
    const int KMAX = 64;
    __nv_bfloat16 *hA = new __nv_bfloat16[M * KMAX], *hB = new __nv_bfloat16[ROWS * KMAX];
    float* hD = new float[M * N];
    __nv_bfloat16 *dA, *dB; float* dD;
    cudaMalloc(&dA, M * KMAX * 2); cudaMalloc(&dB, ROWS * KMAX * 2); cudaMalloc(&dD, M * N * 4);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    for (int RB : {128, 64, 32}) {
        const int K = RB / 2;
        srand(RB);
        for (int i = 0; i < M * K; ++i) hA[i] = __float2bfloat16((rand() % 17 - 8) / 8.f);
        for (int i = 0; i < ROWS * K; ++i) hB[i] = __float2bfloat16((rand() % 13 - 6) / 4.f);
        cudaMemcpy(dA, hA, M * K * 2, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB, ROWS * K * 2, cudaMemcpyHostToDevice);
      for (int pitch : {16, 10}) {
        for (int shift : {0, 1, 2, pitch + 1, 2 * pitch + 2}) {            // row shift = kh * pitch + kw
            for (int ubo : {0, 1}) {
                if (ubo && pitch != 16) continue;
                cudaMemset(dD, 0, M * N * 4);
                probe<<<1, 128, 48 * 1024>>>(dA, dB, dD, RB, shift, ubo, pitch);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("RB=%d shift=%d: %s\n", RB, shift, cudaGetErrorString(e)); return 1; }
                cudaMemcpy(hD, dD, M * N * 4, cudaMemcpyDeviceToHost);
                int bad = 0; double maxerr = 0;
                for (int m = 0; m < M; ++m)
                    for (int n = 0; n < N; ++n) {
                        const int row = (n / 8) * pitch + (n % 8) + shift;    // operand row n -> smem row
                        double ref = 0;
                        for (int k = 0; k < K; ++k) ref += (double)__bfloat162float(hA[m * K + k]) * __bfloat162float(hB[row * K + k]);
                        const double err = fabs(ref - hD[m * N + n]);
                        if (err > 1e-3) ++bad;
                        maxerr = fmax(maxerr, err);
                    }
                printf("row bytes %3d  pitch %2d  shift %2d rows  base_offset %s : %s (mismatches %d, max err %g)\n", RB, pitch, shift, ubo ? "set " : "zero", bad ? "WRONG" : "ok", bad, maxerr);
            }
        }
      }
    }
    return 0;
}
