// Functional probe: tcgen05.mma kind::f16 with the A operand in tensor memory ("TS").  Checks the assumed layouts:
//   A[m][k] (bf16)  -> TMEM lane m, 32-bit column k/2 (even k in the low half), written with tcgen05.st.32x32b
//   B[n][k] (bf16)  -> shared memory, K-major rows of 128 bytes, SWIZZLE_128B
//   D[m][n] (fp32)  -> TMEM lane m, column n
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ts_probe tools/ts_probe.cu && tools/ts_probe
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

constexpr int M = 128, N = 128, K = 64;

__global__ void __launch_bounds__(128) probe(const __nv_bfloat16* A, const __nv_bfloat16* B, float* D, int m_instr) {
    extern __shared__ uint8_t raw[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // B tile: row n, 16-byte chunk c (8 bf16) -> byte offset (n/8)*1024 + (n%8)*128 + ((c ^ (n%8)) * 16)
    for (int i = threadIdx.x; i < N * 8; i += blockDim.x) {
        const int n = i / 8, c = i % 8;
        const uint4 v = *reinterpret_cast<const uint4*>(B + (size_t)n * K + c * 8);
        *reinterpret_cast<uint4*>(smem + (n / 8) * 1024 + (n % 8) * 128 + ((c ^ (n % 8)) * 16)) = v;
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
    const uint32_t a_tm = tmem + 128;                 // columns 128 .. 128 + K/2
    // A rows into TMEM: this thread owns lane m = 32*warp + lane
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
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(m_instr >> 4) << 24);
        const uint64_t hi = (uint64_t)(((1024u >> 4) & 0x3FFFu) | (1u << 14) | (2u << 29)) << 32;
        const uint32_t b_lo = ((smem_u32(smem) >> 4) & 0x3FFFu) | (1u << 16);
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
    // D: lane m, columns 0..N-1
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
    __nv_bfloat16 *hA = new __nv_bfloat16[M * K], *hB = new __nv_bfloat16[N * K];
    float* fa = new float[M * K]; float* fb = new float[N * K];
    srand(1);
    for (int i = 0; i < M * K; ++i) { hA[i] = __float2bfloat16((rand() % 17 - 8) / 8.f); fa[i] = __bfloat162float(hA[i]); }
    for (int i = 0; i < N * K; ++i) { hB[i] = __float2bfloat16((rand() % 13 - 6) / 4.f); fb[i] = __bfloat162float(hB[i]); }
    __nv_bfloat16 *dA, *dB; float* dD;
    cudaMalloc(&dA, M * K * 2); cudaMalloc(&dB, N * K * 2); cudaMalloc(&dD, M * N * 4);
    cudaMemcpy(dA, hA, M * K * 2, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB, N * K * 2, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * 1024);
    float* hD = new float[M * N];
    for (int m_instr : {128, 64}) {
        cudaMemset(dD, 0, M * N * 4);
        probe<<<1, 128, 18 * 1024>>>(dA, dB, dD, m_instr);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("M=%d: %s\n", m_instr, cudaGetErrorString(e)); return 1; }
        cudaMemcpy(hD, dD, M * N * 4, cudaMemcpyDeviceToHost);
        // compare per row: which rows of D match the expected product?
        int good_rows = 0; double maxerr = 0;
        int row_ok[M];
        for (int m = 0; m < M; ++m) {
            double err = 0;
            for (int n = 0; n < N; ++n) {
                double ref = 0;
                for (int k = 0; k < K; ++k) ref += (double)fa[m * K + k] * fb[n * K + k];
                err = fmax(err, fabs(ref - hD[m * N + n]));
            }
            row_ok[m] = err < 1e-3; good_rows += row_ok[m];
            if (m < m_instr) maxerr = fmax(maxerr, err);
        }
        printf("M=%d: rows matching = %d / %d, max err over rows < M: %g\n  ok map: ", m_instr, good_rows, M, maxerr);
        for (int m = 0; m < M; ++m) printf("%d", row_ok[m]);
        printf("\n");
        if (m_instr == 64) {
            // where did rows 0..63 of the product land?  search lanes for each expected row
            printf("  M=64 placement (expected row -> lane): ");
            for (int r = 0; r < 64; r += 8) {
                int found = -1;
                for (int m = 0; m < M && found < 0; ++m) {
                    double err = 0;
                    for (int n = 0; n < N; ++n) {
                        double ref = 0;
                        for (int k = 0; k < K; ++k) ref += (double)fa[r * K + k] * fb[n * K + k];
                        err = fmax(err, fabs(ref - hD[m * N + n]));
                    }
                    if (err < 1e-3) found = m;
                }
                printf("%d->%d ", r, found);
            }
            printf("\n");
        }
    }
    return 0;
}
