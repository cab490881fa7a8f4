// Probe: A operand staged smem -> TMEM with tcgen05.cp (.128x256b = one K16 step of 128 rows), then tcgen05.mma with A in
// tensor memory and B in shared memory.  (1) functional check against the CPU, (2) cycles per {cp, mma} pair vs plain SS mma.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/cp_probe tools/cp_probe.cu && tools/cp_probe
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

constexpr int M = 128, K = 64;

// mode 0: functional (N = n, K = 64: 4 x {cp, mma}); mode 1: timing cp+mma; mode 2: timing SS mma; mode 3: timing cp only
__global__ void __launch_bounds__(128) probe(const __nv_bfloat16* A, const __nv_bfloat16* Bm, float* D, int N, int mode, int iters, long long* cyc) {
    extern __shared__ uint8_t raw[];
    __shared__ __align__(8) uint64_t bar2[2];
    uint64_t& bar = bar2[threadIdx.x == 32 ? 1 : 0];
    __shared__ uint32_t tmem_base_s;
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem;                  // 128 rows x 128 B, SWIZZLE_128B
    uint8_t* sB = smem + 16384;          // N rows x 128 B
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < M * 8; i += blockDim.x) {
        const int r = i / 8, c = i % 8;
        *reinterpret_cast<uint4*>(sA + (r / 8) * 1024 + (r % 8) * 128 + ((c ^ (r % 8)) * 16)) = *reinterpret_cast<const uint4*>(A + (size_t)r * K + c * 8);
    }
    for (int i = threadIdx.x; i < N * 8; i += blockDim.x) {
        const int r = i / 8, c = i % 8;
        *reinterpret_cast<uint4*>(sB + (r / 8) * 1024 + (r % 8) * 128 + ((c ^ (r % 8)) * 16)) = *reinterpret_cast<const uint4*>(Bm + (size_t)r * K + c * 8);
    }
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar2[1])));
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
    const uint32_t a_tm = tmem + 128;                    // ring of 8-column slots (N <= 128 in modes that use it)
    if (mode == 5 && warp == 0) {
        uint32_t is_leader;
        asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(is_leader));
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        const uint64_t hi = (uint64_t)(((1024u >> 4) & 0x3FFFu) | (1u << 14) | (2u << 29)) << 32;
        const uint32_t a_lo = ((smem_u32(sA) >> 4) & 0x3FFFu) | (1u << 16);
        const uint32_t b_lo = ((smem_u32(sB) >> 4) & 0x3FFFu) | (1u << 16);
        const long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            const uint32_t j = (uint32_t)(i & 3) * 2u, acc = i > 0;
            asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.b32 p, %4, 0;\n\tsetp.ne.b32 q, %5, 0;\n\t@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                         ::"r"(tmem), "l"(hi | (a_lo + j)), "l"(hi | (b_lo + j)), "r"(idesc), "r"(acc), "r"(is_leader) : "memory");
        }
        if (is_leader) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar2[0])) : "memory");
        uint32_t ok = 0;
        while (!ok)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar2[0])) : "memory");
        if (lane == 0) cyc[blockIdx.x] = clock64() - t0;
    } else
    if (threadIdx.x == 0 || (mode == 4 && threadIdx.x == 32)) {
        const uint32_t dcol = threadIdx.x ? 128u : 0u;     // mode 4: second issuer accumulates into columns 128..
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        const uint64_t hi = (uint64_t)(((1024u >> 4) & 0x3FFFu) | (1u << 14) | (2u << 29)) << 32;
        const uint32_t a_lo = ((smem_u32(sA) >> 4) & 0x3FFFu) | (1u << 16);
        const uint32_t b_lo = ((smem_u32(sB) >> 4) & 0x3FFFu) | (1u << 16);
        const int n = mode == 0 ? 4 : iters;
        const long long t0 = clock64();
        for (int i = 0; i < n; ++i) {
            const uint32_t j = (uint32_t)(i & 3) * 2u, slot = a_tm + (uint32_t)(i & 15) * 8u, acc = i > 0;
            if (mode != 2)
                asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(slot), "l"(hi | (a_lo + j)) : "memory");
            if (mode == 0 || mode == 1)
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                             ::"r"(tmem), "r"(slot), "l"(hi | (b_lo + j)), "r"(idesc), "r"(acc) : "memory");
            if (mode == 2 || mode == 4)
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                             ::"r"(tmem + dcol), "l"(hi | (a_lo + j)), "l"(hi | (b_lo + j)), "r"(idesc), "r"(acc) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        uint32_t ok = 0;
        while (!ok)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
        if (threadIdx.x == 0) cyc[blockIdx.x] = clock64() - t0; else cyc[512 + blockIdx.x] = clock64() - t0;
    }
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (mode == 0) {
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
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256) : "memory");
}

int main() {
    const int NMAX = 256;
    __nv_bfloat16 *hA = new __nv_bfloat16[M * K], *hB = new __nv_bfloat16[NMAX * K];
    float* hD = new float[M * NMAX];
    srand(3);
    for (int i = 0; i < M * K; ++i) hA[i] = __float2bfloat16((rand() % 17 - 8) / 8.f);
    for (int i = 0; i < NMAX * K; ++i) hB[i] = __float2bfloat16((rand() % 13 - 6) / 4.f);
    __nv_bfloat16 *dA, *dB; float* dD; long long* dC;
    cudaMalloc(&dA, M * K * 2); cudaMalloc(&dB, NMAX * K * 2); cudaMalloc(&dD, M * NMAX * 4); cudaMalloc(&dC, 1024 * 8);
    cudaMemcpy(dA, hA, M * K * 2, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB, NMAX * K * 2, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    {
        const int N = 64;
        cudaMemset(dD, 0, M * NMAX * 4);
        probe<<<1, 128, 56 * 1024>>>(dA, dB, dD, N, 0, 0, dC);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("functional: %s\n", cudaGetErrorString(e)); return 1; }
        cudaMemcpy(hD, dD, M * N * 4, cudaMemcpyDeviceToHost);
        int bad = 0; double maxerr = 0;
        for (int m = 0; m < M; ++m)
            for (int n = 0; n < N; ++n) {
                double ref = 0;
                for (int k = 0; k < K; ++k) ref += (double)__bfloat162float(hA[m * K + k]) * __bfloat162float(hB[n * K + k]);
                const double err = fabs(ref - hD[m * N + n]);
                if (err > 1e-3) ++bad;
                maxerr = fmax(maxerr, err);
            }
        printf("functional cp(128x256b)+mma(TS), N=64, K=64: %s (mismatches %d, max err %g)\n", bad ? "WRONG" : "ok", bad, maxerr);
    }
    long long h[296];
    const int iters = 2048;
    printf("mode N blocks cycles/iter\n");
    for (int blocks : {1, 148, 296})
        for (int mode : {2, 5})
            for (int N : {32, 64, 128}) {
                probe<<<blocks, 128, 56 * 1024>>>(dA, dB, dD, N, mode, iters, dC);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("mode %d N %d: %s\n", mode, N, cudaGetErrorString(e)); return 1; }
                cudaMemcpy(h, dC, blocks * 8, cudaMemcpyDeviceToHost);
                double mx = 0;
                for (int b = 0; b < blocks; ++b) mx = h[b] > mx ? (double)h[b] : mx;
                printf("%s %3d %3d %8.1f\n", mode == 1 ? "cp+mmaTS" : mode == 2 ? "mmaSS   " : mode == 4 ? "2warpsSS" : mode == 5 ? "uniformSS" : "cp only ", N, blocks, mx / iters);
            }
    return 0;
}
