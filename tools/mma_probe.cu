// Micro-benchmark: cycles per tcgen05.mma (kind::f16, bf16 in, fp32 accumulate) as a function of the N extent and of
// where the A operand lives (shared memory "SS" vs tensor memory "TS").  Operand contents are irrelevant (zeros).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mma_probe tools/mma_probe.cu && tools/mma_probe
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128) probe(int mode, int M, int N, int iters, int a_step, long long* out) {
    extern __shared__ uint8_t raw[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    if (threadIdx.x == 0) {
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        // K-major, SWIZZLE_128B: rows of 128 bytes, 8-row groups 1024 B apart
        const uint64_t hi = (uint64_t)(((1024u >> 4) & 0x3FFFu) | (1u << 14) | (2u << 29)) << 32;
        const uint32_t a_lo = ((smem_u32(smem) >> 4) & 0x3FFFu) | (1u << 16);
        const uint32_t b_lo = (((smem_u32(smem) + 16384) >> 4) & 0x3FFFu) | (1u << 16);
        const uint32_t a_tm = tmem + 256;          // TS: A tile [M lanes][8 columns per K16 step]
        long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            const uint32_t j = (uint32_t)(i & 3) * 2u;
            const uint32_t acc = i > 0;
            if (mode == 0) {
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                             "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                             ::"r"(tmem), "l"(hi | (a_lo + j + (uint32_t)((i >> 2) & 1) * (uint32_t)a_step)), "l"(hi | (b_lo + j)), "r"(idesc), "r"(acc) : "memory");
            } else {
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                             "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                             ::"r"(tmem), "r"(a_tm + (uint32_t)(i & 3) * 8u), "l"(hi | (b_lo + j)), "r"(idesc), "r"(acc) : "memory");
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        uint32_t ok = 0;
        while (!ok) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
        }
        long long t1 = clock64();
        out[blockIdx.x] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

int main() {
    long long* d;
    cudaMalloc(&d, 1024 * sizeof(long long));
    long long h[1024];
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    const int iters = 2048;
    printf("mode M N blocks cycles/MMA floor(=max(M,128)*N/256) ratio\n");
    for (int blocks : {1, 148}) {
        for (int mode : {0, 1}) {
            for (int M : {128, 64}) {
                for (int N : {16, 32, 64, 80, 128, 256}) {
                    if (M == 64 && N % 8) continue;
                    if (mode == 1 && N > 256) continue;
                    probe<<<blocks, 128, 50 * 1024>>>(mode, M, N, iters, 0, d);
                    cudaError_t e = cudaDeviceSynchronize();
                    if (e != cudaSuccess) { printf("mode %d M %d N %d: %s\n", mode, M, N, cudaGetErrorString(e)); return 1; }
                    cudaMemcpy(h, d, blocks * sizeof(long long), cudaMemcpyDeviceToHost);
                    double mx = 0;
                    for (int b = 0; b < blocks; ++b) mx = h[b] > mx ? (double)h[b] : mx;
                    const double floor_ = (double)(M > 128 ? M : 128) * N / 256.0;
                    printf("%s %3d %3d %3d %8.1f %6.1f %5.2f\n", mode ? "TS" : "SS", M, N, blocks, mx / iters, floor_, mx / iters / floor_);
                }
            }
        }
    }
    return 0;
}
