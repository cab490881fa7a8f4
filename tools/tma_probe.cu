// Probe: what one SM can pull through TMA, by box shape.  Persistent CTAs (1 or 2 per SM), one thread keeps DEPTH boxes in
// flight into a shared-memory ring and does nothing else (no MMA, no stores): the number is the load path alone.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tma_probe tools/tma_probe.cu && tools/tma_probe
// Shapes (NHWC bf16 activation [256][128][160][C], 671 MB at C = 64: larger than L2):
//   a  2-D map (C, pixels)            box (64, 128)        128 rows of 128 B   -- 1x1 conv over a run of pixels
//   b  4-D map (C, W, H, B)           box (64, 16, 8, 1)   128 rows of 128 B   -- the same bytes as an (8 x 16) patch
//   c  4-D map                        box (64, 18, 10, 1)  180 rows of 128 B   -- 3x3 halo box
//   d  2-D map C = 32                 box (32, 128)        128 rows of 64 B
//   e  4-D map C = 32                 box (32, 18, 10, 1)  180 rows of 64 B
//   f  2-D map (C, pixels)            box (64, 256)        256 rows of 128 B (two tiles per instruction)
//   g  3-D map (16, pixels, 5) C = 80 box (16, 128, 5)     640 rows of 32 B    -- 80 channels as five SWIZZLE_32B blocks
//   h  2-D maps C = 80: box (64,128) + box (16,128) at channel 64             -- what conv_tc_kernel does today
//   i  1-D bulk copy (cp.async.bulk, no tensor map) of 16 KB runs
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}

struct P {
    CUtensorMap m0, m1;
    int mode;          // 2, 3, 4: tensor dims of m0; 1: bulk copy; 24: two 2-D maps (h)
    int tiles;         // boxes in the whole tensor
    int tw, th, hh;    // 4-D: tiles per row / per image column / halo (0 or 1)
    int TW, TH;
    uint32_t bytes, bytes1;
    uint32_t stage_bytes;
    int depth;
    int issuers;       // warps that issue (each its own ring of `depth` stages and every issuers-th box of the CTA)
    int lanes;         // 1: the issuers are lanes 0.. of warp 0 instead of lane 0 of warps 0..
    int rr;            // > 0: ONE issuing loop run by the whole warp 0, box i issued by lane i % rr
    int twomaps;       // 1: alternate between two copies of the same tensor map
    const char* base;
};

__global__ void __launch_bounds__(128) probe(const __grid_constant__ P p, unsigned long long* cyc) {
    extern __shared__ uint8_t raw[];
    __shared__ __align__(8) uint64_t bar_all[32];
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    if (threadIdx.x == 0) {
        for (int i = 0; i < p.depth * p.issuers; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar_all[i])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (p.rr ? threadIdx.x >= 32 : p.lanes ? (int)threadIdx.x >= p.issuers : ((threadIdx.x & 31) != 0 || (int)(threadIdx.x >> 5) >= p.issuers)) return;
    const int iw = p.rr ? 0 : p.lanes ? threadIdx.x : threadIdx.x >> 5;
    uint64_t* bar = bar_all + iw * p.depth;
    smem += (size_t)iw * p.depth * p.stage_bytes;
    const long long t0 = clock64();
    int issued = 0, done = 0;
    long long c_issue = 0, c_wait = 0;
    int n_mine = 0;
    const int tstep = gridDim.x * p.issuers;
    for (int t = blockIdx.x + iw * gridDim.x; t < p.tiles; t += tstep) ++n_mine;
    int t_next = blockIdx.x + iw * gridDim.x;
    int dshift = 0; while ((1 << dshift) < p.depth) ++dshift;
    int tx = t_next % p.tw, ty = (t_next / p.tw) % p.th, tn = t_next / (p.tw * p.th);      // 4-D: advanced without divisions
    const int sx = tstep % p.tw, sy = (tstep / p.tw) % p.th, sn = tstep / (p.tw * p.th);
    while (done < n_mine) {
        while (issued < n_mine && issued - done < p.depth) {
            const int s = issued & (p.depth - 1);
            uint64_t* b = &bar[s];
            uint8_t* dst = smem + (size_t)s * p.stage_bytes;
            const int t = t_next; t_next += tstep;
            if (p.rr && (int)threadIdx.x != (issued & (p.rr - 1))) { ++issued; tx += sx; ty += sy; tn += sn; if (tx >= p.tw) { tx -= p.tw; ++ty; } if (ty >= p.th) { ty -= p.th; ++tn; } continue; }
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(p.bytes + p.bytes1) : "memory");
            if (p.mode == 2) {
                asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                             ::"r"(smem_u32(dst)), "l"(&p.m0), "r"(smem_u32(b)), "r"(0), "r"(t * p.TW) : "memory");
            } else if (p.mode == 24) {
                asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                             ::"r"(smem_u32(dst)), "l"(&p.m0), "r"(smem_u32(b)), "r"(0), "r"(t * p.TW) : "memory");
                asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                             ::"r"(smem_u32(dst + 16384)), "l"(&p.m1), "r"(smem_u32(b)), "r"(64), "r"(t * p.TW) : "memory");
            } else if (p.mode == 3) {
                asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                             ::"r"(smem_u32(dst)), "l"(&p.m0), "r"(smem_u32(b)), "r"(0), "r"(t * p.TW), "r"(0) : "memory");
            } else if (p.mode == 4) {
                const int n = tn;
                const long long ci0 = clock64();
                asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                             ::"r"(smem_u32(dst)), "l"((p.twomaps && (issued & 1)) ? &p.m1 : &p.m0), "r"(smem_u32(b)), "r"(0), "r"(tx * p.TW - p.hh), "r"(ty * p.TH - p.hh), "r"(n) : "memory");
                c_issue += clock64() - ci0;
                tx += sx; ty += sy; tn += sn;
                if (tx >= p.tw) { tx -= p.tw; ++ty; }
                if (ty >= p.th) { ty -= p.th; ++tn; }
            } else {
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(smem_u32(dst)), "l"(p.base + (size_t)t * p.bytes), "r"(p.bytes), "r"(smem_u32(b)) : "memory");
            }
            ++issued;
        }
        const int s = done & (p.depth - 1);
        const long long cw0 = clock64();
        while (!try_wait(&bar[s], (uint32_t)(done >> dshift) & 1u)) {}
        c_wait += clock64() - cw0;
        ++done;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) { cyc[0] = (unsigned long long)(clock64() - t0); cyc[1] = c_issue; cyc[2] = c_wait; }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaFree(0));
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q));
    EncodeTiledFn enc = (EncodeTiledFn)f;
    const int B = 256, H = 128, W = 160;
    const size_t npix = (size_t)B * H * W;
    char* buf;
    CK(cudaMalloc(&buf, npix * 80 * 2));
    CK(cudaMemset(buf, 1, npix * 80 * 2));
    unsigned long long* cyc;
    CK(cudaMalloc(&cyc, 24));
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    const cuuint32_t es[5] = {1, 1, 1, 1, 1};
    for (const char* v = "abcdefghi"; *v; ++v) {
        for (int ctas = 1; ctas <= 2; ++ctas) {
            for (int depth = 2; depth <= 8; depth *= 2)
            for (int iss = 1; iss <= 10; ++iss) {      // 8, 9, 10: round robin over 2 / 4 / 8 lanes of one issuing loop       // 6, 7: 2 / 4 issuing LANES of warp 0       // 1, 2, 4 issuing warps; 5 = one issuer alternating between two copies of the map
                if (iss == 3) continue;
                if (iss > 1 && !(*v == 'b' || *v == 'c' || *v == 'e' || (*v == 'a' && iss != 5))) continue;
                P p;
                memset(&p, 0, sizeof(p)); p.tw = p.th = 1;
                p.depth = depth; p.bytes1 = 0; p.base = buf; p.issuers = iss == 5 ? 1 : iss == 6 ? 2 : iss == 7 ? 4 : iss; p.twomaps = iss == 5; p.lanes = iss == 6 || iss == 7; if (iss >= 8) { p.issuers = 1; p.rr = 2 << (iss - 8); }
                if (p.issuers * depth > 32) continue;
                int C = 64; CUresult r = CUDA_SUCCESS;
                if (*v == 'a' || *v == 'd' || *v == 'f') {
                    C = *v == 'd' ? 32 : 64;
                    const int rows = *v == 'f' ? 256 : 128;
                    cuuint64_t dims[2] = {(cuuint64_t)C, npix}, str[1] = {(cuuint64_t)C * 2};
                    cuuint32_t box[2] = {(cuuint32_t)C, (cuuint32_t)rows};
                    r = enc(&p.m0, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, buf, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            C == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                    p.mode = 2; p.TW = rows; p.tiles = (int)(npix / rows); p.bytes = rows * C * 2;
                } else if (*v == 'b' || *v == 'c' || *v == 'e') {
                    C = *v == 'e' ? 32 : 64;
                    const int hh = *v == 'b' ? 0 : 1;
                    cuuint64_t dims[4] = {(cuuint64_t)C, W, H, B}, str[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
                    cuuint32_t box[4] = {(cuuint32_t)C, (cuuint32_t)(16 + 2 * hh), (cuuint32_t)(8 + 2 * hh), 1};
                    r = enc(&p.m0, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, buf, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            C == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                    p.m1 = p.m0;
                    p.mode = 4; p.TW = 16; p.TH = 8; p.hh = hh; p.tw = W / 16; p.th = H / 8; p.tiles = p.tw * p.th * B;
                    p.bytes = (16 + 2 * hh) * (8 + 2 * hh) * C * 2;
                } else if (*v == 'g') {
                    C = 80;
                    cuuint64_t dims[3] = {16, npix, 5}, str[2] = {160, 32};
                    cuuint32_t box[3] = {16, 128, 5};
                    r = enc(&p.m0, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, buf, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                    p.mode = 3; p.TW = 128; p.tiles = (int)(npix / 128); p.bytes = 128 * 160;
                } else if (*v == 'h') {
                    C = 80;
                    cuuint64_t dims[2] = {80, npix}, str[1] = {160};
                    cuuint32_t box0[2] = {64, 128}, box1[2] = {16, 128};
                    r = enc(&p.m0, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, buf, dims, str, box0, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                    if (r == CUDA_SUCCESS)
                        r = enc(&p.m1, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, buf, dims, str, box1, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                    p.mode = 24; p.TW = 128; p.tiles = (int)(npix / 128); p.bytes = 128 * 128; p.bytes1 = 128 * 32;
                } else {
                    p.mode = 1; p.tiles = (int)(npix / 128); p.bytes = 16384;
                }
                if (r != CUDA_SUCCESS) { printf("%c: encode failed %d\n", *v, (int)r); continue; }
                p.stage_bytes = (p.bytes + p.bytes1 + 1023) & ~1023u;
                const size_t smem = (size_t)p.stage_bytes * depth * p.issuers + 1024;
                if (smem * ctas > 220 * 1024) continue;
                cudaEvent_t e0, e1;
                CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
                probe<<<sms * ctas, 128, smem>>>(p, cyc);
                CK(cudaDeviceSynchronize());
                CK(cudaEventRecord(e0));
                probe<<<sms * ctas, 128, smem>>>(p, cyc);
                CK(cudaEventRecord(e1));
                CK(cudaDeviceSynchronize());
                float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
                unsigned long long hcs[3]; CK(cudaMemcpy(hcs, cyc, 24, cudaMemcpyDeviceToHost)); const unsigned long long hc = hcs[0];
                const double total = (double)p.tiles * (p.bytes + p.bytes1);
                const double rows_per_box = *v == 'g' ? 640 : *v == 'h' ? 256 : (double)(p.bytes) / (C * 2);
                printf("%c ctas/SM=%d issuers=%d%s depth=%d  %.3f ms  %.2f TB/s  %.1f B/clk/SM  %.2f clk/row (cta0 %llu clk, in TMA issue %llu, in wait %llu)\n", *v, ctas, p.issuers, p.twomaps ? "(2 maps)" : p.lanes ? "(lanes)" : p.rr == 2 ? "(rr2)" : p.rr == 4 ? "(rr4)" : p.rr ? "(rr8)" : "", depth, ms, total / ms / 1e9,
                       total / sms / (double)hc, (double)hc / ((double)p.tiles / sms * rows_per_box), hc, hcs[1], hcs[2]);
            }
        }
    }
    return 0;
}
