// Bandwidth-shaped layer kernels of the detect path (sm_100a): stem conv on raw uint8 frames, SPPF pooling,
// nearest-2x upsample / concat slice copy, stand-alone preprocess.  All NHWC bf16, 16-byte vector accesses.
#include "common.cuh"

void b2_count_launch(int n);

namespace {

// ------------------------------------------------------------------------------------------------
// Stem: LetterBox pad (114) + BGR->RGB + /255 (engine/predictor.py:152-175, data/augment.py:1717-1728)
// fused with model.0 = Conv(3 -> C0, k3, s2) + SiLU.  fp32 weights on exact uint8 pixel values.
// One thread per output pixel, all C0 channels; weights broadcast from shared memory.
// ------------------------------------------------------------------------------------------------
template <int C0, typename Loader>
__global__ void __launch_bounds__(256) stem_kernel(Loader ld, int H, int W, int Ho, int Wo, const float* __restrict__ w,
                                                   const float* __restrict__ bias, float in_scale,
                                                   __nv_bfloat16* __restrict__ out, int out_cstride, int out_coff) {
    __shared__ __align__(16) float ws[27 * C0];
    __shared__ float bs[C0];
    // w: [C0][3][3][3] (o, kh, kw, c_rgb)  ->  ws[(kh*3+kw)*3+c][o]
    for (int i = threadIdx.x; i < 27 * C0; i += blockDim.x) {
        const int o = i / 27, t = i % 27;
        ws[t * C0 + o] = w[i];
    }
    for (int i = threadIdx.x; i < C0; i += blockDim.x) bs[i] = bias[i];
    __syncthreads();
    const int wo = blockIdx.x * 32 + (threadIdx.x & 31);
    const int ho = blockIdx.y * 8 + (threadIdx.x >> 5);
    const int b = blockIdx.z;
    if (wo >= Wo || ho >= Ho) return;
    float acc[C0];
#pragma unroll
    for (int o = 0; o < C0; ++o) acc[o] = 0.f;
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
        const int y = 2 * ho + kh - 1;
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
            const int x = 2 * wo + kw - 1;
            float rgb[3];
            ld.load(b, y, x, H, W, rgb);
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float v = rgb[c];
                const float4* wp = reinterpret_cast<const float4*>(&ws[((kh * 3 + kw) * 3 + c) * C0]);
#pragma unroll
                for (int o4 = 0; o4 < C0 / 4; ++o4) {
                    const float4 ww = wp[o4];
                    acc[4 * o4 + 0] = fmaf(v, ww.x, acc[4 * o4 + 0]);
                    acc[4 * o4 + 1] = fmaf(v, ww.y, acc[4 * o4 + 1]);
                    acc[4 * o4 + 2] = fmaf(v, ww.z, acc[4 * o4 + 2]);
                    acc[4 * o4 + 3] = fmaf(v, ww.w, acc[4 * o4 + 3]);
                }
            }
        }
    }
    __nv_bfloat16* op = out + (((size_t)b * Ho + ho) * Wo + wo) * out_cstride + out_coff;
#pragma unroll
    for (int o8 = 0; o8 < C0 / 8; ++o8) {
        uint4 v;
        uint32_t* vp = &v.x;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int o = 8 * o8 + 2 * i;
            vp[i] = pack_bf16x2(silu_f(fmaf(acc[o], in_scale, bs[o])), silu_f(fmaf(acc[o + 1], in_scale, bs[o + 1])));
        }
        *reinterpret_cast<uint4*>(op + 8 * o8) = v;
    }
}

struct LoadU8 {   // [B][src_h][src_w][3] uint8 BGR placed at (pad_top, pad_left) of the canvas, border 114, outside canvas 0
    const uint8_t* p; int sh, sw, pt, pl;
    __device__ __forceinline__ void load(int b, int y, int x, int H, int W, float (&rgb)[3]) const {
        if (y < 0 || x < 0 || y >= H || x >= W) { rgb[0] = rgb[1] = rgb[2] = 0.f; return; }
        const int fy = y - pt, fx = x - pl;
        if (fy < 0 || fx < 0 || fy >= sh || fx >= sw) { rgb[0] = rgb[1] = rgb[2] = 114.f; return; }
        const uint8_t* q = p + (((size_t)b * sh + fy) * sw + fx) * 3;
        rgb[2] = (float)__ldg(q); rgb[1] = (float)__ldg(q + 1); rgb[0] = (float)__ldg(q + 2);
    }
};
template <typename T>
struct LoadPlanar {   // [B][3][H][W] RGB in [0,1]
    const T* p;
    __device__ __forceinline__ void load(int b, int y, int x, int H, int W, float (&rgb)[3]) const {
        if (y < 0 || x < 0 || y >= H || x >= W) { rgb[0] = rgb[1] = rgb[2] = 0.f; return; }
        const size_t plane = (size_t)H * W;
        const T* q = p + (size_t)b * 3 * plane + (size_t)y * W + x;
        rgb[0] = (float)q[0]; rgb[1] = (float)q[plane]; rgb[2] = (float)q[2 * plane];
    }
};

template <typename Loader>
int launch_stem(Loader ld, int B, int H, int W, const float* w, const float* bias, int C0, float in_scale,
                void* out, int out_cstride, int out_coff, cudaStream_t st) {
    const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
    dim3 grid(b2_ceil_div(Wo, 32), b2_ceil_div(Ho, 8), B), block(256);
    __nv_bfloat16* o = (__nv_bfloat16*)out;
#define B2_STEM_CASE(C) case C: stem_kernel<C, Loader><<<grid, block, 0, st>>>(ld, H, W, Ho, Wo, w, bias, in_scale, o, out_cstride, out_coff); break;
    switch (C0) {
        B2_STEM_CASE(16) B2_STEM_CASE(24) B2_STEM_CASE(32) B2_STEM_CASE(40) B2_STEM_CASE(48) B2_STEM_CASE(64) B2_STEM_CASE(80)
        default: b2_set_error("stem: C0=%d not instantiated", C0); return B2_ERR_UNSUPPORTED;
    }
#undef B2_STEM_CASE
    B2_CUDA(cudaGetLastError());
    b2_count_launch(1);
    return B2_OK;
}

// ------------------------------------------------------------------------------------------------
// SPPF pooling: three chained MaxPool2d(5,1,2) == windows of 5, 9, 13 with -inf padding (block.py:237-241)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 max_bf16x8(uint4 a, uint4 b) {
    uint4 r;
    const __nv_bfloat162* pa = reinterpret_cast<const __nv_bfloat162*>(&a);
    const __nv_bfloat162* pb = reinterpret_cast<const __nv_bfloat162*>(&b);
    __nv_bfloat162* pr = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) pr[i] = __hmax2(pa[i], pb[i]);
    return r;
}

__global__ void __launch_bounds__(256) sppf_pool_kernel(__nv_bfloat16* __restrict__ buf, int B, int H, int W, int cstride, int coff, int C) {
    const int chunks = C / 8;
    const size_t total = (size_t)B * H * W * chunks;
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int ck = (int)(idx % chunks);
    size_t pix = idx / chunks;
    const int x = (int)(pix % W), y = (int)((pix / W) % H), b = (int)(pix / ((size_t)W * H));
    const uint32_t ninf = 0xFF80FF80u;   // bf16 -inf pair
    uint4 m5 = make_uint4(ninf, ninf, ninf, ninf), m9 = m5, m13 = m5;
    for (int dy = -6; dy <= 6; ++dy) {
        const int yy = y + dy;
        if (yy < 0 || yy >= H) continue;
        for (int dx = -6; dx <= 6; ++dx) {
            const int xx = x + dx;
            if (xx < 0 || xx >= W) continue;
            const uint4 v = *reinterpret_cast<const uint4*>(buf + (((size_t)b * H + yy) * W + xx) * cstride + coff + ck * 8);
            m13 = max_bf16x8(m13, v);
            if (abs(dy) <= 4 && abs(dx) <= 4) m9 = max_bf16x8(m9, v);
            if (abs(dy) <= 2 && abs(dx) <= 2) m5 = max_bf16x8(m5, v);
        }
    }
    __nv_bfloat16* o = buf + (((size_t)b * H + y) * W + x) * cstride + coff + ck * 8;
    *reinterpret_cast<uint4*>(o + C) = m5;
    *reinterpret_cast<uint4*>(o + 2 * C) = m9;
    *reinterpret_cast<uint4*>(o + 3 * C) = m13;
}

// ------------------------------------------------------------------------------------------------
// nearest upsample (x1 / x2) of a channel slice into a channel slice (Upsample + Concat)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) upsample_slice_kernel(const __nv_bfloat16* __restrict__ in, int B, int H, int W, int ics, int ico,
                                                             int C, int scale, __nv_bfloat16* __restrict__ out, int ocs, int oco) {
    const int chunks = C / 8;
    const int Ho = H * scale, Wo = W * scale;
    const size_t total = (size_t)B * Ho * Wo * chunks;
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int ck = (int)(idx % chunks);
    size_t pix = idx / chunks;
    const int x = (int)(pix % Wo), y = (int)((pix / Wo) % Ho), b = (int)(pix / ((size_t)Wo * Ho));
    const uint4 v = *reinterpret_cast<const uint4*>(in + (((size_t)b * H + y / scale) * W + x / scale) * ics + ico + ck * 8);
    *reinterpret_cast<uint4*>(out + (((size_t)b * Ho + y) * Wo + x) * ocs + oco + ck * 8) = v;
}

__global__ void __launch_bounds__(256) preprocess_u8_kernel(const uint8_t* __restrict__ f, int B, int sh, int sw, int H, int W, int pt, int pl,
                                                            float* __restrict__ out) {
    const size_t total = (size_t)B * H * W;
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int x = (int)(idx % W), y = (int)((idx / W) % H), b = (int)(idx / ((size_t)W * H));
    const int fy = y - pt, fx = x - pl;
    float r = 114.f, g = 114.f, bl = 114.f;
    if (fy >= 0 && fx >= 0 && fy < sh && fx < sw) {
        const uint8_t* q = f + (((size_t)b * sh + fy) * sw + fx) * 3;
        bl = (float)q[0]; g = (float)q[1]; r = (float)q[2];
    }
    const size_t plane = (size_t)H * W;
    float* o = out + (size_t)b * 3 * plane + (size_t)y * W + x;
    o[0] = r / 255.f; o[plane] = g / 255.f; o[2 * plane] = bl / 255.f;
}

// cv2.resize(INTER_LINEAR) for uint8: 11-bit fixed-point coefficients (INTER_RESIZE_COEF_BITS = 11),
// coefficients rounded with saturate_cast<short>(rint(f * 2048)), vertical pass result
// (x >> 4 * beta >> 16 ...) reproduced with the 8u formula  ((b0*(S0>>4))>>16 + (b1*(S1>>4))>>16 + 2) >> 2.
__global__ void __launch_bounds__(256) resize_bilinear_u8_kernel(const uint8_t* __restrict__ src, int B, int sh, int sw,
                                                                 uint8_t* __restrict__ dst, int dh, int dw, double fy, double fx) {
    const size_t total = (size_t)B * dh * dw;
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int x = (int)(idx % dw), y = (int)((idx / dw) % dh), b = (int)(idx / ((size_t)dw * dh));
    float sxf = (float)((x + 0.5) * fx - 0.5);
    int sx = (int)floorf(sxf); sxf -= sx;
    if (sx < 0) { sxf = 0; sx = 0; }
    if (sx >= sw - 1) { sxf = 0; sx = sw - 1; }
    float syf = (float)((y + 0.5) * fy - 0.5);
    int sy = (int)floorf(syf); syf -= sy;
    int sy0 = min(max(sy, 0), sh - 1), sy1 = min(max(sy + 1, 0), sh - 1);
    const int a0 = __float2int_rn((1.f - sxf) * 2048.f), a1 = __float2int_rn(sxf * 2048.f);
    const int b0 = __float2int_rn((1.f - syf) * 2048.f), b1 = __float2int_rn(syf * 2048.f);
    const int sx1 = min(sx + 1, sw - 1);
    const uint8_t* r0 = src + ((size_t)b * sh + sy0) * sw * 3;
    const uint8_t* r1 = src + ((size_t)b * sh + sy1) * sw * 3;
    uint8_t* o = dst + idx * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const int S0 = r0[sx * 3 + c] * a0 + r0[sx1 * 3 + c] * a1;
        const int S1 = r1[sx * 3 + c] * a0 + r1[sx1 * 3 + c] * a1;
        const int v = (((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2;
        o[c] = (uint8_t)min(max(v, 0), 255);
    }
}

}  // namespace

extern "C" int b2_stem_u8(const uint8_t* frames, int B, int src_h, int src_w, int H, int W, int pad_top, int pad_left,
                          const float* w, const float* bias, int C0, void* out, int out_cstride, int out_coff, void* stream) {
    B2_REQUIRE(frames && w && bias && out, "stem: null pointer");
    B2_REQUIRE(C0 % 8 == 0 && out_cstride % 8 == 0 && out_coff % 8 == 0, "stem: channel counts must be multiples of 8");
    B2_REQUIRE(pad_top >= 0 && pad_left >= 0 && pad_top + src_h <= H && pad_left + src_w <= W, "stem: frame does not fit the canvas");
    LoadU8 ld{frames, src_h, src_w, pad_top, pad_left};
    return launch_stem(ld, B, H, W, w, bias, C0, 1.f / 255.f, out, out_cstride, out_coff, (cudaStream_t)stream);
}

extern "C" int b2_stem_f32(const void* bchw, int dtype, int B, int H, int W, const float* w, const float* bias, int C0,
                           void* out, int out_cstride, int out_coff, void* stream) {
    B2_REQUIRE(bchw && w && bias && out, "stem: null pointer");
    B2_REQUIRE(C0 % 8 == 0 && out_cstride % 8 == 0 && out_coff % 8 == 0, "stem: channel counts must be multiples of 8");
    if (dtype == 0) return launch_stem(LoadPlanar<float>{(const float*)bchw}, B, H, W, w, bias, C0, 1.f, out, out_cstride, out_coff, (cudaStream_t)stream);
    if (dtype == 1) return launch_stem(LoadPlanar<__nv_bfloat16>{(const __nv_bfloat16*)bchw}, B, H, W, w, bias, C0, 1.f, out, out_cstride, out_coff, (cudaStream_t)stream);
    b2_set_error("stem: dtype %d unsupported", dtype);
    return B2_ERR_UNSUPPORTED;
}

extern "C" int b2_sppf_pool(void* buf, int B, int H, int W, int cstride, int coff, int C, void* stream) {
    B2_REQUIRE(C % 8 == 0 && cstride % 8 == 0 && coff % 8 == 0 && coff + 4 * C <= cstride, "sppf_pool: bad channel layout");
    const size_t total = (size_t)B * H * W * (C / 8);
    sppf_pool_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>((__nv_bfloat16*)buf, B, H, W, cstride, coff, C);
    B2_CUDA(cudaGetLastError());
    b2_count_launch(1);
    return B2_OK;
}

extern "C" int b2_upsample_slice(const void* in, int B, int H, int W, int in_cstride, int in_coff, int C, int scale,
                                 void* out, int out_cstride, int out_coff, void* stream) {
    B2_REQUIRE(scale == 1 || scale == 2, "upsample: scale must be 1 or 2");
    B2_REQUIRE(C % 8 == 0 && in_cstride % 8 == 0 && in_coff % 8 == 0 && out_cstride % 8 == 0 && out_coff % 8 == 0, "upsample: channels must be multiples of 8");
    const size_t total = (size_t)B * H * scale * W * scale * (C / 8);
    upsample_slice_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)in, B, H, W, in_cstride, in_coff, C, scale, (__nv_bfloat16*)out, out_cstride, out_coff);
    B2_CUDA(cudaGetLastError());
    b2_count_launch(1);
    return B2_OK;
}

extern "C" int b2_preprocess_u8(const uint8_t* frames, int B, int src_h, int src_w, int H, int W, int pad_top, int pad_left,
                                float* out_bchw, void* stream) {
    B2_REQUIRE(pad_top >= 0 && pad_left >= 0 && pad_top + src_h <= H && pad_left + src_w <= W, "preprocess: frame does not fit the canvas");
    const size_t total = (size_t)B * H * W;
    preprocess_u8_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(frames, B, src_h, src_w, H, W, pad_top, pad_left, out_bchw);
    B2_CUDA(cudaGetLastError());
    b2_count_launch(1);
    return B2_OK;
}

extern "C" int b2_resize_bilinear_u8(const uint8_t* src, int B, int sh, int sw, uint8_t* dst, int dh, int dw, void* stream) {
    B2_REQUIRE(sh > 0 && sw > 0 && dh > 0 && dw > 0, "resize: bad shape");
    const size_t total = (size_t)B * dh * dw;
    resize_bilinear_u8_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        src, B, sh, sw, dst, dh, dw, (double)sh / dh, (double)sw / dw);
    B2_CUDA(cudaGetLastError());
    b2_count_launch(1);
    return B2_OK;
}
