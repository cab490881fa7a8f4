// Launch-plan executor for the YOLOv8-P2 forward (sm_100a) + C-ABI plumbing shared by all entry points.
//
// Replaces the python layer loop BaseModel._predict_once (ultralytics/nn/tasks.py:159-188) over the fused
// graph that AutoBackend(fuse=True) runs (nn/autobackend.py:196-219, :608-637).  The host (engine.py) lowers the
// resolved YAML graph to a flat int32 program over NHWC bf16 buffers; concat/chunk are channel offsets, so
// the program consists only of: stem, conv (tcgen05 implicit GEMM), SPPF pooling, nearest-upsample slice copy.
// All tensor maps are built once at create time; the layer launches are replayed from a CUDA graph.
#include "common.cuh"

#include <cuda_profiler_api.h>

#include <stdarg.h>

#include <atomic>
#include <new>
#include <vector>

// ---- from conv_tc.cu ----
size_t b2_conv_launch_size();
int b2_conv_prepare(void* storage, const void* in, int B, int H, int W, int in_cstride, int in_coff, int Cin,
                    const void* w, const float* bias, int Cout, int ksize, int stride, int act,
                    void* out, int out_cstride, int out_coff, const void* residual, int res_cstride, int res_coff);
struct B2ConvSrc { const void* ptr; int cstride, coff, C, up; };
int b2_conv_prepare_ms(void* storage, const B2ConvSrc* srcs, int nsrc, int B, int H, int W,
                       const void* w, const float* bias, int Cout, int ksize, int stride, int act,
                       void* out, int out_cstride, int out_coff, const void* residual, int res_cstride, int res_coff);
struct B2ConvChain { const void* w2; const float* bias2; int Cout2, act2; const void* xsrc; int x_cstride, x_coff, xC; };
int b2_conv_prepare_chain(void* storage, const B2ConvSrc* srcs, int nsrc, int B, int H, int W,
                          const void* w, const float* bias, int Cout, int ksize, int stride, int act,
                          void* out, int out_cstride, int out_coff, const void* residual, int res_cstride, int res_coff,
                          const B2ConvChain* chain);
int b2_conv_set_head_epilogue(void* storage, int epi, float* out_f32);
int b2_conv_launch(const void* storage, cudaStream_t stream);

// ------------------------------------------------------------------------------------------------
// library-wide plumbing
// ------------------------------------------------------------------------------------------------
static thread_local char g_err[1024] = "";
static std::atomic<long long> g_launches{0};

void b2_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
void b2_count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
int b2_num_sms() {
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
    }
    return sms;
}
extern "C" const char* b2_last_error(void) { return g_err; }
extern "C" int b2_version(void) { return 100; }
extern "C" long long b2_launch_count(void) { return g_launches.load(); }

// ------------------------------------------------------------------------------------------------
// engine
// ------------------------------------------------------------------------------------------------
namespace {
constexpr int kMagic = 0xB2D7;
constexpr int kOpWords = 28;
enum Op { OP_STEM = 1, OP_CONV = 2, OP_POOL = 3, OP_UP = 4 };

struct Buf { int h, w, c; size_t off; };
struct Step {
    int op;
    int a[kOpWords - 1];
    std::vector<unsigned char> conv;   // B2ConvLaunch storage (64B aligned inside)
    void* conv_ptr() { return (void*)(((uintptr_t)conv.data() + 63) & ~(uintptr_t)63); }
};
}  // namespace

struct b2_engine {
    int B, H, W, nc, lstride, n_levels;
    std::vector<Buf> bufs;
    std::vector<Step> steps;
    std::vector<int> level_buf, level_stride, level_dist, level_cls;   // dist/cls: fused-head fp32 buffers (-1: unfused)
    char* arena = nullptr;
    size_t arena_bytes = 0, weights_off = 0;
    cudaGraphExec_t graph = nullptr;
    cudaStream_t graph_stream = nullptr;
    bool use_graph = true;
    int n_launch = 0;
    char* buf_ptr(int i) { return arena + bufs[i].off; }
};

static int run_step(b2_engine* e, Step& s, cudaStream_t st) {
    const int* a = s.a;
    switch (s.op) {
        case OP_CONV: return b2_conv_launch(s.conv_ptr(), st);
        case OP_POOL: { const Buf& b = e->bufs[a[0]]; return b2_sppf_pool(e->buf_ptr(a[0]), e->B, b.h, b.w, b.c, a[1], a[2], st); }
        case OP_UP: {
            const Buf& bi = e->bufs[a[0]]; const Buf& bo = e->bufs[a[4]];
            return b2_upsample_slice(e->buf_ptr(a[0]), e->B, bi.h, bi.w, bi.c, a[1], a[2], a[3], e->buf_ptr(a[4]), bo.c, a[5], st);
        }
        default: b2_set_error("engine: bad opcode %d", s.op); return B2_ERR_STATE;
    }
}

extern "C" int b2_engine_create(const int32_t* plan, int plan_words, const void* weights_host, size_t weight_bytes,
                                int B, int H, int W, b2_engine_t** out) {
    B2_REQUIRE(plan && weights_host && out, "engine_create: null pointer");
    B2_REQUIRE(plan_words >= 6 && plan[0] == kMagic, "engine_create: bad plan header");
    B2_REQUIRE(B >= 1 && H % 32 == 0 && W % 32 == 0 && H > 0 && W > 0, "engine_create: H and W must be positive multiples of 32 (got %dx%d)", H, W);
    const int n_bufs = plan[1], n_ops = plan[2], n_levels = plan[3];
    B2_REQUIRE(plan_words == 6 + 3 * n_bufs + 4 * n_levels + kOpWords * n_ops, "engine_create: plan length mismatch");
    b2_engine* e = new (std::nothrow) b2_engine();
    if (!e) { b2_set_error("out of host memory"); return B2_ERR_STATE; }
    e->B = B; e->H = H; e->W = W; e->n_levels = n_levels; e->nc = plan[4]; e->lstride = plan[5];
    const int32_t* p = plan + 6;
    size_t off = 0;
    auto up = [](size_t v) { return (v + 1023) & ~(size_t)1023; };
    for (int i = 0; i < n_bufs; ++i, p += 3) {
        Buf b{p[0], p[1], p[2], off};
        off += up((size_t)B * b.h * b.w * b.c * 2);
        e->bufs.push_back(b);
    }
    for (int l = 0; l < n_levels; ++l, p += 4) { e->level_buf.push_back(p[0]); e->level_stride.push_back(p[1]); e->level_dist.push_back(p[2]); e->level_cls.push_back(p[3]); }
    e->weights_off = off;
    off += up(weight_bytes);
    e->arena_bytes = off;
    cudaError_t ce = cudaMalloc((void**)&e->arena, off);
    if (ce != cudaSuccess) { b2_set_error("engine_create: cudaMalloc(%zu bytes) failed: %s", off, cudaGetErrorString(ce)); delete e; return B2_ERR_CUDA; }
    int rc = B2_OK;
    auto fail = [&](int code) { cudaFree(e->arena); delete e; return code; };
    if (cudaMemset(e->arena, 0, off) != cudaSuccess ||
        cudaMemcpy(e->arena + e->weights_off, weights_host, weight_bytes, cudaMemcpyHostToDevice) != cudaSuccess) {
        b2_set_error("engine_create: arena initialisation failed: %s", cudaGetErrorString(cudaGetLastError()));
        return fail(B2_ERR_CUDA);
    }
    char* wbase = e->arena + e->weights_off;
    e->steps.resize(n_ops);
    for (int i = 0; i < n_ops; ++i, p += kOpWords) {
        Step& s = e->steps[i];
        s.op = p[0];
        for (int k = 0; k < kOpWords - 1; ++k) s.a[k] = p[1 + k];
        const int* a = s.a;
        auto buf_ok = [&](int id) { return id >= 0 && id < n_bufs; };
        if (s.op == OP_STEM) {
            if (i != 0 || !buf_ok(a[0])) { b2_set_error("engine_create: stem must be op 0 with a valid buffer"); return fail(B2_ERR_ARG); }
        } else if (s.op == OP_CONV) {
            // a: 0 in buf, 1 in coff, 2 Cin, 3 out buf, 4 out coff, 5 Cout, 6 k, 7 stride, 8 act, 9 res buf, 10 res coff,
            //    11 weights, 12 bias, 13 second input buf (-1: none), 14 its coff, 15 its channels, 16 up of input 0, 17 up of input 1,
            //    18 head epilogue; chained 1x1 conv (conv_tc.cu ConvParams::chain): 19 flag, 20 weights, 21 bias, 22 Cout2, 23 act2,
            //    24 extra-source buf (-1: none), 25 its coff, 26 its channels.  With a chain, 3 / 4 (and 18) describe the CHAINED
            //    conv's output and 5 is the main conv's Cout.
            const bool two = a[13] >= 0;
            const bool chained = a[19] != 0;
            const int c_final = chained ? a[22] : a[5];
            if (!buf_ok(a[0]) || !buf_ok(a[3]) || (a[9] >= 0 && !buf_ok(a[9])) || (two && !buf_ok(a[13]))) { b2_set_error("engine_create: op %d: bad buffer id", i); return fail(B2_ERR_ARG); }
            const Buf& bi = e->bufs[a[0]]; const Buf& bo = e->bufs[a[3]];
            const int up0 = a[16] == 2 ? 2 : 1, up1 = a[17] == 2 ? 2 : 1;
            const int Hin = bi.h * up0, Win = bi.w * up0;
            B2ConvSrc srcs[2] = {{e->buf_ptr(a[0]), bi.c, a[1], a[2], up0}, {nullptr, 0, 0, 0, 1}};
            if (two) {
                const Buf& b2 = e->bufs[a[13]];
                srcs[1] = B2ConvSrc{e->buf_ptr(a[13]), b2.c, a[14], a[15], up1};
                if (b2.h * up1 != Hin || b2.w * up1 != Win || a[14] + a[15] > b2.c) { b2_set_error("engine_create: op %d: second input geometry mismatch", i); return fail(B2_ERR_ARG); }
            }
            s.conv.resize(b2_conv_launch_size() + 64);
            B2ConvChain ch{};
            if (chained) {
                if (a[24] >= 0 && (!buf_ok(a[24]) || a[25] + a[26] > e->bufs[a[24]].c)) { b2_set_error("engine_create: op %d: bad extra source of the chained conv", i); return fail(B2_ERR_ARG); }
                ch = B2ConvChain{wbase + (size_t)(uint32_t)a[20], (const float*)(wbase + (size_t)(uint32_t)a[21]), a[22], a[23],
                                 a[24] >= 0 ? e->buf_ptr(a[24]) : nullptr, a[24] >= 0 ? e->bufs[a[24]].c : 0, a[25], a[26]};
            }
            rc = b2_conv_prepare_chain(s.conv_ptr(), srcs, two ? 2 : 1, B, Hin, Win,
                                       wbase + (size_t)(uint32_t)a[11], (const float*)(wbase + (size_t)(uint32_t)a[12]), a[5], a[6], a[7], a[8],
                                       e->buf_ptr(a[3]), a[18] ? 16 : bo.c, a[18] ? 0 : a[4],   // head epilogues write fp32 records, not a bf16 slice
                                       a[9] >= 0 ? e->buf_ptr(a[9]) : nullptr,
                                       a[9] >= 0 ? e->bufs[a[9]].c : 0, a[10], chained ? &ch : nullptr);
            if (rc != B2_OK) return fail(rc);
            const int pad = a[6] / 2, ho = (Hin + 2 * pad - a[6]) / a[7] + 1, wo = (Win + 2 * pad - a[6]) / a[7] + 1;
            if (a[18] != 0) {      // fused Detect-head epilogue: the output buffer holds fp32 {4 distances} or {logit, class} per pixel
                rc = b2_conv_set_head_epilogue(s.conv_ptr(), a[18], (float*)e->buf_ptr(a[3]));
                if (rc != B2_OK) return fail(rc);
                if (bo.c != (a[18] == 1 ? 8 : 4)) { b2_set_error("engine_create: op %d: head buffer has the wrong width", i); return fail(B2_ERR_ARG); }
            }
            if (ho != bo.h || wo != bo.w || (a[18] == 0 && a[4] + c_final > bo.c) || a[1] + a[2] > bi.c) {
                b2_set_error("engine_create: op %d: conv geometry does not match its buffers", i); return fail(B2_ERR_ARG);
            }
        } else if (s.op == OP_POOL) {
            if (!buf_ok(a[0])) { b2_set_error("engine_create: op %d: bad buffer id", i); return fail(B2_ERR_ARG); }
        } else if (s.op == OP_UP) {
            if (!buf_ok(a[0]) || !buf_ok(a[4])) { b2_set_error("engine_create: op %d: bad buffer id", i); return fail(B2_ERR_ARG); }
        } else { b2_set_error("engine_create: op %d: unknown opcode %d", i, s.op); return fail(B2_ERR_ARG); }
    }
    e->n_launch = n_ops;
    *out = e;
    return B2_OK;
}

extern "C" int b2_engine_destroy(b2_engine_t* e) {
    if (!e) return B2_OK;
    if (e->graph) cudaGraphExecDestroy(e->graph);
    cudaFree(e->arena);
    delete e;
    return B2_OK;
}

static int run_tail(b2_engine* e, cudaStream_t st) {
    if (e->use_graph) {
        if (!e->graph) {
            // capture on a private non-blocking stream (the caller's stream may be the legacy default stream,
            // which cannot be captured); nothing executes during capture
            cudaGraph_t g = nullptr;
            cudaStream_t cap = nullptr;
            B2_CUDA(cudaStreamCreateWithFlags(&cap, cudaStreamNonBlocking));
            cudaError_t ce = cudaStreamBeginCapture(cap, cudaStreamCaptureModeRelaxed);
            if (ce != cudaSuccess) { cudaStreamDestroy(cap); B2_CUDA(ce); }
            int rc = B2_OK;
            for (size_t i = 1; i < e->steps.size() && rc == B2_OK; ++i) rc = run_step(e, e->steps[i], cap);
            ce = cudaStreamEndCapture(cap, &g);
            cudaStreamDestroy(cap);
            if (rc != B2_OK) { if (g) cudaGraphDestroy(g); return rc; }
            B2_CUDA(ce);
            ce = cudaGraphInstantiate(&e->graph, g, 0);
            cudaGraphDestroy(g);
            B2_CUDA(ce);
        } else {
            b2_count_launch((int)e->steps.size() - 1);
        }
        B2_CUDA(cudaGraphLaunch(e->graph, st));
        return B2_OK;
    }
    for (size_t i = 1; i < e->steps.size(); ++i) {
        int rc = run_step(e, e->steps[i], st);
        if (rc != B2_OK) return rc;
    }
    return B2_OK;
}

extern "C" int b2_engine_forward_u8(b2_engine_t* e, const uint8_t* frames, int src_h, int src_w, int pad_top, int pad_left, void* stream) {
    B2_REQUIRE(e && frames, "engine_forward: null pointer");
    const int* a = e->steps[0].a;
    char* wbase = e->arena + e->weights_off;
    const Buf& bo = e->bufs[a[0]];
    int rc = b2_stem_u8(frames, e->B, src_h, src_w, e->H, e->W, pad_top, pad_left, (const void*)(wbase + (size_t)(uint32_t)a[3]),
                        (const float*)(wbase + (size_t)(uint32_t)a[4]), a[2], e->buf_ptr(a[0]), bo.c, a[1], stream);
    if (rc != B2_OK) return rc;
    return run_tail(e, (cudaStream_t)stream);
}

// Eager replay with a CUDA event between consecutive launches (on the launch stream): per-launch device times
// for roofline accounting (bench.py).  Synchronises.  ms_per_op: [b2_engine_num_launches] floats; op 0 = stem.
extern "C" int b2_engine_profile_u8(b2_engine_t* e, const uint8_t* frames, int src_h, int src_w, int pad_top, int pad_left,
                                    float* ms_per_op, void* stream) {
    B2_REQUIRE(e && frames && ms_per_op, "engine_profile: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t n = e->steps.size();
    std::vector<cudaEvent_t> ev(n + 1);
    for (auto& v : ev) B2_CUDA(cudaEventCreate(&v));
    const int* a = e->steps[0].a;
    char* wbase = e->arena + e->weights_off;
    int rc = B2_OK;
    // B2_NCU_OPS="3,57,59": bracket these launches with cudaProfilerStart/Stop (ncu --profile-from-start off captures only them)
    std::vector<char> mark(n, 0);
    if (const char* sel = getenv("B2_NCU_OPS")) {
        for (const char* q = sel; *q;) {
            char* end = nullptr;
            const long v = strtol(q, &end, 10);
            if (end == q) break;
            if (v >= 0 && (size_t)v < n) mark[(size_t)v] = 1;
            q = *end ? end + 1 : end;
        }
    }
    cudaEventRecord(ev[0], st);
    if (mark[0]) cudaProfilerStart();
    rc = b2_stem_u8(frames, e->B, src_h, src_w, e->H, e->W, pad_top, pad_left, (const void*)(wbase + (size_t)(uint32_t)a[3]),
                    (const float*)(wbase + (size_t)(uint32_t)a[4]), a[2], e->buf_ptr(a[0]), e->bufs[a[0]].c, a[1], stream);
    if (mark[0]) { cudaStreamSynchronize(st); cudaProfilerStop(); }
    cudaEventRecord(ev[1], st);
    for (size_t i = 1; i < n && rc == B2_OK; ++i) {
        if (mark[i]) cudaProfilerStart();
        rc = run_step(e, e->steps[i], st);
        if (mark[i]) { cudaStreamSynchronize(st); cudaProfilerStop(); }
        cudaEventRecord(ev[i + 1], st);
    }
    cudaError_t ce = cudaStreamSynchronize(st);
    if (rc == B2_OK && ce == cudaSuccess)
        for (size_t i = 0; i < n; ++i) cudaEventElapsedTime(&ms_per_op[i], ev[i], ev[i + 1]);
    for (auto& v : ev) cudaEventDestroy(v);
    if (rc != B2_OK) return rc;
    B2_CUDA(ce);
    return B2_OK;
}

extern "C" int b2_engine_forward_f32(b2_engine_t* e, const void* bchw, int dtype, void* stream) {
    B2_REQUIRE(e && bchw, "engine_forward: null pointer");
    const int* a = e->steps[0].a;
    char* wbase = e->arena + e->weights_off;
    const Buf& bo = e->bufs[a[0]];
    int rc = b2_stem_f32(bchw, dtype, e->B, e->H, e->W, (const void*)(wbase + (size_t)(uint32_t)a[3]),
                         (const float*)(wbase + (size_t)(uint32_t)a[4]), a[2], e->buf_ptr(a[0]), bo.c, a[1], stream);
    if (rc != B2_OK) return rc;
    return run_tail(e, (cudaStream_t)stream);
}

extern "C" int b2_engine_levels(b2_engine_t* e, int* n_levels, const void** logits, int* h, int* w, int* stride, int* lstride) {
    B2_REQUIRE(e && n_levels, "engine_levels: null pointer");
    *n_levels = e->n_levels;
    for (int l = 0; l < e->n_levels; ++l) {
        const bool fused = e->level_buf[l] < 0;
        const Buf& b = e->bufs[fused ? e->level_dist[l] : e->level_buf[l]];
        if (logits) logits[l] = fused ? nullptr : e->buf_ptr(e->level_buf[l]);
        if (h) h[l] = b.h;
        if (w) w[l] = b.w;
        if (stride) stride[l] = e->level_stride[l];
    }
    if (lstride) *lstride = e->lstride;
    return B2_OK;
}

extern "C" int b2_engine_head(b2_engine_t* e, const float** dist, const float** cls) {
    B2_REQUIRE(e && dist && cls, "engine_head: null pointer");
    for (int l = 0; l < e->n_levels; ++l) {
        if (e->level_dist[l] < 0 || e->level_cls[l] < 0) { b2_set_error("engine_head: this engine was lowered without the fused Detect head"); return B2_ERR_STATE; }
        dist[l] = (const float*)e->buf_ptr(e->level_dist[l]);
        cls[l] = (const float*)e->buf_ptr(e->level_cls[l]);
    }
    return B2_OK;
}

extern "C" int b2_engine_buffer(b2_engine_t* e, int buf, const void** ptr, int* h, int* w, int* c) {
    B2_REQUIRE(e && buf >= 0 && buf < (int)e->bufs.size(), "engine_buffer: bad buffer id %d", buf);
    if (ptr) *ptr = e->buf_ptr(buf);
    if (h) *h = e->bufs[buf].h;
    if (w) *w = e->bufs[buf].w;
    if (c) *c = e->bufs[buf].c;
    return B2_OK;
}

extern "C" size_t b2_engine_arena_bytes(b2_engine_t* e) { return e ? e->arena_bytes : 0; }
extern "C" int b2_engine_num_launches(b2_engine_t* e) { return e ? e->n_launch : 0; }
extern "C" int b2_engine_use_graph(b2_engine_t* e, int on) {
    B2_REQUIRE(e, "engine_use_graph: null handle");
    e->use_graph = on != 0;
    return B2_OK;
}
