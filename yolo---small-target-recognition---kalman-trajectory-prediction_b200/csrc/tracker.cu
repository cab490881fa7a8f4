// Structure-of-arrays Kalman track bank for sm_100a: one independent multi-target tracker per video stream.
//
// Replaces kalman/enhanced_multi_target_tracker.py:42-132 (EnhancedMultiTargetTracker.update) and
// kalman/enhanced_aircraft_kalman_tracker.py (AircraftKalmanTracker: predict :184-203, update :249-297,
// analyze_motion_pattern :137-182, mark_as_lost :299-317, enhanced_long_term_predict :205-247,
// get_track_info :335-383, should_delete :385-405).
//
// Layout: field-major SoA over S*C slots (slot g = stream*C + i), fp32 state.  With the reference's
// F, H, Q, R, P0 the 8x8 covariance keeps the pattern P[i,j] != 0 <=> i == j (mod 4); x/y share one
// (pos,vel) covariance triple and w/h another, so the bank stores 8 state floats + 2x3 covariance floats
// per track and evaluates the exact closed-form 2x2 recurrences (verified in
// tests/test_oracle_vs_golden.py::test_tracker_known_answer).
//
// Per frame: (1) bank_predict  -- every live slot, grid-wide, coalesced (HBM-bound)
//            (2) associate     -- one CTA per stream: IoU rows on the fly, greedy matching by repeated
//                                 mutual-best (== the reference's descending-IoU greedy, ties -> lowest
//                                 (det, track id))
//            (3) finish        -- one CTA per stream: Kalman update / mark lost / delete / create /
//                                 emit rows (with the reference's extra predict on the first lost frame)
// List order of the reference (= ascending track id) is carried by the id column; slots are recycled.
#include "common.cuh"

#include <new>

void b2_count_launch(int n);

namespace {

enum FField { X0 = 0, X1, X2, X3, X4, X5, X6, X7, PPX, PPV, PVV, PSX, PSV, PSVV, VAVGX, VAVGY, DIRN, SPEED, STAB, PCONF, NFF };
enum IField { ID = 0, AGE, HITS, STREAK, TSU, LOSTF, ISLOST, NVEL, VHEAD, TLEN, THEAD, NIF };
constexpr int kVelRing = 50;       // deque(maxlen=50)  enhanced_aircraft_kalman_tracker.py:79
constexpr int kTraj = B2_TRAJ_LEN; // only the last 30 trajectory points are ever read (:377)

// noise constants, enhanced_aircraft_kalman_tracker.py:44-71
constexpr float P0_POS = 50.f, P0_VEL = 100.f, P0_SVEL = 1.f;
constexpr float Q_POS = 0.1f, Q_SIZE = 0.01f, Q_VEL = 0.1f, Q_SVEL = 0.001f, R_MEAS = 10.f;

struct Bank {
    float* f;       // [NFF][N]
    int32_t* i;     // [NIF][N]
    float* vel;     // [kVelRing*2][N]
    float* traj;    // [kTraj*2][N]
    float4* pbox;   // [N] predicted boxes of this frame
    int32_t* match; // [N] matched detection index or -1
    int32_t* det_match;   // [S][max_dets]
    int32_t* next_id;     // [S]
    int32_t* frame_count; // [S]
    long long* stats;     // [S][8]: created, terminated, active, long_term, recoveries, overflow
    int S, C, N, max_dets;
    int max_lost, min_hits; float iou_thr;
};

struct b2_tracker_impl {
    Bank b;
    void* arena;
    size_t arena_bytes;
};

__device__ __forceinline__ float& FF(const Bank& b, int field, int g) { return b.f[(size_t)field * b.N + g]; }
__device__ __forceinline__ int32_t& II(const Bank& b, int field, int g) { return b.i[(size_t)field * b.N + g]; }

__device__ __forceinline__ void push_traj(const Bank& b, int g, float cx, float cy) {
    int head = II(b, THEAD, g), len = II(b, TLEN, g);
    b.traj[(size_t)(2 * head) * b.N + g] = cx;
    b.traj[(size_t)(2 * head + 1) * b.N + g] = cy;
    head = head + 1 == kTraj ? 0 : head + 1;
    II(b, THEAD, g) = head;
    II(b, TLEN, g) = min(len + 1, kTraj);
}

// x <- F x, P <- F P F^T + Q, age++, tsu++, trajectory push  (:184-203)
__device__ __forceinline__ void predict_slot(const Bank& b, int g) {
    const float cx = FF(b, X0, g) + FF(b, X4, g), cy = FF(b, X1, g) + FF(b, X5, g);
    FF(b, X0, g) = cx; FF(b, X1, g) = cy;
    FF(b, X2, g) += FF(b, X6, g); FF(b, X3, g) += FF(b, X7, g);
    {
        const float pxx = FF(b, PPX, g), pxv = FF(b, PPV, g), pvv = FF(b, PVV, g);
        FF(b, PPX, g) = pxx + 2.f * pxv + pvv + Q_POS; FF(b, PPV, g) = pxv + pvv; FF(b, PVV, g) = pvv + Q_VEL;
    }
    {
        const float pxx = FF(b, PSX, g), pxv = FF(b, PSV, g), pvv = FF(b, PSVV, g);
        FF(b, PSX, g) = pxx + 2.f * pxv + pvv + Q_SIZE; FF(b, PSV, g) = pxv + pvv; FF(b, PSVV, g) = pvv + Q_SVEL;
    }
    II(b, AGE, g) += 1; II(b, TSU, g) += 1;
    push_traj(b, g, cx, cy);
}

__global__ void __launch_bounds__(256) bank_predict_kernel(const Bank b) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= b.N) return;
    b.match[g] = -1;
    if (II(b, ID, g) == 0) return;
    predict_slot(b, g);
    const float cx = FF(b, X0, g), cy = FF(b, X1, g), w = FF(b, X2, g), h = FF(b, X3, g);
    b.pbox[g] = make_float4(cx - w / 2.f, cy - h / 2.f, cx + w / 2.f, cy + h / 2.f);   // state_to_bbox :121-135
}

// Four consecutive slots per thread with 16-byte accesses (N % 4 == 0): the same arithmetic as predict_slot, element-wise;
// dead slots (id == 0) are written back unchanged.  512 bytes per warp per access instead of 128.
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, const float4& v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ float sel(bool c, float a, float b) { return c ? a : b; }

__global__ void __launch_bounds__(256) bank_predict4_kernel(const Bank b) {
    const int g = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (g >= b.N) return;
    const size_t N = b.N;
    const int4 id = *reinterpret_cast<const int4*>(b.i + (size_t)ID * N + g);
    *reinterpret_cast<int4*>(b.match + g) = make_int4(-1, -1, -1, -1);
    const bool l0 = id.x != 0, l1 = id.y != 0, l2 = id.z != 0, l3 = id.w != 0;
    if (!(l0 || l1 || l2 || l3)) return;
    float* F = b.f + g;
    float4 x0 = ld4(F + X0 * N), x1 = ld4(F + X1 * N), x2 = ld4(F + X2 * N), x3 = ld4(F + X3 * N);
    const float4 v0 = ld4(F + X4 * N), v1 = ld4(F + X5 * N), v2 = ld4(F + X6 * N), v3 = ld4(F + X7 * N);
#define B2_ADD4(a, v) a.x = sel(l0, a.x + v.x, a.x); a.y = sel(l1, a.y + v.y, a.y); a.z = sel(l2, a.z + v.z, a.z); a.w = sel(l3, a.w + v.w, a.w);
    B2_ADD4(x0, v0) B2_ADD4(x1, v1) B2_ADD4(x2, v2) B2_ADD4(x3, v3)
#undef B2_ADD4
    st4(F + X0 * N, x0); st4(F + X1 * N, x1); st4(F + X2 * N, x2); st4(F + X3 * N, x3);
#define B2_COV4(PX, PV, VV, QP, QV)                                                                                      \
    {                                                                                                                    \
        float4 pxx = ld4(F + PX * N), pxv = ld4(F + PV * N), pvv = ld4(F + VV * N);                                       \
        pxx.x = sel(l0, pxx.x + 2.f * pxv.x + pvv.x + QP, pxx.x); pxv.x = sel(l0, pxv.x + pvv.x, pxv.x); pvv.x = sel(l0, pvv.x + QV, pvv.x); \
        pxx.y = sel(l1, pxx.y + 2.f * pxv.y + pvv.y + QP, pxx.y); pxv.y = sel(l1, pxv.y + pvv.y, pxv.y); pvv.y = sel(l1, pvv.y + QV, pvv.y); \
        pxx.z = sel(l2, pxx.z + 2.f * pxv.z + pvv.z + QP, pxx.z); pxv.z = sel(l2, pxv.z + pvv.z, pxv.z); pvv.z = sel(l2, pvv.z + QV, pvv.z); \
        pxx.w = sel(l3, pxx.w + 2.f * pxv.w + pvv.w + QP, pxx.w); pxv.w = sel(l3, pxv.w + pvv.w, pxv.w); pvv.w = sel(l3, pvv.w + QV, pvv.w); \
        st4(F + PX * N, pxx); st4(F + PV * N, pxv); st4(F + VV * N, pvv);                                                 \
    }
    B2_COV4(PPX, PPV, PVV, Q_POS, Q_VEL)
    B2_COV4(PSX, PSV, PSVV, Q_SIZE, Q_SVEL)
#undef B2_COV4
    int32_t* I = b.i + g;
    int4 age = *reinterpret_cast<const int4*>(I + (size_t)AGE * N), tsu = *reinterpret_cast<const int4*>(I + (size_t)TSU * N);
    age.x += l0; age.y += l1; age.z += l2; age.w += l3;
    tsu.x += l0; tsu.y += l1; tsu.z += l2; tsu.w += l3;
    *reinterpret_cast<int4*>(I + (size_t)AGE * N) = age; *reinterpret_cast<int4*>(I + (size_t)TSU * N) = tsu;
    int4 head = *reinterpret_cast<const int4*>(I + (size_t)THEAD * N), len = *reinterpret_cast<const int4*>(I + (size_t)TLEN * N);
    const float cx[4] = {x0.x, x0.y, x0.z, x0.w}, cy[4] = {x1.x, x1.y, x1.z, x1.w}, w[4] = {x2.x, x2.y, x2.z, x2.w}, h[4] = {x3.x, x3.y, x3.z, x3.w};
    const bool live[4] = {l0, l1, l2, l3};
    int hd[4] = {head.x, head.y, head.z, head.w}, ln[4] = {len.x, len.y, len.z, len.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (live[k]) {
            b.traj[(size_t)(2 * hd[k]) * N + g + k] = cx[k];                       // push_traj
            b.traj[(size_t)(2 * hd[k] + 1) * N + g + k] = cy[k];
            hd[k] = hd[k] + 1 == kTraj ? 0 : hd[k] + 1;
            ln[k] = min(ln[k] + 1, kTraj);
            b.pbox[g + k] = make_float4(cx[k] - w[k] / 2.f, cy[k] - h[k] / 2.f, cx[k] + w[k] / 2.f, cy[k] + h[k] / 2.f);
        }
    }
    *reinterpret_cast<int4*>(I + (size_t)THEAD * N) = make_int4(hd[0], hd[1], hd[2], hd[3]);
    *reinterpret_cast<int4*>(I + (size_t)TLEN * N) = make_int4(ln[0], ln[1], ln[2], ln[3]);
}

static inline void launch_bank_predict(const Bank& b, cudaStream_t st) {
    if (b.N % 4 == 0) bank_predict4_kernel<<<b2_ceil_div(b.N / 4, 256), 256, 0, st>>>(b);
    else bank_predict_kernel<<<b2_ceil_div(b.N, 256), 256, 0, st>>>(b);
}

// _calculate_iou (enhanced_multi_target_tracker.py:200-232)
__device__ __forceinline__ float iou_ref(const float4& d, const float4& t) {
    const float x1 = fmaxf(d.x, t.x), y1 = fmaxf(d.y, t.y), x2 = fminf(d.z, t.z), y2 = fminf(d.w, t.w);
    if (x2 <= x1 || y2 <= y1) return 0.f;
    const float inter = (x2 - x1) * (y2 - y1);
    const float a1 = (d.z - d.x) * (d.w - d.y), a2 = (t.z - t.x) * (t.w - t.y);
    const float uni = a1 + a2 - inter;
    return uni <= 0.f ? 0.f : inter / uni;
}

__device__ __forceinline__ int block_exclusive_scan(int v, int* s_warp, int* total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        int w = lane < (int)(blockDim.x >> 5) ? s_warp[lane] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += t; }
        s_warp[lane] = w;   // inclusive over warps
    }
    __syncthreads();
    const int warp_off = warp ? s_warp[warp - 1] : 0;
    *total = s_warp[(blockDim.x >> 5) - 1];
    const int r = warp_off + inc - v;
    __syncthreads();
    return r;
}

constexpr int kAssocThreads = 512;
constexpr int kMaxDetsSmem = 1024;

// One CTA per stream.  Greedy descending-IoU matching == repeat { every free detection picks its best free
// track (row max); a pair is accepted iff no other free detection beats it on that track (column max) }.
__global__ void __launch_bounds__(kAssocThreads) associate_global_kernel(const Bank b, const float* __restrict__ dets, int det_cols,
                                                                const int32_t* __restrict__ det_counts) {
    __shared__ float4 s_det[kMaxDetsSmem];
    __shared__ int s_dmatch[kMaxDetsSmem];
    __shared__ int s_best_t[kMaxDetsSmem];
    __shared__ float s_best_iou[kMaxDetsSmem];
    __shared__ int s_progress;
    const int s = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = kAssocThreads / 32;
    const int D = min(det_counts[s], b.max_dets);
    const int g0 = s * b.C;
    for (int d = tid; d < D; d += kAssocThreads) {
        const float* r = dets + ((size_t)s * b.max_dets + d) * det_cols;
        s_det[d] = make_float4(r[0], r[1], r[2], r[3]);
        s_dmatch[d] = -1;
    }
    __syncthreads();
    const float thr = b.iou_thr;
    while (true) {
        if (tid == 0) s_progress = 0;
        // ---- row pass: best free track per free detection (warp per detection) ----
        for (int d = warp; d < D; d += nwarps) {
            if (s_dmatch[d] >= 0) continue;
            const float4 db = s_det[d];
            float best = -1.f; int bt = -1, bid = 0x7fffffff;
            for (int t = lane; t < b.C; t += 32) {
                const int g = g0 + t;
                const int id = II(b, ID, g);
                if (id == 0 || b.match[g] >= 0) continue;
                const float v = iou_ref(db, b.pbox[g]);
                if (v >= thr && (v > best || (v == best && id < bid))) { best = v; bt = t; bid = id; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ob = __shfl_xor_sync(0xffffffffu, best, o);
                const int ot = __shfl_xor_sync(0xffffffffu, bt, o), oid = __shfl_xor_sync(0xffffffffu, bid, o);
                if (ob > best || (ob == best && oid < bid)) { best = ob; bt = ot; bid = oid; }
            }
            if (lane == 0) { s_best_t[d] = bt; s_best_iou[d] = best; }
        }
        __syncthreads();
        // ---- column check: is (d, t) also the best free detection for t? ----
        for (int d = tid; d < D; d += kAssocThreads) {
            if (s_dmatch[d] >= 0) continue;
            const int t = s_best_t[d];
            if (t < 0) continue;
            const float v = s_best_iou[d];
            const float4 tb = b.pbox[g0 + t];
            bool dominated = false;
            for (int e = 0; e < D && !dominated; ++e) {
                if (e == d || s_dmatch[e] >= 0) continue;
                const float ve = iou_ref(s_det[e], tb);
                if (ve >= thr && (ve > v || (ve == v && e < d))) dominated = true;
            }
            if (!dominated) { b.match[g0 + t] = d; s_dmatch[d] = t; s_progress = 1; }
        }
        __syncthreads();
        if (!s_progress) break;
        __syncthreads();
    }
    for (int d = tid; d < b.max_dets; d += kAssocThreads) b.det_match[(size_t)s * b.max_dets + d] = d < D ? s_dmatch[d] : -2;
}

// Same algorithm with the stream's LIVE tracks compacted into shared memory first (box, id, slot): the rounds then
// touch no global memory.  Dynamic shared memory: C * 28 bytes.
__global__ void __launch_bounds__(kAssocThreads) associate_kernel(const Bank b, const float* __restrict__ dets, int det_cols,
                                                                const int32_t* __restrict__ det_counts, int cand_cap) {
    extern __shared__ uint8_t s_raw[];
    float4* l_box = reinterpret_cast<float4*>(s_raw);                    // [C]
    int* l_id = reinterpret_cast<int*>(l_box + b.C);                      // [C]
    int* l_slot = l_id + b.C;                                             // [C]
    int* l_match = l_slot + b.C;                                          // [C]
    __shared__ float4 s_det[kMaxDetsSmem];
    __shared__ int s_dmatch[kMaxDetsSmem];
    __shared__ int s_best_t[kMaxDetsSmem];
    __shared__ float s_best_iou[kMaxDetsSmem];
    __shared__ int s_warp[32];
    __shared__ int s_progress, s_ncand;
    const int s = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = kAssocThreads / 32;
    const int D = min(det_counts[s], b.max_dets);
    const int g0 = s * b.C;
    for (int d = tid; d < D; d += kAssocThreads) {
        const float* r = dets + ((size_t)s * b.max_dets + d) * det_cols;
        s_det[d] = make_float4(r[0], r[1], r[2], r[3]);
        s_dmatch[d] = -1;
    }
    // ---- compact the live tracks (slot order) ----
    int base_live = 0;
    for (int base = 0; base < b.C; base += kAssocThreads) {
        const int t = base + tid;
        const int id = t < b.C ? II(b, ID, g0 + t) : 0;
        int total;
        const int off = block_exclusive_scan(id != 0 ? 1 : 0, s_warp, &total);
        if (id != 0) {
            const int k = base_live + off;
            l_box[k] = b.pbox[g0 + t]; l_id[k] = id; l_slot[k] = t; l_match[k] = -1;
        }
        base_live += total;
    }
    const int L = base_live;
    __syncthreads();
    const float thr = b.iou_thr;
    // ---- fast path: list every pair with IoU >= thr ONCE (sparse: a detection overlaps a handful of tracks), then run the
    //      mutual-best rounds on that list with 64-bit shared-memory atomicMax keys instead of re-evaluating D x L IoUs per
    //      round.  A pair is accepted when the track is the detection's best free candidate by (IoU, lowest track id) and
    //      the detection is the track's best free candidate by (IoU, lowest detection index): exactly the pairs the
    //      reference's descending-IoU greedy walk takes (enhanced_multi_target_tracker.py:234-270).  Falls back to the
    //      dense rounds below if the list overflows. ----
    uint8_t* dyn = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(l_match + b.C) + 7) & ~uintptr_t(7));
    unsigned long long* t_best = reinterpret_cast<unsigned long long*>(dyn);               // [C]
    unsigned long long* d_best = t_best + b.C;                                              // [kMaxDetsSmem]
    float* c_iou = reinterpret_cast<float*>(d_best + kMaxDetsSmem);                         // [cand_cap]
    uint32_t* c_pair = reinterpret_cast<uint32_t*>(c_iou + cand_cap);                       // [cand_cap]: det << 16 | track
    if (tid == 0) s_ncand = 0;
    __syncthreads();
    if (D > 0 && L > 0 && cand_cap > 0) {
        for (int d0 = 0; d0 < D; d0 += nwarps) {
            const int d = d0 + warp;
            if (d < D) {
                const float4 db = s_det[d];
                for (int t = lane; t < L; t += 32) {
                    const float v = iou_ref(db, l_box[t]);
                    if (v >= thr) {
                        const int pos = atomicAdd(&s_ncand, 1);
                        if (pos < cand_cap) { c_iou[pos] = v; c_pair[pos] = ((uint32_t)d << 16) | (uint32_t)t; }
                    }
                }
            }
            if (d0 == 0) {
                // dense scene (the first nwarps detections already project past the list capacity): do not finish a list
                // that will be thrown away, go to the dense rounds
                __syncthreads();
                const int seen = min(D, nwarps);
                if ((long long)s_ncand * D > (long long)cand_cap * seen) { if (tid == 0) s_ncand = cand_cap + 1; break; }
            }
        }
    }
    __syncthreads();
    const int M = s_ncand;
    if (M <= cand_cap && b.C <= 65536) {
        while (M > 0) {
            for (int d = tid; d < D; d += kAssocThreads) d_best[d] = 0ull;
            for (int t = tid; t < L; t += kAssocThreads) t_best[t] = 0ull;
            if (tid == 0) s_progress = 0;
            __syncthreads();
            for (int e = tid; e < M; e += kAssocThreads) {
                const uint32_t pr = c_pair[e];
                const int d = (int)(pr >> 16), t = (int)(pr & 0xFFFFu);
                if (s_dmatch[d] >= 0 || l_match[t] >= 0) continue;
                const unsigned long long hi = (unsigned long long)__float_as_uint(c_iou[e]) << 32;   // IoU > 0: bits order like the float
                atomicMax(&d_best[d], hi | (unsigned)(0xFFFFFFFFu - (unsigned)l_id[t]));
                atomicMax(&t_best[t], hi | (unsigned)(0xFFFFFFFFu - (unsigned)d));
            }
            __syncthreads();
            for (int e = tid; e < M; e += kAssocThreads) {
                const uint32_t pr = c_pair[e];
                const int d = (int)(pr >> 16), t = (int)(pr & 0xFFFFu);
                if (s_dmatch[d] >= 0 || l_match[t] >= 0) continue;
                const unsigned long long hi = (unsigned long long)__float_as_uint(c_iou[e]) << 32;
                if (d_best[d] == (hi | (unsigned)(0xFFFFFFFFu - (unsigned)l_id[t])) && t_best[t] == (hi | (unsigned)(0xFFFFFFFFu - (unsigned)d))) {
                    // unique per d and per t within a round: no two entries share (d, best t) or (t, best d)
                    s_dmatch[d] = t; l_match[t] = d; s_progress = 1;
                }
            }
            __syncthreads();
            if (!s_progress) break;
            __syncthreads();
        }
    } else
    if (D > 0 && L > 0) {
        while (true) {
            if (tid == 0) s_progress = 0;
            // ---- row pass: best free track per free detection (warp per detection) ----
            for (int d = warp; d < D; d += nwarps) {
                if (s_dmatch[d] >= 0) continue;
                const float4 db = s_det[d];
                float best = -1.f; int bt = -1, bid = 0x7fffffff;
                for (int t = lane; t < L; t += 32) {
                    if (l_match[t] >= 0) continue;
                    const float v = iou_ref(db, l_box[t]);
                    const int id = l_id[t];
                    if (v >= thr && (v > best || (v == best && id < bid))) { best = v; bt = t; bid = id; }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
                    const int ot = __shfl_xor_sync(0xffffffffu, bt, o), oid = __shfl_xor_sync(0xffffffffu, bid, o);
                    if (ob > best || (ob == best && oid < bid)) { best = ob; bt = ot; bid = oid; }
                }
                if (lane == 0) { s_best_t[d] = bt; s_best_iou[d] = best; }
            }
            __syncthreads();
            // ---- column check: is (d, t) also the best free detection for t? ----
            for (int d = tid; d < D; d += kAssocThreads) {
                if (s_dmatch[d] >= 0) continue;
                const int t = s_best_t[d];
                if (t < 0) continue;
                const float v = s_best_iou[d];
                const float4 tb = l_box[t];
                bool dominated = false;
                for (int e = 0; e < D && !dominated; ++e) {
                    if (e == d || s_dmatch[e] >= 0) continue;
                    const float ve = iou_ref(s_det[e], tb);
                    if (ve >= thr && (ve > v || (ve == v && e < d))) dominated = true;
                }
                if (!dominated) { l_match[t] = d; s_dmatch[d] = t; s_progress = 1; }
            }
            __syncthreads();
            if (!s_progress) break;
            __syncthreads();
        }
    }
    for (int t = tid; t < L; t += kAssocThreads) if (l_match[t] >= 0) b.match[g0 + l_slot[t]] = l_match[t];
    for (int d = tid; d < b.max_dets; d += kAssocThreads) b.det_match[(size_t)s * b.max_dets + d] = d < D ? (s_dmatch[d] >= 0 ? l_slot[s_dmatch[d]] : -1) : -2;
}

// analyze_motion_pattern (:137-163) + _calculate_direction_consistency (:165-182) over the velocity ring
__device__ void analyze_slot(const Bank& b, int g) {
    const int n = II(b, NVEL, g);
    if (n < 5) return;
    const int head = II(b, VHEAD, g);                      // next write position; oldest = head - n
    float sx = 0.f, sy = 0.f;
    for (int k = 0; k < n; ++k) { sx += b.vel[(size_t)(2 * k) * b.N + g]; sy += b.vel[(size_t)(2 * k + 1) * b.N + g]; }
    const float mx = sx / n, my = sy / n;
    float vx2 = 0.f, vy2 = 0.f;
    for (int k = 0; k < n; ++k) {
        const float dx = b.vel[(size_t)(2 * k) * b.N + g] - mx, dy = b.vel[(size_t)(2 * k + 1) * b.N + g] - my;
        vx2 += dx * dx; vy2 += dy * dy;
    }
    const float sdx = sqrtf(vx2 / n), sdy = sqrtf(vy2 / n);
    // direction changes in chronological order
    const float PI = 3.14159265358979323846f;
    float dsum = 0.f, prev = 0.f;
    float dch[kVelRing];
    int start = head - n; if (start < 0) start += kVelRing;
    for (int k = 0; k < n; ++k) {
        int r = start + k; if (r >= kVelRing) r -= kVelRing;
        const float ang = atan2f(b.vel[(size_t)(2 * r + 1) * b.N + g], b.vel[(size_t)(2 * r) * b.N + g]);
        if (k > 0) {
            float c = ang - prev;
            if (!(fabsf(c) < PI)) c = c - 2.f * PI * (c > 0.f ? 1.f : (c < 0.f ? -1.f : 0.f));
            dch[k - 1] = c; dsum += c;
        }
        prev = ang;
    }
    const float dmean = dsum / (n - 1);
    float dvar = 0.f;
    for (int k = 0; k < n - 1; ++k) { const float e = dch[k] - dmean; dvar += e * e; }
    const float dstd = sqrtf(dvar / (n - 1));
    const float speed_stab = 1.f / (1.f + (sdx + sdy) / 2.f);
    const float dir_cons = 1.f / (1.f + dstd * 10.f);
    const float stab = (speed_stab + dir_cons) / 2.f;
    FF(b, VAVGX, g) = mx; FF(b, VAVGY, g) = my;
    FF(b, SPEED, g) = sqrtf(mx * mx + my * my);
    FF(b, DIRN, g) = atan2f(my, mx);
    FF(b, STAB, g) = stab;
    FF(b, PCONF, g) = stab * fminf((float)n / 30.f, 1.f);
}

// The same analysis by one warp (lane <-> ring entry, two entries per lane): the serial form walks the 50-entry ring three
// times with dependent global loads -- ~100 k cycles for ONE matched track, and every 256-slot sweep of finish_kernel that
// held a matched track waited for it.  Sums become shuffle trees (fp32, different summation order than the serial loop:
// within the 2e-4 / 1e-3 gates of tests/test_gpu_tracker.py against the float64 reference).
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ void analyze_slot_warp(const Bank& b, int g, int lane) {
    const int n = II(b, NVEL, g);
    if (n < 5) return;                                      // uniform: every lane reads the same slot
    const int head = II(b, VHEAD, g);
    int start = head - n; if (start < 0) start += kVelRing;
    float vx[2], vy[2]; bool ok[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
        const int k = lane + 32 * e;                        // chronological index
        ok[e] = k < n;
        int r = start + k; if (r >= kVelRing) r -= kVelRing;
        vx[e] = ok[e] ? b.vel[(size_t)(2 * r) * b.N + g] : 0.f;
        vy[e] = ok[e] ? b.vel[(size_t)(2 * r + 1) * b.N + g] : 0.f;
    }
    const float mx = warp_sum(vx[0] + vx[1]) / n, my = warp_sum(vy[0] + vy[1]) / n;
    float qx = 0.f, qy = 0.f;
#pragma unroll
    for (int e = 0; e < 2; ++e) if (ok[e]) { const float dx = vx[e] - mx, dy = vy[e] - my; qx += dx * dx; qy += dy * dy; }
    const float sdx = sqrtf(warp_sum(qx) / n), sdy = sqrtf(warp_sum(qy) / n);
    const float PI = 3.14159265358979323846f;
    float ang[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) ang[e] = ok[e] ? atan2f(vy[e], vx[e]) : 0.f;
    // angle of the chronologically previous entry: lane - 1 of the same half, or lane 31 of the first half for entry 32
    const float up0 = __shfl_up_sync(0xffffffffu, ang[0], 1), up1 = __shfl_up_sync(0xffffffffu, ang[1], 1);
    const float last0 = __shfl_sync(0xffffffffu, ang[0], 31);
    const float prev[2] = {up0, lane == 0 ? last0 : up1};
    float c[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
        const int k = lane + 32 * e;
        float d = ang[e] - prev[e];
        if (!(fabsf(d) < PI)) d = d - 2.f * PI * (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f));
        c[e] = (ok[e] && k > 0) ? d : 0.f;
    }
    const float dmean = warp_sum(c[0] + c[1]) / (n - 1);
    float qv = 0.f;
#pragma unroll
    for (int e = 0; e < 2; ++e) { const int k = lane + 32 * e; if (ok[e] && k > 0) { const float t = c[e] - dmean; qv += t * t; } }
    const float dstd = sqrtf(warp_sum(qv) / (n - 1));
    if (lane == 0) {
        const float speed_stab = 1.f / (1.f + (sdx + sdy) / 2.f);
        const float dir_cons = 1.f / (1.f + dstd * 10.f);
        const float stab = (speed_stab + dir_cons) / 2.f;
        FF(b, VAVGX, g) = mx; FF(b, VAVGY, g) = my;
        FF(b, SPEED, g) = sqrtf(mx * mx + my * my);
        FF(b, DIRN, g) = atan2f(my, mx);
        FF(b, STAB, g) = stab;
        FF(b, PCONF, g) = stab * fminf((float)n / 30.f, 1.f);
    }
}

// Kalman update with measurement z = bbox_to_state(det)  (:249-297)
__device__ __forceinline__ void update_slot(const Bank& b, int g, const float4& d, bool analyze = true) {
    II(b, TSU, g) = 0; II(b, HITS, g) += 1; II(b, STREAK, g) += 1;
    if (II(b, ISLOST, g)) { II(b, ISLOST, g) = 0; II(b, LOSTF, g) = 0; }
    const float z0 = (d.x + d.z) / 2.f, z1 = (d.y + d.w) / 2.f, z2 = d.z - d.x, z3 = d.w - d.y;
    {   // position block (x, y share the covariance triple)
        const float pxx = FF(b, PPX, g), pxv = FF(b, PPV, g), pvv = FF(b, PVV, g);
        // (I - K H) P with 1 - K_x evaluated as R / S: no cancellation when P_xx >> R (long coasting tracks)
        const float S = pxx + R_MEAS, kx = pxx / S, kv = pxv / S, omk = R_MEAS / S;
        const float y0 = z0 - FF(b, X0, g), y1 = z1 - FF(b, X1, g);
        FF(b, X0, g) += kx * y0; FF(b, X4, g) += kv * y0;
        FF(b, X1, g) += kx * y1; FF(b, X5, g) += kv * y1;
        FF(b, PPX, g) = omk * pxx; FF(b, PPV, g) = omk * pxv; FF(b, PVV, g) = pvv - kv * pxv;
    }
    {   // size block (w, h)
        const float pxx = FF(b, PSX, g), pxv = FF(b, PSV, g), pvv = FF(b, PSVV, g);
        const float S = pxx + R_MEAS, kx = pxx / S, kv = pxv / S, omk = R_MEAS / S;
        const float y2 = z2 - FF(b, X2, g), y3 = z3 - FF(b, X3, g);
        FF(b, X2, g) += kx * y2; FF(b, X6, g) += kv * y2;
        FF(b, X3, g) += kx * y3; FF(b, X7, g) += kv * y3;
        FF(b, PSX, g) = omk * pxx; FF(b, PSV, g) = omk * pxv; FF(b, PSVV, g) = pvv - kv * pxv;
    }
    // velocity ring push, trajectory push, motion analysis
    int head = II(b, VHEAD, g);
    b.vel[(size_t)(2 * head) * b.N + g] = FF(b, X4, g);
    b.vel[(size_t)(2 * head + 1) * b.N + g] = FF(b, X5, g);
    II(b, VHEAD, g) = head + 1 == kVelRing ? 0 : head + 1;
    II(b, NVEL, g) = min(II(b, NVEL, g) + 1, kVelRing);
    push_traj(b, g, FF(b, X0, g), FF(b, X1, g));
    if (analyze) analyze_slot(b, g);
}

// AircraftKalmanTracker.__init__ (:23-101)
__device__ __forceinline__ void init_slot(const Bank& b, int g, const float4& d, int id) {
    const float cx = (d.x + d.z) / 2.f, cy = (d.y + d.w) / 2.f;
    FF(b, X0, g) = cx; FF(b, X1, g) = cy; FF(b, X2, g) = d.z - d.x; FF(b, X3, g) = d.w - d.y;
    FF(b, X4, g) = 0.f; FF(b, X5, g) = 0.f; FF(b, X6, g) = 0.f; FF(b, X7, g) = 0.f;
    FF(b, PPX, g) = P0_POS; FF(b, PPV, g) = 0.f; FF(b, PVV, g) = P0_VEL;
    FF(b, PSX, g) = P0_POS; FF(b, PSV, g) = 0.f; FF(b, PSVV, g) = P0_SVEL;
    FF(b, VAVGX, g) = 0.f; FF(b, VAVGY, g) = 0.f; FF(b, DIRN, g) = 0.f; FF(b, SPEED, g) = 0.f; FF(b, STAB, g) = 0.f; FF(b, PCONF, g) = 0.f;
    II(b, ID, g) = id; II(b, AGE, g) = 0; II(b, HITS, g) = 1; II(b, STREAK, g) = 1; II(b, TSU, g) = 0;
    II(b, LOSTF, g) = 0; II(b, ISLOST, g) = 0; II(b, NVEL, g) = 0; II(b, VHEAD, g) = 0; II(b, TLEN, g) = 0; II(b, THEAD, g) = 0;
    push_traj(b, g, cx, cy);
}

// get_track_info (:335-383) incl. get_lost_prediction / enhanced_long_term_predict side effects
__device__ void emit_slot(const Bank& b, int g, float* row, float* traj_out, int32_t* traj_len_out, int slot, int* long_term) {
    float bx, by, bw, bh, conf; int predicted = II(b, TSU, g) > 0;
    if (predicted) {
        if (II(b, ISLOST, g)) {
            const int k = II(b, LOSTF, g);
            if (k <= 1) {                                   // enhanced_long_term_predict(1) -> self.predict(), 1.0 (:216-217)
                predict_slot(b, g);
                bx = FF(b, X0, g); by = FF(b, X1, g); bw = FF(b, X2, g); bh = FF(b, X3, g); conf = 1.f;
            } else if (FF(b, PCONF, g) > 0.3f) {            // high confidence: mean-velocity extrapolation (:224-236)
                bx = FF(b, X0, g) + FF(b, VAVGX, g) * (float)k; by = FF(b, X1, g) + FF(b, VAVGY, g) * (float)k;
                bw = FF(b, X2, g); bh = FF(b, X3, g);
                conf = FF(b, PCONF, g) * fmaxf(0.1f, 1.f - (float)k / (float)b.max_lost);
            } else {                                        // F^k x (:238-245)
                bx = FF(b, X0, g) + (float)k * FF(b, X4, g); by = FF(b, X1, g) + (float)k * FF(b, X5, g);
                bw = FF(b, X2, g) + (float)k * FF(b, X6, g); bh = FF(b, X3, g) + (float)k * FF(b, X7, g);
                conf = fmaxf(0.1f, 1.f - (float)k / ((float)b.max_lost * 0.5f));
            }
        } else {
            bx = FF(b, X0, g); by = FF(b, X1, g); bw = FF(b, X2, g); bh = FF(b, X3, g);
            conf = fmaxf(0.3f, 1.f - (float)II(b, TSU, g) / 60.f);
        }
    } else {
        bx = FF(b, X0, g); by = FF(b, X1, g); bw = FF(b, X2, g); bh = FF(b, X3, g); conf = 1.f;
    }
    const int tsu = II(b, TSU, g);
    int32_t* irow = reinterpret_cast<int32_t*>(row);
    irow[0] = II(b, ID, g);
    row[1] = bx - bw / 2.f; row[2] = by - bh / 2.f; row[3] = bx + bw / 2.f; row[4] = by + bh / 2.f;
    row[5] = conf; irow[6] = predicted;
    irow[7] = II(b, AGE, g); irow[8] = II(b, HITS, g); irow[9] = II(b, STREAK, g); irow[10] = tsu; irow[11] = tsu;
    irow[12] = predicted;
    row[13] = FF(b, X4, g); row[14] = FF(b, X5, g); row[15] = FF(b, PCONF, g);
    irow[16] = FF(b, STAB, g) > 0.5f ? 1 : 0;
    row[17] = FF(b, SPEED, g); row[18] = FF(b, DIRN, g); irow[19] = slot;
    if (predicted && tsu > 30) *long_term += 1;
    if (traj_out) {
        const int len = II(b, TLEN, g), head = II(b, THEAD, g);
        int start = head - len; if (start < 0) start += kTraj;
        for (int k = 0; k < len; ++k) {
            int r = start + k; if (r >= kTraj) r -= kTraj;
            traj_out[2 * k] = b.traj[(size_t)(2 * r) * b.N + g];
            traj_out[2 * k + 1] = b.traj[(size_t)(2 * r + 1) * b.N + g];
        }
        *traj_len_out = len;
    }
}

constexpr int kFinishThreads = 512;

// n staged rows (80 bytes each, contiguous) -> global, 16 bytes per thread and step
__device__ __forceinline__ void flush_rows(const float* s_rows, float* dst, int n, int tid) {
    const float4* src4 = reinterpret_cast<const float4*>(s_rows);
    float4* dst4 = reinterpret_cast<float4*>(dst);
    for (int i = tid; i < n * (B2_TRACK_COLS / 4); i += kFinishThreads) dst4[i] = src4[i];
}

// One CTA per stream: update / mark lost / delete / emit for existing tracks (slot order), then create.
__global__ void __launch_bounds__(kFinishThreads) finish_kernel(const Bank b, const float* __restrict__ dets, int det_cols,
                                                               const int32_t* __restrict__ det_counts, float* __restrict__ out_rows,
                                                               int32_t* __restrict__ out_counts, float* __restrict__ out_traj,
                                                               int32_t* __restrict__ out_traj_len) {
    __shared__ int s_warp[32];
    __shared__ int s_emit_base, s_free_base, s_created, s_an;
    __shared__ int s_alist[kMaxDetsSmem];                  // slots updated this frame (<= detections of the stream)
    __shared__ __align__(16) float s_rows[kFinishThreads * B2_TRACK_COLS];   // emitted rows of one sweep
    __shared__ long long s_stats[6];
    const int s = blockIdx.x, tid = threadIdx.x;
    const int g0 = s * b.C;
    const int D = min(det_counts[s], b.max_dets);
    if (tid == 0) { s_emit_base = 0; s_free_base = 0; s_created = 0; s_an = 0; }
    if (tid < 6) s_stats[tid] = 0;
    const int frame = b.frame_count[s] + 1;
    __syncthreads();
    int terminated = 0, recoveries = 0, long_term = 0;
    float* rows = out_rows + (size_t)s * b.C * B2_TRACK_COLS;

    // ---- pass 1a: existing tracks: Kalman update / mark lost / delete; matched slots are listed for the motion analysis ----
    for (int base = 0; base < b.C; base += kFinishThreads) {
        const int t = base + tid, g = g0 + t;
        if (t < b.C && II(b, ID, g) != 0) {
            const int m = b.match[g];
            if (m >= 0) {
                if (II(b, ISLOST, g)) recoveries++;
                const float* r = dets + ((size_t)s * b.max_dets + m) * det_cols;
                update_slot(b, g, make_float4(r[0], r[1], r[2], r[3]), false);
                s_alist[atomicAdd(&s_an, 1)] = t;
            } else {                                           // mark_as_lost (:299-317)
                if (!II(b, ISLOST, g)) { II(b, ISLOST, g) = 1; II(b, LOSTF, g) = 0; }
                II(b, LOSTF, g) += 1; II(b, STREAK, g) = 0;
            }
            const int tsu = II(b, TSU, g), age = II(b, AGE, g), hs = II(b, STREAK, g);
            const bool del = tsu > b.max_lost || (age < 5 && hs == 0 && tsu > 15) || (age < 10 && hs <= 1 && tsu > 30);   // :385-405
            if (del) { II(b, ID, g) = 0; terminated++; }
        }
    }
    __syncthreads();
    // ---- pass 1b: motion analysis of the updated tracks, one warp per track (analyze_slot_warp) ----
    for (int i = tid >> 5; i < s_an; i += kFinishThreads / 32) analyze_slot_warp(b, g0 + s_alist[i], tid & 31);
    __syncthreads();
    // ---- pass 1c: emit in slot order ----
    for (int base = 0; base < b.C; base += kFinishThreads) {
        const int t = base + tid, g = g0 + t;
        int emit = 0;
        if (t < b.C && II(b, ID, g) != 0)
            emit = (II(b, STREAK, g) >= b.min_hits || frame <= b.min_hits || II(b, ISLOST, g)) ? 1 : 0;   // multi_target_tracker.py:117-126
        int total;
        const int off = block_exclusive_scan(emit, s_warp, &total);
        const int ebase = s_emit_base;
        if (emit) {
            // the row goes to shared memory: the emitted rows of a sweep are contiguous in the output, so the block writes
            // them with coalesced 16-byte stores instead of twenty 4-byte stores per thread at an 80-byte stride
            const int pos = ebase + off;
            emit_slot(b, g, s_rows + off * B2_TRACK_COLS,
                      out_traj ? out_traj + ((size_t)s * b.C + pos) * kTraj * 2 : nullptr,
                      out_traj_len ? out_traj_len + (size_t)s * b.C + pos : nullptr, t, &long_term);
        }
        __syncthreads();
        flush_rows(s_rows, rows + (size_t)ebase * B2_TRACK_COLS, total, tid);
        if (tid == 0) s_emit_base = ebase + total;
        __syncthreads();
    }

    // ---- pass 2: new tracks for unmatched detections, ascending detection index -> ascending ids,
    //      k-th new track takes the k-th free slot ----
    const int32_t* dm = b.det_match + (size_t)s * b.max_dets;
    const int id0 = b.next_id[s];
    // rank of each unmatched detection
    int n_new = 0;
    {
        // D <= max_dets; serial ranks via scan over chunks
        for (int base = 0; base < D; base += kFinishThreads) {
            const int d = base + tid;
            const int um = (d < D && dm[d] == -1) ? 1 : 0;
            int total;
            (void)block_exclusive_scan(um, s_warp, &total);
            n_new += total;
        }
    }
    if (n_new > 0) {
        int det_base = 0;   // rank offset over detection chunks
        // walk free slots chunk by chunk; for every chunk find which ranks its free slots serve
        int free_seen = 0;
        for (int base = 0; base < b.C && free_seen < n_new; base += kFinishThreads) {
            const int t = base + tid, g = g0 + t;
            const int is_free = (t < b.C && II(b, ID, g) == 0) ? 1 : 0;
            int total;
            const int off = block_exclusive_scan(is_free, s_warp, &total);
            const int rank = free_seen + off;
            if (is_free && rank < n_new) {
                // find the rank-th unmatched detection (ascending)
                int cnt = 0, dsel = -1;
                for (int d = 0; d < D; ++d) { if (dm[d] == -1) { if (cnt == rank) { dsel = d; break; } ++cnt; } }
                const float* r = dets + ((size_t)s * b.max_dets + dsel) * det_cols;
                init_slot(b, g, make_float4(r[0], r[1], r[2], r[3]), id0 + rank);
                b.match[g] = -3;   // marks "created this frame" for the emit pass below
            }
            free_seen += total;
        }
        (void)det_base;
        const int created = min(n_new, free_seen);
        // emit new tracks (hit_streak = 1): condition identical to pass 1
        const bool emit_new = (1 >= b.min_hits) || (frame <= b.min_hits);
        if (emit_new) {
            for (int base = 0; base < b.C; base += kFinishThreads) {
                const int t = base + tid, g = g0 + t;
                const int emit = (t < b.C && II(b, ID, g) != 0 && b.match[g] == -3) ? 1 : 0;
                int total;
                const int off = block_exclusive_scan(emit, s_warp, &total);
                const int ebase = s_emit_base;
                if (emit) {
                    const int pos = ebase + off;
                    emit_slot(b, g, s_rows + off * B2_TRACK_COLS,
                              out_traj ? out_traj + ((size_t)s * b.C + pos) * kTraj * 2 : nullptr,
                              out_traj_len ? out_traj_len + (size_t)s * b.C + pos : nullptr, t, &long_term);
                }
                __syncthreads();
                flush_rows(s_rows, rows + (size_t)ebase * B2_TRACK_COLS, total, tid);
                if (tid == 0) s_emit_base = ebase + total;
                __syncthreads();
            }
        }
        if (tid == 0) { s_created = created; s_stats[5] = n_new - created; }
    }
    __syncthreads();

    // ---- stats (enhanced_multi_target_tracker.py:32-38) ----
    atomicAdd((unsigned long long*)&s_stats[1], (unsigned long long)terminated);
    atomicAdd((unsigned long long*)&s_stats[3], (unsigned long long)long_term);
    atomicAdd((unsigned long long*)&s_stats[4], (unsigned long long)recoveries);
    // active count
    int active = 0;
    for (int t = tid; t < b.C; t += kFinishThreads) active += II(b, ID, g0 + t) != 0;
    atomicAdd((unsigned long long*)&s_stats[2], (unsigned long long)active);
    __syncthreads();
    if (tid == 0) {
        long long* st = b.stats + (size_t)s * 8;
        st[0] += s_created; st[1] += s_stats[1]; st[2] = s_stats[2]; st[3] += s_stats[3]; st[4] += s_stats[4]; st[5] += s_stats[5];
        // the reference numbers every unmatched detection; ids keep advancing even if the bank overflowed
        b.next_id[s] = id0 + n_new;
        b.frame_count[s] = frame;
        out_counts[s] = s_emit_base;
    }
}

__global__ void export_kernel(const Bank b, int s, float* x, float* P, int32_t* meta, int32_t* n_out) {
    // single thread block; serialise in ascending id order is done on the host -- here slot order
    __shared__ int cnt;
    if (threadIdx.x == 0) cnt = 0;
    __syncthreads();
    for (int t = threadIdx.x; t < b.C; t += blockDim.x) {
        const int g = s * b.C + t;
        if (II(b, ID, g) == 0) continue;
        const int k = atomicAdd(&cnt, 1);
        for (int j = 0; j < 8; ++j) x[k * 8 + j] = FF(b, X0 + j, g);
        float* Pk = P + (size_t)k * 64;
        for (int j = 0; j < 64; ++j) Pk[j] = 0.f;
        for (int j = 0; j < 2; ++j) {     // x,y then w,h
            Pk[j * 8 + j] = FF(b, PPX, g); Pk[j * 8 + j + 4] = FF(b, PPV, g); Pk[(j + 4) * 8 + j] = FF(b, PPV, g); Pk[(j + 4) * 8 + j + 4] = FF(b, PVV, g);
            const int q = j + 2;
            Pk[q * 8 + q] = FF(b, PSX, g); Pk[q * 8 + q + 4] = FF(b, PSV, g); Pk[(q + 4) * 8 + q] = FF(b, PSV, g); Pk[(q + 4) * 8 + q + 4] = FF(b, PSVV, g);
        }
        int32_t* mk = meta + (size_t)k * 8;
        mk[0] = II(b, ID, g); mk[1] = II(b, AGE, g); mk[2] = II(b, HITS, g); mk[3] = II(b, STREAK, g);
        mk[4] = II(b, TSU, g); mk[5] = II(b, LOSTF, g); mk[6] = II(b, ISLOST, g); mk[7] = II(b, NVEL, g);
    }
    __syncthreads();
    if (threadIdx.x == 0) *n_out = cnt;
}

}  // namespace

struct b2_tracker { b2_tracker_impl impl; };

extern "C" int b2_tracker_create(int n_streams, int capacity, int max_dets, int max_lost_frames, int min_hits,
                                 float iou_threshold, b2_tracker_t** out) {
    B2_REQUIRE(out, "tracker_create: out is null");
    B2_REQUIRE(n_streams >= 1 && capacity >= 1 && max_dets >= 1 && max_dets <= kMaxDetsSmem,
               "tracker_create: need n_streams>=1, capacity>=1, 1<=max_dets<=%d", kMaxDetsSmem);
    b2_tracker* t = new (std::nothrow) b2_tracker();
    if (!t) { b2_set_error("out of host memory"); return B2_ERR_STATE; }
    Bank& b = t->impl.b;
    b.S = n_streams; b.C = capacity; b.N = n_streams * capacity; b.max_dets = max_dets;
    b.max_lost = max_lost_frames; b.min_hits = min_hits; b.iou_thr = iou_threshold;
    const size_t N = (size_t)b.N;
    auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
    size_t off = 0;
    const size_t o_f = off; off += up(N * NFF * 4);
    const size_t o_i = off; off += up(N * NIF * 4);
    const size_t o_v = off; off += up(N * kVelRing * 2 * 4);
    const size_t o_t = off; off += up(N * kTraj * 2 * 4);
    const size_t o_b = off; off += up(N * 16);
    const size_t o_m = off; off += up(N * 4);
    const size_t o_dm = off; off += up((size_t)n_streams * max_dets * 4);
    const size_t o_ni = off; off += up((size_t)n_streams * 4);
    const size_t o_fc = off; off += up((size_t)n_streams * 4);
    const size_t o_st = off; off += up((size_t)n_streams * 8 * 8);
    cudaError_t e = cudaMalloc(&t->impl.arena, off);
    if (e != cudaSuccess) { b2_set_error("tracker_create: cudaMalloc(%zu) failed: %s", off, cudaGetErrorString(e)); delete t; return B2_ERR_CUDA; }
    t->impl.arena_bytes = off;
    char* a = (char*)t->impl.arena;
    b.f = (float*)(a + o_f); b.i = (int32_t*)(a + o_i); b.vel = (float*)(a + o_v); b.traj = (float*)(a + o_t);
    b.pbox = (float4*)(a + o_b); b.match = (int32_t*)(a + o_m); b.det_match = (int32_t*)(a + o_dm);
    b.next_id = (int32_t*)(a + o_ni); b.frame_count = (int32_t*)(a + o_fc); b.stats = (long long*)(a + o_st);
    *out = t;
    return b2_tracker_reset(t, nullptr);
}

extern "C" int b2_tracker_destroy(b2_tracker_t* t) {
    if (!t) return B2_OK;
    cudaFree(t->impl.arena);
    delete t;
    return B2_OK;
}

namespace {
__global__ void fill_i32(int32_t* p, int n, int v) { const int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) p[i] = v; }
}

extern "C" int b2_tracker_reset(b2_tracker_t* t, void* stream) {
    B2_REQUIRE(t, "tracker_reset: null handle");
    cudaStream_t st = (cudaStream_t)stream;
    B2_CUDA(cudaMemsetAsync(t->impl.arena, 0, t->impl.arena_bytes, st));
    fill_i32<<<b2_ceil_div(t->impl.b.S, 256), 256, 0, st>>>(t->impl.b.next_id, t->impl.b.S, 1);   // next_track_id = 1 (:30)
    B2_CUDA(cudaGetLastError());
    b2_count_launch(1);
    return B2_OK;
}

extern "C" int b2_tracker_bank_predict(b2_tracker_t* t, void* stream) {
    B2_REQUIRE(t, "tracker: null handle");
    launch_bank_predict(t->impl.b, (cudaStream_t)stream);
    B2_CUDA(cudaGetLastError());
    b2_count_launch(1);
    return B2_OK;
}

extern "C" int b2_tracker_update(b2_tracker_t* t, const float* dets, int det_cols, const int32_t* det_counts,
                                 float* out_rows, int32_t* out_counts, float* out_traj, int32_t* out_traj_len, void* stream) {
    B2_REQUIRE(t && dets && det_counts && out_rows && out_counts, "tracker_update: null pointer");
    B2_REQUIRE(det_cols >= 4, "tracker_update: det_cols must be >= 4");
    const Bank& b = t->impl.b;
    cudaStream_t st = (cudaStream_t)stream;
    launch_bank_predict(b, st);
    const size_t track_smem = (size_t)b.C * 28;
    if (track_smem <= 160 * 1024) {
        // sparse matching: best-candidate keys (8 B per track, 8 KB for the detections) + the pair list (8 B per pair) in what
        // is left of ~190 KB of shared memory, up to 8192 pairs
        const size_t fixed = track_smem + (size_t)b.C * 8 + kMaxDetsSmem * 8 + 16;
        size_t cap = fixed < 190 * 1024 ? (190 * 1024 - fixed) / 8 : 0;
        cap = cap > 8192 ? 8192 : (cap < 512 ? 0 : cap);
        // the list must not cost occupancy: banks whose track arrays still allow two CTAs per SM (and have more streams than
        // SMs) keep the small footprint and the dense rounds
        if (track_smem + 30 * 1024 <= 113 * 1024 && b.S > b2_num_sms()) cap = 0;
        const size_t assoc_smem = cap ? fixed + cap * 8 : track_smem;
        if (assoc_smem > 16 * 1024) {
            static size_t granted = 0;
            if (assoc_smem > granted) { B2_CUDA(cudaFuncSetAttribute(associate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)assoc_smem)); granted = assoc_smem; }
        }
        associate_kernel<<<b.S, kAssocThreads, assoc_smem, st>>>(b, dets, det_cols, det_counts, (int)cap);
    } else {
        associate_global_kernel<<<b.S, kAssocThreads, 0, st>>>(b, dets, det_cols, det_counts);
    }
    finish_kernel<<<b.S, kFinishThreads, 0, st>>>(b, dets, det_cols, det_counts, out_rows, out_counts, out_traj, out_traj_len);
    B2_CUDA(cudaGetLastError());
    b2_count_launch(3);
    return B2_OK;
}

extern "C" int b2_tracker_export(b2_tracker_t* t, int stream_idx, float* x_host, float* P_host, int32_t* meta_host,
                                 int32_t* n_tracks_host, long long* stats_host) {
    B2_REQUIRE(t && stream_idx >= 0 && stream_idx < t->impl.b.S, "tracker_export: bad stream index");
    const Bank& b = t->impl.b;
    B2_CUDA(cudaDeviceSynchronize());
    if (x_host || P_host || meta_host || n_tracks_host) {
        float *dx = nullptr, *dP = nullptr; int32_t *dm = nullptr, *dn = nullptr;
        B2_CUDA(cudaMalloc(&dx, (size_t)b.C * 8 * 4)); B2_CUDA(cudaMalloc(&dP, (size_t)b.C * 64 * 4));
        B2_CUDA(cudaMalloc(&dm, (size_t)b.C * 8 * 4)); B2_CUDA(cudaMalloc(&dn, 4));
        export_kernel<<<1, 256>>>(b, stream_idx, dx, dP, dm, dn);
        b2_count_launch(1);
        int n = 0;
        cudaError_t e = cudaMemcpy(&n, dn, 4, cudaMemcpyDeviceToHost);
        if (e == cudaSuccess && x_host) e = cudaMemcpy(x_host, dx, (size_t)n * 8 * 4, cudaMemcpyDeviceToHost);
        if (e == cudaSuccess && P_host) e = cudaMemcpy(P_host, dP, (size_t)n * 64 * 4, cudaMemcpyDeviceToHost);
        if (e == cudaSuccess && meta_host) e = cudaMemcpy(meta_host, dm, (size_t)n * 8 * 4, cudaMemcpyDeviceToHost);
        cudaFree(dx); cudaFree(dP); cudaFree(dm); cudaFree(dn);
        B2_CUDA(e);
        if (n_tracks_host) *n_tracks_host = n;
    }
    if (stats_host) {
        long long st[8];
        B2_CUDA(cudaMemcpy(st, b.stats + (size_t)stream_idx * 8, sizeof(st), cudaMemcpyDeviceToHost));
        int32_t fc = 0, nid = 0;
        B2_CUDA(cudaMemcpy(&fc, b.frame_count + stream_idx, 4, cudaMemcpyDeviceToHost));
        B2_CUDA(cudaMemcpy(&nid, b.next_id + stream_idx, 4, cudaMemcpyDeviceToHost));
        for (int k = 0; k < 5; ++k) stats_host[k] = st[k];
        stats_host[5] = fc; stats_host[6] = nid; stats_host[7] = st[5];
    }
    return B2_OK;
}

extern "C" int b2_tracker_bytes_per_track(int* predict_bytes, int* update_bytes) {
    // predict: read x[8] P[6] id age tsu thead tlen (19 words) ; write x[4] P[6] age tsu thead tlen + 2 traj + pbox(4) + match (21 words)
    if (predict_bytes) *predict_bytes = (19 + 21) * 4;
    // update (matched): read match, id, islost, x[8], P[6], counters(4), vhead, nvel, thead, tlen, det(4) + ring re-scan 2*50*... (up to 3 passes of 100 floats)
    if (update_bytes) *update_bytes = (1 + 2 + 14 + 4 + 4 + 4) * 4 + (14 + 6 + 4 + 2 + 2 + 6) * 4 + 100 * 4;
    return B2_OK;
}

extern "C" int b2_tracker_seed(b2_tracker_t* t, const float* boxes, const int32_t* counts, int max_rows, void* stream) {
    // Seeding == one update on an empty bank with min-hits semantics untouched: every row creates a track.
    (void)t; (void)boxes; (void)counts; (void)max_rows; (void)stream;
    b2_set_error("b2_tracker_seed: use b2_tracker_update on a reset bank (every detection creates a track)");
    return B2_ERR_UNSUPPORTED;
}
