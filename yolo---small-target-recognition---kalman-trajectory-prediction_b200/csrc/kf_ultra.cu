// Batched Ultralytics ByteTrack/BoT-SORT Kalman filters for sm_100a: one thread per track, dense 8x8 fp32
// covariance held in registers, state arrays in the caller's (N,8) / (N,8,8) row-major layout.
//
// Replaces ultralytics/trackers/utils/kalman_filter.py: KalmanFilterXYAH (:39-286) and KalmanFilterXYWH
// (:289-493): initiate, predict / multi_predict, project, update (Cholesky solve of the 4x4 innovation
// covariance; scipy.linalg.cho_factor / cho_solve in the reference), gating_distance.
#include "common.cuh"

void b2_count_launch(int n);

namespace {

constexpr float W_POS = 1.f / 20.f, W_VEL = 1.f / 160.f;

// length scales the std-devs are proportional to: XYAH -> h for all; XYWH -> (w, h, w, h)
__device__ __forceinline__ void scales(int kind, const float* m, float (&s)[4]) {
    if (kind == 0) { s[0] = s[1] = s[2] = s[3] = m[3]; }
    else { s[0] = m[2]; s[1] = m[3]; s[2] = m[2]; s[3] = m[3]; }
}

__global__ void __launch_bounds__(128) kf_initiate_kernel(int kind, const float* __restrict__ meas, float* __restrict__ mean,
                                                          float* __restrict__ cov, int N) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    float m[8];
    for (int i = 0; i < 4; ++i) { m[i] = meas[n * 4 + i]; m[i + 4] = 0.f; }
    float s[4]; scales(kind, m, s);
    float sd[8];
    for (int i = 0; i < 4; ++i) { sd[i] = 2.f * W_POS * s[i]; sd[i + 4] = 10.f * W_VEL * s[i]; }
    if (kind == 0) { sd[2] = 1e-2f; sd[6] = 1e-5f; }
    for (int i = 0; i < 8; ++i) mean[n * 8 + i] = m[i];
    float* P = cov + (size_t)n * 64;
    for (int i = 0; i < 64; ++i) P[i] = 0.f;
    for (int i = 0; i < 8; ++i) P[i * 9] = sd[i] * sd[i];
}

__global__ void __launch_bounds__(128) kf_predict_kernel(int kind, float* __restrict__ mean, float* __restrict__ cov, int N) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    float m[8], P[64];
    for (int i = 0; i < 8; ++i) m[i] = mean[n * 8 + i];
    float* Pg = cov + (size_t)n * 64;
#pragma unroll
    for (int i = 0; i < 64; ++i) P[i] = Pg[i];
    float s[4]; scales(kind, m, s);
    float q[8];
    for (int i = 0; i < 4; ++i) { const float a = W_POS * s[i], b = W_VEL * s[i]; q[i] = a * a; q[i + 4] = b * b; }
    if (kind == 0) { q[2] = 1e-2f * 1e-2f; q[6] = 1e-5f * 1e-5f; }
    for (int i = 0; i < 4; ++i) m[i] += m[i + 4];
    // F P: rows 0..3 += rows 4..7 ; then (F P) F^T: cols 0..3 += cols 4..7
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) P[i * 8 + j] += P[(i + 4) * 8 + j];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) P[i * 8 + j] += P[i * 8 + j + 4];
#pragma unroll
    for (int i = 0; i < 8; ++i) P[i * 9] += q[i];
    for (int i = 0; i < 8; ++i) mean[n * 8 + i] = m[i];
#pragma unroll
    for (int i = 0; i < 64; ++i) Pg[i] = P[i];
}

__device__ __forceinline__ void project_dev(int kind, const float* m, const float* P, float (&pm)[4], float (&S)[16]) {
    float s[4]; scales(kind, m, s);
    for (int i = 0; i < 4; ++i) {
        pm[i] = m[i];
        for (int j = 0; j < 4; ++j) S[i * 4 + j] = P[i * 8 + j];
    }
    for (int i = 0; i < 4; ++i) { float sd = W_POS * s[i]; if (kind == 0 && i == 2) sd = 1e-1f; S[i * 5] += sd * sd; }
}

// lower Cholesky of an SPD d x d matrix (row-major, leading dim ld)
__device__ __forceinline__ void chol(const float* A, int d, int ld, float* L) {
    for (int i = 0; i < d; ++i)
        for (int j = 0; j <= i; ++j) {
            float v = A[i * ld + j];
            for (int k = 0; k < j; ++k) v -= L[i * 4 + k] * L[j * 4 + k];
            L[i * 4 + j] = (i == j) ? sqrtf(v) : v / L[j * 4 + j];
        }
}

__global__ void __launch_bounds__(128) kf_project_kernel(int kind, const float* __restrict__ mean, const float* __restrict__ cov,
                                                         float* __restrict__ pmean, float* __restrict__ pcov, int N) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    float pm[4], S[16];
    project_dev(kind, mean + n * 8, cov + (size_t)n * 64, pm, S);
    for (int i = 0; i < 4; ++i) pmean[n * 4 + i] = pm[i];
    for (int i = 0; i < 16; ++i) pcov[n * 16 + i] = S[i];
}

__global__ void __launch_bounds__(128) kf_update_kernel(int kind, float* __restrict__ mean, float* __restrict__ cov,
                                                        const float* __restrict__ meas, const uint8_t* __restrict__ mask, int N) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N || (mask && !mask[n])) return;
    float m[8], P[64];
    for (int i = 0; i < 8; ++i) m[i] = mean[n * 8 + i];
    float* Pg = cov + (size_t)n * 64;
#pragma unroll
    for (int i = 0; i < 64; ++i) P[i] = Pg[i];
    // Covariances that come out of initiate / predict / update keep the pattern P[i][j] != 0 <=> i == j (mod 4) (diagonal P0, Q, R
    // and H = [I 0]: four independent (position, velocity) filters).  Then S is diagonal and the update has a closed
    // form per coordinate, with 1 - K_x evaluated as r / S: no cancellation, no Cholesky.  Anything else takes the dense path.
    bool structured = true;
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (((i - j) & 3) != 0 && P[i * 8 + j] != 0.f) structured = false;
    if (structured) {
        float s[4]; scales(kind, m, s);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float sd = W_POS * s[i]; if (kind == 0 && i == 2) sd = 1e-1f;
            const float r = sd * sd;
            const float pxx = P[i * 9], pxv = P[i * 8 + i + 4], pvv = P[(i + 4) * 9];
            const float S = pxx + r, kx = pxx / S, kv = pxv / S, omk = r / S;
            const float y = meas[n * 4 + i] - m[i];
            mean[n * 8 + i] = m[i] + kx * y;
            mean[n * 8 + i + 4] = m[i + 4] + kv * y;
            Pg[i * 9] = omk * pxx;
            Pg[i * 8 + i + 4] = omk * pxv; Pg[(i + 4) * 8 + i] = omk * pxv;
            Pg[(i + 4) * 9] = pvv - kv * pxv;
        }
        return;
    }
    float pm[4], S[16], L[16];
    project_dev(kind, m, P, pm, S);
    chol(S, 4, 4, L);
    // K (8x4): row r solves S k = P[r, :4]^T  (K = P H^T S^-1)
    float K[32];
    for (int r = 0; r < 8; ++r) {
        float y[4];
        for (int i = 0; i < 4; ++i) { float v = P[r * 8 + i]; for (int k = 0; k < i; ++k) v -= L[i * 4 + k] * y[k]; y[i] = v / L[i * 5]; }
        for (int i = 3; i >= 0; --i) { float v = y[i]; for (int k = i + 1; k < 4; ++k) v -= L[k * 4 + i] * K[r * 4 + k]; K[r * 4 + i] = v / L[i * 5]; }
    }
    float innov[4];
    for (int i = 0; i < 4; ++i) innov[i] = meas[n * 4 + i] - pm[i];
    for (int r = 0; r < 8; ++r) { float v = m[r]; for (int i = 0; i < 4; ++i) v += K[r * 4 + i] * innov[i]; mean[n * 8 + r] = v; }
    // P -= K S K^T
    float KS[32];
    for (int r = 0; r < 8; ++r)
        for (int j = 0; j < 4; ++j) { float v = 0.f; for (int i = 0; i < 4; ++i) v += K[r * 4 + i] * S[i * 4 + j]; KS[r * 4 + j] = v; }
    for (int r = 0; r < 8; ++r)
        for (int c = 0; c < 8; ++c) { float v = 0.f; for (int j = 0; j < 4; ++j) v += KS[r * 4 + j] * K[c * 4 + j]; Pg[r * 8 + c] = P[r * 8 + c] - v; }
}

__global__ void __launch_bounds__(128) kf_gating_kernel(int kind, const float* __restrict__ mean, const float* __restrict__ cov, int N,
                                                        const float* __restrict__ meas, int M, int only_position, int metric,
                                                        float* __restrict__ out) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= N * M) return;
    const int n = idx / M, mi = idx % M;
    float pm[4], S[16], L[16];
    project_dev(kind, mean + n * 8, cov + (size_t)n * 64, pm, S);
    const int d = only_position ? 2 : 4;
    float dv[4];
    for (int i = 0; i < d; ++i) dv[i] = meas[mi * 4 + i] - pm[i];
    float acc = 0.f;
    if (metric == 1) { for (int i = 0; i < d; ++i) acc += dv[i] * dv[i]; }
    else {
        chol(S, d, 4, L);
        float y[4];
        for (int i = 0; i < d; ++i) { float v = dv[i]; for (int k = 0; k < i; ++k) v -= L[i * 4 + k] * y[k]; y[i] = v / L[i * 5]; acc += y[i] * y[i]; }
    }
    out[idx] = acc;
}

}  // namespace

#define KF_CHECK(kind, N) B2_REQUIRE((kind) == 0 || (kind) == 1, "kf: kind must be 0 (XYAH) or 1 (XYWH)"); B2_REQUIRE((N) >= 0, "kf: N < 0"); if ((N) == 0) return B2_OK

extern "C" int b2_kf_initiate(int kind, const float* meas, float* mean, float* cov, int N, void* stream) {
    KF_CHECK(kind, N);
    kf_initiate_kernel<<<b2_ceil_div(N, 128), 128, 0, (cudaStream_t)stream>>>(kind, meas, mean, cov, N);
    B2_CUDA(cudaGetLastError()); b2_count_launch(1);
    return B2_OK;
}
extern "C" int b2_kf_predict(int kind, float* mean, float* cov, int N, void* stream) {
    KF_CHECK(kind, N);
    kf_predict_kernel<<<b2_ceil_div(N, 128), 128, 0, (cudaStream_t)stream>>>(kind, mean, cov, N);
    B2_CUDA(cudaGetLastError()); b2_count_launch(1);
    return B2_OK;
}
extern "C" int b2_kf_project(int kind, const float* mean, const float* cov, float* pmean, float* pcov, int N, void* stream) {
    KF_CHECK(kind, N);
    kf_project_kernel<<<b2_ceil_div(N, 128), 128, 0, (cudaStream_t)stream>>>(kind, mean, cov, pmean, pcov, N);
    B2_CUDA(cudaGetLastError()); b2_count_launch(1);
    return B2_OK;
}
extern "C" int b2_kf_update(int kind, float* mean, float* cov, const float* meas, const uint8_t* mask, int N, void* stream) {
    KF_CHECK(kind, N);
    kf_update_kernel<<<b2_ceil_div(N, 128), 128, 0, (cudaStream_t)stream>>>(kind, mean, cov, meas, mask, N);
    B2_CUDA(cudaGetLastError()); b2_count_launch(1);
    return B2_OK;
}
extern "C" int b2_kf_gating(int kind, const float* mean, const float* cov, int N, const float* meas, int M,
                            int only_position, int metric, float* out, void* stream) {
    KF_CHECK(kind, N);
    B2_REQUIRE(metric == 0 || metric == 1, "Invalid distance metric");
    if (M == 0) return B2_OK;
    kf_gating_kernel<<<b2_ceil_div(N * M, 128), 128, 0, (cudaStream_t)stream>>>(kind, mean, cov, N, meas, M, only_position, metric, out);
    B2_CUDA(cudaGetLastError()); b2_count_launch(1);
    return B2_OK;
}
