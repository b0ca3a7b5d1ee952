// Device-resident frame loop of vidExample.py (SURVEY.md 8(f) rank 4 + the C3 pipeline): everything the reference's
// per-frame Python does BETWEEN its cv2 / solver calls, kept in HBM so a whole sequence runs without a host round trip.
//
//   vel_klt_sequence   <- vidExample.py:134-135   p, v = KLT(im, im0, p);  vg[vg] = v      (track propagation)
//   vel_seq_pose_t     <- vidExample.py:139-146   estimateWorldCameraPose(findR=False) per frame (fcnNLS_t,
//                                                 utils/NLS.py:102-129), world2image reprojection, rms residual,
//                                                 B[i,0:6], S[i,2:4]
//   vel_seq_stats      <- vidExample.py:142-146,164   dt, dr, cumulative distance, speed = dr/dt*3.6  (S rows)
//   vel_seq_select     <- utils/NLS.py:190-191    v = isfinite(P[4]).sum(1) == nframes  (full-length tracks), compacted
//   vel_seq_rays       <- utils/MSV.py:13-15      U[:, j] = pixel2uvec(K, P[0:2, vg, j].T).T
//   vel_seq_pack_ba    <- utils/NLS.py:198-203    z = [all x | all y] (track fastest), x0 = [points | cam pos | rpy 0]
//   vel_seq_export_P   <- vidExample.py:128,151-153   P[5, npts, n] float32, NaN = invalid
//
// Layout on the device is FRAME-MAJOR: tracks [n][npts][2] float32 (== P[0:2] transposed), alive [n][npts] uint8
// (== vg after frame k), proj [n][npts][2] float32 (== P[2:4]).  Frame-major is what every consumer streams (the pose
// solve reads one frame, the BA measurement vector is "all x, camera slowest, track fastest", utils/NLS.py:198-199);
// vel_seq_export_P produces the reference's [5][npts][n] array for callers that want it (plots).
// A track that failed stays in the arrays at a far out-of-frame sentinel, where K2's bounds test drops it at every level
// (the reference compacts p[v]; LK treats points independently, so surviving tracks are bit-identical either way).
#include "common.cuh"

// lk_track.cu: the whole frame run in one launch (15x15 window, word-aligned pitches); 1 = launched, 0 = not applicable
int vel_lk_sequence_w15h(const uint8_t* frames, int64_t frame_stride, int32_t pitch, const uint8_t* pyr, int64_t pyr_stride,
                         const vel_pyr_layout* layout, int32_t nframes, int32_t npts, const vel_lk_params* params, float* tracks,
                         uint8_t* alive, float* err, vel_stream_t stream);

namespace {

constexpr float kDeadXY = -1.0e5f;   // K2: floor(p * 2^-level - halfWin) < -win  at every level -> status 0, no memory access

__global__ void seq_propagate_kernel(const uint8_t* __restrict__ alive_prev, const uint8_t* __restrict__ status,
                                     uint8_t* __restrict__ alive_next, float2* __restrict__ next_pts, int npts)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npts) return;
    const uint8_t a = (alive_prev[i] != 0 && status[i] != 0) ? 1 : 0;
    alive_next[i] = a;
    if (!a) next_pts[i] = make_float2(kDeadXY, kDeadXY);
}

__global__ void seq_seed_kernel(const uint8_t* __restrict__ alive0, float2* __restrict__ pts0, int npts)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < npts && alive0[i] == 0) pts0[i] = make_float2(kDeadXY, kDeadXY);
}

// ---- per-frame translation solve (fcnNLS_t) + reprojection + residual ------------------------------------------------
constexpr int POSE_THREADS = 256;
constexpr int POSE_MAX_ITER = 30;

__device__ __forceinline__ double warp_sum_d(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ void project_rk(const double* K, double ax, double ay, double az, double& u, double& v)
{
    const double q0 = ax * K[0] + ay * K[3] + az * K[6];
    const double q1 = ax * K[1] + ay * K[4] + az * K[7];
    const double q2 = ax * K[2] + ay * K[5] + az * K[8];
    // one reciprocal instead of two divisions (an FP64 division is ~25 instructions on this part and the pose fit is bound by them);
    // u, v move by at most an ulp, far below the forward-difference noise of the Jacobian they feed
    const double r = 1.0 / q2;
    u = q0 * r;
    v = q1 * r;
}

// (H + I) d = g, 3x3 symmetric, Gaussian elimination with partial pivoting (same as nls.cu)
__device__ void solve3_damped(const double* Hu /*6: 00 01 02 11 12 22*/, const double* g, double* d)
{
    double M[3][4] = {{Hu[0] + 1.0, Hu[1], Hu[2], g[0]}, {Hu[1], Hu[3] + 1.0, Hu[4], g[1]}, {Hu[2], Hu[4], Hu[5] + 1.0, g[2]}};
    for (int k = 0; k < 3; ++k) {
        int piv = k;
        double best = fabs(M[k][k]);
        for (int r = k + 1; r < 3; ++r)
            if (fabs(M[r][k]) > best) { best = fabs(M[r][k]); piv = r; }
        if (piv != k)
            for (int c = k; c < 4; ++c) { const double t = M[k][c]; M[k][c] = M[piv][c]; M[piv][c] = t; }
        for (int r = k + 1; r < 3; ++r) {
            const double f = M[r][k] / M[k][k];
            for (int c = k; c < 4; ++c) M[r][c] -= f * M[k][c];
        }
    }
    for (int r = 2; r >= 0; --r) {
        double s = M[r][3];
        for (int c = r + 1; c < 3; ++c) s -= M[r][c] * d[c];
        d[r] = s / M[r][r];
    }
}

// One CTA per frame f = 1 + blockIdx.x.  Points that are alive at frame f (and in `subset`, if given) enter the fit.
// 3 CTAs per SM: the 299 frames of a C3 sequence then run as ONE wave on 148 SMs (at 2 per SM, 296 slots, three CTAs ran alone afterwards)
__global__ void __launch_bounds__(POSE_THREADS, 3)
seq_pose_t_kernel(const double* __restrict__ Kg, const float2* __restrict__ tracks, const uint8_t* __restrict__ alive,
                  const uint8_t* __restrict__ subset, const double* __restrict__ p3, int npts, double x0a, double x0b, double x0c,
                  const float* B0, float* B, float* __restrict__ S, float2* __restrict__ proj,
                  int32_t* __restrict__ iters)
{
    constexpr int NACC = 9, NW = POSE_THREADS / 32;
    __shared__ double sK[9];
    __shared__ double sx[3];
    __shared__ double sred[NW][NACC];
    __shared__ int s_done;
    __shared__ float st32[3];

    const int f = 1 + blockIdx.x, tid = threadIdx.x;
    const float2* P = tracks + (size_t)f * npts;
    const uint8_t* A = alive + (size_t)f * npts;
    if (tid < 9) sK[tid] = Kg[tid];
    if (tid == 0) { sx[0] = x0a; sx[1] = x0b; sx[2] = x0c; s_done = 0; }
    __syncthreads();

    const double dx = 1e-6, inv_dx = 1e6;
    int it = 0;
    for (; it < POSE_MAX_ITER; ++it) {
        double acc[NACC];
#pragma unroll
        for (int k = 0; k < NACC; ++k) acc[k] = 0.0;
        for (int i = tid; i < npts; i += POSE_THREADS) {
            if (A[i] == 0 || (subset && subset[i] == 0)) continue;
            const float2 pz = P[i];
            const double bx = p3[3 * i] + sx[0], by = p3[3 * i + 1] + sx[1], bz = p3[3 * i + 2] + sx[2];
            double u0, v0, u, v, ju[3], jv[3];
            project_rk(sK, bx, by, bz, u0, v0);
            // (u - u0) * 1e6 for the reference's (u - u0) / 1e-6 (utils/NLS.py:113-118): 1e6 is exact, the quotient differs in the last bit
            project_rk(sK, bx + dx, by, bz, u, v); ju[0] = (u - u0) * inv_dx; jv[0] = (v - v0) * inv_dx;
            project_rk(sK, bx, by + dx, bz, u, v); ju[1] = (u - u0) * inv_dx; jv[1] = (v - v0) * inv_dx;
            project_rk(sK, bx, by, bz + dx, u, v); ju[2] = (u - u0) * inv_dx; jv[2] = (v - v0) * inv_dx;
            const double ru = (double)pz.x - u0, rv = (double)pz.y - v0;
            int k = 0;
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int c = r; c < 3; ++c) acc[k++] += ju[r] * ju[c] + jv[r] * jv[c];
#pragma unroll
            for (int r = 0; r < 3; ++r) acc[6 + r] += ju[r] * ru + jv[r] * rv;
        }
#pragma unroll
        for (int k = 0; k < NACC; ++k) acc[k] = warp_sum_d(acc[k]);
        if ((tid & 31) == 0) {
#pragma unroll
            for (int k = 0; k < NACC; ++k) sred[tid >> 5][k] = acc[k];
        }
        __syncthreads();
        if (tid == 0) {
            double tot[NACC], d[3];
            for (int k = 0; k < NACC; ++k) {
                double s = 0.0;
                for (int w = 0; w < NW; ++w) s += sred[w][k];
                tot[k] = s;
            }
            solve3_damped(tot, tot + 6, d);
            const double ramp = (double)(it + 1) * 0.2;
            const double sc = fmin(ramp * ramp, 1.0);
            double ss = 0.0;
            for (int r = 0; r < 3; ++r) {
                d[r] *= sc;
                sx[r] = sx[r] + d[r];
                ss += d[r] * d[r];
            }
            if (sqrt(ss / 3.0) < 1e-8) s_done = 1;
        }
        __syncthreads();
        if (s_done) break;
    }
    // fcnNLS_t returns float32 (utils/NLS.py:129); the reprojection uses that rounded t (utils/NLS.py:31-32)
    if (tid < 3) st32[tid] = (float)sx[tid];
    __syncthreads();
    const double t0 = (double)st32[0], t1 = (double)st32[1], t2 = (double)st32[2];
    double ss = 0.0;
    int cnt_fit = 0, cnt_alive = 0;
    for (int i = tid; i < npts; i += POSE_THREADS) {
        const bool a = A[i] != 0;
        cnt_alive += a ? 1 : 0;
        float2 pr = make_float2(__int_as_float(0x7fc00000), __int_as_float(0x7fc00000));
        if (a && !(subset && subset[i] == 0)) {
            double u, v;
            project_rk(sK, p3[3 * i] + t0, p3[3 * i + 1] + t1, p3[3 * i + 2] + t2, u, v);
            const float2 pz = P[i];
            const double ru = (double)pz.x - u, rv = (double)pz.y - v;
            ss += ru * ru + rv * rv;
            ++cnt_fit;
            pr = make_float2((float)u, (float)v);
        }
        if (proj) proj[(size_t)f * npts + i] = pr;
    }
    ss = warp_sum_d(ss);
    cnt_fit = __reduce_add_sync(0xffffffffu, cnt_fit);
    cnt_alive = __reduce_add_sync(0xffffffffu, cnt_alive);
    __shared__ double s_ss[NW];
    __shared__ int s_cf[NW], s_ca[NW];
    if ((tid & 31) == 0) { s_ss[tid >> 5] = ss; s_cf[tid >> 5] = cnt_fit; s_ca[tid >> 5] = cnt_alive; }
    __syncthreads();
    if (tid == 0) {
        double tot = 0.0;
        int cf = 0, ca = 0;
        for (int w = 0; w < NW; ++w) { tot += s_ss[w]; cf += s_cf[w]; ca += s_ca[w]; }
        float* Bf = B + (size_t)f * 14;
        for (int k = 0; k < 3; ++k) {
            Bf[3 + k] = st32[k];                    // B[i, 3:6] = t
            Bf[k] = __fadd_rn(B0[k], st32[k]);      // B[i, 0:3] = B[0, 0:3] + t   (float32 array arithmetic)
        }
        S[(size_t)f * 9 + 2] = (float)ca;                                   // vg.sum()
        S[(size_t)f * 9 + 3] = cf > 0 ? (float)sqrt(tot / (2.0 * cf)) : __int_as_float(0x7fc00000);   // rms(p - p_proj)
        iters[f] = s_done ? it + 1 : -POSE_MAX_ITER;
    }
}

// frame 0: P[2:4, vp, 0] = p_.T with p_ = p[vp] (vidExample.py:125,152): the reprojection row of frame 0 is the seeds
__global__ void seq_proj0_kernel(const float2* __restrict__ tracks, const uint8_t* __restrict__ alive, const uint8_t* __restrict__ subset,
                                 int npts, float2* __restrict__ proj)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npts) return;
    const float nanf_ = __int_as_float(0x7fc00000);
    proj[i] = (alive[i] != 0 && !(subset && subset[i] == 0)) ? tracks[i] : make_float2(nanf_, nanf_);
}

// frame 0 rows + everything that chains frames: dt, dr, cumulative distance, speed (vidExample.py:142-146,164)
__global__ void seq_stats_kernel(const float* __restrict__ B, const uint8_t* __restrict__ alive, int nframes, int npts,
                                 float* __restrict__ S)
{
    const int tid = threadIdx.x;
    const float nanf_ = __int_as_float(0x7fc00000);
    __shared__ int s_cnt[32];
    // frame 0: track count
    int c = 0;
    for (int i = tid; i < npts; i += blockDim.x) c += alive[i] != 0 ? 1 : 0;
    c = __reduce_add_sync(0xffffffffu, c);
    if ((tid & 31) == 0) s_cnt[tid >> 5] = c;
    __syncthreads();
    if (tid == 0) {
        int tot = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += s_cnt[w];
        S[2] = (float)tot;
    }
    const float t0 = B[12];
    for (int i = tid; i < nframes; i += blockDim.x) {
        float* Si = S + (size_t)i * 9;
        Si[0] = (float)i;
        Si[5] = __fsub_rn(B[(size_t)i * 14 + 12], t0);
        if (i == 0) {
            Si[4] = nanf_; Si[6] = 0.f; Si[8] = nanf_;   // dt = nan, dr = 0, speed = 0/nan*3.6
        } else {
            const float* Bi = B + (size_t)i * 14;
            const float* Bp = Bi - 14;
            const float dt = __fsub_rn(Bi[12], Bp[12]);
            float s2 = 0.f;
            for (int k = 0; k < 3; ++k) {   // norm(t + B[0,0:3] - B[i-1,0:3]), float32
                const float d = __fsub_rn(__fadd_rn(Bi[3 + k], B[k]), Bp[k]);
                s2 = __fadd_rn(s2, __fmul_rn(d, d));
            }
            const float dr = __fsqrt_rn(s2);
            Si[4] = dt;
            Si[6] = dr;
            Si[8] = __fmul_rn(__fdiv_rn(dr, dt), 3.6f);
        }
    }
    __syncthreads();
    if (tid == 0) {   // r += dr: a sequential float32 sum in the reference
        float r = 0.f;
        S[7] = 0.f;
        for (int i = 1; i < nframes; ++i) {
            r = __fadd_rn(r, S[(size_t)i * 9 + 6]);
            S[(size_t)i * 9 + 7] = r;
        }
    }
}

// order-preserving compaction of the tracks that are alive in `alive_last` (and in subset): one CTA, chunked scan
__global__ void seq_select_kernel(const uint8_t* __restrict__ alive_last, const uint8_t* __restrict__ subset, int npts,
                                  int32_t* __restrict__ idx, int32_t* __restrict__ count)
{
    __shared__ int s_warp[32];
    __shared__ int s_base;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nw = blockDim.x >> 5;
    if (tid == 0) s_base = 0;
    __syncthreads();
    for (int start = 0; start < npts; start += blockDim.x) {
        const int i = start + tid;
        const bool keep = i < npts && alive_last[i] != 0 && !(subset && subset[i] == 0);
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        const int before = __popc(m & ((1u << lane) - 1u));
        if (lane == 0) s_warp[wid] = __popc(m);
        __syncthreads();
        int off = s_base;
        for (int w = 0; w < wid; ++w) off += s_warp[w];
        if (keep) idx[off + before] = i;
        __syncthreads();
        if (tid == 0) {
            int t = 0;
            for (int w = 0; w < nw; ++w) t += s_warp[w];
            s_base += t;
        }
        __syncthreads();
    }
    if (tid == 0) *count = s_base;
}

// U[c][f][j] = pixel2uvec(K, p)  (utils/common.py:122-126):  q = (p - K[2,0:2], K[0,0]) / |q|
__global__ void seq_rays_kernel(const double* __restrict__ Kg, const float2* __restrict__ tracks, const int32_t* __restrict__ idx,
                                int nframes, int npts, int nsel, double* __restrict__ U)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x, f = blockIdx.y;
    if (j >= nsel) return;
    const float2 p = tracks[(size_t)f * npts + (idx ? idx[j] : j)];
    const double qx = (double)p.x - Kg[6], qy = (double)p.y - Kg[7], qz = Kg[0];
    const double n = sqrt(qx * qx + qy * qy + qz * qz);
    const size_t plane = (size_t)nframes * nsel;
    double* o = U + (size_t)f * nsel + j;
    o[0] = qx / n;
    o[plane] = qy / n;
    o[2 * plane] = qz / n;
}

// z [2][n][nsel] and x0 = [pw | cw[1:] | 0]
__global__ void seq_pack_z_kernel(const float2* __restrict__ tracks, const int32_t* __restrict__ idx, int nframes, int npts, int nsel,
                                  double* __restrict__ z)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x, f = blockIdx.y;
    if (j >= nsel) return;
    const float2 p = tracks[(size_t)f * npts + (idx ? idx[j] : j)];
    z[(size_t)f * nsel + j] = (double)p.x;
    z[(size_t)nframes * nsel + (size_t)f * nsel + j] = (double)p.y;
}

__global__ void seq_pack_x_kernel(const double* __restrict__ pw, const float* __restrict__ B, int nframes, int nsel, double* __restrict__ x)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int nc = nframes - 1, np3 = 3 * nsel;
    if (i < np3) x[i] = pw[i];
    else if (i < np3 + 3 * nc) {
        const int c = (i - np3) / 3, k = (i - np3) % 3;
        x[i] = (double)B[(size_t)(c + 1) * 14 + 3 + k];    // cw = B[:, 3:6]
    } else if (i < np3 + 6 * nc) x[i] = 0.0;
}

// A[f] = -(B[f,3:6])  (ray origins of fcnNvintercept: u0 = B[0,0:3] - B[:nf,0:3], utils/MSV.py:16)
__global__ void seq_origins_kernel(const float* __restrict__ B, int nframes, double* __restrict__ A)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 3 * nframes) return;
    const int f = i / 3, k = i % 3;
    A[i] = (double)__fsub_rn(B[k], B[(size_t)f * 14 + k]);
}

// B_ba = B with B[:, 3:6] = cw (the bundle-adjusted camera positions, vidExample.py:157) and B[:, 0:3] = B[0, 0:3] + cw;
// S_ba starts as a copy of S (track counts, residuals), its chained columns are then rewritten by seq_stats_kernel
__global__ void seq_ba_cameras_kernel(const double* __restrict__ x, int nsel, int nframes, const float* __restrict__ B,
                                      const float* __restrict__ S, float* __restrict__ B_ba, float* __restrict__ S_ba)
{
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nframes) return;
    const float* Bi = B + (size_t)f * 14;
    float* Bo = B_ba + (size_t)f * 14;
    for (int k = 6; k < 14; ++k) Bo[k] = Bi[k];
    for (int k = 0; k < 3; ++k) {
        const float c = f == 0 ? 0.f : (float)x[(size_t)3 * nsel + 3 * (size_t)(f - 1) + k];
        Bo[3 + k] = c;
        Bo[k] = __fadd_rn(B[k], c);
    }
    for (int k = 0; k < 9; ++k) S_ba[(size_t)f * 9 + k] = S[(size_t)f * 9 + k];
}

// P[r][i][f] (float32, NaN = invalid): tiled transpose of the frame-major device arrays
__global__ void seq_export_P_kernel(const float2* __restrict__ tracks, const float2* __restrict__ proj, const uint8_t* __restrict__ alive,
                                    int nframes, int npts, float* __restrict__ P)
{
    __shared__ float tile[5][32][33];
    const float nanf_ = __int_as_float(0x7fc00000);
    const int i0 = blockIdx.x * 32, f0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int f = f0 + r, i = i0 + threadIdx.x;
        float v[5] = {nanf_, nanf_, nanf_, nanf_, nanf_};
        if (f < nframes && i < npts && alive[(size_t)f * npts + i] != 0) {
            const float2 p = tracks[(size_t)f * npts + i];
            v[0] = p.x; v[1] = p.y; v[4] = (float)f;
            if (proj) { const float2 q = proj[(size_t)f * npts + i]; v[2] = q.x; v[3] = q.y; }
        }
#pragma unroll
        for (int k = 0; k < 5; ++k) tile[k][r][threadIdx.x] = v[k];
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int i = i0 + r, f = f0 + threadIdx.x;
        if (i < npts && f < nframes) {
#pragma unroll
            for (int k = 0; k < 5; ++k) P[((size_t)k * npts + i) * nframes + f] = tile[k][threadIdx.x][r];
        }
    }
}

}  // namespace

VEL_API int vel_klt_sequence(const uint8_t* frames, int64_t frame_stride, int32_t pitch, const uint8_t* pyr, int64_t pyr_stride,
                             const vel_pyr_layout* layout, int32_t nframes, int32_t npts, const vel_lk_params* params, float* tracks,
                             uint8_t* alive, float* err, uint8_t* status, vel_stream_t stream)
{
    VEL_CHECK_ARG(nframes >= 1 && npts >= 0, "vel_klt_sequence: nframes %d / npts %d", nframes, npts);
    if (npts == 0 || nframes == 1) return VEL_OK;
    VEL_CHECK_ARG(frames && layout && params && tracks && alive && err && status, "vel_klt_sequence: NULL argument");
    cudaStream_t st = (cudaStream_t)stream;
    const int nb = (npts + 255) / 256;
    seq_seed_kernel<<<nb, 256, 0, st>>>(alive, (float2*)tracks, npts);
    // the reference's window (15x15): each half-warp carries its track through every frame inside ONE kernel
    const int rc_seq = vel_lk_sequence_w15h(frames, frame_stride, pitch, pyr, pyr_stride, layout, nframes, npts, params, tracks, alive, err, stream);
    if (rc_seq < 0) return rc_seq;
    if (rc_seq == 1) return VEL_OK;
    // any other window: one K2 launch per pair, the track state chained on the device
    for (int k = 0; k + 1 < nframes; ++k) {
        float* prev = tracks + (size_t)k * npts * 2;
        float* next = prev + (size_t)npts * 2;
        int rc = vel_lk_track(frames + (size_t)k * frame_stride, 0, pitch, pyr ? pyr + (size_t)k * pyr_stride : NULL, 0,
                              frames + (size_t)(k + 1) * frame_stride, 0, pitch, pyr ? pyr + (size_t)(k + 1) * pyr_stride : NULL, 0,
                              layout, 1, prev, 0, npts, params, next, status, err + (size_t)k * npts, NULL, stream);
        if (rc != VEL_OK) return rc;
        seq_propagate_kernel<<<nb, 256, 0, st>>>(alive + (size_t)k * npts, status, alive + (size_t)(k + 1) * npts, (float2*)next, npts);
    }
    VEL_LAUNCH_CHECK("seq_propagate_kernel");
    return VEL_OK;
}

VEL_API int vel_seq_pose_t(const double* K, const float* tracks, const uint8_t* alive, const uint8_t* subset, const double* p3,
                           int32_t nframes, int32_t npts, const double* x0_host, float* B, float* S, float* proj, int32_t* iters,
                           vel_stream_t stream)
{
    VEL_CHECK_ARG(K && tracks && alive && p3 && x0_host && B && S && iters, "vel_seq_pose_t: NULL argument");
    VEL_CHECK_ARG(nframes >= 1 && npts >= 1, "vel_seq_pose_t: nframes %d / npts %d", nframes, npts);
    if (proj) seq_proj0_kernel<<<(npts + 255) / 256, 256, 0, (cudaStream_t)stream>>>((const float2*)tracks, alive, subset, npts, (float2*)proj);
    if (nframes == 1) return VEL_OK;
    seq_pose_t_kernel<<<nframes - 1, POSE_THREADS, 0, (cudaStream_t)stream>>>(K, (const float2*)tracks, alive, subset, p3, npts, x0_host[0],
                                                                            x0_host[1], x0_host[2], B, B, S, (float2*)proj, iters);
    VEL_LAUNCH_CHECK("seq_pose_t_kernel");
    return VEL_OK;
}

VEL_API int vel_seq_stats(const float* B, const uint8_t* alive, int32_t nframes, int32_t npts, float* S, vel_stream_t stream)
{
    VEL_CHECK_ARG(B && alive && S && nframes >= 1 && npts >= 0, "vel_seq_stats: bad argument");
    seq_stats_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(B, alive, nframes, npts, S);
    VEL_LAUNCH_CHECK("seq_stats_kernel");
    return VEL_OK;
}

VEL_API int vel_seq_select(const uint8_t* alive_last, const uint8_t* subset, int32_t npts, int32_t* idx, int32_t* count, vel_stream_t stream)
{
    VEL_CHECK_ARG(alive_last && idx && count && npts >= 0, "vel_seq_select: bad argument");
    seq_select_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(alive_last, subset, npts, idx, count);
    VEL_LAUNCH_CHECK("seq_select_kernel");
    return VEL_OK;
}

VEL_API int vel_seq_rays(const double* K, const float* tracks, const int32_t* idx, int32_t nframes, int32_t npts, int32_t nsel,
                         const float* B, double* U, double* A, vel_stream_t stream)
{
    VEL_CHECK_ARG(K && tracks && U && nframes >= 1 && npts >= 1 && nsel >= 0 && nsel <= npts, "vel_seq_rays: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    if (nsel > 0) {
        dim3 grid((nsel + 255) / 256, nframes);
        seq_rays_kernel<<<grid, 256, 0, st>>>(K, (const float2*)tracks, idx, nframes, npts, nsel, U);
    }
    if (A) {
        VEL_CHECK_ARG(B, "vel_seq_rays: origins requested without B");
        seq_origins_kernel<<<(3 * nframes + 255) / 256, 256, 0, st>>>(B, nframes, A);
    }
    VEL_LAUNCH_CHECK("seq_rays_kernel");
    return VEL_OK;
}

VEL_API int vel_seq_pack_ba(const float* tracks, const int32_t* idx, int32_t nframes, int32_t npts, int32_t nsel, const double* pw,
                            const float* B, double* z, double* x, vel_stream_t stream)
{
    VEL_CHECK_ARG(tracks && pw && B && z && x && nframes >= 1 && npts >= 1 && nsel >= 1 && nsel <= npts, "vel_seq_pack_ba: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    dim3 grid((nsel + 255) / 256, nframes);
    seq_pack_z_kernel<<<grid, 256, 0, st>>>((const float2*)tracks, idx, nframes, npts, nsel, z);
    const int nx = 3 * nsel + 6 * (nframes - 1);
    seq_pack_x_kernel<<<(nx + 255) / 256, 256, 0, st>>>(pw, B, nframes, nsel, x);
    VEL_LAUNCH_CHECK("seq_pack_ba kernels");
    return VEL_OK;
}

VEL_API int vel_seq_export_P(const float* tracks, const float* proj, const uint8_t* alive, int32_t nframes, int32_t npts, float* P,
                             vel_stream_t stream)
{
    VEL_CHECK_ARG(tracks && alive && P && nframes >= 1 && npts >= 1, "vel_seq_export_P: bad argument");
    dim3 grid((npts + 31) / 32, (nframes + 31) / 32), block(32, 8);
    seq_export_P_kernel<<<grid, block, 0, (cudaStream_t)stream>>>((const float2*)tracks, (const float2*)proj, alive, nframes, npts, P);
    VEL_LAUNCH_CHECK("seq_export_P_kernel");
    return VEL_OK;
}

VEL_API int vel_seq_ba_cameras(const double* x, int32_t nsel, int32_t nframes, const float* B, const float* S, float* B_ba, float* S_ba,
                               vel_stream_t stream)
{
    VEL_CHECK_ARG(x && B && S && B_ba && S_ba && nsel >= 1 && nframes >= 1, "vel_seq_ba_cameras: bad argument");
    seq_ba_cameras_kernel<<<(nframes + 127) / 128, 128, 0, (cudaStream_t)stream>>>(x, nsel, nframes, B, S, B_ba, S_ba);
    VEL_LAUNCH_CHECK("seq_ba_cameras_kernel");
    return VEL_OK;
}
