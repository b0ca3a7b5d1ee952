// K10: robust affine fit -- cv2.estimateAffine2D(from, to, method=cv2.RANSAC) as the tracker calls it after every
// LK stage (utils/KLT.py:116,127 and :33), with cv2's defaults (3 px, 2000 iterations,
// confidence 0.99, refinement on the inliers).
//
// OpenCV's algorithm (calib3d ptsetreg.cpp, restated in oracle/ransac_oracle.py and pinned against cv2 4.13): a
// sequential hypothesise-and-verify loop driven by cv::RNG(2^64-1).  One CTA runs it:
//   * thread 0 draws the 3-point subset (multiply-with-carry RNG, re-draw on repeats, collinearity rejection) and
//     solves the closed-form affine in float64 -- every product and sum with explicit rounding intrinsics, because an
//     FMA contraction would change a residual by an ulp and, now and then, an inlier decision;
//   * all threads evaluate the float32 residuals of their points against float32(thr^2) and the CTA counts the inliers
//     (integer sums: deterministic);
//   * an improving model's inlier flags are written out and the iteration budget shrinks exactly as in
//     RANSACUpdateNumIters (double log/pow; the budget is an integer, so a last-ulp difference between CUDA's and
//     glibc's log cannot show unless num/denom sits within 1e-15 of a half-integer).
// The refinement cv2 performs (10 Levenberg-Marquardt steps from the kept model over the inliers) converges to the
// least-squares affine of the inliers -- the problem is linear -- so it is computed directly: centred float64 normal
// equations, fixed-order reductions.  Inlier masks are bit-identical to cv2; T agrees to ~1e-12 (identical after the
// float32 cast KLTregional applies, utils/KLT.py:58).
#include "common.cuh"
#include <float.h>

namespace {

constexpr int RS_THREADS = 512;
constexpr int RS_MAX_PER_THREAD = 16;              // up to 8192 correspondences

struct RsShared {
    float F[6];
    int go;                                         // 1: evaluate F, 0: stop
    int new_best;
    int warp_count[RS_THREADS / 32];
    double red[RS_THREADS / 32][7];
    double stats[11];
};

__device__ __forceinline__ unsigned rng_next(unsigned long long& s)
{
    s = (unsigned long long)(unsigned)s * 4164903690ull + (unsigned)(s >> 32);
    return (unsigned)s;
}

__device__ __forceinline__ bool third_collinear(const float2 p0, const float2 p1, const float2 p2)
{
    const double dx1 = (double)__fsub_rn(p1.x, p2.x), dy1 = (double)__fsub_rn(p1.y, p2.y);
    const double dx2 = (double)__fsub_rn(p0.x, p2.x), dy2 = (double)__fsub_rn(p0.y, p2.y);
    const double lhs = fabs(__dsub_rn(__dmul_rn(dx2, dy1), __dmul_rn(dy2, dx1)));
    const double rhs = __dmul_rn((double)FLT_EPSILON, __dadd_rn(__dadd_rn(__dadd_rn(fabs(dx1), fabs(dy1)), fabs(dx2)), fabs(dy2)));
    return lhs <= rhs;
}

// a*b + c*d + e*f, left to right, unfused
__device__ __forceinline__ double dot3(double a, double b, double c, double d, double e, double f)
{
    return __dadd_rn(__dadd_rn(__dmul_rn(a, b), __dmul_rn(c, d)), __dmul_rn(e, f));
}

__device__ void affine_from_3(const float2* fr, const float2* to, double (&M)[6])
{
    const double x1 = fr[0].x, y1 = fr[0].y, x2 = fr[1].x, y2 = fr[1].y, x3 = fr[2].x, y3 = fr[2].y;
    const double X1 = to[0].x, Y1 = to[0].y, X2 = to[1].x, Y2 = to[1].y, X3 = to[2].x, Y3 = to[2].y;
    const double y23 = __dsub_rn(y2, y3), y31 = __dsub_rn(y3, y1), y12 = __dsub_rn(y1, y2);
    const double x32 = __dsub_rn(x3, x2), x13 = __dsub_rn(x1, x3), x21 = __dsub_rn(x2, x1);
    const double c1 = __dsub_rn(__dmul_rn(x2, y3), __dmul_rn(x3, y2)), c2 = __dsub_rn(__dmul_rn(x3, y1), __dmul_rn(x1, y3)),
                 c3 = __dsub_rn(__dmul_rn(x1, y2), __dmul_rn(x2, y1));
    const double d = __ddiv_rn(1., dot3(x1, y23, x2, y31, x3, y12));
    M[0] = __dmul_rn(d, dot3(X1, y23, X2, y31, X3, y12));
    M[1] = __dmul_rn(d, dot3(X1, x32, X2, x13, X3, x21));
    M[2] = __dmul_rn(d, dot3(X1, c1, X2, c2, X3, c3));
    M[3] = __dmul_rn(d, dot3(Y1, y23, Y2, y31, Y3, y12));
    M[4] = __dmul_rn(d, dot3(Y1, x32, Y2, x13, Y3, x21));
    M[5] = __dmul_rn(d, dot3(Y1, c1, Y2, c2, Y3, c3));
}

__device__ int update_iters(double p, double ep, int model_points, int max_iters)
{
    p = fmin(fmax(p, 0.), 1.);
    ep = fmin(fmax(ep, 0.), 1.);
    double num = fmax(1. - p, DBL_MIN);
    double denom = 1. - pow(1. - ep, (double)model_points);
    if (denom < DBL_MIN) return 0;
    num = log(num);
    denom = log(denom);
    return (denom >= 0 || -num >= max_iters * (-denom)) ? max_iters : (int)rint(num / denom);
}

// fixed-order CTA sum of K doubles per thread; every thread receives the totals
template <int K>
__device__ void block_sum(double (&v)[K], RsShared& S, int tid)
{
#pragma unroll
    for (int k = 0; k < K; ++k)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
    __syncthreads();
    if ((tid & 31) == 0)
#pragma unroll
        for (int k = 0; k < K; ++k) S.red[tid >> 5][k] = v[k];
    __syncthreads();
#pragma unroll
    for (int k = 0; k < K; ++k) {
        double s = 0.;
        for (int w = 0; w < RS_THREADS / 32; ++w) s += S.red[w][k];
        v[k] = s;
    }
}

// the whole CTA; from / to may have been written earlier in the same kernel (the masked form compacts into scratch): plain loads
__device__ void ransac_affine2d_body(RsShared& S, const float2* from, const float2* to, int n, double thresh, double conf, int max_iters,
                                     int refine, uint8_t* inliers, double* T, int* info)
{
    const int tid = threadIdx.x;
    float2 f[RS_MAX_PER_THREAD], t[RS_MAX_PER_THREAD];
#pragma unroll
    for (int q = 0; q < RS_MAX_PER_THREAD; ++q) {
        const int i = tid + q * RS_THREADS;
        f[q] = i < n ? from[i] : make_float2(0.f, 0.f);
        t[q] = i < n ? to[i] : make_float2(0.f, 0.f);
    }
    for (int i = tid; i < n; i += RS_THREADS) inliers[i] = 0;
    const float thr2 = (float)(thresh * thresh);

    // thread-0 state of the sequential loop
    unsigned long long rng = 0xFFFFFFFFFFFFFFFFull;
    int niters = max(max_iters, 1), iter = 0, max_good = 0;
    double best[6] = {0., 0., 0., 0., 0., 0.};

    if (n == 3) {            // cv2: exactly three correspondences are fitted directly, all inliers, no refinement
        if (tid == 0) {
            const float2 a[3] = {from[0], from[1], from[2]}, b[3] = {to[0], to[1], to[2]};
            affine_from_3(a, b, best);
            for (int k = 0; k < 6; ++k) T[k] = best[k];
            inliers[0] = inliers[1] = inliers[2] = 1;
            info[0] = 1; info[1] = 3; info[2] = 0;
        }
        return;
    }

    for (;;) {
        if (tid == 0) {
            int go = 0;
            if (iter < niters) {
                for (int attempt = 0; attempt < 10000 && !go; ++attempt) {
                    int idx[3];
                    for (int i = 0; i < 3; ++i) {
                        int v;
                        bool rep;
                        do {
                            v = (int)(rng_next(rng) % (unsigned)n);
                            rep = false;
                            for (int k = 0; k < i; ++k) rep |= (idx[k] == v);
                        } while (rep);
                        idx[i] = v;
                    }
                    const float2 a[3] = {from[idx[0]], from[idx[1]], from[idx[2]]}, b[3] = {to[idx[0]], to[idx[1]], to[idx[2]]};
                    if (third_collinear(a[0], a[1], a[2]) || third_collinear(b[0], b[1], b[2])) continue;
                    double M[6];
                    affine_from_3(a, b, M);
                    for (int k = 0; k < 6; ++k) { S.F[k] = (float)M[k]; S.stats[k] = M[k]; }
                    go = 1;
                }
            }
            S.go = go;
        }
        __syncthreads();
        if (!S.go) break;
        const float F0 = S.F[0], F1 = S.F[1], F2 = S.F[2], F3 = S.F[3], F4 = S.F[4], F5 = S.F[5];
        unsigned flags = 0;
#pragma unroll
        for (int q = 0; q < RS_MAX_PER_THREAD; ++q) {
            if (tid + q * RS_THREADS < n) {
                const float a = __fsub_rn(__fadd_rn(__fadd_rn(__fmul_rn(F0, f[q].x), __fmul_rn(F1, f[q].y)), F2), t[q].x);
                const float b = __fsub_rn(__fadd_rn(__fadd_rn(__fmul_rn(F3, f[q].x), __fmul_rn(F4, f[q].y)), F5), t[q].y);
                const float e = __fadd_rn(__fmul_rn(a, a), __fmul_rn(b, b));
                flags |= (e <= thr2 ? 1u : 0u) << q;
            }
        }
        int cnt = __popc(flags);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        if ((tid & 31) == 0) S.warp_count[tid >> 5] = cnt;
        __syncthreads();
        if (tid == 0) {
            int total = 0;
            for (int w = 0; w < RS_THREADS / 32; ++w) total += S.warp_count[w];
            S.new_best = 0;
            if (total > max(max_good, 2)) {
                S.new_best = 1;
                max_good = total;
                for (int k = 0; k < 6; ++k) best[k] = S.stats[k];
                niters = update_iters(conf, (double)(n - total) / n, 3, niters);
            }
            ++iter;
        }
        __syncthreads();
        if (S.new_best) {
#pragma unroll
            for (int q = 0; q < RS_MAX_PER_THREAD; ++q) {
                const int i = tid + q * RS_THREADS;
                if (i < n) inliers[i] = (uint8_t)((flags >> q) & 1u);
            }
        }
    }

    // ---- refinement on the inliers: centred least squares (what cv2's LM iterations converge to) -------------------------
    if (tid == 0) { S.stats[0] = (double)max_good; for (int k = 0; k < 6; ++k) S.stats[1 + k] = best[k]; S.stats[7] = (double)iter; }
    __syncthreads();
    const int ninl = (int)S.stats[0];
    if (ninl == 0) {
        if (tid == 0) { info[0] = 0; info[1] = 0; info[2] = (int)S.stats[7]; }
        return;
    }
    __syncthreads();
    unsigned flags = 0;
#pragma unroll
    for (int q = 0; q < RS_MAX_PER_THREAD; ++q) {
        const int i = tid + q * RS_THREADS;
        if (i < n && inliers[i]) flags |= 1u << q;
    }
    double Tm[6];
    for (int k = 0; k < 6; ++k) Tm[k] = S.stats[1 + k];
    if (refine && n > 3) {
        double m[4] = {0., 0., 0., 0.};
#pragma unroll
        for (int q = 0; q < RS_MAX_PER_THREAD; ++q)
            if ((flags >> q) & 1u) { m[0] += f[q].x; m[1] += f[q].y; m[2] += t[q].x; m[3] += t[q].y; }
        block_sum<4>(m, S, tid);
        const double mx = m[0] / ninl, my = m[1] / ninl, mX = m[2] / ninl, mY = m[3] / ninl;
        double s[7] = {0., 0., 0., 0., 0., 0., 0.};     // Sxx Sxy Syy SxX SyX SxY SyY
#pragma unroll
        for (int q = 0; q < RS_MAX_PER_THREAD; ++q)
            if ((flags >> q) & 1u) {
                const double x = f[q].x - mx, y = f[q].y - my, X = t[q].x - mX, Y = t[q].y - mY;
                s[0] += x * x; s[1] += x * y; s[2] += y * y; s[3] += x * X; s[4] += y * X; s[5] += x * Y; s[6] += y * Y;
            }
        block_sum<7>(s, S, tid);
        const double det = s[0] * s[2] - s[1] * s[1];
        if (fabs(det) > 0.) {
            const double a00 = (s[3] * s[2] - s[4] * s[1]) / det, a01 = (s[4] * s[0] - s[3] * s[1]) / det;
            const double a10 = (s[5] * s[2] - s[6] * s[1]) / det, a11 = (s[6] * s[0] - s[5] * s[1]) / det;
            Tm[0] = a00; Tm[1] = a01; Tm[2] = mX - (a00 * mx + a01 * my);
            Tm[3] = a10; Tm[4] = a11; Tm[5] = mY - (a10 * mx + a11 * my);
        }
    }
    if (tid == 0) {
        for (int k = 0; k < 6; ++k) T[k] = Tm[k];
        info[0] = 1; info[1] = ninl; info[2] = (int)S.stats[7];
    }
}

__global__ void __launch_bounds__(RS_THREADS)
ransac_affine2d_kernel(const float2* __restrict__ from, const float2* __restrict__ to, int n, double thresh, double conf, int max_iters,
                       int refine, uint8_t* __restrict__ inliers, double* __restrict__ T, int* __restrict__ info)
{
    __shared__ RsShared S;
    ransac_affine2d_body(S, from, to, n, thresh, conf, max_iters, refine, inliers, T, info);
}

// The tracker's own call shape (utils/KLT.py:116-117, :127): `T, inl = estimateAffine2D(p0[v], p[v]); v[v] = inl` with v the
// status mask of the LK stage before it -- compaction, fit and scatter in one launch, so the LK output never leaves the device:
// from_c / to_c = the rows where mask != 0 in order (to first mapped as to * to_scale + to_off in float32: the `p /= scale` of
// :114 and the map-back of the translated ROI, :89), the fit on those, mask_out[i] = mask[i] & inlier.  info[3] = rows kept.
__global__ void __launch_bounds__(RS_THREADS)
ransac_affine2d_masked_kernel(const float2* __restrict__ from, const float2* __restrict__ to, const uint8_t* mask, int n_full,
                              float to_scale, float to_off_x, float to_off_y, double thresh, double conf, int max_iters, int refine,
                              float2* cf, float2* ct, int* cidx, uint8_t* cinl, uint8_t* mask_out, float2* __restrict__ to_out,
                              double* __restrict__ T, int* __restrict__ info)
{
    __shared__ RsShared S;
    __shared__ int wsum[RS_THREADS / 32];
    __shared__ int s_total;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int per = (n_full + RS_THREADS - 1) / RS_THREADS, lo = min(tid * per, n_full), hi = min(lo + per, n_full);
    int cnt = 0;
    for (int i = lo; i < hi; ++i) cnt += mask[i] != 0;
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    if (tid == 0) {
        int run = 0;
        for (int w = 0; w < RS_THREADS / 32; ++w) { const int v = wsum[w]; wsum[w] = run; run += v; }
        s_total = run;
    }
    __syncthreads();
    int pos = wsum[warp] + incl - cnt;
    for (int i = lo; i < hi; ++i) {
        const float2 tv = to[i];
        const float2 tm = make_float2(__fadd_rn(__fmul_rn(tv.x, to_scale), to_off_x), __fadd_rn(__fmul_rn(tv.y, to_scale), to_off_y));
        if (to_out) to_out[i] = tm;                       // the mapped points of ALL rows (the caller's `p`)
        if (mask[i] != 0) { cf[pos] = from[i]; ct[pos] = tm; cidx[pos] = i; ++pos; }
    }
    const int n = s_total;
    __syncthreads();
    if (tid == 0) info[3] = n;
    if (n < 3) {                                          // cv2 returns (None, None): nothing to scatter, the caller decides
        for (int i = tid; i < n_full; i += RS_THREADS) mask_out[i] = 0;
        if (tid == 0) { info[0] = 0; info[1] = 0; info[2] = 0; }
        return;
    }
    ransac_affine2d_body(S, cf, ct, n, thresh, conf, max_iters, refine, cinl, T, info);
    __syncthreads();
    for (int i = tid; i < n_full; i += RS_THREADS) mask_out[i] = 0;
    __syncthreads();
    for (int i = tid; i < n; i += RS_THREADS) mask_out[cidx[i]] = cinl[i];
}

}  // namespace

VEL_API int vel_estimate_affine2d_ransac(const float* from_xy, const float* to_xy, int32_t npts, double threshold, double confidence,
                                         int32_t max_iters, int32_t refine, uint8_t* inliers, double* T, int32_t* info,
                                         vel_stream_t stream)
{
    VEL_CHECK_ARG(from_xy && to_xy && inliers && T && info, "vel_estimate_affine2d_ransac: NULL argument");
    VEL_CHECK_ARG(npts >= 3 && npts <= RS_THREADS * RS_MAX_PER_THREAD, "vel_estimate_affine2d_ransac: npts %d outside [3,%d]", npts,
                  RS_THREADS * RS_MAX_PER_THREAD);
    VEL_CHECK_ARG(threshold > 0. && max_iters >= 1, "vel_estimate_affine2d_ransac: bad threshold / iteration budget");
    ransac_affine2d_kernel<<<1, RS_THREADS, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float2*>(from_xy),
                                                                      reinterpret_cast<const float2*>(to_xy), npts, threshold, confidence,
                                                                      max_iters, refine, inliers, T, info);
    VEL_LAUNCH_CHECK("ransac_affine2d_kernel");
    return VEL_OK;
}

VEL_API size_t vel_estimate_affine2d_ransac_masked_workspace(int32_t npts)
{
    if (npts <= 0) return 0;
    return (size_t)npts * (2 * sizeof(float2) + sizeof(int) + 1) + 64;
}

VEL_API int vel_estimate_affine2d_ransac_masked(const float* from_xy, const float* to_xy, const uint8_t* mask, int32_t npts, float to_scale,
                                                float to_off_x, float to_off_y, double threshold, double confidence, int32_t max_iters,
                                                int32_t refine, void* work, size_t work_bytes, uint8_t* mask_out, float* to_mapped, double* T,
                                                int32_t* info, vel_stream_t stream)
{
    VEL_CHECK_ARG(from_xy && to_xy && mask && mask_out && T && info && work, "vel_estimate_affine2d_ransac_masked: NULL argument");
    VEL_CHECK_ARG(npts >= 1 && npts <= RS_THREADS * RS_MAX_PER_THREAD, "vel_estimate_affine2d_ransac_masked: npts %d outside [1,%d]", npts,
                  RS_THREADS * RS_MAX_PER_THREAD);
    VEL_CHECK_ARG(threshold > 0. && max_iters >= 1, "vel_estimate_affine2d_ransac_masked: bad threshold / iteration budget");
    VEL_CHECK_ARG(work_bytes >= vel_estimate_affine2d_ransac_masked_workspace(npts) && ((size_t)work & 7) == 0,
                  "vel_estimate_affine2d_ransac_masked: workspace too small or misaligned");
    float2* cf = (float2*)work;
    float2* ct = cf + npts;
    int* cidx = (int*)(ct + npts);
    uint8_t* cinl = (uint8_t*)(cidx + npts);
    ransac_affine2d_masked_kernel<<<1, RS_THREADS, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float2*>(from_xy), reinterpret_cast<const float2*>(to_xy), mask, npts, to_scale, to_off_x, to_off_y, threshold,
        confidence, max_iters, refine, cf, ct, cidx, cinl, mask_out, reinterpret_cast<float2*>(to_mapped), T, info);
    VEL_LAUNCH_CHECK("ransac_affine2d_masked_kernel");
    return VEL_OK;
}
