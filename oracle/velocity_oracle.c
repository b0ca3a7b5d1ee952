/*
 * velocity_oracle.c -- CPU ORACLE (test infrastructure, NOT the product).
 *
 * Scalar C restatement of the third-party arithmetic the reference's KLT path delegates to
 * opencv-python (unpinned in /root/reference/requirements.txt:5; the version installed in this
 * image and used to pin this oracle is opencv-python-headless 4.13.0.92).  The reference call
 * sites this file stands in for:
 *     cv2.calcOpticalFlowPyrLK   utils/KLT.py:45,48     -> orc_calc_optical_flow_pyr_lk
 *     cv2.pyrDown (inside LK)    (implicit)             -> orc_pyrdown_u8
 *     cv2.Scharr  (inside LK)    (implicit)             -> orc_scharr_s16
 *     cv2.remap(INTER_LINEAR)    utils/KLT.py:73        -> orc_remap_affine_u8 (+ the float32
 *                                                          meshgrid arithmetic of utils/KLT.py:70-72)
 *     cv2.resize(.., 1/4, NEAREST) utils/KLT.py:111,113 -> orc_decimate4_u8
 *     cv2.BFMatcher().knnMatch(k=2) utils/KLT.py:16,25  -> orc_knn2_hamming / orc_knn2_l2
 *     cv2.cvtColor(BGR2GRAY)     vidExample.py:91       -> orc_bgr2gray_u8
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product (velocity_b200/) never does.
 *
 * Parity pin: tests/golden/*.npz were produced by running the unmodified reference functions
 * (which call cv2 4.13.0) in the authoring container via tests/golden/make_golden.py;
 * tests/test_oracle_*.py check this file against them (status masks bit-exact, points <= 2e-3 px:
 * cv2's SIMD build accumulates the 2x2 system in float32 lanes, this file accumulates exactly in
 * int64 and converts once).
 *
 * Build: see oracle/build_oracle.py  (gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC).
 * -ffp-contract=off matters: the CUDA kernels use explicit *_rn intrinsics so that oracle and
 * device float32 results are bit-identical.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_API __attribute__((visibility("default")))

/* BORDER_REFLECT_101 for a single overshoot (callers guarantee |overshoot| < n). */
static inline int reflect101(int i, int n)
{
    if (n == 1) return 0;
    while (i < 0 || i >= n) {
        if (i < 0) i = -i;
        else i = 2 * n - 2 - i;
    }
    return i;
}

/* round-half-to-even of a float (cvRound on x86 == cvtss2si in the default rounding mode). */
static inline int round_rne(float v) { return (int)lrintf(v); }

/* ------------------------------------------------------------------------------------------ */
/* cv2.pyrDown for CV_8UC1: separable [1 4 6 4 1], integer sums, (s + 128) >> 8, REFLECT_101,   */
/* destination size ((w+1)/2, (h+1)/2).                                                         */
ORC_API void orc_pyrdown_u8(const uint8_t* src, int sw, int sh, int spitch, uint8_t* dst, int dpitch)
{
    const int dw = (sw + 1) / 2, dh = (sh + 1) / 2;
    int* hrow = (int*)malloc(sizeof(int) * (size_t)dw * 5);
    for (int y = 0; y < dh; ++y) {
        for (int k = 0; k < 5; ++k) {
            const int sy = reflect101(2 * y - 2 + k, sh);
            const uint8_t* s = src + (size_t)sy * spitch;
            int* h = hrow + (size_t)k * dw;
            for (int x = 0; x < dw; ++x) {
                const int c = 2 * x;
                h[x] = s[reflect101(c - 2, sw)] + s[reflect101(c + 2, sw)] +
                       4 * (s[reflect101(c - 1, sw)] + s[reflect101(c + 1, sw)]) + 6 * s[reflect101(c, sw)];
            }
        }
        uint8_t* d = dst + (size_t)y * dpitch;
        for (int x = 0; x < dw; ++x) {
            const int v = hrow[x] + hrow[4 * dw + x] + 4 * (hrow[dw + x] + hrow[3 * dw + x]) + 6 * hrow[2 * dw + x];
            d[x] = (uint8_t)((v + 128) >> 8);
        }
    }
    free(hrow);
}

/* cv2.resize(fx=fy=1/4, INTER_NEAREST) == src[::4, ::4] with size (round(w/4), round(h/4)).   */
ORC_API void orc_decimate4_u8(const uint8_t* src, int sw, int sh, int spitch, uint8_t* dst, int dw, int dh, int dpitch)
{
    for (int y = 0; y < dh; ++y) {
        int sy = y * 4; if (sy > sh - 1) sy = sh - 1;
        for (int x = 0; x < dw; ++x) {
            int sx = x * 4; if (sx > sw - 1) sx = sw - 1;
            dst[(size_t)y * dpitch + x] = src[(size_t)sy * spitch + sx];
        }
    }
}

/* Unnormalised Scharr pair as the LK code computes it (calcSharrDeriv): REFLECT_101 at the edge. */
static inline void scharr_at(const uint8_t* img, int w, int h, int pitch, int x, int y, int* gx, int* gy)
{
    const int xm = reflect101(x - 1, w), xp = reflect101(x + 1, w);
    const int ym = reflect101(y - 1, h), yp = reflect101(y + 1, h);
    const uint8_t* r0 = img + (size_t)ym * pitch;
    const uint8_t* r1 = img + (size_t)y * pitch;
    const uint8_t* r2 = img + (size_t)yp * pitch;
    /* vertical smoothing / difference, then horizontal difference / smoothing */
    const int s_m = 3 * (r0[xm] + r2[xm]) + 10 * r1[xm];
    const int s_p = 3 * (r0[xp] + r2[xp]) + 10 * r1[xp];
    const int d_m = r2[xm] - r0[xm];
    const int d_c = r2[x] - r0[x];
    const int d_p = r2[xp] - r0[xp];
    *gx = s_p - s_m;
    *gy = 3 * (d_m + d_p) + 10 * d_c;
}

ORC_API void orc_scharr_s16(const uint8_t* img, int w, int h, int pitch, int16_t* dxy /* [h][w][2] */)
{
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            int gx, gy;
            scharr_at(img, w, h, pitch, x, y, &gx, &gy);
            dxy[((size_t)y * w + x) * 2 + 0] = (int16_t)gx;
            dxy[((size_t)y * w + x) * 2 + 1] = (int16_t)gy;
        }
}

/* ------------------------------------------------------------------------------------------ */
/* Pyramidal Lucas-Kanade                                                                       */

typedef struct {
    const uint8_t* p;
    int w, h, pitch;
} orc_img;

/* image sample with the winSize-wide REFLECT_101 padding the LK pyramid carries */
static inline int img_at(const orc_img* im, int x, int y)
{
    return im->p[(size_t)reflect101(y, im->h) * im->pitch + reflect101(x, im->w)];
}

/* derivative sample: Scharr inside the image, constant 0 in the padding */
static inline void deriv_at(const orc_img* im, int x, int y, int* gx, int* gy)
{
    if (x < 0 || y < 0 || x >= im->w || y >= im->h) { *gx = 0; *gy = 0; return; }
    scharr_at(im->p, im->w, im->h, im->pitch, x, y, gx, gy);
}

#define W_BITS 14
static inline void bilin_weights(float a, float b, int* w00, int* w01, int* w10, int* w11)
{
    const float one_a = 1.f - a, one_b = 1.f - b;
    *w00 = round_rne((one_a * one_b) * (float)(1 << W_BITS));
    *w01 = round_rne((a * one_b) * (float)(1 << W_BITS));
    *w10 = round_rne((one_a * b) * (float)(1 << W_BITS));
    *w11 = (1 << W_BITS) - *w00 - *w01 - *w10;
}

static inline int descale(int v, int n) { return (v + (1 << (n - 1))) >> n; }

/* number of pyramid levels calcOpticalFlowPyrLK really uses (buildOpticalFlowPyramid stops at the
 * first level whose width <= win_w or height <= win_h) */
ORC_API int orc_lk_effective_max_level(int w, int h, int win_w, int win_h, int max_level)
{
    int lvl = 0;
    while (lvl < max_level) {
        const int nw = (w + 1) / 2, nh = (h + 1) / 2;
        if (nw <= win_w || nh <= win_h) break;
        w = nw; h = nh; ++lvl;
    }
    return lvl;
}

/* One point, one level.  `next` holds the running estimate (in/out).  Returns nothing; updates
 * status/err only at level 0, exactly as OpenCV's LKTrackerInvoker does. */
static void lk_point_level(const orc_img* I, const orc_img* J, int level, int max_level_eff, float px, float py,
                           float* next_x, float* next_y, uint8_t* status, float* err, int win_w, int win_h,
                           int max_count, float eps2, float min_eig_thr, int16_t* Ibuf, int16_t* dbuf)
{
    const float half_x = (float)(win_w - 1) * 0.5f, half_y = (float)(win_h - 1) * 0.5f;
    const float scale = (float)(1. / (double)(1 << level));
    float prev_x = px * scale, prev_y = py * scale;
    float nx, ny;
    if (level == max_level_eff) { nx = prev_x; ny = prev_y; }
    else { nx = *next_x * 2.f; ny = *next_y * 2.f; }
    *next_x = nx; *next_y = ny;

    prev_x -= half_x; prev_y -= half_y;
    int ipx = (int)floorf(prev_x), ipy = (int)floorf(prev_y);
    if (ipx < -win_w || ipx >= I->w || ipy < -win_h || ipy >= I->h) {
        if (level == 0) { *status = 0; *err = 0.f; }
        return;
    }
    int w00, w01, w10, w11;
    bilin_weights(prev_x - (float)ipx, prev_y - (float)ipy, &w00, &w01, &w10, &w11);

    int64_t sA11 = 0, sA12 = 0, sA22 = 0;
    for (int y = 0; y < win_h; ++y)
        for (int x = 0; x < win_w; ++x) {
            const int X = ipx + x, Y = ipy + y;
            const int iv = descale(img_at(I, X, Y) * w00 + img_at(I, X + 1, Y) * w01 + img_at(I, X, Y + 1) * w10 +
                                       img_at(I, X + 1, Y + 1) * w11, W_BITS - 5);
            int gx00, gy00, gx01, gy01, gx10, gy10, gx11, gy11;
            deriv_at(I, X, Y, &gx00, &gy00);
            deriv_at(I, X + 1, Y, &gx01, &gy01);
            deriv_at(I, X, Y + 1, &gx10, &gy10);
            deriv_at(I, X + 1, Y + 1, &gx11, &gy11);
            const int ix = descale(gx00 * w00 + gx01 * w01 + gx10 * w10 + gx11 * w11, W_BITS);
            const int iy = descale(gy00 * w00 + gy01 * w01 + gy10 * w10 + gy11 * w11, W_BITS);
            Ibuf[y * win_w + x] = (int16_t)iv;
            dbuf[(y * win_w + x) * 2 + 0] = (int16_t)ix;
            dbuf[(y * win_w + x) * 2 + 1] = (int16_t)iy;
            sA11 += (int64_t)ix * ix;
            sA12 += (int64_t)ix * iy;
            sA22 += (int64_t)iy * iy;
        }
    const float FLT_SCALE = 1.f / (float)(1 << 20);
    const float A11 = (float)sA11 * FLT_SCALE, A12 = (float)sA12 * FLT_SCALE, A22 = (float)sA22 * FLT_SCALE;
    float D = A11 * A22 - A12 * A12;
    const float dA = A11 - A22;
    const float min_eig = (A22 + A11 - sqrtf(dA * dA + 4.f * A12 * A12)) / (float)(2 * win_w * win_h);
    if (min_eig < min_eig_thr || D < FLT_EPSILON) {
        if (level == 0) *status = 0;
        return;
    }
    D = 1.f / D;

    nx -= half_x; ny -= half_y;
    float pdx = 0.f, pdy = 0.f;
    for (int j = 0; j < max_count; ++j) {
        const int inx = (int)floorf(nx), iny = (int)floorf(ny);
        if (inx < -win_w || inx >= J->w || iny < -win_h || iny >= J->h) {
            if (level == 0) *status = 0;
            break;
        }
        bilin_weights(nx - (float)inx, ny - (float)iny, &w00, &w01, &w10, &w11);
        int64_t sb1 = 0, sb2 = 0;
        for (int y = 0; y < win_h; ++y)
            for (int x = 0; x < win_w; ++x) {
                const int X = inx + x, Y = iny + y;
                const int diff = descale(img_at(J, X, Y) * w00 + img_at(J, X + 1, Y) * w01 + img_at(J, X, Y + 1) * w10 +
                                             img_at(J, X + 1, Y + 1) * w11, W_BITS - 5) - Ibuf[y * win_w + x];
                sb1 += (int64_t)diff * dbuf[(y * win_w + x) * 2 + 0];
                sb2 += (int64_t)diff * dbuf[(y * win_w + x) * 2 + 1];
            }
        const float b1 = (float)sb1 * FLT_SCALE, b2 = (float)sb2 * FLT_SCALE;
        const float dx = (A12 * b2 - A22 * b1) * D;
        const float dy = (A12 * b1 - A11 * b2) * D;
        nx += dx; ny += dy;
        *next_x = nx + half_x; *next_y = ny + half_y;
        if (dx * dx + dy * dy <= eps2) break;
        if (j > 0 && fabsf(dx + pdx) < 0.01f && fabsf(dy + pdy) < 0.01f) {
            *next_x -= dx * 0.5f; *next_y -= dy * 0.5f;
            break;
        }
        pdx = dx; pdy = dy;
    }

    if (level == 0 && *status) {
        const float fx = *next_x - half_x, fy = *next_y - half_y;
        const int inx = (int)floorf(fx), iny = (int)floorf(fy);
        if (inx < -win_w || inx >= J->w || iny < -win_h || iny >= J->h) { *status = 0; return; }
        bilin_weights(fx - (float)inx, fy - (float)iny, &w00, &w01, &w10, &w11);
        int64_t se = 0;
        for (int y = 0; y < win_h; ++y)
            for (int x = 0; x < win_w; ++x) {
                const int X = inx + x, Y = iny + y;
                const int diff = descale(img_at(J, X, Y) * w00 + img_at(J, X + 1, Y) * w01 + img_at(J, X, Y + 1) * w10 +
                                             img_at(J, X + 1, Y + 1) * w11, W_BITS - 5) - Ibuf[y * win_w + x];
                se += diff < 0 ? -diff : diff;
            }
        *err = (float)se / (float)(32 * win_w * win_h);
    }
}

/* Whole call: build both pyramids, track n points.  pts/out are [n][2] (x, y) float32.
 * eps is the user-facing epsilon (squared internally like OpenCV does); max_count/eps clamped as
 * OpenCV clamps them.  nthreads <= 0: all cores. Returns the effective max level. */
ORC_API int orc_calc_optical_flow_pyr_lk(const uint8_t* imgA, const uint8_t* imgB, int w, int h, int pitchA, int pitchB,
                                         const float* pts, int n, int win_w, int win_h, int max_level, int max_count,
                                         double eps, float min_eig_thr, float* out, uint8_t* status, float* err,
                                         int nthreads)
{
    if (max_count < 0) max_count = 0;
    if (max_count > 100) max_count = 100;
    if (eps < 0.) eps = 0.;
    if (eps > 10.) eps = 10.;
    const float eps2 = (float)(eps * eps);
    const int L = orc_lk_effective_max_level(w, h, win_w, win_h, max_level);

    orc_img A[16], B[16];
    uint8_t* ownedA[16] = {0};
    uint8_t* ownedB[16] = {0};
    A[0] = (orc_img){imgA, w, h, pitchA};
    B[0] = (orc_img){imgB, w, h, pitchB};
    for (int l = 1; l <= L; ++l) {
        const int lw = (A[l - 1].w + 1) / 2, lh = (A[l - 1].h + 1) / 2;
        ownedA[l] = (uint8_t*)malloc((size_t)lw * lh);
        ownedB[l] = (uint8_t*)malloc((size_t)lw * lh);
        orc_pyrdown_u8(A[l - 1].p, A[l - 1].w, A[l - 1].h, A[l - 1].pitch, ownedA[l], lw);
        orc_pyrdown_u8(B[l - 1].p, B[l - 1].w, B[l - 1].h, B[l - 1].pitch, ownedB[l], lw);
        A[l] = (orc_img){ownedA[l], lw, lh, lw};
        B[l] = (orc_img){ownedB[l], lw, lh, lw};
    }
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
    nthreads = 1;
#endif
#pragma omp parallel num_threads(nthreads)
    {
        int16_t* Ibuf = (int16_t*)malloc(sizeof(int16_t) * (size_t)win_w * win_h);
        int16_t* dbuf = (int16_t*)malloc(sizeof(int16_t) * (size_t)win_w * win_h * 2);
#pragma omp for schedule(dynamic, 16)
        for (int i = 0; i < n; ++i) {
            uint8_t st = 1;
            float e = 0.f;
            float nx = 0.f, ny = 0.f;
            for (int l = L; l >= 0; --l)
                lk_point_level(&A[l], &B[l], l, L, pts[2 * i], pts[2 * i + 1], &nx, &ny, &st, &e, win_w, win_h, max_count,
                               eps2, min_eig_thr, Ibuf, dbuf);
            out[2 * i] = nx; out[2 * i + 1] = ny;
            status[i] = st; err[i] = e;
        }
        free(Ibuf); free(dbuf);
    }
    for (int l = 1; l <= L; ++l) { free(ownedA[l]); free(ownedB[l]); }
    return L;
}

/* ------------------------------------------------------------------------------------------ */
/* utils/KLT.py:70-73 -- float32 affine map of the ROI grid, then cv2.remap(INTER_LINEAR,         */
/* BORDER_CONSTANT 0) with OpenCV's fixed point: coordinates to 1/32 px, weights 15 bit.          */
ORC_API void orc_remap_affine_u8(const uint8_t* src, int sw, int sh, int spitch, const float* T /* 3x2 row-major */,
                                 int x0, int y0, int dw, int dh, uint8_t* dst, int dpitch)
{
    const float t00 = T[0], t01 = T[1], t10 = T[2], t11 = T[3], t20 = T[4], t21 = T[5];
    for (int r = 0; r < dh; ++r) {
        const float y = (float)(y0 + r);
        for (int c = 0; c < dw; ++c) {
            const float x = (float)(x0 + c);
            const float mx = (x * t00 + y * t10) + t20;
            const float my = (x * t01 + y * t11) + t21;
            const int sx = round_rne(mx * 32.f), sy = round_rne(my * 32.f);
            /* OpenCV stores the integer part as saturated int16 */
            int ix = sx >> 5, iy = sy >> 5;
            if (ix < -32768) ix = -32768; if (ix > 32767) ix = 32767;
            if (iy < -32768) iy = -32768; if (iy > 32767) iy = 32767;
            const int fx = sx & 31, fy = sy & 31;
            const int w00 = (32 - fx) * (32 - fy) * 32, w01 = fx * (32 - fy) * 32;
            const int w10 = (32 - fx) * fy * 32, w11 = fx * fy * 32;
            int acc = 0;
            const int in_x0 = (ix >= 0 && ix < sw), in_x1 = (ix + 1 >= 0 && ix + 1 < sw);
            const int in_y0 = (iy >= 0 && iy < sh), in_y1 = (iy + 1 >= 0 && iy + 1 < sh);
            if (in_y0 && in_x0) acc += src[(size_t)iy * spitch + ix] * w00;
            if (in_y0 && in_x1) acc += src[(size_t)iy * spitch + ix + 1] * w01;
            if (in_y1 && in_x0) acc += src[(size_t)(iy + 1) * spitch + ix] * w10;
            if (in_y1 && in_x1) acc += src[(size_t)(iy + 1) * spitch + ix + 1] * w11;
            dst[(size_t)r * dpitch + c] = (uint8_t)((acc + (1 << 14)) >> 15);
        }
    }
}

/* ------------------------------------------------------------------------------------------ */
/* cv2.BFMatcher(normType).knnMatch(q, t, k=2): two nearest train rows per query row, ascending   */
/* distance, ties resolved to the lower train index.  idx/dist are [nq][2]; missing -> -1.        */
ORC_API void orc_knn2_hamming(const uint8_t* q, int nq, const uint8_t* t, int nt, int nbytes, int32_t* idx, int32_t* dist)
{
#pragma omp parallel for schedule(static)
    for (int i = 0; i < nq; ++i) {
        int b0 = -1, b1 = -1, d0 = INT32_MAX, d1 = INT32_MAX;
        for (int j = 0; j < nt; ++j) {
            int d = 0;
            for (int k = 0; k < nbytes; ++k) d += __builtin_popcount((unsigned)(q[(size_t)i * nbytes + k] ^ t[(size_t)j * nbytes + k]));
            if (d < d0) { d1 = d0; b1 = b0; d0 = d; b0 = j; }
            else if (d < d1) { d1 = d; b1 = j; }
        }
        idx[2 * i] = b0; idx[2 * i + 1] = b1;
        dist[2 * i] = b0 < 0 ? -1 : d0; dist[2 * i + 1] = b1 < 0 ? -1 : d1;
    }
}

ORC_API void orc_knn2_l2(const float* q, int nq, const float* t, int nt, int dim, int32_t* idx, float* dist)
{
#pragma omp parallel for schedule(static)
    for (int i = 0; i < nq; ++i) {
        int b0 = -1, b1 = -1;
        float d0 = INFINITY, d1 = INFINITY;
        for (int j = 0; j < nt; ++j) {
            float s = 0.f;
            for (int k = 0; k < dim; ++k) {
                const float e = q[(size_t)i * dim + k] - t[(size_t)j * dim + k];
                s += e * e;
            }
            if (s < d0) { d1 = d0; b1 = b0; d0 = s; b0 = j; }
            else if (s < d1) { d1 = s; b1 = j; }
        }
        idx[2 * i] = b0; idx[2 * i + 1] = b1;
        dist[2 * i] = b0 < 0 ? -1.f : sqrtf(d0); dist[2 * i + 1] = b1 < 0 ? -1.f : sqrtf(d1);
    }
}

/* cv2.cvtColor(BGR2GRAY) for CV_8UC3 (vidExample.py:91): OpenCV 4.13 uses 15-bit fixed point. */
ORC_API void orc_bgr2gray_u8(const uint8_t* bgr, int w, int h, int bpitch, uint8_t* gray, int gpitch)
{
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            const uint8_t* p = bgr + (size_t)y * bpitch + 3 * x;
            gray[(size_t)y * gpitch + x] = (uint8_t)((3735 * p[0] + 19235 * p[1] + 9798 * p[2] + 16384) >> 15);
        }
}

/* ---- cv2.cornerSubPix(im, p, (5,5), (-1,-1), criteria) (vidExample.py:113-115), CV_8UC1 input ---------------------------
 * Arithmetic of OpenCV 4.13 (imgproc cornersubpix.cpp + getRectSubPix 8U->32F), pinned by black-box comparison: 36k
 * points on six image sizes and three termination criteria reproduce cv2 bit for bit (tests/golden/subpix.npz holds a
 * committed subset).  Patch elements:
 *   both rows and both columns inside the frame : (p00*a11 + p01*a12) + (p10*a21 + p11*a22), float32, unfused
 *   row beyond the frame (replicated), columns inside : fma(p01, a, p00*(1-a))
 *   column beyond the frame : p0*(1-b) + p1*b with the edge column -- except that rows ABOVE the frame take column
 *                              w-2 instead of w-1 on the right side (a quirk of cv2 4.13's sampler, reproduced)
 * then per iteration the 2x2 system of the gradient products, accumulated in float64 in raster order over the
 * (2*win+1)^2 window with the float32 mask exp(-y^2)*exp(-x^2); an update that leaves the frame is discarded. */
static void rect_subpix_8u32f(const uint8_t* src, int step, int cols, int rows, float* dst, int dw, int dh, float cx, float cy)
{
    const float centx = cx - (dw - 1) * 0.5f, centy = cy - (dh - 1) * 0.5f;
    const int ipx = (int)floorf(centx), ipy = (int)floorf(centy);
    const float a = centx - ipx, b = centy - ipy;
    const float a11 = (1.f - a) * (1.f - b), a12 = a * (1.f - b), a21 = (1.f - a) * b, a22 = a * b, b1 = 1.f - b, b2 = b;
    if (0 <= ipx && ipx < cols - dw && 0 <= ipy && ipy < rows - dh) {
        const uint8_t* p = src + (size_t)ipy * step + ipx;
        for (int i = 0; i < dh; ++i, p += step, dst += dw)
            for (int j = 0; j < dw; ++j)
                dst[j] = (p[j] * a11 + p[j + 1] * a12) + (p[j + step] * a21 + p[j + step + 1] * a22);
        return;
    }
    int rx = ipx >= 0 ? 0 : -ipx; if (rx > dw) rx = dw;
    int rw = ipx < cols - dw ? dw : cols - ipx - 1; if (rw < 0) rw = 0;
    const int ry = ipy >= 0 ? 0 : -ipy;
    int rh = ipy < rows - dh ? dh : rows - ipy - 1; if (rh < 0) rh = 0;
    for (int i = 0; i < dh; ++i, dst += dw) {
        int y0 = ipy + i; y0 = y0 < 0 ? 0 : (y0 >= rows ? rows - 1 : y0);
        const int rep = (i < ry || i >= rh);
        const uint8_t* r0 = src + (size_t)y0 * step;
        const uint8_t* r1 = rep ? r0 : r0 + step;
        for (int j = 0; j < dw; ++j) {
            if (j < rx) dst[j] = r0[0] * b1 + r1[0] * b2;
            else if (j >= rw) {
                int xe = ipx + rw - (i < ry ? 1 : 0);
                xe = xe < 0 ? 0 : (xe >= cols ? cols - 1 : xe);
                dst[j] = r0[xe] * b1 + r1[xe] * b2;
            } else {
                const int x = ipx + j;
                if (rep) dst[j] = fmaf((float)r0[x + 1], a, r0[x] * (1.f - a));
                else dst[j] = (r0[x] * a11 + r0[x + 1] * a12) + (r1[x] * a21 + r1[x + 1] * a22);
            }
        }
    }
}

ORC_API void orc_corner_subpix_u8(const uint8_t* src, int cols, int rows, int step, float* pts, int n, int winw, int winh,
                                  int max_iters, double eps)
{
    const int win_w = winw * 2 + 1, win_h = winh * 2 + 1;
    float* mask = (float*)malloc(sizeof(float) * win_w * win_h);
    max_iters = max_iters < 1 ? 1 : (max_iters > 100 ? 100 : max_iters);
    eps = eps < 0. ? 0. : eps;
    eps *= eps;
    for (int i = 0; i < win_h; i++) {
        const float y = (float)(i - winh) / winh;
        const float vy = expf(-y * y);
        for (int j = 0; j < win_w; j++) {
            const float x = (float)(j - winw) / winw;
            mask[i * win_w + j] = (float)(vy * expf(-x * x));
        }
    }
#pragma omp parallel
    {
        float* buf = (float*)malloc(sizeof(float) * (win_w + 2) * (win_h + 2));
#pragma omp for schedule(static)
        for (int k = 0; k < n; ++k) {
            const float cTx = pts[2 * k], cTy = pts[2 * k + 1];
            float cIx = cTx, cIy = cTy;
            int iter = 0;
            double err = 0;
            do {
                rect_subpix_8u32f(src, step, cols, rows, buf, win_w + 2, win_h + 2, cIx, cIy);
                const float* subpix = buf + (win_w + 2) + 1;
                double a = 0, b = 0, c = 0, bb1 = 0, bb2 = 0;
                for (int i = 0, kk = 0; i < win_h; i++, subpix += win_w + 2) {
                    const double py = i - winh;
                    for (int j = 0; j < win_w; j++, kk++) {
                        const double m = mask[kk];
                        const double tgx = subpix[j + 1] - subpix[j - 1];
                        const double tgy = subpix[j + win_w + 2] - subpix[j - win_w - 2];
                        const double gxx = tgx * tgx * m, gxy = tgx * tgy * m, gyy = tgy * tgy * m;
                        const double px = j - winw;
                        a += gxx; b += gxy; c += gyy;
                        bb1 += gxx * px + gxy * py;
                        bb2 += gxy * px + gyy * py;
                    }
                }
                const double det = a * c - b * b;
                if (fabs(det) <= DBL_EPSILON * DBL_EPSILON) break;
                const double scale = 1.0 / det;
                const float nx = (float)(cIx + c * scale * bb1 - b * scale * bb2);
                const float ny = (float)(cIy - b * scale * bb1 + a * scale * bb2);
                err = (nx - cIx) * (nx - cIx) + (ny - cIy) * (ny - cIy);
                if (nx < 0 || nx >= cols || ny < 0 || ny >= rows) break;   /* an update that leaves the frame is discarded */
                cIx = nx; cIy = ny;
            } while (++iter < max_iters && err > eps);
            if (fabsf(cIx - cTx) > winw || fabsf(cIy - cTy) > winh) { cIx = cTx; cIy = cTy; }
            pts[2 * k] = cIx; pts[2 * k + 1] = cIy;
        }
        free(buf);
    }
    free(mask);
}

ORC_API int orc_version(void) { return 1; }
