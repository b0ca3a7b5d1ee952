// K9: feature initialisation -- cv2.goodFeaturesToTrack(roi, maxCorners, quality, 0, blockSize=5,
// useHarrisDetector=True) as the reference calls it at vidExample.py:110.
//
// The arithmetic is OpenCV's (un-vendored, unpinned: requirements.txt:5; pinned here by black-box comparison
// against opencv-python 4.13.0 and restated on the CPU in oracle/gftt_oracle.py):
//   1. Sobel pair, CV_32F, scale s = 1/(4*blockSize*255), REFLECT_101, float32 kernel (k1, k0, k1) = (1,2,1)*s:
//        Dx = fma(r(y-1) + r(y+1), k1, r(y)*k0),  r = p(x+1) - p(x-1)
//        Dy = t(y+1) - t(y-1),  t = fma(p(x+1), k1, fma(p(x), k0, p(x-1)*k1)) for columns < 32*(w/32),
//             ((p(x-1)*k1 + p(x)*k0) + p(x+1)*k1) in the remaining columns (cv2's scalar row-filter tail)
//   2. cov = (Dx*Dx, Dx*Dy, Dy*Dy), float32
//   3. 5x5 box sums (REFLECT_101) in float64: row sums left to right, then cv2's RUNNING column sum from the top of
//      the image (SUM += entering row; out = float(SUM); SUM -= leaving row).  Double sums are exact -- hence
//      order-free -- except where a cancellation residue of the fused Sobel makes a gradient ~1e-11; replicating
//      the running order keeps even those pixels bit-identical, at the price of a sequential walk down each column.
//   4. R = (a*c - b*b) - k*((a+c)*(a+c)), float32, unfused
//   5. keep R > float(max(R)*quality); 3x3 local maxima inside the 1-px frame; order by value descending, ties by
//      higher raster address; first maxCorners.  (minDistance = 0 as in the reference: no spacing filter.)
//
// HBM streaming: the frame is read once (cov kernel); cov, row-sum and R planes are written and read once.  Every float step uses
// explicit round-to-nearest intrinsics so nothing is contracted differently from the oracle.  Selection is
// deterministic: candidates are appended with an atomic counter (arbitrary order), then the K-th largest 64-bit key
// (value bits << 32 | address, all distinct) is found by an 8-pass radix select and the survivors are ranked by
// counting -- no host round trip, no library sort.
#include "common.cuh"
#include <float.h>
#include <math.h>

namespace {

struct GfttState {           // device-resident control block at the start of the workspace
    unsigned max_bits;       // orderable bits of max(R)
    unsigned n_cand;         // candidates appended
    unsigned n_sel;          // survivors appended
    unsigned n_out;          // min(n_cand, max_corners)
    unsigned long long prefix;   // radix-select prefix; after 8 passes the K-th largest key
    unsigned remaining;      // rank still to descend inside the current prefix
    unsigned hist[256];
};

__device__ __forceinline__ unsigned orderable(float v)
{
    const unsigned u = __float_as_uint(v);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float from_orderable(unsigned u)
{
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

__global__ void gftt_reset_kernel(GfttState* S)
{
    if (threadIdx.x == 0) { S->max_bits = 0u; S->n_cand = 0u; S->n_sel = 0u; S->n_out = 0u; S->prefix = 0ull; S->remaining = 0u; }
    S->hist[threadIdx.x] = 0u;
}

// steps 1-2: one thread per pixel
__global__ void __launch_bounds__(256)
harris_cov_kernel(const uint8_t* __restrict__ img, int w, int h, int pitch, float k1, float k0, int fused_cols,
                  float* __restrict__ cxx, float* __restrict__ cxy, float* __restrict__ cyy)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= w) return;
    const int xm = reflect101(x - 1, w), xp = reflect101(x + 1, w);
    const uint8_t* r0 = img + (size_t)reflect101(y - 1, h) * pitch;
    const uint8_t* r1 = img + (size_t)y * pitch;
    const uint8_t* r2 = img + (size_t)reflect101(y + 1, h) * pitch;
    const float a0 = (float)__ldg(r0 + xm), a1 = (float)__ldg(r0 + x), a2 = (float)__ldg(r0 + xp);
    const float b0 = (float)__ldg(r1 + xm), b2 = (float)__ldg(r1 + xp);
    const float c0 = (float)__ldg(r2 + xm), c1 = (float)__ldg(r2 + x), c2 = (float)__ldg(r2 + xp);
    const float ra = __fsub_rn(a2, a0), rb = __fsub_rn(b2, b0), rc = __fsub_rn(c2, c0);
    const float dx = __fmaf_rn(__fadd_rn(ra, rc), k1, __fmul_rn(rb, k0));
    float ta, tc;
    if (x < fused_cols) {
        ta = __fmaf_rn(a2, k1, __fmaf_rn(a1, k0, __fmul_rn(a0, k1)));
        tc = __fmaf_rn(c2, k1, __fmaf_rn(c1, k0, __fmul_rn(c0, k1)));
    } else {
        ta = __fadd_rn(__fadd_rn(__fmul_rn(a0, k1), __fmul_rn(a1, k0)), __fmul_rn(a2, k1));
        tc = __fadd_rn(__fadd_rn(__fmul_rn(c0, k1), __fmul_rn(c1, k0)), __fmul_rn(c2, k1));
    }
    const float dy = __fsub_rn(tc, ta);
    const size_t o = (size_t)y * w + x;
    cxx[o] = __fmul_rn(dx, dx);
    cxy[o] = __fmul_rn(dx, dy);
    cyy[o] = __fmul_rn(dy, dy);
}

// step 3a: 5-tap row sums in float64, left to right (cv2's RowSum), one thread per pixel
__global__ void __launch_bounds__(256)
harris_rowsum_kernel(const float* __restrict__ cxx, const float* __restrict__ cxy, const float* __restrict__ cyy, int w, int h,
                     double* __restrict__ sxx, double* __restrict__ sxy, double* __restrict__ syy)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= w) return;
    int xs[5];
#pragma unroll
    for (int d = 0; d < 5; ++d) xs[d] = reflect101(x + d - 2, w);
    const size_t ro = (size_t)y * w;
    const float* planes[3] = {cxx + ro, cxy + ro, cyy + ro};
    double* outs[3] = {sxx, sxy, syy};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        double s = __dadd_rn((double)__ldg(planes[c] + xs[0]), (double)__ldg(planes[c] + xs[1]));
        s = __dadd_rn(s, (double)__ldg(planes[c] + xs[2]));
        s = __dadd_rn(s, (double)__ldg(planes[c] + xs[3]));
        outs[c][ro + x] = __dadd_rn(s, (double)__ldg(planes[c] + xs[4]));
    }
}

// steps 3b-4: cv2's running float64 column sums.  One thread per column walks down the rows (the recurrence is what
// makes the result independent of nothing but cv2's own order); the row sums of the NEXT five rows are requested
// before the current five are consumed, so each batch of five steps exposes one memory latency.
constexpr int SCAN_THREADS = 32;

__global__ void __launch_bounds__(SCAN_THREADS)
harris_colscan_response_kernel(const double* __restrict__ sxx, const double* __restrict__ sxy, const double* __restrict__ syy,
                               int w, int h, float kf, float* __restrict__ R, GfttState* S)
{
    const int x = blockIdx.x * SCAN_THREADS + threadIdx.x;
    float vmax = -INFINITY;
    if (x < w) {
        const double* pl[3] = {sxx + x, sxy + x, syy + x};
        double ring[3][5];        // row sums of the five rows inside the window; slot = (row + 2) % 5
        double sum[3] = {0., 0., 0.};
#pragma unroll
        for (int i = -2; i <= 1; ++i) {
            const size_t ro = (size_t)reflect101(i, h) * w;
#pragma unroll
            for (int c = 0; c < 3; ++c) { ring[c][i + 2] = __ldg(pl[c] + ro); sum[c] = __dadd_rn(sum[c], ring[c][i + 2]); }
        }
        double nxt[3][5];         // entering rows y0+2 .. y0+6 of the current batch
#pragma unroll
        for (int j = 0; j < 5; ++j) {
            const size_t ro = (size_t)reflect101(min(j + 2, h + 1), h) * w;
#pragma unroll
            for (int c = 0; c < 3; ++c) nxt[c][j] = __ldg(pl[c] + ro);
        }
        for (int y0 = 0; y0 < h; y0 += 5) {
            double nxt2[3][5];
#pragma unroll
            for (int j = 0; j < 5; ++j) {      // rows (y0 + 5) + j + 2, clamped into the reflect-padded range (unused beyond it)
                const size_t ro = (size_t)reflect101(min(y0 + 7 + j, h + 1), h) * w;
#pragma unroll
                for (int c = 0; c < 3; ++c) nxt2[c][j] = __ldg(pl[c] + ro);
            }
#pragma unroll
            for (int j = 0; j < 5; ++j) {
                const int y = y0 + j;
                if (y < h) {
                    const int enter = (j + 4) % 5, leave = j;       // slots of rows y+2 and y-2 (y0 is a multiple of 5)
                    double s[3];
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        s[c] = __dadd_rn(sum[c], nxt[c][j]);
                        sum[c] = __dsub_rn(s[c], ring[c][leave]);
                        ring[c][enter] = nxt[c][j];
                    }
                    const float a = __double2float_rn(s[0]), b = __double2float_rn(s[1]), c = __double2float_rn(s[2]);
                    const float ac = __fadd_rn(a, c);
                    const float r = __fsub_rn(__fsub_rn(__fmul_rn(a, c), __fmul_rn(b, b)), __fmul_rn(kf, __fmul_rn(ac, ac)));
                    R[(size_t)y * w + x] = r;
                    vmax = fmaxf(vmax, r);
                }
            }
#pragma unroll
            for (int j = 0; j < 5; ++j)
#pragma unroll
                for (int c = 0; c < 3; ++c) nxt[c][j] = nxt2[c][j];
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
    if ((threadIdx.x & 31) == 0 && vmax > -INFINITY) atomicMax(&S->max_bits, orderable(vmax));
}

// step 5a: threshold + 3x3 non-maximum suppression, candidates appended as 64-bit keys
__global__ void __launch_bounds__(256)
gftt_candidates_kernel(const float* __restrict__ R, int w, int h, double quality, GfttState* S,
                       unsigned long long* __restrict__ keys, unsigned capacity)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x + 1, y = blockIdx.y + 1;
    if (x >= w - 1 || y >= h - 1) return;
    const float thr = (float)((double)from_orderable(S->max_bits) * quality);
    const float* c = R + (size_t)y * w + x;
    const float v = __ldg(c);
    if (!(v > thr) || v == 0.f) return;
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
        for (int dx = -1; dx <= 1; ++dx) {
            if (dx == 0 && dy == 0) continue;
            const float n = __ldg(c + dy * w + dx);
            if ((n > thr ? n : 0.f) > v) return;     // a thresholded neighbour is larger: not a local maximum
        }
    const unsigned slot = atomicAdd(&S->n_cand, 1u);
    if (slot < capacity) keys[slot] = ((unsigned long long)orderable(v) << 32) | (unsigned)(y * w + x);
}

// step 5b: radix select of the K-th largest key, 8 bits per pass from the top
__global__ void gftt_hist_kernel(const unsigned long long* __restrict__ keys, GfttState* S, int shift, unsigned capacity)
{
    __shared__ unsigned sh[256];
    sh[threadIdx.x] = 0u;
    __syncthreads();
    const unsigned n = min(S->n_cand, capacity);
    const unsigned long long prefix = S->prefix;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const unsigned long long k = keys[i];
        if (shift == 56 || (k >> (shift + 8)) == (prefix >> (shift + 8))) atomicAdd(&sh[(unsigned)(k >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (sh[threadIdx.x]) atomicAdd(&S->hist[threadIdx.x], sh[threadIdx.x]);
}

__global__ void gftt_pick_kernel(GfttState* S, int shift, unsigned max_corners, unsigned capacity)
{
    if (threadIdx.x == 0) {
        const unsigned n = min(S->n_cand, capacity);
        if (shift == 56) { S->n_out = min(n, max_corners); S->remaining = S->n_out; S->prefix = 0ull; }
        if (n > max_corners) {
            unsigned rem = S->remaining, cum = 0u;
            for (int b = 255; b >= 0; --b) {
                const unsigned c = S->hist[b];
                if (cum + c >= rem) { S->prefix |= (unsigned long long)b << shift; S->remaining = rem - cum; break; }
                cum += c;
            }
        }
    }
    __syncthreads();
    S->hist[threadIdx.x] = 0u;
}

__global__ void gftt_compact_kernel(const unsigned long long* __restrict__ keys, GfttState* S, unsigned long long* __restrict__ sel,
                                    unsigned max_corners, unsigned capacity)
{
    const unsigned n = min(S->n_cand, capacity);
    const unsigned long long kth = n > max_corners ? S->prefix : 0ull;     // keys are distinct: exactly n_out keys are >= kth
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const unsigned long long k = keys[i];
        if (k >= kth) {
            const unsigned slot = atomicAdd(&S->n_sel, 1u);
            if (slot < max_corners) sel[slot] = k;
        }
    }
}

// step 5c: rank the survivors by counting (keys are distinct), write (x, y) in cv2's order
__global__ void gftt_rank_kernel(const unsigned long long* __restrict__ sel, const GfttState* S, int w, float* __restrict__ out_xy,
                                 int* __restrict__ out_count)
{
    const unsigned n = S->n_out;
    if (blockIdx.x == 0 && threadIdx.x == 0) *out_count = (int)n;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const unsigned long long k = sel[i];
        unsigned rank = 0;
        for (unsigned j = 0; j < n; ++j) rank += sel[j] > k;
        const unsigned addr = (unsigned)k;
        out_xy[2 * rank] = (float)(addr % (unsigned)w);
        out_xy[2 * rank + 1] = (float)(addr / (unsigned)w);
    }
}

// ---- cv2.cornerSubPix (vidExample.py:113-115) ------------------------------------------------------------------------
// One thread per corner: the iteration is a short serial recurrence (float64 sums in raster order are part of the
// contract) over <= 1000 corners once per clip, so the kernel is about fidelity, not throughput.  The sampler and the
// update follow oracle/velocity_oracle.c::orc_corner_subpix_u8 operation by operation with explicit rounding
// intrinsics (nothing may be contracted); the float32 window mask is computed on the HOST with the C library's expf,
// which is what cv2 itself evaluates.
constexpr int SUBPIX_MAX_WIN = 7;                                        // half window; the reference uses 5
struct SubpixMask { float m[(2 * SUBPIX_MAX_WIN + 1) * (2 * SUBPIX_MAX_WIN + 1)]; };

__device__ void rect_subpix_8u32f(const uint8_t* __restrict__ src, int step, int cols, int rows, float* dst, int dw, int dh, float cx,
                                  float cy)
{
    const float centx = __fsub_rn(cx, (dw - 1) * 0.5f), centy = __fsub_rn(cy, (dh - 1) * 0.5f);
    const int ipx = __float2int_rd(centx), ipy = __float2int_rd(centy);
    const float a = __fsub_rn(centx, (float)ipx), b = __fsub_rn(centy, (float)ipy);
    const float b1 = __fsub_rn(1.f, b), b2 = b, a1 = __fsub_rn(1.f, a);
    const float a11 = __fmul_rn(a1, b1), a12 = __fmul_rn(a, b1), a21 = __fmul_rn(a1, b), a22 = __fmul_rn(a, b);
    if (0 <= ipx && ipx < cols - dw && 0 <= ipy && ipy < rows - dh) {
        const uint8_t* p = src + (size_t)ipy * step + ipx;
        for (int i = 0; i < dh; ++i, p += step, dst += dw)
            for (int j = 0; j < dw; ++j)
                dst[j] = __fadd_rn(__fadd_rn(__fmul_rn((float)__ldg(p + j), a11), __fmul_rn((float)__ldg(p + j + 1), a12)),
                                   __fadd_rn(__fmul_rn((float)__ldg(p + j + step), a21), __fmul_rn((float)__ldg(p + j + step + 1), a22)));
        return;
    }
    int rx = ipx >= 0 ? 0 : -ipx; if (rx > dw) rx = dw;
    int rw = ipx < cols - dw ? dw : cols - ipx - 1; if (rw < 0) rw = 0;
    const int ry = ipy >= 0 ? 0 : -ipy;
    int rh = ipy < rows - dh ? dh : rows - ipy - 1; if (rh < 0) rh = 0;
    for (int i = 0; i < dh; ++i, dst += dw) {
        int y0 = ipy + i; y0 = y0 < 0 ? 0 : (y0 >= rows ? rows - 1 : y0);
        const bool rep = (i < ry || i >= rh);
        const uint8_t* r0 = src + (size_t)y0 * step;
        const uint8_t* r1 = rep ? r0 : r0 + step;
        for (int j = 0; j < dw; ++j) {
            if (j < rx) dst[j] = __fadd_rn(__fmul_rn((float)__ldg(r0), b1), __fmul_rn((float)__ldg(r1), b2));
            else if (j >= rw) {
                int xe = ipx + rw - (i < ry ? 1 : 0);      // cv2 4.13: rows above the frame take column w-2 on the right side
                xe = xe < 0 ? 0 : (xe >= cols ? cols - 1 : xe);
                dst[j] = __fadd_rn(__fmul_rn((float)__ldg(r0 + xe), b1), __fmul_rn((float)__ldg(r1 + xe), b2));
            } else {
                const int x = ipx + j;
                if (rep) dst[j] = __fmaf_rn((float)__ldg(r0 + x + 1), a, __fmul_rn((float)__ldg(r0 + x), a1));
                else dst[j] = __fadd_rn(__fadd_rn(__fmul_rn((float)__ldg(r0 + x), a11), __fmul_rn((float)__ldg(r0 + x + 1), a12)),
                                        __fadd_rn(__fmul_rn((float)__ldg(r1 + x), a21), __fmul_rn((float)__ldg(r1 + x + 1), a22)));
            }
        }
    }
}

__global__ void __launch_bounds__(64)
corner_subpix_kernel(const uint8_t* __restrict__ src, int cols, int rows, int step, float* __restrict__ pts, int n, int winw, int winh,
                     int max_iters, double eps2, const SubpixMask mask)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int win_w = 2 * winw + 1, win_h = 2 * winh + 1, bw = win_w + 2;
    float buf[(2 * SUBPIX_MAX_WIN + 3) * (2 * SUBPIX_MAX_WIN + 3)];
    const float cTx = pts[2 * k], cTy = pts[2 * k + 1];
    float cIx = cTx, cIy = cTy;
    int iter = 0;
    double err = 0.;
    do {
        rect_subpix_8u32f(src, step, cols, rows, buf, bw, win_h + 2, cIx, cIy);
        double a = 0., b = 0., c = 0., bb1 = 0., bb2 = 0.;
        for (int i = 0, kk = 0; i < win_h; ++i) {
            const float* sp = buf + (i + 1) * bw + 1;
            const double py = (double)(i - winh);
            for (int j = 0; j < win_w; ++j, ++kk) {
                const double m = (double)mask.m[kk];
                const double tgx = (double)__fsub_rn(sp[j + 1], sp[j - 1]);
                const double tgy = (double)__fsub_rn(sp[j + bw], sp[j - bw]);
                const double gxx = __dmul_rn(__dmul_rn(tgx, tgx), m), gxy = __dmul_rn(__dmul_rn(tgx, tgy), m),
                             gyy = __dmul_rn(__dmul_rn(tgy, tgy), m);
                const double px = (double)(j - winw);
                a = __dadd_rn(a, gxx); b = __dadd_rn(b, gxy); c = __dadd_rn(c, gyy);
                bb1 = __dadd_rn(bb1, __dadd_rn(__dmul_rn(gxx, px), __dmul_rn(gxy, py)));
                bb2 = __dadd_rn(bb2, __dadd_rn(__dmul_rn(gxy, px), __dmul_rn(gyy, py)));
            }
        }
        const double det = __dsub_rn(__dmul_rn(a, c), __dmul_rn(b, b));
        if (fabs(det) <= DBL_EPSILON * DBL_EPSILON) break;
        const double scale = __ddiv_rn(1.0, det);
        const float nx = __double2float_rn(__dsub_rn(__dadd_rn((double)cIx, __dmul_rn(__dmul_rn(c, scale), bb1)), __dmul_rn(__dmul_rn(b, scale), bb2)));
        const float ny = __double2float_rn(__dadd_rn(__dsub_rn((double)cIy, __dmul_rn(__dmul_rn(b, scale), bb1)), __dmul_rn(__dmul_rn(a, scale), bb2)));
        const float ex = __fsub_rn(nx, cIx), ey = __fsub_rn(ny, cIy);
        err = (double)__fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey));
        if (nx < 0.f || nx >= (float)cols || ny < 0.f || ny >= (float)rows) break;     // an update that leaves the frame is discarded
        cIx = nx; cIy = ny;
    } while (++iter < max_iters && err > eps2);
    if (fabsf(__fsub_rn(cIx, cTx)) > (float)winw || fabsf(__fsub_rn(cIy, cTy)) > (float)winh) { cIx = cTx; cIy = cTy; }
    pts[2 * k] = cIx;
    pts[2 * k + 1] = cIy;
}

size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

}  // namespace

VEL_API size_t vel_good_features_workspace(int32_t width, int32_t height, int32_t max_corners)
{
    if (width <= 0 || height <= 0 || max_corners <= 0) return 0;
    const size_t px = (size_t)width * height;
    // control block | 3 cov planes + R (float) | 3 row-sum planes (double; the candidate keys reuse the first) | survivors
    return align256(sizeof(GfttState)) + 4 * align256(px * sizeof(float)) + 3 * align256(px * sizeof(double)) +
           align256((size_t)max_corners * sizeof(unsigned long long));
}

VEL_API int vel_good_features_harris_u8(const uint8_t* img, int32_t width, int32_t height, int32_t pitch, int32_t max_corners,
                                        double quality, int32_t block_size, double k, void* work, size_t work_bytes,
                                        float* response, float* out_xy, int32_t* out_count, vel_stream_t stream)
{
    VEL_CHECK_ARG(img && work && out_xy && out_count, "vel_good_features_harris_u8: NULL argument");
    VEL_CHECK_ARG(width >= 8 && height >= 8 && pitch >= width, "vel_good_features_harris_u8: image %dx%d (pitch %d) too small", width,
                  height, pitch);
    VEL_CHECK_ARG((long long)width * height < (1ll << 31), "vel_good_features_harris_u8: image too large");
    VEL_CHECK_ARG(max_corners > 0 && max_corners <= 65536, "vel_good_features_harris_u8: max_corners %d outside [1,65536]", max_corners);
    VEL_CHECK_ARG(quality > 0. && quality < 1., "vel_good_features_harris_u8: quality must be in (0,1)");
    if (block_size != 5) {
        vel_set_error("vel_good_features_harris_u8: only blockSize 5 (the reference's value, vidExample.py:110) is implemented");
        return VEL_ERR_UNSUPPORTED;
    }
    VEL_CHECK_ARG(work_bytes >= vel_good_features_workspace(width, height, max_corners), "vel_good_features_harris_u8: workspace too small");

    const size_t px = (size_t)width * height;
    char* base = static_cast<char*>(work);
    GfttState* S = reinterpret_cast<GfttState*>(base); base += align256(sizeof(GfttState));
    float* cxx = reinterpret_cast<float*>(base); base += align256(px * sizeof(float));
    float* cxy = reinterpret_cast<float*>(base); base += align256(px * sizeof(float));
    float* cyy = reinterpret_cast<float*>(base); base += align256(px * sizeof(float));
    float* R = reinterpret_cast<float*>(base); base += align256(px * sizeof(float));
    double* sxx = reinterpret_cast<double*>(base); base += align256(px * sizeof(double));
    double* sxy = reinterpret_cast<double*>(base); base += align256(px * sizeof(double));
    double* syy = reinterpret_cast<double*>(base); base += align256(px * sizeof(double));
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(sxx);     // the row sums are dead once R exists
    unsigned long long* sel = reinterpret_cast<unsigned long long*>(base);
    if (response) R = response;     // caller wants the response plane (float32 [height][width])

    cudaStream_t st = (cudaStream_t)stream;
    const float k1 = (float)(1.0 / (4.0 * block_size * 255.0)), k0 = 2.f * k1;
    gftt_reset_kernel<<<1, 256, 0, st>>>(S);
    harris_cov_kernel<<<dim3((width + 255) / 256, height), 256, 0, st>>>(img, width, height, pitch, k1, k0, (width / 32) * 32, cxx, cxy, cyy);
    harris_rowsum_kernel<<<dim3((width + 255) / 256, height), 256, 0, st>>>(cxx, cxy, cyy, width, height, sxx, sxy, syy);
    harris_colscan_response_kernel<<<(width + SCAN_THREADS - 1) / SCAN_THREADS, SCAN_THREADS, 0, st>>>(sxx, sxy, syy, width, height, (float)k, R, S);
    const unsigned capacity = (unsigned)px;
    gftt_candidates_kernel<<<dim3((width - 2 + 255) / 256, height - 2), 256, 0, st>>>(R, width, height, quality, S, keys, capacity);
    const int sel_grid = 4 * kNumSMs;
    for (int shift = 56; shift >= 0; shift -= 8) {
        gftt_hist_kernel<<<sel_grid, 256, 0, st>>>(keys, S, shift, capacity);
        gftt_pick_kernel<<<1, 256, 0, st>>>(S, shift, (unsigned)max_corners, capacity);
    }
    gftt_compact_kernel<<<sel_grid, 256, 0, st>>>(keys, S, sel, (unsigned)max_corners, capacity);
    gftt_rank_kernel<<<(max_corners + 255) / 256, 256, 0, st>>>(sel, S, width, out_xy, out_count);
    VEL_LAUNCH_CHECK("good-features kernels");
    return VEL_OK;
}

VEL_API int vel_corner_subpix_u8(const uint8_t* img, int32_t width, int32_t height, int32_t pitch, float* pts, int32_t npts, int32_t win_w,
                                 int32_t win_h, int32_t max_iters, double eps, vel_stream_t stream)
{
    VEL_CHECK_ARG(npts >= 0, "vel_corner_subpix_u8: npts < 0");
    if (npts == 0) return VEL_OK;
    VEL_CHECK_ARG(img && pts, "vel_corner_subpix_u8: NULL argument");
    VEL_CHECK_ARG(width > 0 && height > 0 && pitch >= width, "vel_corner_subpix_u8: bad image %dx%d pitch %d", width, height, pitch);
    VEL_CHECK_ARG(win_w >= 1 && win_h >= 1 && win_w <= SUBPIX_MAX_WIN && win_h <= SUBPIX_MAX_WIN,
                  "vel_corner_subpix_u8: half window %dx%d outside [1,%d]", win_w, win_h, SUBPIX_MAX_WIN);
    VEL_CHECK_ARG(width >= 2 * win_w + 5 && height >= 2 * win_h + 5, "vel_corner_subpix_u8: image smaller than the sampling window");
    max_iters = max_iters < 1 ? 1 : (max_iters > 100 ? 100 : max_iters);      // cv2: MIN(MAX(maxCount, 1), 100)
    eps = eps < 0. ? 0. : eps;
    SubpixMask mask;
    const int ww = 2 * win_w + 1, wh = 2 * win_h + 1;
    for (int i = 0; i < wh; ++i) {
        const float y = (float)(i - win_h) / win_h;
        const float vy = expf(-y * y);
        for (int j = 0; j < ww; ++j) {
            const float x = (float)(j - win_w) / win_w;
            mask.m[i * ww + j] = (float)(vy * expf(-x * x));
        }
    }
    corner_subpix_kernel<<<(npts + 63) / 64, 64, 0, (cudaStream_t)stream>>>(img, width, height, pitch, pts, npts, win_w, win_h, max_iters,
                                                                          eps * eps, mask);
    VEL_LAUNCH_CHECK("corner_subpix_kernel");
    return VEL_OK;
}
