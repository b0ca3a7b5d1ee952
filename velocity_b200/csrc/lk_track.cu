// K2: pyramidal Lucas-Kanade tracker with fused forward-backward check.
//
// Replaces cv2.calcOpticalFlowPyrLK as the reference drives it through cv2calcOpticalFlowPyrLK
// (utils/KLT.py:37-51): forward pass, optional backward pass from the forward result, and
// v = st_fwd & st_bwd & (||p1 - p1'||_2 < fbt).  Arithmetic follows OpenCV 4.13's LKTrackerInvoker
// (restated on the CPU in oracle/velocity_oracle.c, which is pinned against cv2 by tests/golden):
//   * bilinear weights in 14-bit fixed point (round-half-even of float32 products),
//   * template I with 5 fractional bits, Scharr derivatives (REFLECT_101 inside the image,
//     constant 0 in the window padding) interpolated with the same weights,
//   * G = sum [Ix^2, IxIy; IxIy, Iy^2] and b = sum diff*[Ix, Iy] accumulated EXACTLY in int64 and
//     converted to float32 once (cv2 accumulates float32 SIMD lanes; <= 2e-3 px apart, see tests),
//   * float32 2x2 solve, eps^2 / oscillation stopping rules, minEig and bounds status rules.
// All float32 steps use explicit round-to-nearest intrinsics so no FMA contraction can make the
// device differ from the oracle: device == oracle bit for bit.
//
// Kernels in this file (all produce identical bits; the dispatcher at the bottom picks by window size and pitch):
//   lk_track_w15h_kernel   (lk_w15h.cuh) 15x15, pitches % 4 == 0: two points per warp, word gathers + DP2A  -- default
//   lk_track_w15_kernel    15x15, any pitch: warp per point, byte gathers, REDUX sums (VEL_LK_W15=bytes forces it)
//   lk_track_cols_kernel   16..63 px wide windows (51x51 lk_fine, cv2's default 21x21): CTA per point, template in smem
//   lk_track_kernel        every other window up to 127x127: group of 32 / 128 threads per point, int64 sums
// J is gathered from global memory through L1/L2 (the whole pyramid of a 1080p frame is 2.7 MB, i.e. L2 resident).
#include <cuda.h>

#include "common.cuh"
#include <stdlib.h>
#include <string.h>

namespace {

struct LkLevels {
    int max_level;
    int w[VEL_MAX_LEVELS], h[VEL_MAX_LEVELS], pitch[VEL_MAX_LEVELS];
    long long off[VEL_MAX_LEVELS];
};

struct LkArgs {
    const uint8_t* prev0; long long prev_stride; int prev_pitch;
    const uint8_t* prev_pyr; long long prev_pyr_stride;
    const uint8_t* next0; long long next_stride; int next_pitch;
    const uint8_t* next_pyr; long long next_pyr_stride;
    LkLevels lv;
    const float* pts; long long pts_stride; int npts;
    float* out; uint8_t* status; float* err; float* back;
    int win_w, win_h, max_count;
    float eps2, min_eig, fbt;
    int force_bytes;   // VEL_LK_W15=bytes: byte-gather paths only (cross-check of the word-gather paths)
    // sequence form (lk_track_w15h_kernel<true>): pairs k = 0..seq_pairs-1 are frame k -> k+1 of ONE frame run, row k+1 of
    // out / alive is produced from row k inside the kernel; alive is [seq_pairs+1][npts] with row 0 given
    int seq_pairs;
    uint8_t* alive;
};

struct Img {
    const uint8_t* p;
    int w, h, pitch;
};

__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }

__device__ __forceinline__ int px_reflect(const Img& im, int x, int y)
{
    return (int)ldg_u8(im.p + (long long)reflect101_1(y, im.h) * im.pitch + reflect101_1(x, im.w));
}

// unnormalised Scharr pair at an in-image pixel, REFLECT_101 at the edge (OpenCV calcSharrDeriv)
__device__ __forceinline__ void scharr_at(const Img& im, int x, int y, int& gx, int& gy)
{
    const int xm = x > 0 ? x - 1 : (im.w > 1 ? 1 : 0);
    const int xp = x < im.w - 1 ? x + 1 : (im.w > 1 ? im.w - 2 : 0);
    const int ym = y > 0 ? y - 1 : (im.h > 1 ? 1 : 0);
    const int yp = y < im.h - 1 ? y + 1 : (im.h > 1 ? im.h - 2 : 0);
    const uint8_t* r0 = im.p + (long long)ym * im.pitch;
    const uint8_t* r1 = im.p + (long long)y * im.pitch;
    const uint8_t* r2 = im.p + (long long)yp * im.pitch;
    const int a0 = ldg_u8(r0 + xm), a1 = ldg_u8(r0 + x), a2 = ldg_u8(r0 + xp);
    const int b0 = ldg_u8(r1 + xm), b2 = ldg_u8(r1 + xp);
    const int c0 = ldg_u8(r2 + xm), c1 = ldg_u8(r2 + x), c2 = ldg_u8(r2 + xp);
    gx = 3 * (a2 + c2) + 10 * b2 - 3 * (a0 + c0) - 10 * b0;
    gy = 3 * ((c0 - a0) + (c2 - a2)) + 10 * (c1 - a1);
}

struct Weights {
    int w00, w01, w10, w11;
};

__device__ __forceinline__ Weights bilin_weights(float a, float b)
{
    const float one_a = fsub(1.f, a), one_b = fsub(1.f, b);
    Weights w;
    w.w00 = __float2int_rn(fmul(fmul(one_a, one_b), 16384.f));
    w.w01 = __float2int_rn(fmul(fmul(a, one_b), 16384.f));
    w.w10 = __float2int_rn(fmul(fmul(one_a, b), 16384.f));
    w.w11 = 16384 - w.w00 - w.w01 - w.w10;
    return w;
}

// ---- group-wide exact reductions --------------------------------------------------------------
template <int NT>
__device__ __forceinline__ void group_sync()
{
    if (NT == 32) __syncwarp();
    else __syncthreads();
}

__device__ __forceinline__ long long warp_sum_ll(long long v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Sums K int64 values over the group; every thread receives the totals.  `red` is a per-group
// shared scratch of 2 * (NT/32) * K int64 (double-buffered by `parity` so one barrier suffices).
template <int NT, int K>
__device__ __forceinline__ void group_sum(long long (&v)[K], long long* red, int& parity, int tid)
{
#pragma unroll
    for (int k = 0; k < K; ++k) v[k] = warp_sum_ll(v[k]);
    if (NT > 32) {
        constexpr int NW = NT / 32;
        long long* buf = red + parity * (NW * K);
        if ((tid & 31) == 0) {
#pragma unroll
            for (int k = 0; k < K; ++k) buf[(tid >> 5) * K + k] = v[k];
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < K; ++k) {
            long long s = 0;
#pragma unroll
            for (int w = 0; w < NW; ++w) s += buf[w * K + k];
            v[k] = s;
        }
        parity ^= 1;
    }
}

// ---- one pass (all levels) for one point --------------------------------------------------------
// I: template pyramid, J: search pyramid.  Returns status (1/0); writes next position and err.
template <int NT>
__device__ void track_point(const LkArgs& A, const uint8_t* I0, int I0_pitch, const uint8_t* Ipyr, const uint8_t* J0,
                            int J0_pitch, const uint8_t* Jpyr, float px, float py, float& out_x, float& out_y, int& out_status,
                            float& out_err, short2* sD, short* sI, short2* sG, long long* red, int& parity, int tid)
{
    const int ww = A.win_w, wh = A.win_h, npx = ww * wh;
    const int tw = ww + 1, th = wh + 1;
    const float half_x = fmul((float)(ww - 1), 0.5f), half_y = fmul((float)(wh - 1), 0.5f);
    const float FLT_SCALE = 1.f / 1048576.f;

    int status = 1;
    float err = 0.f;
    float next_x = 0.f, next_y = 0.f;

    for (int level = A.lv.max_level; level >= 0; --level) {
        Img I, J;
        I.w = J.w = A.lv.w[level];
        I.h = J.h = A.lv.h[level];
        if (level == 0) { I.p = I0; I.pitch = I0_pitch; J.p = J0; J.pitch = J0_pitch; }
        else { I.p = Ipyr + A.lv.off[level]; J.p = Jpyr + A.lv.off[level]; I.pitch = J.pitch = A.lv.pitch[level]; }

        const float scale = 1.f / (float)(1 << level);
        float prev_x = fmul(px, scale), prev_y = fmul(py, scale);
        float nx, ny;
        if (level == A.lv.max_level) { nx = prev_x; ny = prev_y; }
        else { nx = fmul(next_x, 2.f); ny = fmul(next_y, 2.f); }
        next_x = nx; next_y = ny;

        prev_x = fsub(prev_x, half_x); prev_y = fsub(prev_y, half_y);
        const int ipx = __float2int_rd(prev_x), ipy = __float2int_rd(prev_y);
        if (ipx < -ww || ipx >= I.w || ipy < -wh || ipy >= I.h) {
            if (level == 0) { status = 0; err = 0.f; }
            continue;
        }
        Weights w = bilin_weights(fsub(prev_x, (float)ipx), fsub(prev_y, (float)ipy));

        group_sync<NT>();  // previous level's readers of sD/sI/sG are done
        // (1) Scharr derivative tile over the (ww+1) x (wh+1) bilinear footprint
        for (int t = tid; t < tw * th; t += NT) {
            const int ty = t / tw, tx = t - ty * tw;
            const int X = ipx + tx, Y = ipy + ty;
            int gx = 0, gy = 0;
            if (X >= 0 && Y >= 0 && X < I.w && Y < I.h) scharr_at(I, X, Y, gx, gy);
            sD[t] = make_short2((short)gx, (short)gy);
        }
        group_sync<NT>();
        // (2) template patch + gradient matrix
        long long acc[3] = {0, 0, 0};
        {
            const bool inside = ipx >= 0 && ipy >= 0 && ipx + ww < I.w && ipy + wh < I.h;
            for (int i = tid; i < npx; i += NT) {
                const int y = i / ww, x = i - y * ww;
                int i00, i01, i10, i11;
                if (inside) {
                    const uint8_t* r = I.p + (long long)(ipy + y) * I.pitch + ipx + x;
                    i00 = ldg_u8(r); i01 = ldg_u8(r + 1); i10 = ldg_u8(r + I.pitch); i11 = ldg_u8(r + I.pitch + 1);
                } else {
                    i00 = px_reflect(I, ipx + x, ipy + y); i01 = px_reflect(I, ipx + x + 1, ipy + y);
                    i10 = px_reflect(I, ipx + x, ipy + y + 1); i11 = px_reflect(I, ipx + x + 1, ipy + y + 1);
                }
                const int ival = (i00 * w.w00 + i01 * w.w01 + i10 * w.w10 + i11 * w.w11 + (1 << 8)) >> 9;
                const short2 d00 = sD[y * tw + x], d01 = sD[y * tw + x + 1], d10 = sD[(y + 1) * tw + x],
                             d11 = sD[(y + 1) * tw + x + 1];
                const int ix = (d00.x * w.w00 + d01.x * w.w01 + d10.x * w.w10 + d11.x * w.w11 + (1 << 13)) >> 14;
                const int iy = (d00.y * w.w00 + d01.y * w.w01 + d10.y * w.w10 + d11.y * w.w11 + (1 << 13)) >> 14;
                sI[i] = (short)ival;
                sG[i] = make_short2((short)ix, (short)iy);
                acc[0] += (long long)(ix * ix);
                acc[1] += (long long)(ix * iy);
                acc[2] += (long long)(iy * iy);
            }
        }
        if (NT == 32) __syncwarp();  // template visible to the whole warp (NT > 32: barrier inside group_sum)
        group_sum<NT, 3>(acc, red, parity, tid);
        const float A11 = fmul(__ll2float_rn(acc[0]), FLT_SCALE), A12 = fmul(__ll2float_rn(acc[1]), FLT_SCALE),
                    A22 = fmul(__ll2float_rn(acc[2]), FLT_SCALE);
        float D = fsub(fmul(A11, A22), fmul(A12, A12));
        const float dA = fsub(A11, A22);
        const float disc = fadd(fmul(dA, dA), fmul(fmul(4.f, A12), A12));
        const float min_eig = __fdiv_rn(fsub(fadd(A22, A11), __fsqrt_rn(disc)), (float)(2 * ww * wh));
        if (min_eig < A.min_eig || D < 1.1920928955078125e-07f) {
            if (level == 0) status = 0;
            continue;
        }
        D = __fdiv_rn(1.f, D);

        nx = fsub(nx, half_x); ny = fsub(ny, half_y);
        float pdx = 0.f, pdy = 0.f;
        for (int j = 0; j < A.max_count; ++j) {
            const int inx = __float2int_rd(nx), iny = __float2int_rd(ny);
            if (inx < -ww || inx >= J.w || iny < -wh || iny >= J.h) {
                if (level == 0) status = 0;
                break;
            }
            w = bilin_weights(fsub(nx, (float)inx), fsub(ny, (float)iny));
            long long b[2] = {0, 0};
            const bool inside = inx >= 0 && iny >= 0 && inx + ww < J.w && iny + wh < J.h;
            if (inside) {
                const uint8_t* base = J.p + (long long)iny * J.pitch + inx;
                for (int i = tid; i < npx; i += NT) {
                    const int y = i / ww, x = i - y * ww;
                    const uint8_t* r = base + (long long)y * J.pitch + x;
                    const int j00 = ldg_u8(r), j01 = ldg_u8(r + 1), j10 = ldg_u8(r + J.pitch), j11 = ldg_u8(r + J.pitch + 1);
                    const int diff = ((j00 * w.w00 + j01 * w.w01 + j10 * w.w10 + j11 * w.w11 + (1 << 8)) >> 9) - (int)sI[i];
                    const short2 g = sG[i];
                    b[0] += (long long)(diff * (int)g.x);
                    b[1] += (long long)(diff * (int)g.y);
                }
            } else {
                for (int i = tid; i < npx; i += NT) {
                    const int y = i / ww, x = i - y * ww;
                    const int j00 = px_reflect(J, inx + x, iny + y), j01 = px_reflect(J, inx + x + 1, iny + y);
                    const int j10 = px_reflect(J, inx + x, iny + y + 1), j11 = px_reflect(J, inx + x + 1, iny + y + 1);
                    const int diff = ((j00 * w.w00 + j01 * w.w01 + j10 * w.w10 + j11 * w.w11 + (1 << 8)) >> 9) - (int)sI[i];
                    const short2 g = sG[i];
                    b[0] += (long long)(diff * (int)g.x);
                    b[1] += (long long)(diff * (int)g.y);
                }
            }
            group_sum<NT, 2>(b, red, parity, tid);
            const float b1 = fmul(__ll2float_rn(b[0]), FLT_SCALE), b2 = fmul(__ll2float_rn(b[1]), FLT_SCALE);
            const float dx = fmul(fsub(fmul(A12, b2), fmul(A22, b1)), D);
            const float dy = fmul(fsub(fmul(A12, b1), fmul(A11, b2)), D);
            nx = fadd(nx, dx); ny = fadd(ny, dy);
            next_x = fadd(nx, half_x); next_y = fadd(ny, half_y);
            if (fadd(fmul(dx, dx), fmul(dy, dy)) <= A.eps2) break;
            if (j > 0 && fabsf(fadd(dx, pdx)) < 0.01f && fabsf(fadd(dy, pdy)) < 0.01f) {
                next_x = fsub(next_x, fmul(dx, 0.5f));
                next_y = fsub(next_y, fmul(dy, 0.5f));
                break;
            }
            pdx = dx; pdy = dy;
        }

        if (level == 0 && status) {
            const float fx = fsub(next_x, half_x), fy = fsub(next_y, half_y);
            const int inx = __float2int_rd(fx), iny = __float2int_rd(fy);
            if (inx < -ww || inx >= J.w || iny < -wh || iny >= J.h) {
                status = 0;
            } else {
                w = bilin_weights(fsub(fx, (float)inx), fsub(fy, (float)iny));
                long long e[1] = {0};
                for (int i = tid; i < npx; i += NT) {
                    const int y = i / ww, x = i - y * ww;
                    const int j00 = px_reflect(J, inx + x, iny + y), j01 = px_reflect(J, inx + x + 1, iny + y);
                    const int j10 = px_reflect(J, inx + x, iny + y + 1), j11 = px_reflect(J, inx + x + 1, iny + y + 1);
                    const int diff = ((j00 * w.w00 + j01 * w.w01 + j10 * w.w10 + j11 * w.w11 + (1 << 8)) >> 9) - (int)sI[i];
                    e[0] += (long long)abs(diff);
                }
                group_sum<NT, 1>(e, red, parity, tid);
                err = __fdiv_rn(__ll2float_rn(e[0]), (float)(32 * ww * wh));
            }
        }
    }
    out_x = next_x; out_y = next_y; out_status = status; out_err = err;
}

template <int NT, int GROUPS>
__global__ void __launch_bounds__(NT * GROUPS)
lk_track_kernel(const LkArgs A)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int group = threadIdx.x / NT, tid = threadIdx.x % NT;
    const int pair = blockIdx.y;
    const int pt = blockIdx.x * GROUPS + group;
    if (NT == 32 && pt >= A.npts) return;  // warp-granular groups may exit independently

    const int npx = A.win_w * A.win_h, ntile = (A.win_w + 1) * (A.win_h + 1);
    // per-group shared slices: sD (short2 x ntile) | sG (short2 x npx) | sI (short x npx) | red
    const size_t bytes_D = ((size_t)ntile * 4 + 15) & ~(size_t)15;
    const size_t bytes_G = ((size_t)npx * 4 + 15) & ~(size_t)15;
    const size_t bytes_I = ((size_t)npx * 2 + 15) & ~(size_t)15;
    const size_t bytes_R = (NT > 32) ? (size_t)2 * (NT / 32) * 3 * sizeof(long long) : 0;
    unsigned char* base = smem_raw + (size_t)group * (bytes_D + bytes_G + bytes_I + bytes_R);
    short2* sD = reinterpret_cast<short2*>(base);
    short2* sG = reinterpret_cast<short2*>(base + bytes_D);
    short* sI = reinterpret_cast<short*>(base + bytes_D + bytes_G);
    long long* red = reinterpret_cast<long long*>(base + bytes_D + bytes_G + bytes_I);
    int parity = 0;

    const uint8_t* P0 = A.prev0 + (long long)pair * A.prev_stride;
    const uint8_t* Pp = A.prev_pyr ? A.prev_pyr + (long long)pair * A.prev_pyr_stride : nullptr;
    const uint8_t* N0 = A.next0 + (long long)pair * A.next_stride;
    const uint8_t* Np = A.next_pyr ? A.next_pyr + (long long)pair * A.next_pyr_stride : nullptr;

    const float* pin = A.pts + (long long)pair * A.pts_stride + 2ll * pt;
    const float px = __ldg(pin), py = __ldg(pin + 1);

    float fx, fy, ferr;
    int fst;
    track_point<NT>(A, P0, A.prev_pitch, Pp, N0, A.next_pitch, Np, px, py, fx, fy, fst, ferr, sD, sI, sG, red, parity, tid);
    int st = fst;
    float bx = 0.f, by = 0.f;
    if (A.fbt >= 0.f && fst) {
        float berr;
        int bst;
        track_point<NT>(A, N0, A.next_pitch, Np, P0, A.prev_pitch, Pp, fx, fy, bx, by, bst, berr, sD, sI, sG, red, parity, tid);
        const float ddx = fsub(px, bx), ddy = fsub(py, by);
        const float fbe = __fsqrt_rn(fadd(fmul(ddx, ddx), fmul(ddy, ddy)));
        st = bst && (fbe < A.fbt);
    }
    if (tid == 0) {
        const long long o = (long long)pair * A.npts + pt;
        A.out[2 * o] = fx;
        A.out[2 * o + 1] = fy;
        A.status[o] = (uint8_t)st;
        A.err[o] = fst ? ferr : 0.f;
        if (A.back) { A.back[2 * o] = bx; A.back[2 * o + 1] = by; }
    }
}

// =====================================================================================================
// Specialised path for the 15x15 window (the reference's lk_coarse, utils/KLT.py:106, and BASELINE C2).
//
// One warp per point.  lane = (h, c): c = lane & 15 is a column of the 16x16 bilinear footprint,
// h = lane >> 4 selects rows 8h..8h+7.  The template (I, Ix, Iy of the lane's 8 pixels) lives in
// REGISTERS for the whole level; per iteration each lane gathers its 9 J bytes, forms the left- and
// right-column halves of the bilinear sum (the right half travels one lane down by shuffle) and the
// exact integer sums are reduced with REDUX (two 16-bit halves per value).  The Scharr tile is built
// from a shared-memory staged 18x18 region with dp4a on packed 4-byte windows.  Arithmetic is the
// same integer/float32 sequence as the generic path and the oracle: results are bit-identical.
constexpr int W15 = 15;
constexpr int W15_REGION_PITCH = 20;
constexpr int W15_REGION_BYTES = 384;  // 19 rows x 20 B, padded

__device__ __forceinline__ int dp4a_us(unsigned a, int b, int c)
{
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

// Exact 32-lane sum of per-lane int32 partials (the total may need 34 bits), returned as
// float(sum) * 2^-20 with a single rounding -- bit-identical to (float)(int64 sum) * 2^-20:
// the two REDUX halves (hi * 65536 and lo) are each exactly representable in float32, and one FMA
// rounds their exact sum once (power-of-two scaling commutes with the rounding).
__device__ __forceinline__ float warp_sum_scaled(int v)
{
    const unsigned lo = (unsigned)v & 0xffffu;
    const int hi = v >> 16;
    const int shi = __reduce_add_sync(0xffffffffu, hi);        // |shi| < 2^18
    const unsigned slo = __reduce_add_sync(0xffffffffu, lo);   // < 2^21
    return __fmaf_rn((float)shi, 0.0625f, __fmul_rn((float)slo, 1.f / 1048576.f));
}

struct W15Patch {
    int I[8], gx[8], gy[8];
};

// single-overshoot REFLECT_101 clamped into range (far corners of the staged region can overshoot
// twice on 16-pixel-wide top levels; those entries are never used but must stay in bounds)
__device__ __forceinline__ unsigned reflect_safe(int i, int n)
{
    i = i < 0 ? -i : i;
    i = i >= n ? 2 * n - 2 - i : i;
    return (unsigned)max(0, min(i, n - 1));
}

// the lane's 9 J bytes of the 16x16 footprint at integer position (inx, iny); 32-bit offsets from the
// (warp-uniform) level base keep the address arithmetic to one add per load
__device__ __forceinline__ void w15_gather(const Img& J, int inx, int iny, int h, int c, int (&jv)[9])
{
    // 0 <= inx <= w - 16  <=>  (unsigned)inx <= (unsigned)(w - 16)   (w >= 16 on every level)
    const bool inside = (unsigned)inx <= (unsigned)(J.w - 16) && (unsigned)iny <= (unsigned)(J.h - 16);
    const unsigned pitch = (unsigned)J.pitch;
    if (inside) {
        const unsigned off = (unsigned)(iny + 8 * h) * pitch + (unsigned)(inx + c);
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            const unsigned rr = (k == 8) ? (h ? 7u : 8u) : (unsigned)k;   // tile row 16 does not exist: re-read row 15
            jv[k] = (int)__ldg(J.p + (off + rr * pitch));
        }
    } else {
        const unsigned xx = reflect_safe(inx + c, J.w);
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            const int rr = min(8 * h + k, 15);
            jv[k] = (int)__ldg(J.p + (reflect_safe(iny + rr, J.h) * pitch + xx));
        }
    }
}

// diff = ((sum + 256) >> 9) - I  ==  (sum + (256 - (I << 9))) >> 9 ; P.I holds 256 - (I << 9)
__device__ __forceinline__ void w15_diff(const int (&jv)[9], const Weights& w, const W15Patch& P, int (&diff)[8])
{
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int left = jv[i] * w.w00 + (jv[i + 1] * w.w10 + P.I[i]);
        const int right = jv[i] * w.w01 + jv[i + 1] * w.w11;          // belongs to the pixel one column to the left
        diff[i] = (left + __shfl_down_sync(0xffffffffu, right, 1)) >> 9;
    }
}

constexpr int W15_WARPS = 8;

// One warp tracks one point: forward pass, then (fused) the backward pass from the forward result.
// A single instance of the level / iteration code serves both passes and the final error
// evaluation, which keeps the kernel inside the instruction cache.
__global__ void __launch_bounds__(32 * W15_WARPS, 4)
lk_track_w15_kernel(const LkArgs A)
{
    __shared__ __align__(16) uint8_t s_region[W15_WARPS][W15_REGION_BYTES];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int pair = blockIdx.y;
    const int pt = blockIdx.x * W15_WARPS + warp;
    if (pt >= A.npts) return;
    uint8_t* sreg = s_region[warp];
    const int h = lane >> 4, c = lane & 15;
    const float half = 7.0f;

    const uint8_t* P0 = A.prev0 + (long long)pair * A.prev_stride;
    const uint8_t* Pp = A.prev_pyr ? A.prev_pyr + (long long)pair * A.prev_pyr_stride : nullptr;
    const uint8_t* N0 = A.next0 + (long long)pair * A.next_stride;
    const uint8_t* Np = A.next_pyr ? A.next_pyr + (long long)pair * A.next_pyr_stride : nullptr;
    const float* pin = A.pts + (long long)pair * A.pts_stride + 2ll * pt;
    const float px0 = __ldg(pin), py0 = __ldg(pin + 1);

    float fx = 0.f, fy = 0.f, ferr = 0.f, bx = 0.f, by = 0.f;
    int fst = 0, st = 0;
    const int npass = A.fbt >= 0.f ? 2 : 1;

    for (int pass = 0; pass < npass; ++pass) {
        // pass 0: template = prev, search = next, start at the input point
        // pass 1: roles swapped, start at the forward result
        const uint8_t* I0 = pass ? N0 : P0;
        const uint8_t* Ipyr = pass ? Np : Pp;
        const uint8_t* J0 = pass ? P0 : N0;
        const uint8_t* Jpyr = pass ? Pp : Np;
        const int I0_pitch = pass ? A.next_pitch : A.prev_pitch, J0_pitch = pass ? A.prev_pitch : A.next_pitch;
        const float px = pass ? fx : px0, py = pass ? fy : py0;

        int status = 1;
        float err = 0.f;
        float next_x = 0.f, next_y = 0.f;

        for (int level = A.lv.max_level; level >= 0; --level) {
            Img I, J;
            I.w = J.w = A.lv.w[level];
            I.h = J.h = A.lv.h[level];
            if (level == 0) { I.p = I0; I.pitch = I0_pitch; J.p = J0; J.pitch = J0_pitch; }
            else { I.p = Ipyr + A.lv.off[level]; J.p = Jpyr + A.lv.off[level]; I.pitch = J.pitch = A.lv.pitch[level]; }

            const float scale = 1.f / (float)(1 << level);
            float prev_x = fmul(px, scale), prev_y = fmul(py, scale);
            float nx, ny;
            if (level == A.lv.max_level) { nx = prev_x; ny = prev_y; }
            else { nx = fmul(next_x, 2.f); ny = fmul(next_y, 2.f); }
            next_x = nx; next_y = ny;

            prev_x = fsub(prev_x, half); prev_y = fsub(prev_y, half);
            const int ipx = __float2int_rd(prev_x), ipy = __float2int_rd(prev_y);
            if (ipx < -W15 || ipx >= I.w || ipy < -W15 || ipy >= I.h) {
                if (level == 0) { status = 0; err = 0.f; }
                continue;
            }
            Weights w = bilin_weights(fsub(prev_x, (float)ipx), fsub(prev_y, (float)ipy));

            // ---- stage the 18x18 region (rows ipy-1.., cols ipx-1..) in shared memory ------------------
            const int rx0 = ipx - 1, ry0 = ipy - 1;
            const bool interior = rx0 >= 0 && ry0 >= 0 && rx0 + 17 < I.w && ry0 + 17 < I.h;
            __syncwarp();
            {
                const unsigned pitch = (unsigned)I.pitch;
                // column offsets of this lane's two region columns (c, and 16/17 for the tail entries)
                const int k1 = lane + 32;                       // second tail entry index (>= 36 for lanes >= 4)
                const unsigned trow0 = lane >> 1, tcol0 = 16 + (lane & 1);
                const unsigned trow1 = k1 >> 1, tcol1 = 16 + (k1 & 1);
                if (interior) {
                    const unsigned base = (unsigned)ry0 * pitch + (unsigned)rx0;
#pragma unroll
                    for (int q = 0; q < 9; ++q) {
                        const unsigned row = 2 * q + h;
                        sreg[row * W15_REGION_PITCH + c] = __ldg(I.p + (base + row * pitch + c));
                    }
                    sreg[trow0 * W15_REGION_PITCH + tcol0] = __ldg(I.p + (base + trow0 * pitch + tcol0));
                    if (k1 < 36) sreg[trow1 * W15_REGION_PITCH + tcol1] = __ldg(I.p + (base + trow1 * pitch + tcol1));
                } else {
                    const unsigned xc = reflect_safe(rx0 + c, I.w);
#pragma unroll
                    for (int q = 0; q < 9; ++q) {
                        const unsigned row = 2 * q + h;
                        sreg[row * W15_REGION_PITCH + c] = __ldg(I.p + (reflect_safe(ry0 + (int)row, I.h) * pitch + xc));
                    }
                    sreg[trow0 * W15_REGION_PITCH + tcol0] =
                        __ldg(I.p + (reflect_safe(ry0 + (int)trow0, I.h) * pitch + reflect_safe(rx0 + (int)tcol0, I.w)));
                    if (k1 < 36)
                        sreg[trow1 * W15_REGION_PITCH + tcol1] =
                            __ldg(I.p + (reflect_safe(ry0 + (int)trow1, I.h) * pitch + reflect_safe(rx0 + (int)tcol1, I.w)));
                }
            }
            __syncwarp();

            // ---- template: Scharr of the lane's 9 tile rows at columns c and c+1, streamed row by row ---
            W15Patch P;
            int a11 = 0, a12 = 0, a22 = 0;
            {
                const unsigned* rowp = reinterpret_cast<const unsigned*>(sreg + (8 * h) * W15_REGION_PITCH + (c & ~3));
                const int sh = (c & 3) * 8;
                const bool in_x0 = (ipx + c) >= 0 && (ipx + c) < I.w, in_x1 = (ipx + c + 1) >= 0 && (ipx + c + 1) < I.w;
                unsigned win[11];
                int hd0[11], hs0[11], hd1[11], hs1[11];
#pragma unroll
                for (int t = 0; t < 11; ++t) {
                    win[t] = __funnelshift_r(rowp[t * (W15_REGION_PITCH / 4)], rowp[t * (W15_REGION_PITCH / 4) + 1], sh);
                    hd0[t] = dp4a_us(win[t], 0x000100FF, 0);   // (-1, 0, +1, 0)
                    hs0[t] = dp4a_us(win[t], 0x00030A03, 0);   // ( 3,10,  3, 0)
                    hd1[t] = dp4a_us(win[t], 0x0100FF00, 0);   // ( 0,-1,  0,+1)
                    hs1[t] = dp4a_us(win[t], 0x030A0300, 0);   // ( 0, 3, 10, 3)
                }
                int gx0[9], gy0[9], gx1[9], gy1[9];
#pragma unroll
                for (int y = 0; y < 9; ++y) {
                    gx0[y] = 3 * (hd0[y] + hd0[y + 2]) + 10 * hd0[y + 1];
                    gx1[y] = 3 * (hd1[y] + hd1[y + 2]) + 10 * hd1[y + 1];
                    gy0[y] = hs0[y + 2] - hs0[y];
                    gy1[y] = hs1[y + 2] - hs1[y];
                    if (!interior) {   // derivative is constant 0 in the window padding outside the image
                        const int Y = ipy + 8 * h + y;
                        const bool in_y = Y >= 0 && Y < I.h;
                        if (!(in_y && in_x0)) { gx0[y] = 0; gy0[y] = 0; }
                        if (!(in_y && in_x1)) { gx1[y] = 0; gy1[y] = 0; }
                    }
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const bool active = (c < W15) && (8 * h + i < W15);
                    const unsigned r0 = win[i + 1], r1 = win[i + 2];
                    const int i00 = (r0 >> 8) & 0xff, i01 = (r0 >> 16) & 0xff, i10 = (r1 >> 8) & 0xff, i11 = (r1 >> 16) & 0xff;
                    const int ival = (i00 * w.w00 + i01 * w.w01 + i10 * w.w10 + i11 * w.w11 + (1 << 8)) >> 9;
                    int ix = (gx0[i] * w.w00 + gx1[i] * w.w01 + gx0[i + 1] * w.w10 + gx1[i + 1] * w.w11 + (1 << 13)) >> 14;
                    int iy = (gy0[i] * w.w00 + gy1[i] * w.w01 + gy0[i + 1] * w.w10 + gy1[i + 1] * w.w11 + (1 << 13)) >> 14;
                    if (!active) { ix = 0; iy = 0; }
                    P.I[i] = (1 << 8) - (ival << 9); P.gx[i] = ix; P.gy[i] = iy;
                    a11 += ix * ix; a12 += ix * iy; a22 += iy * iy;
                }
            }
            const float A11 = warp_sum_scaled(a11), A12 = warp_sum_scaled(a12), A22 = warp_sum_scaled(a22);
            float D = fsub(fmul(A11, A22), fmul(A12, A12));
            const float dA = fsub(A11, A22);
            const float disc = fadd(fmul(dA, dA), fmul(fmul(4.f, A12), A12));
            const float min_eig = __fdiv_rn(fsub(fadd(A22, A11), __fsqrt_rn(disc)), (float)(2 * W15 * W15));
            if (min_eig < A.min_eig || D < 1.1920928955078125e-07f) {
                if (level == 0) status = 0;
                continue;
            }
            D = __fdiv_rn(1.f, D);

            // ---- Newton iterations; at level 0 one extra trip through the same code evaluates err -----
            nx = fsub(nx, half); ny = fsub(ny, half);
            float pdx = 0.f, pdy = 0.f;
            bool final_eval = false;
            for (int j = 0;; ++j) {
                if (!final_eval && j >= A.max_count) {
                    if (level == 0 && status) final_eval = true;
                    else break;
                }
                const float qx = final_eval ? fsub(next_x, half) : nx, qy = final_eval ? fsub(next_y, half) : ny;
                const int inx = __float2int_rd(qx), iny = __float2int_rd(qy);
                // -15 <= inx < w  <=>  (unsigned)(inx + 15) < (unsigned)(w + 15)
                if ((unsigned)(inx + W15) >= (unsigned)(J.w + W15) || (unsigned)(iny + W15) >= (unsigned)(J.h + W15)) {
                    if (level == 0) status = 0;
                    break;
                }
                w = bilin_weights(fsub(qx, (float)inx), fsub(qy, (float)iny));
                int jv[9], diff[8];
                w15_gather(J, inx, iny, h, c, jv);
                w15_diff(jv, w, P, diff);
                if (final_eval) {
                    int e = 0;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const bool active = (c < W15) && (8 * h + i < W15);
                        e += active ? abs(diff[i]) : 0;
                    }
                    e = __reduce_add_sync(0xffffffffu, e);
                    err = __fdiv_rn((float)e, (float)(32 * W15 * W15));
                    break;
                }
                int sb1 = 0, sb2 = 0;
#pragma unroll
                for (int i = 0; i < 8; ++i) { sb1 += diff[i] * P.gx[i]; sb2 += diff[i] * P.gy[i]; }
                const float b1 = warp_sum_scaled(sb1), b2 = warp_sum_scaled(sb2);
                const float dx = fmul(fsub(fmul(A12, b2), fmul(A22, b1)), D);
                const float dy = fmul(fsub(fmul(A12, b1), fmul(A11, b2)), D);
                nx = fadd(nx, dx); ny = fadd(ny, dy);
                next_x = fadd(nx, half); next_y = fadd(ny, half);
                bool stop = fadd(fmul(dx, dx), fmul(dy, dy)) <= A.eps2;
                if (!stop && j > 0 && fabsf(fadd(dx, pdx)) < 0.01f && fabsf(fadd(dy, pdy)) < 0.01f) {
                    next_x = fsub(next_x, fmul(dx, 0.5f));
                    next_y = fsub(next_y, fmul(dy, 0.5f));
                    stop = true;
                }
                pdx = dx; pdy = dy;
                if (stop) {
                    if (level == 0 && status) final_eval = true;
                    else break;
                }
            }
        }

        if (pass == 0) {
            fx = next_x; fy = next_y; fst = status; ferr = err; st = status;
            if (!fst) break;   // the backward pass cannot change a failed track
        } else {
            bx = next_x; by = next_y;
            const float ddx = fsub(px0, bx), ddy = fsub(py0, by);
            const float fbe = __fsqrt_rn(fadd(fmul(ddx, ddx), fmul(ddy, ddy)));
            st = status && (fbe < A.fbt);
        }
    }
    if (lane == 0) {
        const long long o = (long long)pair * A.npts + pt;
        A.out[2 * o] = fx;
        A.out[2 * o + 1] = fy;
        A.status[o] = (uint8_t)st;
        A.err[o] = fst ? ferr : 0.f;
        if (A.back) { A.back[2 * o] = bx; A.back[2 * o + 1] = by; }
    }
}

// ---- helpers of the word-gathering 15x15 kernel (lk_w15h.cuh) --------------------------------------------------------
__device__ __forceinline__ int dp2a_lo(int w, unsigned px, int acc)
{
    int d;
    asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(w), "r"(px), "r"(acc));
    return d;
}
__device__ __forceinline__ int dp2a_hi(int w, unsigned px, int acc)
{
    int d;
    asm("dp2a.hi.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(w), "r"(px), "r"(acc));
    return d;
}
__device__ __forceinline__ unsigned ldg_u32(const uint8_t* p) { return __ldg(reinterpret_cast<const unsigned*>(p)); }

#include "lk_w15h.cuh"
#include "lk_w15s.cuh"

// =====================================================================================================
// Column-streaming path for mid-size windows (16 <= win_w <= COLS-1, win_h <= 63; the reference's
// 51x51 lk_fine, cv2's default 21x21).  One CTA of 128 threads per point; the template lives in shared
// memory (I pre-folded as 256 - (I << 9), Ix, Iy as int rows of 4-pixel groups).  The template is built with
// thread = (column slot c, row group g) walking DOWN its column (two bytes of the previous row stay in
// registers).  Search iterations use thread = (4-column group, row group): per row two aligned word loads
// re-aligned by funnel shifts, per pixel two DP2A (bilinear sample minus template) + two IMAD (b sums),
// template operands by 128-bit shared loads -- ~7 instructions per pixel instead of ~11.5 for the byte
// walk, which remains for odd pitches and windows hanging over the frame border.  Per-thread sums fit
// int32 (<= 8 rows x 4 pixels x 2^25), the CTA-wide sums are exact int64.  Same arithmetic as every other
// path => bit-identical results.
template <int COLS>
__global__ void __launch_bounds__(128)
lk_track_cols_kernel(const LkArgs A)
{
    constexpr int NT = 128, NG = NT / COLS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x;
    const int pair = blockIdx.y, pt = blockIdx.x;
    const int ww = A.win_w, wh = A.win_h, npx = ww * wh;
    const int tw = ww + 1, th = wh + 1;
    const int wp = (ww + 3) & ~3;                       // template row pitch: 4-pixel groups are 16-byte aligned
    const size_t bytes_D = ((size_t)tw * th * 4 + 15) & ~(size_t)15;
    const size_t bytes_T = (size_t)wp * wh * 4;
    short2* sD = reinterpret_cast<short2*>(smem_raw);
    int* sI = reinterpret_cast<int*>(smem_raw + bytes_D);                 // 256 - (I << 9)
    int* sGx = reinterpret_cast<int*>(smem_raw + bytes_D + bytes_T);
    int* sGy = reinterpret_cast<int*>(smem_raw + bytes_D + 2 * bytes_T);
    long long* red = reinterpret_cast<long long*>(smem_raw + bytes_D + 3 * bytes_T);
    for (int i = tid; i < wp * wh; i += NT) { sGx[i] = 0; sGy[i] = 0; sI[i] = 0; }   // padding columns stay zero for good
    int parity = 0;

    const int c = tid % COLS, g = tid / COLS;
    const int rows_per = (wh + NG - 1) / NG;
    const int r0 = g * rows_per, r1 = min(wh, r0 + rows_per);
    const bool col_active = c < ww && r0 < r1;
    // search iterations: thread = (4-column group cgw, row group rgw), aligned word gathers + DP2A (pitch % 4 == 0)
    const int ncg = wp >> 2, nrg = NT / ncg;
    const int cgw = tid % ncg, rgw = tid / ncg;
    const int rows_w = (wh + nrg - 1) / nrg;
    const int wr0 = rgw * rows_w, wr1 = min(wh, wr0 + rows_w);
    const bool w_active = rgw < nrg && wr0 < wr1;
    bool words = !A.force_bytes && (A.prev_pitch % 4 == 0) && (A.next_pitch % 4 == 0);
    for (int l = 1; l <= A.lv.max_level; ++l) words = words && (A.lv.pitch[l] % 4 == 0);

    const float half_x = fmul((float)(ww - 1), 0.5f), half_y = fmul((float)(wh - 1), 0.5f);
    const float FLT_SCALE = 1.f / 1048576.f;

    const uint8_t* P0 = A.prev0 + (long long)pair * A.prev_stride;
    const uint8_t* Pp = A.prev_pyr ? A.prev_pyr + (long long)pair * A.prev_pyr_stride : nullptr;
    const uint8_t* N0 = A.next0 + (long long)pair * A.next_stride;
    const uint8_t* Np = A.next_pyr ? A.next_pyr + (long long)pair * A.next_pyr_stride : nullptr;
    const float* pin = A.pts + (long long)pair * A.pts_stride + 2ll * pt;
    const float px0 = __ldg(pin), py0 = __ldg(pin + 1);

    float fx = 0.f, fy = 0.f, ferr = 0.f, bx = 0.f, by = 0.f;
    int fst = 0, st = 0;
    const int npass = A.fbt >= 0.f ? 2 : 1;

    for (int pass = 0; pass < npass; ++pass) {
        const uint8_t* I0 = pass ? N0 : P0;
        const uint8_t* Ipyr = pass ? Np : Pp;
        const uint8_t* J0 = pass ? P0 : N0;
        const uint8_t* Jpyr = pass ? Pp : Np;
        const int I0_pitch = pass ? A.next_pitch : A.prev_pitch, J0_pitch = pass ? A.prev_pitch : A.next_pitch;
        const float px = pass ? fx : px0, py = pass ? fy : py0;
        int status = 1;
        float err = 0.f;
        float next_x = 0.f, next_y = 0.f;

        for (int level = A.lv.max_level; level >= 0; --level) {
            Img I, J;
            I.w = J.w = A.lv.w[level];
            I.h = J.h = A.lv.h[level];
            if (level == 0) { I.p = I0; I.pitch = I0_pitch; J.p = J0; J.pitch = J0_pitch; }
            else { I.p = Ipyr + A.lv.off[level]; J.p = Jpyr + A.lv.off[level]; I.pitch = J.pitch = A.lv.pitch[level]; }

            const float scale = 1.f / (float)(1 << level);
            float prev_x = fmul(px, scale), prev_y = fmul(py, scale);
            float nx, ny;
            if (level == A.lv.max_level) { nx = prev_x; ny = prev_y; }
            else { nx = fmul(next_x, 2.f); ny = fmul(next_y, 2.f); }
            next_x = nx; next_y = ny;
            prev_x = fsub(prev_x, half_x); prev_y = fsub(prev_y, half_y);
            const int ipx = __float2int_rd(prev_x), ipy = __float2int_rd(prev_y);
            if (ipx < -ww || ipx >= I.w || ipy < -wh || ipy >= I.h) {
                if (level == 0) { status = 0; err = 0.f; }
                continue;
            }
            Weights w = bilin_weights(fsub(prev_x, (float)ipx), fsub(prev_y, (float)ipy));

            __syncthreads();
            // (1) Scharr tile over the (ww+1) x (wh+1) footprint, one column per thread, streamed down the rows
            for (int tx = tid; tx < tw; tx += NT) {
                const int X = ipx + tx;
                const bool in_x = X >= 0 && X < I.w;
                const unsigned xm = reflect_safe(X - 1, I.w), xc = reflect_safe(X, I.w), xp = reflect_safe(X + 1, I.w);
                // rows Y-1, Y, Y+1 sliding
                unsigned ry = reflect_safe(ipy - 1, I.h) * (unsigned)I.pitch;
                int a0 = __ldg(I.p + ry + xm), a1 = __ldg(I.p + ry + xc), a2 = __ldg(I.p + ry + xp);
                ry = reflect_safe(ipy, I.h) * (unsigned)I.pitch;
                int b0 = __ldg(I.p + ry + xm), b1 = __ldg(I.p + ry + xc), b2 = __ldg(I.p + ry + xp);
                for (int ty = 0; ty < th; ++ty) {
                    const int Y = ipy + ty;
                    ry = reflect_safe(Y + 1, I.h) * (unsigned)I.pitch;
                    const int c0 = __ldg(I.p + ry + xm), c1 = __ldg(I.p + ry + xc), c2 = __ldg(I.p + ry + xp);
                    int gx = 0, gy = 0;
                    if (in_x && Y >= 0 && Y < I.h) {
                        gx = 3 * (a2 + c2) + 10 * b2 - 3 * (a0 + c0) - 10 * b0;
                        gy = 3 * ((c0 - a0) + (c2 - a2)) + 10 * (c1 - a1);
                    }
                    sD[ty * tw + tx] = make_short2((short)gx, (short)gy);
                    a0 = b0; a1 = b1; a2 = b2; b0 = c0; b1 = c1; b2 = c2;
                }
            }
            __syncthreads();
            // (2) template patch + gradient matrix, column-streamed
            long long acc[3] = {0, 0, 0};
            if (col_active) {
                const unsigned x0 = reflect_safe(ipx + c, I.w), x1 = reflect_safe(ipx + c + 1, I.w);
                unsigned ry = reflect_safe(ipy + r0, I.h) * (unsigned)I.pitch;
                int i00 = __ldg(I.p + ry + x0), i01 = __ldg(I.p + ry + x1);
                int a11 = 0, a12 = 0, a22 = 0;
                for (int y = r0; y < r1; ++y) {
                    ry = reflect_safe(ipy + y + 1, I.h) * (unsigned)I.pitch;
                    const int i10 = __ldg(I.p + ry + x0), i11 = __ldg(I.p + ry + x1);
                    const int ival = (i00 * w.w00 + i01 * w.w01 + i10 * w.w10 + i11 * w.w11 + (1 << 8)) >> 9;
                    const short2 d00 = sD[y * tw + c], d01 = sD[y * tw + c + 1], d10 = sD[(y + 1) * tw + c], d11 = sD[(y + 1) * tw + c + 1];
                    const int ix = (d00.x * w.w00 + d01.x * w.w01 + d10.x * w.w10 + d11.x * w.w11 + (1 << 13)) >> 14;
                    const int iy = (d00.y * w.w00 + d01.y * w.w01 + d10.y * w.w10 + d11.y * w.w11 + (1 << 13)) >> 14;
                    sI[y * wp + c] = (1 << 8) - (ival << 9);
                    sGx[y * wp + c] = ix;
                    sGy[y * wp + c] = iy;
                    a11 += ix * ix; a12 += ix * iy; a22 += iy * iy;   // <= 32 rows x 2^24: fits int32
                    i00 = i10; i01 = i11;
                }
                acc[0] = a11; acc[1] = a12; acc[2] = a22;
            }
            group_sum<NT, 3>(acc, red, parity, tid);   // contains the barrier that publishes sI / sG
            const float A11 = fmul(__ll2float_rn(acc[0]), FLT_SCALE), A12 = fmul(__ll2float_rn(acc[1]), FLT_SCALE),
                        A22 = fmul(__ll2float_rn(acc[2]), FLT_SCALE);
            float D = fsub(fmul(A11, A22), fmul(A12, A12));
            const float dA = fsub(A11, A22);
            const float disc = fadd(fmul(dA, dA), fmul(fmul(4.f, A12), A12));
            const float min_eig = __fdiv_rn(fsub(fadd(A22, A11), __fsqrt_rn(disc)), (float)(2 * ww * wh));
            if (min_eig < A.min_eig || D < 1.1920928955078125e-07f) {
                if (level == 0) status = 0;
                continue;
            }
            D = __fdiv_rn(1.f, D);

            nx = fsub(nx, half_x); ny = fsub(ny, half_y);
            float pdx = 0.f, pdy = 0.f;
            bool final_eval = false;
            for (int j = 0;; ++j) {
                if (!final_eval && j >= A.max_count) {
                    if (level == 0 && status) final_eval = true;
                    else break;
                }
                const float qx = final_eval ? fsub(next_x, half_x) : nx, qy = final_eval ? fsub(next_y, half_y) : ny;
                const int inx = __float2int_rd(qx), iny = __float2int_rd(qy);
                if (inx < -ww || inx >= J.w || iny < -wh || iny >= J.h) {
                    if (level == 0) status = 0;
                    break;
                }
                w = bilin_weights(fsub(qx, (float)inx), fsub(qy, (float)iny));
                int sb1 = 0, sb2 = 0, se = 0;
                const bool inside = inx >= 0 && iny >= 0 && inx + ww < J.w && iny + wh < J.h;
                const unsigned pitch = (unsigned)J.pitch;
                if (words && inside) {
                    if (w_active) {
                        const int W0 = (int)__byte_perm((unsigned)w.w00, (unsigned)w.w01, 0x5410);
                        const int W1 = (int)__byte_perm((unsigned)w.w10, (unsigned)w.w11, 0x5410);
                        const unsigned off = (unsigned)(iny + wr0) * pitch + (unsigned)(inx + 4 * cgw);
                        const unsigned mis = ((unsigned)(size_t)J.p + off) & 3u, sh = mis * 8u;
                        const uint8_t* r = J.p + (int)(off - mis);
                        // the second word starts at footprint column 4cgw + 4 - mis: beyond column ww nothing active needs it
                        const bool skip_hi = 4 * cgw + 4 - (int)mis > ww;
                        // all rows of the thread are requested before the first is consumed (this kernel's shared-memory
                        // footprint leaves no L1: every gather is an L2 round trip, exposed once instead of once per row)
                        constexpr int MAXR = 8;                      // rows_w <= ceil(63 / 8)
                        unsigned lo[MAXR + 1], hi[MAXR + 1];
#pragma unroll
                        for (int q = 0; q <= MAXR; ++q) {
                            lo[q] = 0; hi[q] = 0;
                            if (wr0 + q <= wr1) {
                                lo[q] = ldg_u32(r + q * pitch);
                                hi[q] = lo[q];
                                if (!skip_hi) hi[q] = ldg_u32(r + q * pitch + 4);
                            }
                        }
                        const int o0 = wr0 * wp + 4 * cgw;
                        const int4* pI = reinterpret_cast<const int4*>(sI + o0);
                        const int4* pX = reinterpret_cast<const int4*>(sGx + o0);
                        const int4* pY = reinterpret_cast<const int4*>(sGy + o0);
                        const int nact = min(4, ww - 4 * cgw);      // active pixels of this group (the padding has Ix = Iy = 0)
                        unsigned t0 = __funnelshift_r(lo[0], hi[0], sh), t1 = __funnelshift_rc(lo[0], hi[0], sh + 8u);
#pragma unroll
                        for (int q = 0; q < MAXR; ++q) {
                            if (wr0 + q < wr1) {
                                const unsigned b0 = __funnelshift_r(lo[q + 1], hi[q + 1], sh), b1 = __funnelshift_rc(lo[q + 1], hi[q + 1], sh + 8u);
                                const int4 ti = pI[q * (wp >> 2)];
                                const int d0 = dp2a_lo(W1, b0, dp2a_lo(W0, t0, ti.x)) >> 9;
                                const int d1 = dp2a_lo(W1, b1, dp2a_lo(W0, t1, ti.y)) >> 9;
                                const int d2 = dp2a_hi(W1, b0, dp2a_hi(W0, t0, ti.z)) >> 9;
                                const int d3 = dp2a_hi(W1, b1, dp2a_hi(W0, t1, ti.w)) >> 9;
                                if (final_eval) {
                                    se += abs(d0) + (nact > 1 ? abs(d1) : 0) + (nact > 2 ? abs(d2) : 0) + (nact > 3 ? abs(d3) : 0);
                                } else {
                                    const int4 gx = pX[q * (wp >> 2)], gy = pY[q * (wp >> 2)];
                                    sb1 += d0 * gx.x + d1 * gx.y + d2 * gx.z + d3 * gx.w;
                                    sb2 += d0 * gy.x + d1 * gy.y + d2 * gy.z + d3 * gy.w;
                                }
                                t0 = b0; t1 = b1;
                            }
                        }
                    }
                } else if (col_active) {
                    const int* pI = sI + r0 * wp + c;
                    const int* pX = sGx + r0 * wp + c;
                    const int* pY = sGy + r0 * wp + c;
                    if (inside) {
                        const uint8_t* p = J.p + ((unsigned)(iny + r0) * pitch + (unsigned)(inx + c));
                        int a = __ldg(p), b = __ldg(p + 1);
#pragma unroll 4
                        for (int y = r0; y < r1; ++y) {
                            p += pitch;
                            const int an = __ldg(p), bn = __ldg(p + 1);
                            const int diff = (a * w.w00 + b * w.w01 + an * w.w10 + bn * w.w11 + *pI) >> 9;
                            sb1 += diff * *pX; sb2 += diff * *pY; se += abs(diff);
                            a = an; b = bn; pI += wp; pX += wp; pY += wp;
                        }
                    } else {
                        const unsigned x0 = reflect_safe(inx + c, J.w), x1 = reflect_safe(inx + c + 1, J.w);
                        unsigned ry = reflect_safe(iny + r0, J.h) * pitch;
                        int a = __ldg(J.p + ry + x0), b = __ldg(J.p + ry + x1);
                        for (int y = r0; y < r1; ++y) {
                            ry = reflect_safe(iny + y + 1, J.h) * pitch;
                            const int an = __ldg(J.p + ry + x0), bn = __ldg(J.p + ry + x1);
                            const int diff = (a * w.w00 + b * w.w01 + an * w.w10 + bn * w.w11 + *pI) >> 9;
                            sb1 += diff * *pX; sb2 += diff * *pY; se += abs(diff);
                            a = an; b = bn; pI += wp; pX += wp; pY += wp;
                        }
                    }
                }
                if (final_eval) {
                    long long e[1] = {se};
                    group_sum<NT, 1>(e, red, parity, tid);
                    err = __fdiv_rn(__ll2float_rn(e[0]), (float)(32 * ww * wh));
                    break;
                }
                long long bsum[2] = {sb1, sb2};
                group_sum<NT, 2>(bsum, red, parity, tid);
                const float b1 = fmul(__ll2float_rn(bsum[0]), FLT_SCALE), b2 = fmul(__ll2float_rn(bsum[1]), FLT_SCALE);
                const float dx = fmul(fsub(fmul(A12, b2), fmul(A22, b1)), D);
                const float dy = fmul(fsub(fmul(A12, b1), fmul(A11, b2)), D);
                nx = fadd(nx, dx); ny = fadd(ny, dy);
                next_x = fadd(nx, half_x); next_y = fadd(ny, half_y);
                bool stop = fadd(fmul(dx, dx), fmul(dy, dy)) <= A.eps2;
                if (!stop && j > 0 && fabsf(fadd(dx, pdx)) < 0.01f && fabsf(fadd(dy, pdy)) < 0.01f) {
                    next_x = fsub(next_x, fmul(dx, 0.5f));
                    next_y = fsub(next_y, fmul(dy, 0.5f));
                    stop = true;
                }
                pdx = dx; pdy = dy;
                if (stop) {
                    if (level == 0 && status) final_eval = true;
                    else break;
                }
            }
        }
        if (pass == 0) {
            fx = next_x; fy = next_y; fst = status; ferr = err; st = status;
            if (!fst) break;
        } else {
            bx = next_x; by = next_y;
            const float ddx = fsub(px0, bx), ddy = fsub(py0, by);
            const float fbe = __fsqrt_rn(fadd(fmul(ddx, ddx), fmul(ddy, ddy)));
            st = status && (fbe < A.fbt);
        }
    }
    if (tid == 0) {
        const long long o = (long long)pair * A.npts + pt;
        A.out[2 * o] = fx;
        A.out[2 * o + 1] = fy;
        A.status[o] = (uint8_t)st;
        A.err[o] = fst ? ferr : 0.f;
        if (A.back) { A.back[2 * o] = bx; A.back[2 * o + 1] = by; }
    }
}

size_t cols_smem_bytes(int win_w, int win_h)
{
    const size_t ntile = (size_t)(win_w + 1) * (win_h + 1), wp = (size_t)((win_w + 3) & ~3);
    return ((ntile * 4 + 15) & ~(size_t)15) + 3 * wp * win_h * 4 + (size_t)2 * 4 * 3 * sizeof(long long);
}

size_t group_smem_bytes(int win_w, int win_h, int nt)
{
    const size_t npx = (size_t)win_w * win_h, ntile = (size_t)(win_w + 1) * (win_h + 1);
    size_t b = ((ntile * 4 + 15) & ~(size_t)15) + ((npx * 4 + 15) & ~(size_t)15) + ((npx * 2 + 15) & ~(size_t)15);
    if (nt > 32) b += (size_t)2 * (nt / 32) * 3 * sizeof(long long);
    return b;
}

}  // namespace

VEL_API int vel_lk_track(const uint8_t* prev_frames, int64_t prev_frame_stride, int32_t prev_pitch, const uint8_t* prev_pyr,
                         int64_t prev_pyr_stride, const uint8_t* next_frames, int64_t next_frame_stride, int32_t next_pitch,
                         const uint8_t* next_pyr, int64_t next_pyr_stride, const vel_pyr_layout* layout, int32_t npairs,
                         const float* prev_pts, int64_t pts_stride, int32_t npts, const vel_lk_params* params, float* next_pts,
                         uint8_t* status, float* err, float* back_pts, vel_stream_t stream)
{
    VEL_CHECK_ARG(npts >= 0, "vel_lk_track: npts < 0");
    if (npts == 0) return VEL_OK;   // empty point set: nothing to do (buffers may be NULL)
    VEL_CHECK_ARG(prev_frames && next_frames && layout && prev_pts && params && next_pts && status && err,
                  "vel_lk_track: NULL argument");
    VEL_CHECK_ARG(npairs > 0 && npairs <= 65535, "vel_lk_track: npairs %d outside [1,65535]", npairs);
    const int ww = params->win_w, wh = params->win_h;
    VEL_CHECK_ARG(ww >= 3 && wh >= 3 && ww <= 127 && wh <= 127, "vel_lk_track: window %dx%d outside [3,127]", ww, wh);
    VEL_CHECK_ARG(layout->max_level >= 0 && layout->max_level < VEL_MAX_LEVELS, "vel_lk_track: bad layout");
    VEL_CHECK_ARG(layout->width[0] > ww && layout->height[0] > wh,
                  "vel_lk_track: image %dx%d must be larger than the window %dx%d", layout->width[0], layout->height[0], ww, wh);
    VEL_CHECK_ARG(layout->max_level == 0 || (prev_pyr && next_pyr), "vel_lk_track: pyramid buffers required for max_level > 0");
    VEL_CHECK_ARG(prev_pitch >= layout->width[0] && next_pitch >= layout->width[0], "vel_lk_track: pitch < width");

    LkArgs A;
    A.prev0 = prev_frames; A.prev_stride = prev_frame_stride; A.prev_pitch = prev_pitch;
    A.prev_pyr = prev_pyr; A.prev_pyr_stride = prev_pyr_stride;
    A.next0 = next_frames; A.next_stride = next_frame_stride; A.next_pitch = next_pitch;
    A.next_pyr = next_pyr; A.next_pyr_stride = next_pyr_stride;
    A.lv.max_level = layout->max_level;
    for (int l = 0; l < VEL_MAX_LEVELS; ++l) {
        A.lv.w[l] = layout->width[l]; A.lv.h[l] = layout->height[l]; A.lv.pitch[l] = layout->pitch[l]; A.lv.off[l] = layout->offset[l];
    }
    A.pts = prev_pts; A.pts_stride = pts_stride; A.npts = npts;
    A.out = next_pts; A.status = status; A.err = err; A.back = back_pts;
    A.seq_pairs = 0; A.alive = nullptr;
    A.win_w = ww; A.win_h = wh;
    int mc = params->max_count; mc = mc < 0 ? 0 : (mc > 100 ? 100 : mc);
    double eps = params->eps; eps = eps < 0. ? 0. : (eps > 10. ? 10. : eps);
    A.max_count = mc;
    A.eps2 = (float)(eps * eps);
    A.min_eig = params->min_eig_threshold;
    A.fbt = params->fb_threshold;

    const char* force = getenv("VEL_LK_W15");
    A.force_bytes = (force && strcmp(force, "bytes") == 0) ? 1 : 0;

    cudaStream_t st = (cudaStream_t)stream;
    if (ww == W15 && wh == W15) {
        // word-gathering kernel (two points per warp): needs every row pitch to be a multiple of 4 bytes;
        // VEL_LK_W15=bytes forces the byte-gather kernel (kept for odd pitches and as an independent cross-check)
        bool words = (prev_pitch % 4 == 0) && (next_pitch % 4 == 0);
        for (int l = 1; l <= layout->max_level; ++l) words = words && (layout->pitch[l] % 4 == 0);
        if (A.force_bytes) words = false;
        if (words) {
            dim3 grid((npts + 2 * WH_WARPS - 1) / (2 * WH_WARPS), npairs);
            lk_track_w15h_kernel<false><<<grid, 32 * WH_WARPS, 0, st>>>(A);
        } else {
            dim3 grid((npts + W15_WARPS - 1) / W15_WARPS, npairs);
            lk_track_w15_kernel<<<grid, 32 * W15_WARPS, 0, st>>>(A);
        }
    } else if (ww >= 16 && ww <= 63 && wh >= 8 && wh <= 63) {
        const size_t smem = cols_smem_bytes(ww, wh);
        dim3 grid(npts, npairs);
        if (ww <= 31) {
            auto kern = lk_track_cols_kernel<32>;
            if (smem > 48 * 1024) VEL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            kern<<<grid, 128, smem, st>>>(A);
        } else {
            auto kern = lk_track_cols_kernel<64>;
            if (smem > 48 * 1024) VEL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            kern<<<grid, 128, smem, st>>>(A);
        }
    } else if (ww * wh <= 1024) {
        constexpr int NT = 32, GROUPS = 8;
        const size_t smem = GROUPS * group_smem_bytes(ww, wh, NT);
        auto kern = lk_track_kernel<NT, GROUPS>;
        if (smem > 48 * 1024) VEL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        dim3 grid((npts + GROUPS - 1) / GROUPS, npairs);
        kern<<<grid, NT * GROUPS, smem, st>>>(A);
    } else {
        constexpr int NT = 128, GROUPS = 1;
        const size_t smem = group_smem_bytes(ww, wh, NT);
        VEL_CHECK_ARG(smem <= 227 * 1024, "vel_lk_track: window %dx%d needs %zu B of shared memory", ww, wh, smem);
        auto kern = lk_track_kernel<NT, GROUPS>;
        if (smem > 48 * 1024) VEL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        dim3 grid(npts, npairs);
        kern<<<grid, NT, smem, st>>>(A);
    }
    VEL_LAUNCH_CHECK("lk_track_kernel");
    return VEL_OK;
}

// Internal entry of vel_klt_sequence (sequence.cu): the whole frame run in ONE launch of lk_track_w15h_kernel<true> when the
// configuration is the one that kernel serves (15x15 window, every row pitch a multiple of 4 bytes).  Returns 1 when launched,
// 0 when the caller has to take the per-pair path, < 0 on error.
namespace {
typedef CUresult (*SeqEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                     const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                     CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// one 3-D byte map (x, row, frame) per pyramid level, box 32 x 32 x 1
bool seq_make_maps(SeqTma* tm, const uint8_t* frames, int64_t frame_stride, int pitch, const uint8_t* pyr, int64_t pyr_stride,
                   const vel_pyr_layout* L, int nframes)
{
    static SeqEncodeTiledFn enc = nullptr;
    if (!enc) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess || qres != cudaDriverEntryPointSuccess)
            return false;
        enc = (SeqEncodeTiledFn)p;
    }
    for (int l = 0; l <= L->max_level; ++l) {
        const uint8_t* base = l == 0 ? frames : pyr + L->offset[l];
        const long long fs = l == 0 ? frame_stride : pyr_stride;
        const int pt = l == 0 ? pitch : L->pitch[l];
        if ((((size_t)base) & 15) != 0 || (fs & 15) != 0 || (pt & 15) != 0 || L->width[l] < TMA_BOX || L->height[l] < TMA_BOX) return false;
        cuuint64_t dims[3] = {(cuuint64_t)L->width[l], (cuuint64_t)L->height[l], (cuuint64_t)nframes};
        cuuint64_t strides[2] = {(cuuint64_t)pt, (cuuint64_t)fs};
        cuuint32_t box[3] = {TMA_BOXW, TMA_BOX, 1};
        cuuint32_t estr[3] = {1, 1, 1};
        if (enc(&tm->map[l], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return false;
    }
    return true;
}
}  // namespace

int vel_lk_sequence_w15h(const uint8_t* frames, int64_t frame_stride, int32_t pitch, const uint8_t* pyr, int64_t pyr_stride,
                         const vel_pyr_layout* layout, int32_t nframes, int32_t npts, const vel_lk_params* params, float* tracks,
                         uint8_t* alive, float* err, vel_stream_t stream)
{
    if (params->win_w != W15 || params->win_h != W15 || pitch % 4 != 0 || frame_stride % 4 != 0) return 0;
    for (int l = 1; l <= layout->max_level; ++l)
        if (layout->pitch[l] % 4 != 0) return 0;
    if (layout->max_level > 0 && (!pyr || pyr_stride % 4 != 0)) return 0;
    const char* force = getenv("VEL_LK_W15");
    if (force && strcmp(force, "bytes") == 0) return 0;
    const char* seq = getenv("VEL_LK_SEQ");
    if (seq && strcmp(seq, "pairs") == 0) return 0;      // cross-check switch: one launch per pair
    if (!(layout->width[0] > W15 && layout->height[0] > W15 && pitch >= layout->width[0])) return 0;   // vel_lk_track reports it
    LkArgs A;
    A.prev0 = frames; A.prev_stride = frame_stride; A.prev_pitch = pitch;
    A.prev_pyr = pyr; A.prev_pyr_stride = pyr_stride;
    A.next0 = frames + frame_stride; A.next_stride = frame_stride; A.next_pitch = pitch;
    A.next_pyr = pyr ? pyr + pyr_stride : nullptr; A.next_pyr_stride = pyr_stride;
    A.lv.max_level = layout->max_level;
    for (int l = 0; l < VEL_MAX_LEVELS; ++l) {
        A.lv.w[l] = layout->width[l]; A.lv.h[l] = layout->height[l]; A.lv.pitch[l] = layout->pitch[l]; A.lv.off[l] = layout->offset[l];
    }
    A.pts = tracks; A.pts_stride = 0; A.npts = npts;
    A.out = tracks + 2ll * npts; A.status = nullptr; A.err = err; A.back = nullptr;
    A.win_w = W15; A.win_h = W15;
    int mc = params->max_count; mc = mc < 0 ? 0 : (mc > 100 ? 100 : mc);
    double eps = params->eps; eps = eps < 0. ? 0. : (eps > 10. ? 10. : eps);
    A.max_count = mc;
    A.eps2 = (float)(eps * eps);
    A.min_eig = params->min_eig_threshold;
    A.fbt = params->fb_threshold;
    A.force_bytes = 0;
    A.seq_pairs = nframes - 1; A.alive = alive;
    if (seq && strcmp(seq, "twice") == 0) {               // cross-check switch: frame loop inside the batch kernel, template built per pass
        constexpr int SEQ_WARPS = 2;
        dim3 grid((npts + 2 * SEQ_WARPS - 1) / (2 * SEQ_WARPS), 1);
        lk_track_w15h_kernel<true><<<grid, 32 * SEQ_WARPS, 0, (cudaStream_t)stream>>>(A);
        VEL_LAUNCH_CHECK("lk_track_w15h_kernel<seq>");
        return 1;
    }
    dim3 grid((npts + 2 * WS_WARPS - 1) / (2 * WS_WARPS), 1);
    SeqTma tm;
    memset(&tm, 0, sizeof(tm));
    // VEL_LK_SEQ=tma: the search neighbourhoods staged by TMA (the A/B of that design; results identical, see profiles/README.md)
    if (seq && strcmp(seq, "tma") == 0 && layout->max_level <= 2 && seq_make_maps(&tm, frames, frame_stride, pitch, pyr, pyr_stride, layout, nframes)) {
        lk_seq_w15h_kernel<true><<<grid, 32 * WS_WARPS, 0, (cudaStream_t)stream>>>(A, tm);
    } else {
        lk_seq_w15h_kernel<false><<<grid, 32 * WS_WARPS, 0, (cudaStream_t)stream>>>(A, tm);
    }
    VEL_LAUNCH_CHECK("lk_seq_w15h_kernel");
    return 1;
}
