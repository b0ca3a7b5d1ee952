// K1: image pyramid (cv2.pyrDown chain) and the 1/4 nearest-neighbour decimation.
//
// Reference call sites: the pyramid cv2.calcOpticalFlowPyrLK builds internally (utils/KLT.py:45,48)
// and cv2.resize(.., fx=fy=1/4, INTER_NEAREST) (utils/KLT.py:111,113).
//
// pyrDown semantics for CV_8UC1 (verified bit-exact against cv2 4.13 by the oracle tests):
//   dst(x,y) = (sum_{i,j} k[i] k[j] src(2x-2+i, 2y-2+j) + 128) >> 8, k = [1 4 6 4 1],
//   BORDER_REFLECT_101, dst size ((w+1)/2, (h+1)/2).
//
// Streaming, HBM-bound: algorithmic bytes per level = src bytes read once + dst bytes written once.
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"

namespace {

// ---- vectorised interior kernel -------------------------------------------------------------
// One CTA produces a TILE_W x TILE_H block of the destination.  The (2*TILE_W+4+pad) x (2*TILE_H+4)
// source footprint is staged in shared memory with 16-byte loads, filtered horizontally into
// uint16 rows (max 16*255 fits), then vertically; four output pixels are packed per 32-bit store.
constexpr int TILE_W = 128;   // destination columns per CTA (multiple of 8)
constexpr int TILE_H = 16;    // destination rows per CTA
constexpr int SRC_ROWS = 2 * TILE_H + 3;          // rows 2*y0-2 .. 2*(y0+TILE_H-1)+2
constexpr int SRC_COLS = 2 * TILE_W + 32;         // cols 2*x0-16 .. 2*x0+2*TILE_W+15 (16-byte aligned both ends)
constexpr int PYR_THREADS = 256;

__device__ __forceinline__ unsigned dp4a_uu(unsigned a, unsigned b, unsigned c)
{
    unsigned d;
    asm("dp4a.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

__global__ void __launch_bounds__(PYR_THREADS)
pyrdown_tiled_kernel(const uint8_t* __restrict__ src_base, long long src_stride, int sw, int sh, int spitch,
                     uint8_t* __restrict__ dst_base, long long dst_stride, int dw, int dh, int dpitch)
{
    __shared__ __align__(16) uint8_t s_src[SRC_ROWS][SRC_COLS];
    __shared__ __align__(16) uint16_t s_h[SRC_ROWS][TILE_W];

    const uint8_t* __restrict__ src = src_base + (long long)blockIdx.z * src_stride;
    uint8_t* __restrict__ dst = dst_base + (long long)blockIdx.z * dst_stride;
    const int ox0 = blockIdx.x * TILE_W, oy0 = blockIdx.y * TILE_H;
    const int sx0 = 2 * ox0 - 16, sy0 = 2 * oy0 - 2;
    const int tid = threadIdx.x;

    // stage the source footprint: 16-byte chunks that lie inside the row are vector loads (rows are
    // reflected as a whole); only the bytes of edge-straddling chunks that a valid output really taps
    // go byte by byte, and chunks/rows no valid output needs are skipped altogether
    const bool aligned = ((((uintptr_t)src) | (uintptr_t)spitch) & 15) == 0;
    constexpr int VEC_PER_ROW = SRC_COLS / 16;
    const int need_x0 = 2 * ox0 - 2, need_x1 = 2 * (min(ox0 + TILE_W, dw) - 1) + 2;   // inclusive source columns
    const int need_rows = 2 * (min(oy0 + TILE_H, dh) - 1 - oy0) + 5;                   // rows 0 .. need_rows-1 of the tile
    for (int i = tid; i < need_rows * VEC_PER_ROW; i += PYR_THREADS) {
        const int r = i / VEC_PER_ROW, v = i - r * VEC_PER_ROW;
        const int xs = sx0 + 16 * v;
        if (xs + 15 < need_x0 || xs > need_x1) continue;
        int yy = sy0 + r;
        yy = yy < 0 ? -yy : (yy >= sh ? 2 * sh - 2 - yy : yy);   // |overshoot| <= 2 here
        yy = max(0, min(yy, sh - 1));
        const uint8_t* row = src + (long long)yy * spitch;
        if (aligned && xs >= 0 && xs + 16 <= sw) {
            reinterpret_cast<uint4*>(&s_src[r][0])[v] = __ldg(reinterpret_cast<const uint4*>(row + xs));
        } else {
            const int b0 = max(0, need_x0 - xs), b1 = min(15, need_x1 - xs);
            for (int b = b0; b <= b1; ++b) s_src[r][16 * v + b] = row[reflect101(xs + b, sw)];
        }
    }
    __syncthreads();

    // horizontal pass: item = 4 adjacent outputs of one source row; output m is centred on byte 16+2m
    // of the 16-byte group starting at 8j+8:  h_m = [1 4 6 4] . bytes(14+2m .. 17+2m) + byte(18+2m)
    for (int i = tid; i < SRC_ROWS * (TILE_W / 4); i += PYR_THREADS) {
        const int r = i / (TILE_W / 4), j = i - r * (TILE_W / 4);
        const uint2 a = *reinterpret_cast<const uint2*>(&s_src[r][8 * j + 8]);    // bytes  8..15
        const uint2 b = *reinterpret_cast<const uint2*>(&s_src[r][8 * j + 16]);   // bytes 16..23
        const unsigned c24 = s_src[r][8 * j + 24];
        const unsigned K = 0x04060401u;  // weights for bytes k, k+1, k+2, k+3 (the fifth tap has weight 1)
        const unsigned w0 = __funnelshift_r(a.y, b.x, 16);   // bytes 14..17
        const unsigned w1 = b.x;                             // bytes 16..19
        const unsigned w2 = __funnelshift_r(b.x, b.y, 16);   // bytes 18..21
        const unsigned w3 = b.y;                             // bytes 20..23
        const unsigned h0 = dp4a_uu(w0, K, (b.x >> 16) & 0xff);   // + byte 18
        const unsigned h1 = dp4a_uu(w1, K, b.y & 0xff);           // + byte 20
        const unsigned h2 = dp4a_uu(w2, K, (b.y >> 16) & 0xff);   // + byte 22
        const unsigned h3 = dp4a_uu(w3, K, c24);                  // + byte 24
        *reinterpret_cast<uint2*>(&s_h[r][4 * j]) = make_uint2(h0 | (h1 << 16), h2 | (h3 << 16));
    }
    __syncthreads();

    // vertical pass on packed 16-bit pairs: every partial sum is <= 16 * 4080 = 65280, so the two
    // halves of a register never carry into each other
    for (int i = tid; i < TILE_H * (TILE_W / 4); i += PYR_THREADS) {
        const int y = i / (TILE_W / 4), j = i - y * (TILE_W / 4);
        const int oy = oy0 + y, ox = ox0 + 4 * j;
        if (oy >= dh || ox >= dw) continue;
        const uint2 r0 = *reinterpret_cast<const uint2*>(&s_h[2 * y][4 * j]);
        const uint2 r1 = *reinterpret_cast<const uint2*>(&s_h[2 * y + 1][4 * j]);
        const uint2 r2 = *reinterpret_cast<const uint2*>(&s_h[2 * y + 2][4 * j]);
        const uint2 r3 = *reinterpret_cast<const uint2*>(&s_h[2 * y + 3][4 * j]);
        const uint2 r4 = *reinterpret_cast<const uint2*>(&s_h[2 * y + 4][4 * j]);
        const unsigned ax = r0.x + r4.x + ((r1.x + r3.x) << 2) + r2.x * 6u + 0x00800080u;
        const unsigned ay = r0.y + r4.y + ((r1.y + r3.y) << 2) + r2.y * 6u + 0x00800080u;
        // bytes 1 and 3 of ax/ay are the four outputs ((sum + 128) >> 8)
        const unsigned out = __byte_perm(ax, ay, 0x7531);
        uint8_t* d = dst + (long long)oy * dpitch + ox;
        if (ox + 3 < dw && ((((uintptr_t)d) & 3) == 0)) {
            *reinterpret_cast<unsigned*>(d) = out;
        } else {
            d[0] = (uint8_t)(out & 0xff);
            if (ox + 1 < dw) d[1] = (uint8_t)((out >> 8) & 0xff);
            if (ox + 2 < dw) d[2] = (uint8_t)((out >> 16) & 0xff);
            if (ox + 3 < dw) d[3] = (uint8_t)(out >> 24);
        }
    }
}

// ---- register-streaming kernel (widths that are multiples of 8, 8-byte aligned rows) ----------------
// Thread t owns 4 output columns (source columns 8t..8t+7) for STRIP_H output rows and walks DOWN the
// source rows: one 8-byte load per row, the 3 halo bytes come from the neighbouring lanes by shuffle
// (lane 0 / 31 fetch them), the horizontal [1 4 6 4 1] pass is 4 dp4a on funnel-shifted words, and the
// five most recent filtered rows stay in registers as packed 16-bit pairs for the vertical pass.
// No shared memory, no barriers; 128-byte coalesced stores.
constexpr int STRIP_THREADS = 64;

// STRIP_H output rows per CTA: 16 for the big level-1 grids; 8 for the smaller upper levels, whose grids would
// otherwise not fill the machine (one partial wave of latency-bound threads)
template <int STRIP_H>
__global__ void __launch_bounds__(STRIP_THREADS)
pyrdown_strip_kernel(const uint8_t* __restrict__ src_base, long long src_stride, int sw, int sh, int spitch,
                     uint8_t* __restrict__ dst_base, long long dst_stride, int dw, int dh, int dpitch)
{
    const uint8_t* __restrict__ src = src_base + (long long)blockIdx.z * src_stride;
    uint8_t* __restrict__ dst = dst_base + (long long)blockIdx.z * dst_stride;
    const int lane = threadIdx.x & 31;
    const int t = blockIdx.x * STRIP_THREADS + threadIdx.x;
    const int nt_row = sw >> 3;                       // threads that own real columns
    const int tc = min(t, nt_row - 1);                // clamped: idle threads still feed the shuffles
    const bool last = tc == nt_row - 1;
    const int oy0 = blockIdx.y * STRIP_H;
    const unsigned K = 0x04060401u;

    constexpr int NROWS = 2 * STRIP_H + 3;
    constexpr int BATCH = 7;   // source rows loaded ahead of the arithmetic (memory-level parallelism)
    uint2 win[NROWS];
#pragma unroll
    for (int r0 = 0; r0 < NROWS; r0 += BATCH) {
        uint2 a[BATCH];
        unsigned edge_lo[BATCH], edge_hi[BATCH];
#pragma unroll
        for (int k = 0; k < BATCH; ++k) {
            if (r0 + k < NROWS) {
                int yy = 2 * oy0 - 2 + r0 + k;
                yy = yy < 0 ? -yy : (yy >= sh ? 2 * sh - 2 - yy : yy);
                yy = max(0, min(yy, sh - 1));
                const uint8_t* row = src + (long long)yy * spitch + 8 * tc;
                a[k] = __ldg(reinterpret_cast<const uint2*>(row));
                edge_lo[k] = (lane == 0 && tc > 0) ? (unsigned)__ldg(reinterpret_cast<const unsigned short*>(row - 2)) : 0u;
                edge_hi[k] = (lane == 31 && !last) ? (unsigned)__ldg(row + 8) : 0u;
            }
        }
#pragma unroll
        for (int k = 0; k < BATCH; ++k) {
            const int r = r0 + k;
            if (r < NROWS) {
                const uint2 v = a[k];
                unsigned lo2 = __shfl_up_sync(0xffffffffu, v.y, 1) >> 16;      // source columns 8t-2, 8t-1
                unsigned b0 = __shfl_down_sync(0xffffffffu, v.x, 1) & 0xffu;   // source column 8t+8
                if (lane == 0) lo2 = tc > 0 ? edge_lo[k] : (((v.x >> 16) & 0xffu) | (((v.x >> 8) & 0xffu) << 8));   // REFLECT_101 at column 0
                if (last) b0 = (v.y >> 16) & 0xffu;                                                                // REFLECT_101 at column sw
                else if (lane == 31) b0 = edge_hi[k];
                const unsigned h0 = dp4a_uu(lo2 | (v.x << 16), K, (v.x >> 16) & 0xffu);
                const unsigned h1 = dp4a_uu(v.x, K, v.y & 0xffu);
                const unsigned h2 = dp4a_uu(__funnelshift_r(v.x, v.y, 16), K, (v.y >> 16) & 0xffu);
                const unsigned h3 = dp4a_uu(v.y, K, b0);
                win[r] = make_uint2(h0 | (h1 << 16), h2 | (h3 << 16));
                if (r >= 4 && (r & 1) == 0) {
                    const int y = (r - 4) >> 1, oy = oy0 + y;
                    const unsigned ax = win[r - 4].x + win[r].x + ((win[r - 3].x + win[r - 1].x) << 2) + win[r - 2].x * 6u + 0x00800080u;
                    const unsigned ay = win[r - 4].y + win[r].y + ((win[r - 3].y + win[r - 1].y) << 2) + win[r - 2].y * 6u + 0x00800080u;
                    if (t < nt_row && oy < dh)
                        *reinterpret_cast<unsigned*>(dst + (long long)oy * dpitch + 4 * t) = __byte_perm(ax, ay, 0x7531);
                }
            }
        }
    }
}

// ---- fused two-level kernel: levels 1 AND 2 from one pass over the frame ------------------------------------------------
// (source width a multiple of 16, 16-byte aligned rows: 1080p, 4K, ...)
// Level 2 taps level-1 rows 2y-2..2y+2, so a thread that walks down the frame producing level-1 rows can emit level 2 from
// the five most recent ones without level 1 ever being read back (67 MB per 129 1080p frames, and one launch instead of two).
// Thread t owns 16 source columns = 8 level-1 columns = 4 level-2 columns and walks DOWN: per iteration four new source rows
// (one 16-byte load each, the next four already in flight) -> two level-1 rows -> one level-2 row.  Horizontal halos come from
// the neighbouring lanes by shuffle at BOTH levels; a warp therefore overlaps its neighbours by one lane on each side (lanes
// 1..30 own columns and store, lanes 0 / 31 recompute the neighbours' level-1 columns; their own outer halo bytes are fetched
// from memory).  Rows: a CTA produces H2 level-2 rows and the 2*H2 level-1 rows under them, plus three level-1 halo rows
// (two above, one below) that are computed and not stored.  REFLECT_101: columns by byte selection at t = 0 / t = last; rows
// by reflecting the (virtual) source row index -- which commutes with the symmetric filter at the top edge; below the bottom
// edge it does not (even heights), so a level-1 row r >= h1 is recomputed from the five source rows of row 2*h1-2-r.
constexpr int F2_WARPS = 2;                 // warps per CTA, side by side along x
constexpr int F2_OWN = 30;                  // owner lanes per warp

// horizontally filtered row of 16 source bytes: 8 outputs as packed 16-bit pairs.  pw = the word LEFT of v (its two top bytes are
// source columns -2, -1), nx = the word RIGHT of v (its low byte is column 16).  Output m taps bytes 2m-2 .. 2m+2:
// [1 4 6 4] on the word starting at byte 2m-2 plus byte 2 of the word starting at byte 2m -- two dp4a, no byte extraction.
__device__ __forceinline__ uint4 hfilt16(const uint4 v, unsigned pw, unsigned nx)
{
    const unsigned K = 0x04060401u, K5 = 0x00010000u;
    const unsigned wm2 = __funnelshift_r(pw, v.x, 16), w2 = __funnelshift_r(v.x, v.y, 16), w6 = __funnelshift_r(v.y, v.z, 16);
    const unsigned w10 = __funnelshift_r(v.z, v.w, 16), w14 = __funnelshift_r(v.w, nx, 16);
    const unsigned h0 = dp4a_uu(v.x, K5, dp4a_uu(wm2, K, 0u));
    const unsigned h1 = dp4a_uu(w2, K5, dp4a_uu(v.x, K, 0u));
    const unsigned h2 = dp4a_uu(v.y, K5, dp4a_uu(w2, K, 0u));
    const unsigned h3 = dp4a_uu(w6, K5, dp4a_uu(v.y, K, 0u));
    const unsigned h4 = dp4a_uu(v.z, K5, dp4a_uu(w6, K, 0u));
    const unsigned h5 = dp4a_uu(w10, K5, dp4a_uu(v.z, K, 0u));
    const unsigned h6 = dp4a_uu(v.w, K5, dp4a_uu(w10, K, 0u));
    const unsigned h7 = dp4a_uu(w14, K5, dp4a_uu(v.w, K, 0u));
    return make_uint4(__byte_perm(h0, h1, 0x5410), __byte_perm(h2, h3, 0x5410), __byte_perm(h4, h5, 0x5410), __byte_perm(h6, h7, 0x5410));
}

// the same for a level-1 row of 8 bytes (o0, o1): 4 outputs
__device__ __forceinline__ uint2 hfilt8(unsigned o0, unsigned o1, unsigned pw, unsigned nx)
{
    const unsigned K = 0x04060401u, K5 = 0x00010000u;
    const unsigned wm2 = __funnelshift_r(pw, o0, 16), w2 = __funnelshift_r(o0, o1, 16), w6 = __funnelshift_r(o1, nx, 16);
    const unsigned g0 = dp4a_uu(o0, K5, dp4a_uu(wm2, K, 0u));
    const unsigned g1 = dp4a_uu(w2, K5, dp4a_uu(o0, K, 0u));
    const unsigned g2 = dp4a_uu(o1, K5, dp4a_uu(w2, K, 0u));
    const unsigned g3 = dp4a_uu(w6, K5, dp4a_uu(o1, K, 0u));
    return make_uint2(__byte_perm(g0, g1, 0x5410), __byte_perm(g2, g3, 0x5410));
}

// vertical [1 4 6 4 1] on packed 16-bit pairs, + 128, >> 8: two words of pairs -> four output bytes
__device__ __forceinline__ unsigned vfilt4(unsigned a0, unsigned a1, unsigned a2, unsigned a3, unsigned a4, unsigned b0, unsigned b1, unsigned b2,
                                           unsigned b3, unsigned b4)
{
    const unsigned ax = a0 + a4 + ((a1 + a3) << 2) + a2 * 6u + 0x00800080u;
    const unsigned ay = b0 + b4 + ((b1 + b3) << 2) + b2 * 6u + 0x00800080u;
    return __byte_perm(ax, ay, 0x7531);
}

struct F2Args {
    const uint8_t* src; long long src_stride; int sw, sh, spitch;
    uint8_t* d1; long long d_stride; int w1, h1, p1;
    uint8_t* d2; int w2, h2, p2;
};

struct F2Row { uint4 v; unsigned elo, ehi; };      // one source row in flight: the thread's 16 bytes + the warp-edge halo words

// INTERIOR: every source row of the strip lies inside the frame (row pointers are stepped, no reflection, no bottom-edge rows)
template <int H2, bool INTERIOR>
__device__ __forceinline__ void f2_strip(const F2Args& A, const uint8_t* __restrict__ src, uint8_t* __restrict__ d1, uint8_t* __restrict__ d2)
{
    const int lane = threadIdx.x & 31, wx = blockIdx.x * F2_WARPS + (threadIdx.x >> 5);
    const int nt = A.sw >> 4;                                   // threads that own real columns
    const int t = wx * F2_OWN - 1 + lane;
    const int tc = max(0, min(t, nt - 1));
    const bool owner = lane >= 1 && lane <= F2_OWN && t < nt;   // t >= 0 follows from lane >= 1
    const bool first = tc == 0, last = tc == nt - 1;
    const bool edge_warp = __any_sync(0xffffffffu, first || last);
    const bool ld_lo = lane == 0 && !first, ld_hi = lane == 31 && !last;
    const int y2_0 = blockIdx.y * H2;
    const int R0 = 2 * y2_0 - 2;                                // first (virtual) level-1 row of the strip
    const int sh = A.sh, spitch = A.spitch;
    const uint8_t* __restrict__ col = src + 16 * tc;
    const uint8_t* __restrict__ rowp = col + (long long)(2 * R0 - 2) * spitch;      // INTERIOR: the next row to load

    auto load_row = [&](int sv, F2Row& r) {
        const uint8_t* row;
        if (INTERIOR) {
            row = rowp;
            rowp += spitch;
        } else {
            int yy = sv < 0 ? -sv : (sv >= sh ? 2 * sh - 2 - sv : sv);
            yy = max(0, min(yy, sh - 1));
            row = col + (long long)yy * spitch;
        }
        r.v = __ldg(reinterpret_cast<const uint4*>(row));
        r.elo = 0u; r.ehi = 0u;
        if (ld_lo) r.elo = __ldg(reinterpret_cast<const unsigned*>(row - 4));     // columns 16t-4 .. 16t-1
        if (ld_hi) r.ehi = __ldg(reinterpret_cast<const unsigned*>(row + 16));    // columns 16t+16 .. 16t+19
    };
    auto filt_row = [&](const F2Row& r) -> uint4 {
        unsigned pw = __shfl_up_sync(0xffffffffu, r.v.w, 1);      // top bytes: source columns 16t-2, 16t-1
        unsigned nx = __shfl_down_sync(0xffffffffu, r.v.x, 1);    // low byte: source column 16t+16
        if (lane == 0) pw = r.elo;
        if (lane == 31) nx = r.ehi;
        if (edge_warp) {
            if (first) pw = __byte_perm(r.v.x, 0u, 0x1200);       // REFLECT_101 at column 0: columns 2, 1
            if (last) nx = __byte_perm(r.v.w, 0u, 0x0002);        // REFLECT_101 at column sw: column sw-2
        }
        return hfilt16(r.v, pw, nx);
    };
    // level-1 row (8 bytes in o0, o1) -> its horizontally filtered form for level 2
    auto hfilt_l1 = [&](unsigned o0, unsigned o1) -> uint2 {
        unsigned pw = __shfl_up_sync(0xffffffffu, o1, 1);         // top bytes: level-1 columns 8t-2, 8t-1
        unsigned nx = __shfl_down_sync(0xffffffffu, o0, 1);       // low byte: level-1 column 8t+8
        if (edge_warp) {
            if (t <= 0) pw = __byte_perm(o0, 0u, 0x1200);
            if (t >= nt - 1) nx = __byte_perm(o1, 0u, 0x0002);
        }
        return hfilt8(o0, o1, pw, nx);
    };
    // a level-1 row computed from scratch (five source rows): only for virtual rows at or below the bottom edge
    auto l1_row_direct = [&](int r1, unsigned& o0, unsigned& o1) {
        uint4 f[5];
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            int yy = 2 * r1 - 2 + k;
            yy = yy < 0 ? -yy : (yy >= sh ? 2 * sh - 2 - yy : yy);
            yy = max(0, min(yy, sh - 1));
            const uint8_t* row = col + (long long)yy * spitch;
            F2Row r;
            r.v = __ldg(reinterpret_cast<const uint4*>(row));
            r.elo = ld_lo ? __ldg(reinterpret_cast<const unsigned*>(row - 4)) : 0u;
            r.ehi = ld_hi ? __ldg(reinterpret_cast<const unsigned*>(row + 16)) : 0u;
            f[k] = filt_row(r);
        }
        o0 = vfilt4(f[0].x, f[1].x, f[2].x, f[3].x, f[4].x, f[0].y, f[1].y, f[2].y, f[3].y, f[4].y);
        o1 = vfilt4(f[0].z, f[1].z, f[2].z, f[3].z, f[4].z, f[0].w, f[1].w, f[2].w, f[3].w, f[4].w);
    };

    // prologue: the three carried source rows 2*R0-2 .. 2*R0, and the first four rows of the loop in flight
    uint4 c0, c1, c2;
    {
        F2Row r[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) load_row(2 * R0 - 2 + k, r[k]);
        c0 = filt_row(r[0]); c1 = filt_row(r[1]); c2 = filt_row(r[2]);
    }
    F2Row bufA[4], bufB[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) load_row(2 * R0 + 1 + k, bufA[k]);
    uint2 q0 = make_uint2(0, 0), q1 = q0, q2 = q0, q3 = q0;        // level-2-filtered level-1 rows a_{j-2}, b_{j-2}, a_{j-1}, b_{j-1}

    // iteration j: four source rows (cur; the next four are requested into nxt first) -> level-1 rows a_j, b_j -> level-2 row j-2
    auto iteration = [&](int j, F2Row (&cur)[4], F2Row (&nxt)[4]) {
        const int ra = R0 + 2 * j, rb = ra + 1;
        if (j + 1 < H2 + 2) {
#pragma unroll
            for (int k = 0; k < 4; ++k) load_row(2 * ra + 5 + k, nxt[k]);
        }
        const uint4 f0 = filt_row(cur[0]), f1 = filt_row(cur[1]), f2 = filt_row(cur[2]), f3 = filt_row(cur[3]);
        unsigned a0 = vfilt4(c0.x, c1.x, c2.x, f0.x, f1.x, c0.y, c1.y, c2.y, f0.y, f1.y);
        unsigned a1 = vfilt4(c0.z, c1.z, c2.z, f0.z, f1.z, c0.w, c1.w, c2.w, f0.w, f1.w);
        unsigned b0 = vfilt4(c2.x, f0.x, f1.x, f2.x, f3.x, c2.y, f0.y, f1.y, f2.y, f3.y);
        unsigned b1 = vfilt4(c2.z, f0.z, f1.z, f2.z, f3.z, c2.w, f0.w, f1.w, f2.w, f3.w);
        c0 = f1; c1 = f2; c2 = f3;
        if (!INTERIOR) {                                           // rows at / below the bottom edge: level-1 row 2*h1-2-r instead
            if (ra >= A.h1) l1_row_direct(2 * A.h1 - 2 - ra, a0, a1);
            if (rb >= A.h1) l1_row_direct(2 * A.h1 - 2 - rb, b0, b1);
        }
        if (owner && j >= 1 && j <= H2) {
            uint8_t* o = d1 + (long long)ra * A.p1 + 8 * t;
            if (INTERIOR || ra < A.h1) *reinterpret_cast<uint2*>(o) = make_uint2(a0, a1);
            if (INTERIOR || rb < A.h1) *reinterpret_cast<uint2*>(o + A.p1) = make_uint2(b0, b1);
        }
        const uint2 ga = hfilt_l1(a0, a1), gb = hfilt_l1(b0, b1);
        if (j >= 2) {
            const int y2 = y2_0 + j - 2;
            const unsigned o = vfilt4(q0.x, q1.x, q2.x, q3.x, ga.x, q0.y, q1.y, q2.y, q3.y, ga.y);
            if (owner && (INTERIOR || y2 < A.h2)) *reinterpret_cast<unsigned*>(d2 + (long long)y2 * A.p2 + 4 * t) = o;
        }
        q0 = q2; q1 = q3; q2 = ga; q3 = gb;
    };
    static_assert((H2 & 1) == 0, "the loop is unrolled by two (ping-pong row buffers)");
    const int jend = min(H2 + 2, A.h2 - y2_0 + 2);               // a strip that hangs over the frame stops after its last real level-2 row
#pragma unroll 1
    for (int j = 0; j < jend; j += 2) {
        iteration(j, bufA, bufB);
        iteration(j + 1, bufB, bufA);
    }
}

// ---- the same strip with the source rows staged by TMA ---------------------------------------------------------------
// The LDG form keeps at most eight 16-byte loads per thread in flight (registers), ~4.7 MB over the chip at 16 warps per SM --
// short of what 6.4 TB/s needs at HBM latency.  Here lane 0 of each warp issues one cp.async.bulk.tensor (3-D map: x in 32-bit
// words, row, frame) per GROUP of four source rows: box = 136 words x 4 rows = the warp's 32 x 16 bytes plus 16 bytes either side
// (the halo words of lanes 0 / 31), into a ring of F2_STAGES groups per warp, completion on one mbarrier per stage.  Bytes in
// flight no longer cost registers: 4 stages x 2176 B per warp.  Out-of-frame columns are zero-filled by the TMA unit and replaced
// by the REFLECT_101 bytes exactly as in the LDG form; strips that touch the top / bottom edge take the LDG form (row reflection).
constexpr int F2_STAGES = 4;
constexpr int F2_BOXW = 136;                        // 32-bit words per box row: 16 B + 32 lanes x 16 B + 16 B
constexpr int F2_GROUP_BYTES = 4 * F2_BOXW * 4;     // 2176 = 17 x 128

__device__ __forceinline__ uint32_t f2_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void f2_mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void f2_mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void f2_mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "F2_WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra F2_DONE_%=;\n\t"
        "bra F2_WAIT_%=;\n\t"
        "F2_DONE_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void f2_tma_load_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint32_t bar)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar) : "memory");
}

template <int H2>
__device__ __forceinline__ void f2_strip_tma(const F2Args& A, const CUtensorMap* tm, uint8_t* __restrict__ d1, uint8_t* __restrict__ d2,
                                             unsigned (*ring)[4][F2_BOXW], unsigned long long* bars)
{
    const int lane = threadIdx.x & 31, wx = blockIdx.x * F2_WARPS + (threadIdx.x >> 5);
    const int nt = A.sw >> 4;
    const int t = wx * F2_OWN - 1 + lane;
    const int tc = max(0, min(t, nt - 1));
    const bool owner = lane >= 1 && lane <= F2_OWN && t < nt;
    const bool first = tc == 0, last = tc == nt - 1;
    const bool edge_warp = __any_sync(0xffffffffu, first || last);
    const int y2_0 = blockIdx.y * H2;
    const int R0 = 2 * y2_0 - 2;
    const int jend = min(H2 + 2, A.h2 - y2_0 + 2);
    const int ngroups = 1 + jend;                               // group 0 = the carried rows, group j+1 = the rows of iteration j
    const int cx = 4 * (wx * F2_OWN - 1) - 4;                   // box origin in 32-bit words: 16 bytes left of lane 0
    const int row_g0 = 2 * R0 - 3;                              // group 0 = rows 2*R0-3 .. 2*R0 (the first is not used)
    const uint32_t ring0 = f2_smem_u32(&ring[0][0][0]), bar0 = f2_smem_u32(bars);

    auto issue = [&](int g) {                                   // lane 0 only
        const int sidx = g % F2_STAGES;
        const uint32_t bar = bar0 + 8u * sidx;
        f2_mbar_expect_tx(bar, F2_GROUP_BYTES);
        f2_tma_load_3d(ring0 + (uint32_t)sidx * F2_GROUP_BYTES, tm, cx, g == 0 ? row_g0 : 2 * R0 + 1 + 4 * (g - 1), (int)blockIdx.z, bar);
    };
    if (lane == 0) {
        for (int s2 = 0; s2 < F2_STAGES; ++s2) f2_mbar_init(bar0 + 8u * s2, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        for (int g = 0; g < F2_STAGES && g < ngroups; ++g) issue(g);
    }
    __syncwarp();

    // the four rows of group g -> registers (the stage is re-armed for group g + F2_STAGES as soon as every lane has read it)
    auto fetch = [&](int g, F2Row (&r)[4]) {
        const int sidx = g % F2_STAGES;
        f2_mbar_wait(bar0 + 8u * sidx, (unsigned)(g / F2_STAGES) & 1u);
        const unsigned (*rows)[F2_BOXW] = ring[sidx];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            r[k].v = *reinterpret_cast<const uint4*>(&rows[k][4 + 4 * lane]);
            r[k].elo = 0u; r[k].ehi = 0u;
            if (lane == 0) r[k].elo = rows[k][3];
            if (lane == 31) r[k].ehi = rows[k][4 + 128];
        }
        __syncwarp();
        if (lane == 0 && g + F2_STAGES < ngroups) issue(g + F2_STAGES);
    };
    auto filt_row = [&](const F2Row& r) -> uint4 {
        unsigned pw = __shfl_up_sync(0xffffffffu, r.v.w, 1);
        unsigned nx = __shfl_down_sync(0xffffffffu, r.v.x, 1);
        if (lane == 0) pw = r.elo;
        if (lane == 31) nx = r.ehi;
        if (edge_warp) {
            if (first) pw = __byte_perm(r.v.x, 0u, 0x1200);
            if (last) nx = __byte_perm(r.v.w, 0u, 0x0002);
        }
        return hfilt16(r.v, pw, nx);
    };
    auto hfilt_l1 = [&](unsigned o0, unsigned o1) -> uint2 {
        unsigned pw = __shfl_up_sync(0xffffffffu, o1, 1);
        unsigned nx = __shfl_down_sync(0xffffffffu, o0, 1);
        if (edge_warp) {
            if (t <= 0) pw = __byte_perm(o0, 0u, 0x1200);
            if (t >= nt - 1) nx = __byte_perm(o1, 0u, 0x0002);
        }
        return hfilt8(o0, o1, pw, nx);
    };

    uint4 c0, c1, c2;
    {
        F2Row r[4];
        fetch(0, r);
        c0 = filt_row(r[1]); c1 = filt_row(r[2]); c2 = filt_row(r[3]);
    }
    uint2 q0 = make_uint2(0, 0), q1 = q0, q2 = q0, q3 = q0;
#pragma unroll 1
    for (int j = 0; j < jend; ++j) {
        const int ra = R0 + 2 * j;
        F2Row cur[4];
        fetch(j + 1, cur);
        const uint4 f0 = filt_row(cur[0]), f1 = filt_row(cur[1]), f2 = filt_row(cur[2]), f3 = filt_row(cur[3]);
        const unsigned a0 = vfilt4(c0.x, c1.x, c2.x, f0.x, f1.x, c0.y, c1.y, c2.y, f0.y, f1.y);
        const unsigned a1 = vfilt4(c0.z, c1.z, c2.z, f0.z, f1.z, c0.w, c1.w, c2.w, f0.w, f1.w);
        const unsigned b0 = vfilt4(c2.x, f0.x, f1.x, f2.x, f3.x, c2.y, f0.y, f1.y, f2.y, f3.y);
        const unsigned b1 = vfilt4(c2.z, f0.z, f1.z, f2.z, f3.z, c2.w, f0.w, f1.w, f2.w, f3.w);
        c0 = f1; c1 = f2; c2 = f3;
        if (owner && j >= 1 && j <= H2) {
            uint8_t* o = d1 + (long long)ra * A.p1 + 8 * t;
            *reinterpret_cast<uint2*>(o) = make_uint2(a0, a1);
            *reinterpret_cast<uint2*>(o + A.p1) = make_uint2(b0, b1);
        }
        const uint2 ga = hfilt_l1(a0, a1), gb = hfilt_l1(b0, b1);
        if (j >= 2) {
            const unsigned o = vfilt4(q0.x, q1.x, q2.x, q3.x, ga.x, q0.y, q1.y, q2.y, q3.y, ga.y);
            if (owner) *reinterpret_cast<unsigned*>(d2 + (long long)(y2_0 + j - 2) * A.p2 + 4 * t) = o;
        }
        q0 = q2; q1 = q3; q2 = ga; q3 = gb;
    }
}

template <int H2>
__global__ void __launch_bounds__(32 * F2_WARPS)
pyrdown2_fused_tma_kernel(const F2Args A, const __grid_constant__ CUtensorMap tm)
{
    __shared__ __align__(128) unsigned ring[F2_WARPS][F2_STAGES][4][F2_BOXW];
    __shared__ __align__(8) unsigned long long bars[F2_WARPS][F2_STAGES];
    const uint8_t* __restrict__ src = A.src + (long long)blockIdx.z * A.src_stride;
    uint8_t* __restrict__ d1 = A.d1 + (long long)blockIdx.z * A.d_stride;
    uint8_t* __restrict__ d2 = A.d2 + (long long)blockIdx.z * A.d_stride;
    const int R0 = 2 * (int)blockIdx.y * H2 - 2;
    const bool interior = 2 * R0 - 2 >= 0 && 2 * (R0 + 2 * (H2 + 1) + 1) + 2 < A.sh;
    const int w = threadIdx.x >> 5;
    if (interior) f2_strip_tma<H2>(A, &tm, d1, d2, ring[w], bars[w]);
    else f2_strip<H2, false>(A, src, d1, d2);
}

template <int H2>
__global__ void __launch_bounds__(32 * F2_WARPS)
pyrdown2_fused_kernel(const F2Args A)
{
    const uint8_t* __restrict__ src = A.src + (long long)blockIdx.z * A.src_stride;
    uint8_t* __restrict__ d1 = A.d1 + (long long)blockIdx.z * A.d_stride;
    uint8_t* __restrict__ d2 = A.d2 + (long long)blockIdx.z * A.d_stride;
    const int R0 = 2 * (int)blockIdx.y * H2 - 2;
    // rows never need more than one reflection: |overshoot| <= 6 at the top (sh >= 16 checked by the host); rows far below the frame
    // (a strip that hangs over it) are clamped, their results are never stored
    const bool interior = 2 * R0 - 2 >= 0 && 2 * (R0 + 2 * (H2 + 1) + 1) + 2 < A.sh;
    if (interior) f2_strip<H2, true>(A, src, d1, d2);
    else f2_strip<H2, false>(A, src, d1, d2);
}

typedef CUresult (*F2EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// 3-D map over the frames as 32-bit words: (x word, row, frame); box = 136 words x 4 rows x 1 frame; zero fill outside
bool f2_make_map(CUtensorMap* tm, const uint8_t* frames, int sw, int sh, int pitch, long long frame_stride, int nframes)
{
    static F2EncodeTiledFn enc = nullptr;
    if (!enc) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess || qres != cudaDriverEntryPointSuccess)
            return false;
        enc = (F2EncodeTiledFn)p;
    }
    const long long fs = nframes > 1 ? frame_stride : (long long)pitch * sh;
    if (fs <= 0 || (fs & 15) != 0) return false;
    cuuint64_t dims[3] = {(cuuint64_t)(sw / 4), (cuuint64_t)sh, (cuuint64_t)nframes};
    cuuint64_t strides[2] = {(cuuint64_t)pitch, (cuuint64_t)fs};
    cuuint32_t box[3] = {(cuuint32_t)F2_BOXW, 4, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    return enc(tm, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, (void*)frames, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

__global__ void decimate4_kernel(const uint8_t* __restrict__ src, int sw, int sh, int spitch, uint8_t* __restrict__ dst,
                                 int dw, int dh, int dpitch)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= dw || y >= dh) return;
    const int sx = min(4 * x, sw - 1), sy = min(4 * y, sh - 1);
    dst[(long long)y * dpitch + x] = __ldg(src + (long long)sy * spitch + sx);
}

}  // namespace

VEL_API int vel_pyr_layout_make(int32_t width, int32_t height, int32_t win_w, int32_t win_h, int32_t max_level,
                                vel_pyr_layout* out)
{
    VEL_CHECK_ARG(out != nullptr, "vel_pyr_layout_make: out is NULL");
    VEL_CHECK_ARG(width > 0 && height > 0 && win_w > 0 && win_h > 0, "vel_pyr_layout_make: bad size %dx%d win %dx%d", width,
                  height, win_w, win_h);
    VEL_CHECK_ARG(max_level >= 0 && max_level < VEL_MAX_LEVELS, "vel_pyr_layout_make: max_level %d outside [0,%d)", max_level,
                  VEL_MAX_LEVELS);
    int w = width, h = height, lvl = 0;
    int64_t off = 0;
    out->width[0] = w; out->height[0] = h; out->pitch[0] = 0; out->offset[0] = 0;
    while (lvl < max_level) {
        const int nw = (w + 1) / 2, nh = (h + 1) / 2;
        if (nw <= win_w || nh <= win_h) break;
        ++lvl;
        w = nw; h = nh;
        const int pitch = (w + 15) & ~15;
        out->width[lvl] = w; out->height[lvl] = h; out->pitch[lvl] = pitch; out->offset[lvl] = off;
        off += ((int64_t)pitch * h + 255) & ~(int64_t)255;
    }
    for (int l = lvl + 1; l < VEL_MAX_LEVELS; ++l) { out->width[l] = out->height[l] = out->pitch[l] = 0; out->offset[l] = 0; }
    out->max_level = lvl;
    out->bytes = off;
    return VEL_OK;
}

VEL_API int vel_pyramid_u8(const uint8_t* frames, int64_t frame_stride, int32_t pitch, int32_t nframes,
                           const vel_pyr_layout* L, uint8_t* pyr, int64_t pyr_stride, vel_stream_t stream)
{
    VEL_CHECK_ARG(frames && L, "vel_pyramid_u8: NULL argument");
    VEL_CHECK_ARG(nframes > 0 && nframes <= 65535, "vel_pyramid_u8: nframes %d outside [1,65535]", nframes);
    VEL_CHECK_ARG(L->max_level >= 0 && L->max_level < VEL_MAX_LEVELS, "vel_pyramid_u8: bad layout");
    VEL_CHECK_ARG(pitch >= L->width[0], "vel_pyramid_u8: pitch %d < width %d", pitch, L->width[0]);
    if (L->max_level == 0) return VEL_OK;
    VEL_CHECK_ARG(pyr != nullptr && pyr_stride >= L->bytes, "vel_pyramid_u8: pyramid buffer missing or stride too small");
    cudaStream_t st = (cudaStream_t)stream;
    int l_first = 1;
    {   // levels 1 and 2 in one pass over the frames (VEL_PYR_FUSED=0 keeps the level-by-level kernels: the cross-check of the tests)
        const char* env = getenv("VEL_PYR_FUSED");
        const bool fused_on = !(env && env[0] == '0');
        const int sw = L->width[0], sh = L->height[0];
        const bool ok = fused_on && L->max_level >= 2 && sw % 16 == 0 && sw >= 32 && sh >= 16 && pitch % 16 == 0 &&
                        ((((uintptr_t)frames) | (uintptr_t)frame_stride) & 15) == 0 &&
                        ((((uintptr_t)(pyr + L->offset[1])) | (uintptr_t)pyr_stride | (uintptr_t)L->pitch[1]) & 7) == 0 &&
                        ((((uintptr_t)(pyr + L->offset[2])) | (uintptr_t)L->pitch[2]) & 3) == 0;
        if (ok) {
            F2Args A;
            A.src = frames; A.src_stride = frame_stride; A.sw = sw; A.sh = sh; A.spitch = pitch;
            A.d1 = pyr + L->offset[1]; A.d_stride = pyr_stride; A.w1 = L->width[1]; A.h1 = L->height[1]; A.p1 = L->pitch[1];
            A.d2 = pyr + L->offset[2]; A.w2 = L->width[2]; A.h2 = L->height[2]; A.p2 = L->pitch[2];
            const int nt = sw / 16, warps = (nt + F2_OWN - 1) / F2_OWN;
            // rows per strip: 32 level-2 rows keep the halo rows (11 source rows per strip) below 9 % of the loads; small images
            // take 16 so that the grid still fills the machine
            const int gx = (warps + F2_WARPS - 1) / F2_WARPS;
            const char* eh = getenv("VEL_PYR_H2");
            const int h2sel = eh ? atoi(eh) : ((long long)gx * ((A.h2 + 31) / 32) * nframes >= 8ll * kNumSMs ? 32 : 16);
            const char* et = getenv("VEL_PYR_TMA");
            CUtensorMap tm;
            const bool tma = !(et && et[0] == '0') && h2sel == 32 && f2_make_map(&tm, frames, sw, sh, pitch, frame_stride, nframes);
            if (tma) {
                dim3 grid(gx, (A.h2 + 31) / 32, nframes);
                pyrdown2_fused_tma_kernel<32><<<grid, 32 * F2_WARPS, 0, st>>>(A, tm);
            } else if (h2sel == 32) {
                dim3 grid(gx, (A.h2 + 31) / 32, nframes);
                pyrdown2_fused_kernel<32><<<grid, 32 * F2_WARPS, 0, st>>>(A);
            } else if (h2sel == 24) {
                dim3 grid(gx, (A.h2 + 23) / 24, nframes);
                pyrdown2_fused_kernel<24><<<grid, 32 * F2_WARPS, 0, st>>>(A);
            } else {
                dim3 grid(gx, (A.h2 + 15) / 16, nframes);
                pyrdown2_fused_kernel<16><<<grid, 32 * F2_WARPS, 0, st>>>(A);
            }
            VEL_LAUNCH_CHECK("pyrdown2_fused_kernel");
            l_first = 3;
        }
    }
    for (int l = l_first; l <= L->max_level; ++l) {
        const uint8_t* src = l == 1 ? frames : pyr + L->offset[l - 1];
        const long long sstride = l == 1 ? frame_stride : pyr_stride;
        const int spitch = l == 1 ? pitch : L->pitch[l - 1];
        const int sw = L->width[l - 1];
        const bool strip_ok = (sw % 8 == 0) && sw >= 16 && (spitch % 8 == 0) && ((((uintptr_t)src) | (uintptr_t)sstride) & 7) == 0 &&
                              ((((uintptr_t)(pyr + L->offset[l])) | (uintptr_t)pyr_stride | (uintptr_t)L->pitch[l]) & 3) == 0;
        if (strip_ok) {
            const int gx = (sw / 8 + STRIP_THREADS - 1) / STRIP_THREADS;
            const long long ctas16 = (long long)gx * ((L->height[l] + 15) / 16) * nframes;
            if (ctas16 >= 8ll * 4 * kNumSMs) {      // >= 8 CTAs of 2 warps per scheduler: enough loads in flight
                dim3 grid(gx, (L->height[l] + 15) / 16, nframes);
                pyrdown_strip_kernel<16><<<grid, STRIP_THREADS, 0, st>>>(src, sstride, sw, L->height[l - 1], spitch, pyr + L->offset[l],
                                                                        pyr_stride, L->width[l], L->height[l], L->pitch[l]);
            } else {
                dim3 grid(gx, (L->height[l] + 7) / 8, nframes);
                pyrdown_strip_kernel<8><<<grid, STRIP_THREADS, 0, st>>>(src, sstride, sw, L->height[l - 1], spitch, pyr + L->offset[l],
                                                                       pyr_stride, L->width[l], L->height[l], L->pitch[l]);
            }
            VEL_LAUNCH_CHECK("pyrdown_strip_kernel");
        } else {
            dim3 grid((L->width[l] + TILE_W - 1) / TILE_W, (L->height[l] + TILE_H - 1) / TILE_H, nframes);
            pyrdown_tiled_kernel<<<grid, PYR_THREADS, 0, st>>>(src, sstride, sw, L->height[l - 1], spitch, pyr + L->offset[l],
                                                              pyr_stride, L->width[l], L->height[l], L->pitch[l]);
            VEL_LAUNCH_CHECK("pyrdown_tiled_kernel");
        }
    }
    return VEL_OK;
}

VEL_API int vel_decimate4_u8(const uint8_t* src, int32_t width, int32_t height, int32_t pitch, uint8_t* dst, int32_t dst_width,
                             int32_t dst_height, int32_t dst_pitch, vel_stream_t stream)
{
    VEL_CHECK_ARG(src && dst, "vel_decimate4_u8: NULL argument");
    VEL_CHECK_ARG(width > 0 && height > 0 && dst_width > 0 && dst_height > 0 && dst_height <= 65535,
                  "vel_decimate4_u8: bad size");
    dim3 grid((dst_width + 255) / 256, dst_height);
    decimate4_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, width, height, pitch, dst, dst_width, dst_height, dst_pitch);
    VEL_LAUNCH_CHECK("decimate4_kernel");
    return VEL_OK;
}
