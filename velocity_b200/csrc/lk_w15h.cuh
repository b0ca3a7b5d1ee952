// 15x15 "half-warp" kernel (included by lk_track.cu inside its anonymous namespace): the default path for the
// reference's lk_coarse window (utils/KLT.py:106) whenever every row pitch is a multiple of 4 bytes (always true for the
// pyramid levels; true for level 0 of ordinary frames and of ROI views cut from them).
//
// TWO points per warp: lanes 0-15 own one point, lanes 16-31 the next.  Within a half, lane = (rq, cg) owns the 4x4
// pixel block rows 4rq..4rq+3, columns 4cg..4cg+3 of the 16x16 bilinear footprint.
//   * pixels are fetched as aligned 32-bit words (2 per row) and re-aligned with one funnel shift per window;
//   * a bilinear sample is two DP2A instructions: the packed 16-bit weight pairs (w00,w01) / (w10,w11) against two
//     adjacent pixel bytes, with 256 - (I << 9) as the accumulator input: diff = dp2a(dp2a(I', top, W0), bottom, W1) >> 9;
//   * the template's Scharr tile comes from the lane's own 7 x 7 byte neighbourhood (DP4A horizontal taps, vertical
//     taps in registers, streamed row by row); no shared memory, no barriers besides the level-top __syncwarp();
//   * the per-point SCALAR program -- floor / weights / bounds / 2x2 solve / stopping rules, ~70% of a search
//     iteration's instructions when a whole warp serves one point -- is executed by both halves in one instruction
//     stream.  The halves are ordinary divergent SIMT code: each leaves its iteration loop when ITS point has converged
//     and waits at the __syncwarp() on top of the next pyramid level; every collective is a 16-lane xor-shuffle
//     reduction under the half's own member mask.
// Measured (B200, C2 workload): 21.6 us/pair vs 32.6 for the byte-gather kernel; variants tried and dropped: whole warp
// per point with REDUX sums (27.6), template patch in shared memory for 24 warps/SM (L1 thrashes: 28.0), staged search
// region in shared memory (27.3), next-level prefetch (+1.4 us).  Arithmetic and float32 operation order are those of
// the oracle: results are bit-identical (tests/test_klt_gpu.py runs both kernels against it).
constexpr int WH_WARPS = 4;                   // 8 points per CTA

// exact 16-lane sum (caller guarantees it fits int32)
__device__ __forceinline__ int half_sum(int v, unsigned hmask)
{
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(hmask, v, o);
    return v;
}

// Four exact 16-lane sums for the price of nine shuffles (instead of sixteen): the first two butterfly steps also
// TRANSPOSE -- after xor 8 a lane carries only two of the four quantities, after xor 4 only one -- so steps xor 2 and
// xor 1 reduce a single register; lane (b3, b2) of the half then owns the total of q[2*b3 + b2] and four indexed
// shuffles hand every total to every lane.  Wrap-around uint32 arithmetic: exact whenever each total fits 32 bits.
__device__ __forceinline__ void half_sum4(unsigned (&q)[4], unsigned hmask, int lane)
{
    const bool b3 = lane & 8, b2 = lane & 4;
    unsigned k0 = b3 ? q[2] : q[0], k1 = b3 ? q[3] : q[1];
    const unsigned s0 = b3 ? q[0] : q[2], s1 = b3 ? q[1] : q[3];
    k0 += __shfl_xor_sync(hmask, s0, 8);
    k1 += __shfl_xor_sync(hmask, s1, 8);
    unsigned k = b2 ? k1 : k0;
    const unsigned s = b2 ? k0 : k1;
    k += __shfl_xor_sync(hmask, s, 4);
    k += __shfl_xor_sync(hmask, k, 2);
    k += __shfl_xor_sync(hmask, k, 1);
    const int base = lane & 16;
    q[0] = __shfl_sync(hmask, k, base);
    q[1] = __shfl_sync(hmask, k, base + 4);
    q[2] = __shfl_sync(hmask, k, base + 8);
    q[3] = __shfl_sync(hmask, k, base + 12);
}

// float(hi * 65536 + lo) * 2^-20 with a single rounding: both limbs are exact in float32 (|hi| < 2^19, lo < 2^20) and
// one FMA rounds their exact sum once (same argument as warp_sum_scaled)
__device__ __forceinline__ float limbs_scaled(unsigned lo, unsigned hi)
{
    return __fmaf_rn((float)(int)hi, 0.0625f, __fmul_rn((float)lo, 1.f / 1048576.f));
}

// the lane's five J rows (tile rows 4rq .. 4rq+4) as byte windows: w0 = tile columns 4cg..4cg+3, w1 = 4cg+1..4cg+4
__device__ __forceinline__ void wh_gather(const Img& J, int inx, int iny, int rq, int cg, unsigned (&w0)[5], unsigned (&w1)[5])
{
    const bool inside = (unsigned)inx <= (unsigned)(J.w - 16) && (unsigned)iny <= (unsigned)(J.h - 16);
    const unsigned pitch = (unsigned)J.pitch;
    if (inside) {
        const unsigned off = (unsigned)(iny + 4 * rq) * pitch + (unsigned)(inx + 4 * cg);
        const unsigned mis = ((unsigned)(size_t)J.p + off) & 3u, sh = mis * 8u;
        const uint8_t* r = J.p + (int)(off - mis);              // may point up to 3 bytes before J.p (inside the parent allocation)
        // the second word of the last column group may lie wholly beyond column 15: do not touch it
        const bool skip_hi = (cg == 3) && (mis == 0);
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            const unsigned l = ldg_u32(r);
            unsigned h = l;
            if (!skip_hi) h = ldg_u32(r + 4);
            w0[k] = __funnelshift_r(l, h, sh);
            w1[k] = __funnelshift_rc(l, h, sh + 8u);
            if (k < 3 || (k == 3 && rq != 3)) r += pitch;          // tile row 16 does not exist: re-read row 15
        }
    } else {
        unsigned xo[5];
#pragma unroll
        for (int b = 0; b < 5; ++b) xo[b] = reflect_safe(inx + 4 * cg + b, J.w);
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            const unsigned yo = reflect_safe(iny + min(4 * rq + k, 15), J.h) * pitch;
            const unsigned v0 = ldg_u8(J.p + (yo + xo[0])), v1 = ldg_u8(J.p + (yo + xo[1])), v2 = ldg_u8(J.p + (yo + xo[2])),
                           v3 = ldg_u8(J.p + (yo + xo[3])), v4 = ldg_u8(J.p + (yo + xo[4]));
            const unsigned mid = v1 | (v2 << 8) | (v3 << 16);
            w0[k] = v0 | (mid << 8);
            w1[k] = mid | (v4 << 24);
        }
    }
}

// 96 registers -> five 4-warp CTAs (20 warps, 40 points) per SM: their 128-byte-line neighbourhoods still fit L1
//
// SEQ = true is the SEQUENCE form (vel_klt_sequence, vidExample.py:134-135): a point's track through frame k+1 depends only
// on its own position in frame k, never on another point, so the frame loop moves INSIDE the kernel -- each half-warp
// carries its point through all A.seq_pairs consecutive pairs (frame k -> k+1, forward + backward + gate), writes row
// k+1 of the track / alive arrays as it goes, and parks a failed track at the out-of-frame sentinel.  No launch and no
// grid-wide dependency per frame; 2-warp CTAs spread the 2048 warps of 4096 tracks evenly over the 148 SMs.
constexpr float kSeqDeadXY = -1.0e5f;     // same sentinel as sequence.cu

template <bool SEQ>
__global__ void __launch_bounds__(32 * WH_WARPS, 5)
lk_track_w15h_kernel(const LkArgs A)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int4 rP[12];   // template patch: [0..3] 256 - (I << 9), [4..7] Ix, [8..11] Iy of pixel rows 0..3 (x = column 4cg .. w = 4cg+3)
    const int slot = lane >> 4, hl = lane & 15;
    const int pt = blockIdx.x * (2 * (blockDim.x >> 5)) + 2 * warp + slot;
    if (pt - slot >= A.npts) return;                   // both points of this warp are beyond the set
    const bool in_set = pt < A.npts;
    const unsigned hmask = slot ? 0xffff0000u : 0x0000ffffu;
    const int cg = hl & 3, rq = hl >> 2;
    const float half = 7.0f;

    bool valid = in_set;                               // SEQ: "this track is still alive"
    float sx = 0.f, sy = 0.f;                          // SEQ: the point's position in the current frame
    if (SEQ && in_set) {
        valid = A.alive[pt] != 0;
        sx = __ldg(A.pts + 2ll * pt); sy = __ldg(A.pts + 2ll * pt + 1);
    }
    const int npairs_loop = SEQ ? A.seq_pairs : 1;
  for (int seqk = 0; seqk < npairs_loop; ++seqk) {
    const int pair = SEQ ? seqk : blockIdx.y;
    const uint8_t* P0 = A.prev0 + (long long)pair * A.prev_stride;
    const uint8_t* Pp = A.prev_pyr ? A.prev_pyr + (long long)pair * A.prev_pyr_stride : nullptr;
    const uint8_t* N0 = A.next0 + (long long)pair * A.next_stride;
    const uint8_t* Np = A.next_pyr ? A.next_pyr + (long long)pair * A.next_pyr_stride : nullptr;
    float px0 = sx, py0 = sy;
    if (!SEQ && valid) {
        const float* pin = A.pts + (long long)pair * A.pts_stride + 2ll * pt;
        px0 = __ldg(pin); py0 = __ldg(pin + 1);
    }

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
        const bool alive = pass ? (fst != 0) : valid;    // the backward pass cannot change a failed track

        int status = 1;
        float err = 0.f;
        float next_x = 0.f, next_y = 0.f;

        for (int level = A.lv.max_level; level >= 0; --level) {
            __syncwarp();                                 // the halves re-join here after their own iteration counts
            if (!alive) continue;
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
            int W0 = (int)__byte_perm((unsigned)w.w00, (unsigned)w.w01, 0x5410);
            int W1 = (int)__byte_perm((unsigned)w.w10, (unsigned)w.w11, 0x5410);

            // ---- template: the lane's 7 image rows (tile rows 4rq-1 .. 4rq+5) x 7 bytes (tile columns 4cg-1 .. 4cg+5),
            //      streamed row by row: horizontal Scharr taps -> vertical taps -> bilinear template pixels ----------
            int a11 = 0, a12 = 0, a22 = 0;
            {
                const bool interior = ipx >= 1 && ipy >= 1 && ipx + 16 < I.w && ipy + 16 < I.h;
                const unsigned pitch = (unsigned)I.pitch;
                // all 7 rows are requested before the first one is consumed (one exposed memory latency, not seven)
                unsigned lo7[7], hi7[7];
                if (interior) {
                    const unsigned off = (unsigned)(ipy + 4 * rq - 1) * pitch + (unsigned)(ipx + 4 * cg - 1);
                    const unsigned mis = ((unsigned)(size_t)I.p + off) & 3u, sh = mis * 8u;
                    const uint8_t* r = I.p + (int)(off - mis);
                    const bool skip_w2 = (cg == 3) && (mis < 3);     // never touch a word that lies wholly beyond tile column 16
                    unsigned L0[7], L1[7], L2[7];
#pragma unroll
                    for (int q = 0; q < 7; ++q) {
                        L0[q] = ldg_u32(r); L1[q] = ldg_u32(r + 4);
                        L2[q] = 0;
                        if (!skip_w2) L2[q] = ldg_u32(r + 8);
                        if (q < 5 || (q == 5 && rq != 3)) r += pitch;      // tile row 17 is never needed: re-read row 16
                    }
#pragma unroll
                    for (int q = 0; q < 7; ++q) {
                        lo7[q] = __funnelshift_r(L0[q], L1[q], sh);
                        hi7[q] = __funnelshift_r(L1[q], L2[q], sh);
                    }
                } else {
                    unsigned xo[7];
#pragma unroll
                    for (int b = 0; b < 7; ++b) xo[b] = reflect_safe(ipx + 4 * cg - 1 + b, I.w);
#pragma unroll
                    for (int q = 0; q < 7; ++q) {
                        const unsigned yo = reflect_safe(ipy + 4 * rq - 1 + q, I.h) * pitch;
                        lo7[q] = ldg_u8(I.p + (yo + xo[0])) | (ldg_u8(I.p + (yo + xo[1])) << 8) | (ldg_u8(I.p + (yo + xo[2])) << 16) |
                                 (ldg_u8(I.p + (yo + xo[3])) << 24);
                        hi7[q] = ldg_u8(I.p + (yo + xo[4])) | (ldg_u8(I.p + (yo + xo[5])) << 8) | (ldg_u8(I.p + (yo + xo[6])) << 16);
                    }
                }
                if (interior) {
                    // Interior windows: the bilinear sample and the Scharr pair are both linear in the image, so the
                    // interpolated derivative equals the Scharr pair of the interpolated image B = w00 I(y,x) + w01 I(y,x+1)
                    // + w10 I(y+1,x) + w11 I(y+1,x+1) EXACTLY (integers, before the >> 14) -- one DP2A grid instead of a
                    // DP4A derivative grid followed by four multiplies per derivative per pixel.  (Outside the frame the
                    // derivative image is zero-padded, which is not linear in I: those windows take the branch below.)
                    int Bt[6];                   // top-row contributions of the B row in flight
                    int Br[3][6];                // ring: the last three complete B rows (B row y <-> tile row 4rq - 1 + y)
                    int hdr[3][4], hsr[3][4];    // their horizontal difference / smoothing taps at the lane's 4 columns
#pragma unroll
                    for (int q = 0; q < 7; ++q) {
                        const unsigned lo = lo7[q], hi = hi7[q];
                        const unsigned v1 = __funnelshift_r(lo, hi, 8), v5 = hi >> 8;
                        if (q >= 1) {            // bottom row of B row y = q - 1
                            const int y = q - 1, c = y % 3;
                            Br[c][0] = dp2a_lo(W1, lo, Bt[0]); Br[c][1] = dp2a_lo(W1, v1, Bt[1]); Br[c][2] = dp2a_hi(W1, lo, Bt[2]);
                            Br[c][3] = dp2a_hi(W1, v1, Bt[3]); Br[c][4] = dp2a_lo(W1, hi, Bt[4]); Br[c][5] = dp2a_lo(W1, v5, Bt[5]);
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                hdr[c][j] = Br[c][j + 2] - Br[c][j];
                                hsr[c][j] = 3 * (Br[c][j] + Br[c][j + 2]) + 10 * Br[c][j + 1];
                            }
                            if (y >= 2) {        // pixel row rr = y - 2: B rows y-2 (above), y-1 (centre), y (below)
                                const int rr = y - 2, ca = (y - 2) % 3, cc = (y - 1) % 3;
                                int pI[4], pgx[4], pgy[4];
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    const bool active = (cg < 3 || j < 3) && (rq < 3 || rr < 3);
                                    int ix = (3 * (hdr[ca][j] + hdr[c][j]) + 10 * hdr[cc][j] + (1 << 13)) >> 14;
                                    int iy = (hsr[c][j] - hsr[ca][j] + (1 << 13)) >> 14;
                                    if (!active) { ix = 0; iy = 0; }
                                    const int ival = (Br[cc][j + 1] + (1 << 8)) >> 9;
                                    pI[j] = (1 << 8) - (ival << 9); pgx[j] = ix; pgy[j] = iy;
                                    a11 += ix * ix; a12 += ix * iy; a22 += iy * iy;
                                }
                                rP[rr] = make_int4(pI[0], pI[1], pI[2], pI[3]);
                                rP[4 + rr] = make_int4(pgx[0], pgx[1], pgx[2], pgx[3]);
                                rP[8 + rr] = make_int4(pgy[0], pgy[1], pgy[2], pgy[3]);
                            }
                        }
                        if (q < 6) {             // top row of B row y = q
                            Bt[0] = dp2a_lo(W0, lo, 0); Bt[1] = dp2a_lo(W0, v1, 0); Bt[2] = dp2a_hi(W0, lo, 0);
                            Bt[3] = dp2a_hi(W0, v1, 0); Bt[4] = dp2a_lo(W0, hi, 0); Bt[5] = dp2a_lo(W0, v5, 0);
                        }
                    }
                } else {
                    int hd[3][5], hs[3][5];      // rings: horizontal taps of the last three image rows
                    int gx[2][5], gy[2][5];      //        Scharr pair of the last two tile rows
                    unsigned win1[3], win2[3];   //        byte windows 1..4 / 2..5 of the last three image rows
    #pragma unroll
                    for (int q = 0; q < 7; ++q) {
                        const unsigned lo = lo7[q], hi = hi7[q];
                        const int c = q % 3;
                        win1[c] = __funnelshift_r(lo, hi, 8);
                        win2[c] = __funnelshift_r(lo, hi, 16);
                        const unsigned win3 = __funnelshift_r(lo, hi, 24);
                        const unsigned wins[5] = {lo, win1[c], win2[c], win3, hi};
    #pragma unroll
                        for (int t = 0; t < 5; ++t) {
                            hd[c][t] = dp4a_us(wins[t], 0x000100FF, 0);   // (-1, 0, +1, 0)
                            hs[c][t] = dp4a_us(wins[t], 0x00030A03, 0);   // ( 3,10,  3, 0)
                        }
                        if (q >= 2) {
                            const int y = q - 2;                       // tile row 4rq + y, centred on image row q - 1
                            const int d = y & 1, c0 = (q - 2) % 3, c1 = (q - 1) % 3;
    #pragma unroll
                            for (int t = 0; t < 5; ++t) {
                                gx[d][t] = 3 * (hd[c0][t] + hd[c][t]) + 10 * hd[c1][t];
                                gy[d][t] = hs[c][t] - hs[c0][t];
                            }
                            if (!interior) {   // the derivative image is padded with constant 0 outside the frame
                                const bool in_y = (unsigned)(ipy + 4 * rq + y) < (unsigned)I.h;
    #pragma unroll
                                for (int t = 0; t < 5; ++t) {
                                    if (!(in_y && (unsigned)(ipx + 4 * cg + t) < (unsigned)I.w)) { gx[d][t] = 0; gy[d][t] = 0; }
                                }
                            }
                            if (y >= 1) {
                                const int rr = y - 1;                  // pixel row: Scharr rows rr (d ^ 1) and rr + 1 (d); image rows q-2, q-1
                                int pI[4], pgx[4], pgy[4];
    #pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    const bool active = (cg < 3 || j < 3) && (rq < 3 || rr < 3);
                                    const unsigned top = (j & 1) ? win2[c0] : win1[c0], bot = (j & 1) ? win2[c1] : win1[c1];
                                    const int s = (j & 2) ? dp2a_hi(W1, bot, dp2a_hi(W0, top, 1 << 8)) : dp2a_lo(W1, bot, dp2a_lo(W0, top, 1 << 8));
                                    const int ival = s >> 9;
                                    int ix = (gx[d ^ 1][j] * w.w00 + gx[d ^ 1][j + 1] * w.w01 + gx[d][j] * w.w10 + gx[d][j + 1] * w.w11 + (1 << 13)) >> 14;
                                    int iy = (gy[d ^ 1][j] * w.w00 + gy[d ^ 1][j + 1] * w.w01 + gy[d][j] * w.w10 + gy[d][j + 1] * w.w11 + (1 << 13)) >> 14;
                                    if (!active) { ix = 0; iy = 0; }
                                    pI[j] = (1 << 8) - (ival << 9); pgx[j] = ix; pgy[j] = iy;
                                    a11 += ix * ix; a12 += ix * iy; a22 += iy * iy;
                                }
                                rP[rr] = make_int4(pI[0], pI[1], pI[2], pI[3]);
                                rP[4 + rr] = make_int4(pgx[0], pgx[1], pgx[2], pgx[3]);
                                rP[8 + rr] = make_int4(pgy[0], pgy[1], pgy[2], pgy[3]);
                            }
                        }
                    }
                }
            }
            // a11, a22 >= 0 and their totals stay below 2^32 (225 px x 4080^2): one unsigned word each; a12 needs two limbs
            unsigned qa[4] = {(unsigned)a11, (unsigned)a22, (unsigned)a12 & 0xffffu, (unsigned)(a12 >> 16)};
            half_sum4(qa, hmask, lane);
            const float A11 = fmul(__uint2float_rn(qa[0]), 1.f / 1048576.f), A22 = fmul(__uint2float_rn(qa[1]), 1.f / 1048576.f);
            const float A12 = limbs_scaled(qa[2], qa[3]);
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
                if ((unsigned)(inx + W15) >= (unsigned)(J.w + W15) || (unsigned)(iny + W15) >= (unsigned)(J.h + W15)) {
                    if (level == 0) status = 0;
                    break;
                }
                w = bilin_weights(fsub(qx, (float)inx), fsub(qy, (float)iny));
                W0 = (int)__byte_perm((unsigned)w.w00, (unsigned)w.w01, 0x5410);
                W1 = (int)__byte_perm((unsigned)w.w10, (unsigned)w.w11, 0x5410);
                unsigned w0[5], w1[5];
                wh_gather(J, inx, iny, rq, cg, w0, w1);
                int df[16];
#pragma unroll
                for (int rr = 0; rr < 4; ++rr) {
                    const unsigned t0 = w0[rr], t1 = w1[rr], b0 = w0[rr + 1], b1 = w1[rr + 1];
                    const int4 pI = rP[rr];
                    df[4 * rr + 0] = dp2a_lo(W1, b0, dp2a_lo(W0, t0, pI.x)) >> 9;
                    df[4 * rr + 1] = dp2a_lo(W1, b1, dp2a_lo(W0, t1, pI.y)) >> 9;
                    df[4 * rr + 2] = dp2a_hi(W1, b0, dp2a_hi(W0, t0, pI.z)) >> 9;
                    df[4 * rr + 3] = dp2a_hi(W1, b1, dp2a_hi(W0, t1, pI.w)) >> 9;
                }
                if (final_eval) {
                    int e = 0;
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const bool active = (cg < 3 || (i & 3) < 3) && (rq < 3 || i < 12);
                        e += active ? abs(df[i]) : 0;
                    }
                    e = half_sum(e, hmask);
                    err = __fdiv_rn((float)e, (float)(32 * W15 * W15));
                    break;
                }
                int sb1 = 0, sb2 = 0;
#pragma unroll
                for (int rr = 0; rr < 4; ++rr) {
                    const int4 g1 = rP[4 + rr], g2 = rP[8 + rr];
                    sb1 += df[4 * rr + 0] * g1.x + df[4 * rr + 1] * g1.y + df[4 * rr + 2] * g1.z + df[4 * rr + 3] * g1.w;
                    sb2 += df[4 * rr + 0] * g2.x + df[4 * rr + 1] * g2.y + df[4 * rr + 2] * g2.z + df[4 * rr + 3] * g2.w;
                }
                unsigned qb[4] = {(unsigned)sb1 & 0xffffu, (unsigned)(sb1 >> 16), (unsigned)sb2 & 0xffffu, (unsigned)(sb2 >> 16)};
                half_sum4(qb, hmask, lane);
                const float b1 = limbs_scaled(qb[0], qb[1]), b2 = limbs_scaled(qb[2], qb[3]);
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
            fx = next_x; fy = next_y; fst = valid ? status : 0; ferr = err; st = fst;
        } else if (alive) {
            bx = next_x; by = next_y;
            const float ddx = fsub(px0, bx), ddy = fsub(py0, by);
            const float fbe = __fsqrt_rn(fadd(fmul(ddx, ddx), fmul(ddy, ddy)));
            st = status && (fbe < A.fbt);
        }
    }
    if (!SEQ) {
        if (hl == 0 && valid) {
            const long long o = (long long)pair * A.npts + pt;
            A.out[2 * o] = fx;
            A.out[2 * o + 1] = fy;
            A.status[o] = (uint8_t)st;
            A.err[o] = fst ? ferr : 0.f;
            if (A.back) { A.back[2 * o] = bx; A.back[2 * o + 1] = by; }
        }
    } else {
        const bool live = valid && st != 0;            // alive[k+1] = alive[k] & status   (vg[vg] = v)
        if (hl == 0 && in_set) {
            const long long o = (long long)pair * A.npts + pt;
            A.out[2 * o] = live ? fx : kSeqDeadXY;
            A.out[2 * o + 1] = live ? fy : kSeqDeadXY;
            A.err[o] = (valid && fst) ? ferr : 0.f;
            A.alive[(long long)(pair + 1) * A.npts + pt] = live ? 1 : 0;
        }
        valid = live;
        sx = fx; sy = fy;
        if (!__any_sync(0xffffffffu, valid)) {         // both tracks of this warp are gone: park the remaining rows and leave
            if (hl == 0 && in_set) {
                for (int k2 = pair + 1; k2 < npairs_loop; ++k2) {
                    const long long o = (long long)k2 * A.npts + pt;
                    A.out[2 * o] = kSeqDeadXY; A.out[2 * o + 1] = kSeqDeadXY; A.err[o] = 0.f;
                    A.alive[(long long)(k2 + 1) * A.npts + pt] = 0;
                }
            }
            return;
        }
    }
  }
}
