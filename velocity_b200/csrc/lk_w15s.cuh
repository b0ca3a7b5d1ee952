// 15x15 SEQUENCE kernel (included by lk_track.cu after lk_w15h.cuh, whose helpers it shares): vel_klt_sequence's fast path.
//
// In a frame run, the template of the BACKWARD pass of pair j-1 (patch of frame j around the forward result p_j) and the
// template of the FORWARD pass of pair j (patch of frame j around p_j) are the same patch, with the same sub-pixel weights,
// at every pyramid level -- and so are the 2x2 gradient matrix and the minimum-eigenvalue test.  The batch kernel
// (lk_w15h.cuh) builds it twice because pairs are independent launches' worth of work there; here each half-warp carries its
// point through all frames and, per frame j and level, builds the template ONCE and runs the two searches that use it one
// after the other: against frame j-1 (backward pass of pair j-1 -> forward-backward verdict -> alive[j]) and against
// frame j+1 (forward pass of pair j -> tracks[j+1], err[j]).  The forward search runs before the verdict on pair j-1 is
// known; a failed verdict just discards it (the track is parked at the sentinel) -- surviving tracks are bit-identical to
// the per-pair kernels (tests/test_sequence_gpu.py).  Three template builds per pair instead of six: the SURVEY 8(d)
// sequence byte model ("each pyramid read once per role") made literal, and ~25 % fewer instructions per pair.
// The backward search skips the patch-error evaluation (cv2 computes it, the reference wrapper never reads it).

// Newton iterations of one pyramid level against image J (+ the final patch error when want_err); same arithmetic, same order
// as the loop in lk_track_w15h_kernel
// ---- TMA-staged search neighbourhood (VEL_LK_SEQ=tma; the A/B of north_star's "TMA-staged patches") -------------------------------
// Per level and role the half-warp's lane 0 requests ONE 48 x 32-byte box of the search image around the start position
// (cp.async.bulk.tensor, 3-D map x / row / frame per pyramid level) BEFORE the template is built, so the box lands while the
// template's ~400 instructions run; the Newton iterations then gather their 16 x 16 footprint from shared memory as long as it
// stays inside the box (start -8 .. +5 px at least) and fall back to the global gather otherwise (or when the footprint's
// 32 x 32 neighbourhood would leave the frame: the TMA unit zero-fills, LK needs REFLECT_101).
// MEASURED (tools/micro/tma_u8_box.cu): a tiled-mode box must START on a 16-byte boundary of global memory -- with one-byte
// elements the innermost coordinate has to be a multiple of 16, or the request faults ("illegal instruction").  Hence the box
// is 48 bytes wide and starts at the search origin rounded down to 16.
struct SeqTma {
    CUtensorMap map[3];
};
constexpr int TMA_BOX = 32;                        // rows of the box, and the neighbourhood (32 x 32) that must lie inside the frame
constexpr int TMA_BOXW = 48;                       // bytes per box row (16-byte aligned start + 32 + slack)

__device__ __forceinline__ void seq_mbar_init(unsigned bar, unsigned count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void seq_mbar_expect_tx(unsigned bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void seq_mbar_wait(unsigned bar, unsigned parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "SEQ_WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra SEQ_DONE_%=;\n\t"
        "bra SEQ_WAIT_%=;\n\t"
        "SEQ_DONE_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void seq_tma_load_3d(unsigned dst, const CUtensorMap* map, int c0, int c1, int c2, unsigned bar)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar) : "memory");
}

// the lane's five J rows from the staged box (same bytes, same word/funnel-shift form as wh_gather)
__device__ __forceinline__ void tile_gather(const uint8_t* tile, int dx, int dy, int rq, int cg, unsigned (&w0)[5], unsigned (&w1)[5])
{
    const unsigned off = (unsigned)(dy + 4 * rq) * TMA_BOXW + (unsigned)(dx + 4 * cg);
    const unsigned mis = off & 3u, sh = mis * 8u;
    const uint8_t* r = tile + (off - mis);
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        const unsigned l = *reinterpret_cast<const unsigned*>(r), h = *reinterpret_cast<const unsigned*>(r + 4);
        w0[k] = __funnelshift_r(l, h, sh);
        w1[k] = __funnelshift_rc(l, h, sh + 8u);
        if (k < 3 || (k == 3 && rq != 3)) r += TMA_BOXW;         // as wh_gather: tile row 16 is never needed
    }
}

__device__ __forceinline__ void w15_search(const Img& J, const int4 (&rP)[12], float A11, float A12, float A22, float D, int level, int max_count,
                                           float eps2, bool want_err, int rq, int cg, unsigned hmask, int lane, float& next_x, float& next_y,
                                           int& status, float& err, const uint8_t* tile = nullptr, int box_x = 0, int box_y = 0)
{
    const float half = 7.0f;
    float nx = fsub(next_x, half), ny = fsub(next_y, half);
    float pdx = 0.f, pdy = 0.f;
    bool final_eval = false;
    for (int j = 0;; ++j) {
        if (!final_eval && j >= max_count) {
            if (want_err && status) final_eval = true;
            else break;
        }
        const float qx = final_eval ? fsub(next_x, half) : nx, qy = final_eval ? fsub(next_y, half) : ny;
        const int inx = __float2int_rd(qx), iny = __float2int_rd(qy);
        if ((unsigned)(inx + W15) >= (unsigned)(J.w + W15) || (unsigned)(iny + W15) >= (unsigned)(J.h + W15)) {
            if (level == 0) status = 0;
            break;
        }
        const Weights w = bilin_weights(fsub(qx, (float)inx), fsub(qy, (float)iny));
        const int W0 = (int)__byte_perm((unsigned)w.w00, (unsigned)w.w01, 0x5410);
        const int W1 = (int)__byte_perm((unsigned)w.w10, (unsigned)w.w11, 0x5410);
        unsigned w0[5], w1[5];
        // staged box: rows iny .. iny+16, bytes inx .. inx+19 of the footprint (incl. the second word of the last column group)
        if (tile && inx >= box_x && iny >= box_y && inx + 20 <= box_x + TMA_BOXW && iny + 17 <= box_y + TMA_BOX)
            tile_gather(tile, inx - box_x, iny - box_y, rq, cg, w0, w1);
        else
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
        bool stop = fadd(fmul(dx, dx), fmul(dy, dy)) <= eps2;
        if (!stop && j > 0 && fabsf(fadd(dx, pdx)) < 0.01f && fabsf(fadd(dy, pdy)) < 0.01f) {
            next_x = fsub(next_x, fmul(dx, 0.5f));
            next_y = fsub(next_y, fmul(dy, 0.5f));
            stop = true;
        }
        pdx = dx; pdy = dy;
        if (stop) {
            if (want_err && status) final_eval = true;
            else break;
        }
    }
}

constexpr int WS_WARPS = 2;                        // 4 points per CTA: 1024 CTAs for 4096 tracks, 6.9 per SM

template <bool TMA>
__global__ void __launch_bounds__(32 * WS_WARPS, 10)
lk_seq_w15h_kernel(const LkArgs A, const __grid_constant__ SeqTma T)
{
    __shared__ __align__(128) uint8_t s_tile[TMA ? WS_WARPS : 1][2][2][TMA ? TMA_BOXW * TMA_BOX : 16];
    __shared__ __align__(8) unsigned long long s_bar[WS_WARPS][2][2];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int4 rP[12];
    const int slot = lane >> 4, hl = lane & 15;
    const int pt = blockIdx.x * (2 * WS_WARPS) + 2 * warp + slot;
    if (pt - slot >= A.npts) return;                   // both points of this warp are beyond the set
    const bool in_set = pt < A.npts;
    const unsigned hmask = slot ? 0xffff0000u : 0x0000ffffu;
    const int cg = hl & 3, rq = hl >> 2;
    const float half = 7.0f;
    const int npairs = A.seq_pairs;
    const bool use_fb = A.fbt >= 0.f;

    unsigned tph[2] = {0u, 0u};                        // TMA: phase of the role's barrier that the NEXT request completes
    bool tpend[2] = {false, false};                    //      a request is in flight (or landed) and has not been waited for
    if (TMA) {
        if (hl == 0) {
            seq_mbar_init((unsigned)__cvta_generic_to_shared(&s_bar[warp][slot][0]), 1);
            seq_mbar_init((unsigned)__cvta_generic_to_shared(&s_bar[warp][slot][1]), 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        }
        __syncwarp();
    }
    bool live = false;                                 // alive[j-1] verdict (alive[0] is given)
    float px = 0.f, py = 0.f;                          // the point in frame j
    float qx0 = 0.f, qy0 = 0.f;                        // the point in frame j-1 (start of pair j-1)
    int fst_prev = 1;                                  // forward status of pair j-1
    if (in_set) {
        live = A.alive[pt] != 0;
        px = __ldg(A.pts + 2ll * pt); py = __ldg(A.pts + 2ll * pt + 1);
    }

    for (int j = 0; j <= npairs; ++j) {
        const bool do_bwd = j >= 1 && use_fb, do_fwd = j < npairs;
        const bool tent = live && fst_prev != 0;       // alive up to the verdict that this step delivers
        const bool b_act = tent && do_bwd, f_act = tent && do_fwd;
        const uint8_t* I0 = A.prev0 + (long long)j * A.prev_stride;
        const uint8_t* Ipyr = A.prev_pyr ? A.prev_pyr + (long long)j * A.prev_pyr_stride : nullptr;
        float bnx = 0.f, bny = 0.f, fnx = 0.f, fny = 0.f, ferr = 0.f;
        int bst = 1, fst = 1;

        for (int level = A.lv.max_level; level >= 0; --level) {
            __syncwarp();                              // the halves re-join here after their own iteration counts
            if (!(b_act || f_act)) continue;
            Img I;
            I.w = A.lv.w[level];
            I.h = A.lv.h[level];
            if (level == 0) { I.p = I0; I.pitch = A.prev_pitch; }
            else { I.p = Ipyr + A.lv.off[level]; I.pitch = A.lv.pitch[level]; }

            const float scale = 1.f / (float)(1 << level);
            float prev_x = fmul(px, scale), prev_y = fmul(py, scale);
            if (level == A.lv.max_level) { bnx = prev_x; bny = prev_y; fnx = prev_x; fny = prev_y; }
            else { bnx = fmul(bnx, 2.f); bny = fmul(bny, 2.f); fnx = fmul(fnx, 2.f); fny = fmul(fny, 2.f); }

            prev_x = fsub(prev_x, half); prev_y = fsub(prev_y, half);
            const int ipx = __float2int_rd(prev_x), ipy = __float2int_rd(prev_y);
            if (ipx < -W15 || ipx >= I.w || ipy < -W15 || ipy >= I.h) {
                if (level == 0) { bst = 0; fst = 0; ferr = 0.f; }
                continue;
            }
            // TMA: request the search boxes of both roles now; they land while the template is built
            bool thave[2] = {false, false};
            int tbx[2] = {0, 0}, tby[2] = {0, 0};
            if (TMA && level < 3) {
#pragma unroll
                for (int role = 0; role < 2; ++role) {
                    if (!(role ? f_act : b_act)) continue;
                    const float sx0 = role ? fnx : bnx, sy0 = role ? fny : bny;
                    const int ox = __float2int_rd(fsub(sx0, half)) - 8, oy = __float2int_rd(fsub(sy0, half)) - 8;
                    if (ox < 0 || oy < 0 || ox + TMA_BOX > I.w || oy + TMA_BOX > I.h) continue;
                    const unsigned bar = (unsigned)__cvta_generic_to_shared(&s_bar[warp][slot][role]);
                    if (tpend[role]) { seq_mbar_wait(bar, tph[role] ^ 1u); tpend[role] = false; }   // an earlier box nobody consumed
                    __syncwarp(hmask);                      // every lane of the half is done reading the buffer's previous tenant
                    // (lanes 0 and 16 reach the request together with different boxes: the compiler serialises the uniform-datapath
                    // UTMALDG with an ELECT loop, nothing to do here)
                    if (hl == 0) {
                        seq_mbar_expect_tx(bar, TMA_BOXW * TMA_BOX);
                        const unsigned dst = (unsigned)__cvta_generic_to_shared(&s_tile[warp][slot][role][0]);
                        const int fz = role ? j + 1 : j - 1;
                        if (level == 0) seq_tma_load_3d(dst, &T.map[0], ox & ~15, oy, fz, bar);
                        else if (level == 1) seq_tma_load_3d(dst, &T.map[1], ox & ~15, oy, fz, bar);
                        else seq_tma_load_3d(dst, &T.map[2], ox & ~15, oy, fz, bar);
                    }
                    thave[role] = true; tbx[role] = ox & ~15; tby[role] = oy;
                    tpend[role] = true; tph[role] ^= 1u;
                }
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
                if (level == 0) { bst = 0; fst = 0; }
                continue;
            }
            D = __fdiv_rn(1.f, D);

            // the two searches that share this template: role 0 = backward pass of pair j-1 (against frame j-1),
            // role 1 = forward pass of pair j (against frame j+1).  One copy of the search code (instruction cache).
#pragma unroll 1
            for (int role = 0; role < 2; ++role) {
                if (!(role ? f_act : b_act)) continue;
                const long long jf = role ? (long long)(j + 1) : (long long)(j - 1);
                Img J;
                J.w = I.w; J.h = I.h;
                if (level == 0) { J.p = A.prev0 + jf * A.prev_stride; J.pitch = A.prev_pitch; }
                else { J.p = A.prev_pyr + jf * A.prev_pyr_stride + A.lv.off[level]; J.pitch = A.lv.pitch[level]; }
                float sx_ = role ? fnx : bnx, sy_ = role ? fny : bny, e_ = 0.f;
                int st_ = role ? fst : bst;
                const uint8_t* tile = nullptr;
                if (TMA && thave[role]) {
                    if (tpend[role]) {
                        seq_mbar_wait((unsigned)__cvta_generic_to_shared(&s_bar[warp][slot][role]), tph[role] ^ 1u);
                        tpend[role] = false;
                    }
                    tile = &s_tile[warp][slot][role][0];
                }
                w15_search(J, rP, A11, A12, A22, D, level, A.max_count, A.eps2, role == 1 && level == 0, rq, cg, hmask, lane, sx_, sy_, st_, e_,
                           tile, tbx[role], tby[role]);
                if (role) { fnx = sx_; fny = sy_; fst = st_; ferr = e_; }
                else { bnx = sx_; bny = sy_; bst = st_; }
            }
        }

        // ---- verdict on pair j-1: alive[j] = alive[j-1] & st_fwd & st_bwd & (||p_(j-1) - back|| < fbt)   (utils/KLT.py:47-50) ----
        if (j >= 1) {
            bool ok = tent;
            if (use_fb && tent) {
                const float ddx = fsub(qx0, bnx), ddy = fsub(qy0, bny);
                const float fbe = __fsqrt_rn(fadd(fmul(ddx, ddx), fmul(ddy, ddy)));
                ok = bst != 0 && (fbe < A.fbt);
            }
            if (hl == 0 && in_set) {
                A.alive[(long long)j * A.npts + pt] = ok ? 1 : 0;
                if (!ok) {                                   // row j was written as the forward result one step ago: park it
                    const long long o = (long long)(j - 1) * A.npts + pt;
                    A.out[2 * o] = kSeqDeadXY; A.out[2 * o + 1] = kSeqDeadXY;
                }
            }
            live = ok;
        }
        // ---- forward result of pair j: tracks[j+1], err[j] ---------------------------------------------------------------
        if (do_fwd) {
            if (hl == 0 && in_set) {
                const long long o = (long long)j * A.npts + pt;
                A.out[2 * o] = live ? fnx : kSeqDeadXY;
                A.out[2 * o + 1] = live ? fny : kSeqDeadXY;
                A.err[o] = (live && fst) ? ferr : 0.f;
            }
            qx0 = px; qy0 = py;
            px = fnx; py = fny;
            fst_prev = fst;
        }
        if (!__any_sync(0xffffffffu, live)) {               // both tracks of this warp are gone: park the remaining rows and leave
            if (TMA) {                                      // no box may still be on its way into this CTA's shared memory
#pragma unroll
                for (int role = 0; role < 2; ++role)
                    if (tpend[role]) { seq_mbar_wait((unsigned)__cvta_generic_to_shared(&s_bar[warp][slot][role]), tph[role] ^ 1u); tpend[role] = false; }
            }
            if (hl == 0 && in_set) {
                for (int k2 = j + 1; k2 <= npairs; ++k2) {
                    A.alive[(long long)k2 * A.npts + pt] = 0;
                    const long long o = (long long)(k2 - 1) * A.npts + pt;
                    A.out[2 * o] = kSeqDeadXY; A.out[2 * o + 1] = kSeqDeadXY;
                    if (k2 < npairs) A.err[(long long)k2 * A.npts + pt] = 0.f;
                }
            }
            return;
        }
    }
    if (TMA) {
#pragma unroll
        for (int role = 0; role < 2; ++role)
            if (tpend[role]) seq_mbar_wait((unsigned)__cvta_generic_to_shared(&s_bar[warp][slot][role]), tph[role] ^ 1u);
    }
}
