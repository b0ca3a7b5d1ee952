// K8 building blocks: the dense FP64 linear algebra behind `delta = inv(JJ^T + I) @ J(z - zhat)` (utils/NLS.py:236) once
// the point blocks are eliminated -- hand-written for sm_100a, no cuBLAS / cuSOLVER.
//
//   vel_syrk_lower_sub   S(lower) -= E E^T          the Schur product  W' W'^T  (6nc x 3nt x 6nc, 39.5 GFLOP at C3)
//   vel_spd_solve        S = L L^T in place, x = S^-1 b   blocked right-looking Cholesky with the forward substitution
//                                                    folded in (b rides along as an extra row), then the backward substitution
//
// SYRK: FP64 tensor-core MMA (mma.sync m8n8k4 f64, "DMMA"), 128x128 CTA tiles (16 warps of 32x32), operands staged in
// shared memory by cp.async through a 3-deep ring of 32-column k-tiles, row stride padded by 4 doubles so a fragment load (8 rows x 4 k) hits
// 32 distinct 8-byte slots.  The lower triangle has nb(nb+1)/2 tiles -- 120 for M = 1794 -- which do not fill 148 SMs, so
// the K dimension is split into SK chunks and a persistent grid walks the (chunk, tile) items chunk-major (the working set
// of a chunk, M x K/SK doubles, stays in L2; E streams from HBM once).  The SK partial products of a tile are applied to S
// in chunk order through a per-tile turnstile (chunk c waits for c-1): bit-reproducible, no atomics, no partial buffers.
// Tile height is adapted to M (BMe = ceil(M/nb) rounded up to 8), so M = 1794 runs 15 x 120-row blocks instead of 14 full
// ones plus a 2-row sliver, and warps are assigned to 32x32 sub-tiles with a skew that keeps the four SM sub-partitions
// equally loaded when the last sub-tiles are short.
//
// Cholesky: one cooperative persistent kernel, 64-column panels.  Per panel: every CTA factors the 64x64 diagonal block
// redundantly in shared memory (8-column register-blocked steps), CTAs share the rows below it for the triangular solve,
// grid barrier, then the trailing matrix is updated tile by tile with DMMA, grid barrier.  The right-hand side is row n of
// the matrix, so y = L^-1 b falls out of the same sweeps; the backward substitution runs in the same kernel.
#include <cooperative_groups.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace {

int sm_count()
{
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = kNumSMs;
    }
    return n;
}

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src)
{
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// mbarrier helpers (shared::cta): the k-tile ring of the SYRK is synchronised per stage, never CTA-wide
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
// arrive on the barrier when all cp.async of this thread issued so far have landed (does not change the expected count)
__device__ __forceinline__ void cp_async_mbar_arrive(unsigned long long* bar)
{
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity)
{
    const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(a), "r"(parity) : "memory");
}

__device__ __forceinline__ int ld_acquire(const int* p)
{
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(int* p, int v) { asm volatile("st.release.gpu.global.s32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory"); }

// ---- SYRK ---------------------------------------------------------------------------------------------------------------
constexpr int SY_BM = 128, SY_BK = 32, SY_STAGES = 3, SY_LDS = SY_BK + 4, SY_THREADS = 512;
constexpr int SY_STAGE_DOUBLES = 2 * SY_BM * SY_LDS;                       // A tile + B tile
constexpr size_t SY_SMEM = sizeof(double) * SY_STAGES * SY_STAGE_DOUBLES;   // 221,184 B

struct SyrkPlan {
    int nb, bme, ntiles, sk, ktiles, grid, tile0;
};

// blk_lo/blk_hi restrict the product to the tile rows [blk_lo, blk_hi) (a rank's share of the distributed form); hi < 0 = all
SyrkPlan syrk_plan(int m, int k, int blk_lo = 0, int blk_hi = -1)
{
    SyrkPlan p;
    p.nb = (m + SY_BM - 1) / SY_BM;
    p.bme = (((m + p.nb - 1) / p.nb) + 7) & ~7;
    if (blk_hi < 0 || blk_hi > p.nb) blk_hi = p.nb;
    if (blk_lo < 0) blk_lo = 0;
    if (blk_lo > blk_hi) blk_lo = blk_hi;
    p.tile0 = blk_lo * (blk_lo + 1) / 2;
    p.ntiles = blk_hi * (blk_hi + 1) / 2 - p.tile0;
    p.ktiles = (k + SY_BK - 1) / SY_BK;
    const int sms = sm_count();
    // Items (tile, K chunk) are handed out in chunk-major order -- the first grid-ful by position, the rest by a device-side work
    // queue (a CTA that drew cheap diagonal tiles simply takes more; VEL_SYRK_QUEUE=0 = static round-robin, the A/B).  The split
    // is chosen by the round model below: an item costs ~11 us of fill + turnstile + read-modify-write of its 128x128 block of S
    // besides its k-tiles (4.2 us each), measured by a sweep at M = 1794: SK = 6 -> 1.47 ms, SK = 16 -> 1.55 ms with the queue
    // (1.48 / 1.56 without).
    int best = 1;
    double best_cost = 1e30;
    for (int sk = 1; sk <= 16 && sk <= p.ktiles; ++sk) {
        const long long items = (long long)p.ntiles * sk;
        const double rounds = (double)((items + sms - 1) / sms);
        const double cost = rounds * ((p.ktiles + sk - 1) / sk) + 3.0 * rounds;     // k-tiles on the critical path + epilogue/prologue per item
        if (cost < best_cost - 1e-9) { best_cost = cost; best = sk; }
    }
    p.sk = best;
    if (const char* e = getenv("VEL_SYRK_SK")) {          // tuning / experiment override
        const int v = atoi(e);
        if (v >= 1 && v <= p.ktiles) p.sk = v;
    }
    const long long items = (long long)p.ntiles * p.sk;
    p.grid = (int)(items < sms ? (items > 0 ? items : 1) : sms);
    return p;
}

// one k-tile (SY_BK columns) of a warp's 32x32 sub-tile: MT x NT m8n8k4 products per 4 columns
template <int MT, int NT>
__device__ __forceinline__ void syrk_ktile(double (&acc)[4][4][2], const double* pa, const double* pb)
{
#pragma unroll
    for (int kk = 0; kk < SY_BK / 4; ++kk) {
        double a[MT], b[NT];
#pragma unroll
        for (int i = 0; i < MT; ++i) a[i] = pa[i * 8 * SY_LDS + kk * 4];
#pragma unroll
        for (int j = 0; j < NT; ++j) b[j] = pb[j * 8 * SY_LDS + kk * 4];
#pragma unroll
        for (int i = 0; i < MT; ++i)
#pragma unroll
            for (int j = 0; j < NT; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
}

template <int MT>
__device__ __forceinline__ void syrk_ktile_nt(double (&acc)[4][4][2], const double* pa, const double* pb, int nt_cnt)
{
    if (nt_cnt == 4) syrk_ktile<MT, 4>(acc, pa, pb);
    else if (nt_cnt == 3) syrk_ktile<MT, 3>(acc, pa, pb);
    else if (nt_cnt == 2) syrk_ktile<MT, 2>(acc, pa, pb);
    else if (nt_cnt == 1) syrk_ktile<MT, 1>(acc, pa, pb);
}

// E [m][ld] row-major, K contiguous (ld even, rows 16-byte aligned, columns k..ld-1 of the last k-tile readable and ZERO)
__global__ void __launch_bounds__(SY_THREADS, 1)
dsyrk_lower_sub_kernel(const double* __restrict__ E, long long ld, int m, int nb, int bme, int ntiles, int sk, int ktiles,
                       double* __restrict__ S, long long lds, int* __restrict__ flags, int tile0, const int* __restrict__ gate,
                       int* __restrict__ queue)
{
    if (gate && *gate) return;                                   // device-side loop control (vel_ba_iterate): the whole grid leaves
    extern __shared__ __align__(16) double sy_smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t4 = lane & 3;
    const int wr = warp >> 2, wc = ((warp & 3) - wr) & 3;     // (wr + wc) % 4 == warp % 4: balances short edge sub-tiles over the sub-partitions
    const int mt_cnt = max(0, min(4, (bme - wr * 32 + 7) >> 3));
    const int nt_cnt = max(0, min(4, (bme - wc * 32 + 7) >> 3));
    const long long nitems = (long long)ntiles * sk;
    const bool vec_s = (lds & 1) == 0 && ((size_t)S & 15) == 0;   // 16-byte read-modify-write of S
    __shared__ unsigned long long full_bar[SY_STAGES], empty_bar[SY_STAGES];
    if (tid == 0) {
        for (int s2 = 0; s2 < SY_STAGES; ++s2) { mbar_init(&full_bar[s2], SY_THREADS); mbar_init(&empty_bar[s2], SY_THREADS / 32); }
    }
    __syncthreads();
    // ---- work items and the k-tile stream -----------------------------------------------------------------------------------
    // A CTA consumes a STREAM of k-tiles that runs across item boundaries: the producer cursor stays SY_STAGES-1 tiles ahead of
    // the consumer and moves on into the NEXT item (already drawn from the queue) while the current one is still being
    // multiplied, so the pipeline never drains: the first tiles of an item land during the epilogue of the one before.
    struct Item {
        int row0, col0, kt0, kt1, t, c;
        bool diag, valid;
    };
    auto decode = [&](long long item) -> Item {
        Item it;
        it.valid = item < nitems;
        const long long ii = it.valid ? item : 0;
        it.c = (int)(ii / ntiles); it.t = (int)(ii % ntiles);
        const int tg = it.t + tile0;                            // position in the triangular enumeration of ALL tiles
        int bi = (int)((sqrtf(8.f * (float)tg + 1.f) - 1.f) * 0.5f);
        while ((bi + 1) * (bi + 2) / 2 <= tg) ++bi;
        while (bi * (bi + 1) / 2 > tg) --bi;
        const int bj = tg - bi * (bi + 1) / 2;
        it.diag = bi == bj;
        it.row0 = bi * bme; it.col0 = bj * bme;
        it.kt0 = (int)((long long)ktiles * it.c / sk); it.kt1 = (int)((long long)ktiles * (it.c + 1) / sk);
        return it;
    };
    __shared__ long long s_item;
    auto draw = [&]() -> long long {                            // CTA-wide: the next item index (queue, or grid stride)
        if (tid == 0) s_item = (long long)gridDim.x + (long long)atomicAdd(queue, 1);
        __syncthreads();
        const long long v = s_item;
        __syncthreads();
        return v;
    };
    long long stride_next = (long long)blockIdx.x + gridDim.x;  // static form (queue == nullptr)
    Item cur = decode(blockIdx.x);
    Item nxt = decode(queue ? draw() : stride_next);
    stride_next += gridDim.x;

    constexpr int CPR = SY_BK / 2;                              // 16-byte chunks per tile row
    constexpr int NLOAD = SY_BM * CPR / SY_THREADS;             // chunks per thread and operand
    // this thread's chunks of a tile: row r_u, 16-byte chunk ch_u (fixed); the global row is clamped to the matrix per item
    auto issue_tile = [&](const Item& it, int kt, long long gt) {
        const int slot = (int)(gt % SY_STAGES);
        const unsigned use = (unsigned)(gt / SY_STAGES);
        mbar_wait(&empty_bar[slot], (use & 1u) ^ 1u);           // everybody finished the previous tenant of the slot (free at first use)
        double* sA = sy_smem + slot * SY_STAGE_DOUBLES;
        double* sB = sA + SY_BM * SY_LDS;
        const double* Ek = E + (long long)kt * SY_BK;
#pragma unroll
        for (int u = 0; u < NLOAD; ++u) {
            const int q = tid + SY_THREADS * u, r = q / CPR, ch = q % CPR;
            cp_async16(sA + r * SY_LDS + ch * 2, Ek + (long long)min(it.row0 + r, m - 1) * ld + ch * 2);
            if (!it.diag) cp_async16(sB + r * SY_LDS + ch * 2, Ek + (long long)min(it.col0 + r, m - 1) * ld + ch * 2);
        }
        cp_async_mbar_arrive(&full_bar[slot]);
    };
    // producer cursor over the stream: tiles of `cur`, then tiles of `nxt`
    int p_next = 0, p_kt = cur.kt0;                             // p_next = 1: the cursor is inside nxt
    long long p_gt = 0, c_gt = 0;                               // tiles issued / consumed by this CTA over its whole life
    auto produce_one = [&]() {
        if (!p_next && p_kt >= cur.kt1) { p_next = 1; p_kt = nxt.kt0; }
        if (p_next && (!nxt.valid || p_kt >= nxt.kt1)) return;  // never beyond the next item
        issue_tile(p_next ? nxt : cur, p_kt, p_gt);
        ++p_kt; ++p_gt;
    };
    if (cur.valid) {
#pragma unroll
        for (int s2 = 0; s2 < SY_STAGES - 1; ++s2) produce_one();
    }

    while (cur.valid) {
        const bool diag = cur.diag;
        const int row0 = cur.row0, col0 = cur.col0, t = cur.t, c = cur.c;
        const bool skip_warp = diag && wc > wr;                 // sub-tile strictly above the diagonal: never stored

        double acc[4][4][2];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

        for (int kt = cur.kt0; kt < cur.kt1; ++kt) {
            const long long gt = c_gt++;
            const int slot = (int)(gt % SY_STAGES);
            mbar_wait(&full_bar[slot], (unsigned)(gt / SY_STAGES) & 1u);
            if (!skip_warp) {
                const double* sA = sy_smem + slot * SY_STAGE_DOUBLES;
                const double* sB = diag ? sA : sA + SY_BM * SY_LDS;
                const double* pa = sA + (wr * 32 + g) * SY_LDS + t4;
                const double* pb = sB + (wc * 32 + g) * SY_LDS + t4;
                // a predicated mma.sync costs a WARPSYNC + NOP each (ncu: 1 per DMMA): dispatch once per k-tile on the
                // warp-uniform sub-tile counts instead, so that every DMMA in the hot paths is unconditional
                if (mt_cnt == 4) syrk_ktile_nt<4>(acc, pa, pb, nt_cnt);
                else if (mt_cnt == 3) syrk_ktile_nt<3>(acc, pa, pb, nt_cnt);
                else if (mt_cnt == 2) syrk_ktile_nt<2>(acc, pa, pb, nt_cnt);
                else if (mt_cnt == 1) syrk_ktile_nt<1>(acc, pa, pb, nt_cnt);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[slot]);
            // refill: the tile SY_STAGES-1 ahead in the stream (possibly the next item's); every warp has had this whole k-tile to
            // leave the slot it goes into, so the wait inside rarely blocks
            produce_one();
        }

        // turnstile: the partial products of a tile are subtracted from S in chunk order.  Thread 0 waits for its turn and draws the
        // item after next in the same breath, so that one barrier publishes both.
        if (tid == 0) {
            if (c > 0) {
                while (ld_acquire(flags + t) < c) __nanosleep(64);
            }
            if (queue) s_item = (long long)gridDim.x + (long long)atomicAdd(queue, 1);
        }
        __syncthreads();
        const long long drawn = queue ? s_item : 0;
        if (!skip_warp) {
            const int rlim = min(row0 + bme, m), clim = min(col0 + bme, m);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int r = row0 + wr * 32 + i * 8 + g;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int cc = col0 + wc * 32 + j * 8 + 2 * t4;
                    if (i < mt_cnt && j < nt_cnt && r < rlim) {
                        double* p = S + (long long)r * lds + cc;
                        const bool ok0 = cc < clim && (!diag || cc <= r), ok1 = cc + 1 < clim && (!diag || cc + 1 <= r);   // strictly-upper entries stay untouched
                        if (vec_s && ok0 && ok1) {
                            double2 v = __ldcg(reinterpret_cast<const double2*>(p));
                            v.x -= acc[i][j][0]; v.y -= acc[i][j][1];
                            *reinterpret_cast<double2*>(p) = v;
                        } else {
                            if (ok0) p[0] = __ldcg(p) - acc[i][j][0];
                            if (ok1) p[1] = __ldcg(p + 1) - acc[i][j][1];
                        }
                    }
                }
            }
        }
        // release the tile for the next chunk; this barrier also keeps s_item stable until every thread has read it
        __threadfence();
        __syncthreads();
        if (sk > 1 && tid == 0) st_release(flags + t, c + 1);
        // advance: the next item becomes the current one, the drawn item the next (dynamic: a CTA that drew cheap diagonal tiles
        // or finished early simply takes more)
        cur = nxt;
        if (p_next) p_next = 0;                                 // the cursor was already inside it: p_kt stays
        else p_kt = cur.kt0;                                    // (only when the old item had fewer tiles than the look-ahead)
        long long ni;
        if (queue) ni = drawn;
        else { ni = stride_next; stride_next += gridDim.x; }
        nxt = decode(cur.valid ? ni : nitems);
    }
}

// ---- Cholesky + solve -----------------------------------------------------------------------------------------------------
constexpr int CH_NB = 64, CH_NBO = 256, CH_THREADS = 512, CH_LD = CH_NB + 1, CH_LDT = CH_NB + 4;
// diagonal block (stride 65: conflict-free column walks) + two DMMA operand tiles (stride 68: conflict-free fragment loads)
constexpr size_t CH_SMEM = sizeof(double) * (CH_NB * CH_LD + 2 * CH_NB * CH_LDT);      // 102,912 B
constexpr int CH_MAXOWN = 2048;                                                        // own-tile list of the task-graph form
constexpr size_t CH_DAG_SMEM = sizeof(double) * (CH_NB * CH_LD + 4 * CH_NB * CH_LDT);  // + the panel's inverse diagonal tile + the resident sub-diagonal tile

// reciprocal to full double precision from the 20-bit hardware seed and two Newton steps: ~170 cycles of dependent latency
// against ~250 for the library's rsqrt / divide (FP64 instructions have ~38 cycles of latency on this part, so the per-column
// pivot chain of a Cholesky factorisation is what bounds it)
__device__ __forceinline__ double fast_rcp(double x)
{
    // one third-order step from the 20-bit seed: y0 (1 + e + e^2), e = 1 - x y0, |e| <= 2^-20 -> relative error e^3 = 2^-60: three
    // dependent FMAs on the pivot chain instead of the four of two Newton steps
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double e = fma(-x, y, 1.0);
    const double t = fma(e, e, e);
    return fma(y, t, y);
}

#ifdef VEL_CHOL_TIMING
__device__ unsigned long long g_cb_t[4];
#define CB_T(k) do { if (blockIdx.x == 0 && threadIdx.x == 0) { long long now_ = clock64(); g_cb_t[k] += now_ - cb_last; cb_last = now_; } } while (0)
#else
#define CB_T(k)
#endif

// row block of the t-th tile of a lower-triangular tile enumeration (t = ta (ta + 1) / 2 + tb, tb <= ta), t < 28
__constant__ unsigned char kTriRow[28] = {0, 1, 1, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 4, 5, 5, 5, 5, 5, 5, 6, 6, 6, 6, 6, 6, 6};

struct CholScratch {
    double T[CH_NB][8];      // the current 8-column sub-panel scaled by the reciprocal pivots (rows of the block)
    double d[CH_NB];         // pivots of the square-root-free factorisation
};

// 64x64 (w x w) lower Cholesky of sD in place, all CH_THREADS threads of the CTA, organised around the one thing that bounds
// it -- the pivot-to-pivot dependency chain:
//   * square-root free inside (A = M D^-1 M^T, M = unscaled columns): a column costs one reciprocal + one multiply + one FMA on
//     the chain; the 64 square roots are taken at the end, in parallel (L = M D^-1/2);
//   * 8-column steps; every thread that owns a row below the sub-block (and the warp that writes it back) factors the 8x8 diagonal
//     sub-block redundantly in registers (no shuffles, no barrier between that and the forward substitution of the thread's own
//     row); the other warps skip the step -- sixteen warps repeating it made the phase FP64-throughput-bound (0.86 -> 0.76 ms);
//   * the rank-8 update of the rest of the block runs on the tensor cores (DMMA).
// ok is cleared when a pivot is not positive; rdiag receives 1 / L_cc.
__device__ void chol_block(double (*sD)[CH_LD], int w, int* ok, double* rdiag, CholScratch* cs)
{
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t4 = lane & 3;
#ifdef VEL_CHOL_TIMING
    long long cb_last = clock64();
#endif
    for (int jb = 0; jb < w; jb += 8) {
        const int wb = min(8, w - jb);
        CB_T(3);
        const int below = w - jb - wb;
        if (warp * 32 < below || warp == 14) {       // the other warps have no row below the sub-block and do not write it back: no redundant FP64 work on the shared pipe
        // ---- (i) the 8x8 diagonal sub-block, redundantly in every PARTICIPATING thread: a[u(u+1)/2 + c], c <= u --------------------
        double a[36], rp[8];
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int c = 0; c <= u; ++c)
                a[u * (u + 1) / 2 + c] = (u < wb) ? sD[jb + u][jb + c] : (u == c ? 1.0 : 0.0);     // identity padding past the block
        bool bad = false;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const double piv = a[c * (c + 1) / 2 + c];
            bad = bad || !(piv > 0.0);
            rp[c] = fast_rcp(piv);
            double m[8];
#pragma unroll
            for (int u = c + 1; u < 8; ++u) {
                m[u] = a[u * (u + 1) / 2 + c];
                a[u * (u + 1) / 2 + c] = m[u] * rp[c];                       // T[u][c] = M[u][c] / d_c
            }
#pragma unroll
            for (int u = c + 1; u < 8; ++u)
#pragma unroll
                for (int u2 = c + 1; u2 <= u; ++u2) a[u * (u + 1) / 2 + u2] -= m[u] * a[u2 * (u2 + 1) / 2 + c];
        }
        if (bad && tid == 448) *ok = 0;             // warp 14 always takes part (warp 0 only while rows remain below the sub-block)
        // ---- (ii) the thread's own row below the sub-block: M[r][q] = a[r][q] - sum_{p<q} M[r][p] T[q][p] ------------------------
        if (tid < below) {
            const int r = jb + wb + tid;
            double x[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) x[q] = q < wb ? sD[r][jb + q] : 0.0;
#pragma unroll
            for (int q = 0; q < 8; ++q)
#pragma unroll
                for (int u = q + 1; u < 8; ++u) x[u] -= x[q] * a[u * (u + 1) / 2 + q];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                if (q < wb) sD[r][jb + q] = x[q];
                cs->T[r][q] = q < wb ? x[q] * rp[q] : 0.0;
            }
        } else if (tid >= 448 && tid < 448 + 8) {
            // one thread per row of the sub-block writes it back as M (unscaled: T * d) and publishes the pivots
            const int u = tid - 448;
            if (u < wb) {
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    if (c < u) {
                        double tv = 0.0, dv = 1.0;
#pragma unroll
                        for (int uu = 0; uu < 8; ++uu)
                            if (uu == u) { tv = a[uu * (uu + 1) / 2 + c]; }
                        dv = a[c * (c + 1) / 2 + c];
                        sD[jb + u][jb + c] = tv * dv;
                    }
                }
                double du = 0.0;
#pragma unroll
                for (int uu = 0; uu < 8; ++uu)
                    if (uu == u) du = a[uu * (uu + 1) / 2 + uu];
                cs->d[jb + u] = du;
            }
        }
        }
        __syncthreads();
        CB_T(0);
        // ---- (iii) rank-8 update of the rest of the block on the tensor cores: a[r][r'] -= sum_q M[r][q] T[r'][q] ----------------
        if (below > 0) {
            // at most 28 lower 8x8 tiles (below <= 56): a warp takes tile `warp` and tile `warp + 16`, both in flight at once (their
            // loads, two DMMAs and read-modify-write are independent chains); tile -> (row block, column block) from a table
            const int R0 = jb + wb, nt8 = (below + 7) >> 3, ntile = nt8 * (nt8 + 1) / 2;
            const int tt0 = warp, tt1 = warp + CH_THREADS / 32;
            const bool on0 = tt0 < ntile, on1 = tt1 < ntile;
            const int ta0 = kTriRow[tt0 < 28 ? tt0 : 0], tb0 = tt0 - ta0 * (ta0 + 1) / 2;
            const int ta1 = kTriRow[tt1 < 28 ? tt1 : 0], tb1 = tt1 - ta1 * (ta1 + 1) / 2;
            const int ra0 = R0 + ta0 * 8 + g, rb0 = R0 + tb0 * 8 + g, ra1 = R0 + ta1 * 8 + g, rb1 = R0 + tb1 * 8 + g;
            double av0[2], bv0[2], av1[2], bv1[2];
#pragma unroll
            for (int kk = 0; kk < 2; ++kk) {
                const int q = kk * 4 + t4;
                av0[kk] = (on0 && ra0 < w && q < wb) ? sD[ra0][jb + q] : 0.0;
                bv0[kk] = (on0 && rb0 < w) ? cs->T[rb0][q] : 0.0;
                av1[kk] = (on1 && ra1 < w && q < wb) ? sD[ra1][jb + q] : 0.0;
                bv1[kk] = (on1 && rb1 < w) ? cs->T[rb1][q] : 0.0;
            }
            double d00 = 0.0, d01 = 0.0, d10 = 0.0, d11 = 0.0;
            if (on0) { dmma884(d00, d01, av0[0], bv0[0]); }
            if (on1) { dmma884(d10, d11, av1[0], bv1[0]); }
            if (on0) { dmma884(d00, d01, av0[1], bv0[1]); }
            if (on1) { dmma884(d10, d11, av1[1], bv1[1]); }
            const int cc0 = R0 + tb0 * 8 + 2 * t4, cc1 = R0 + tb1 * 8 + 2 * t4;
            if (on0 && ra0 < w) {
                if (cc0 < w && cc0 <= ra0) sD[ra0][cc0] -= d00;
                if (cc0 + 1 < w && cc0 + 1 <= ra0) sD[ra0][cc0 + 1] -= d01;
            }
            if (on1 && ra1 < w) {
                if (cc1 < w && cc1 <= ra1) sD[ra1][cc1] -= d10;
                if (cc1 + 1 < w && cc1 + 1 <= ra1) sD[ra1][cc1 + 1] -= d11;
            }
        }
        __syncthreads();
        CB_T(2);
    }
    // ---- L = M D^-1/2: the square roots, all at once ------------------------------------------------------------------------------
    if (tid < w) {
        const double dv = cs->d[tid];
        const double rs = rsqrt(dv);
        rdiag[tid] = rs;                       // 1 / L_cc
        cs->d[tid] = dv * rs;                  // L_cc
    }
    __syncthreads();
    for (int e = tid; e < CH_NB * CH_NB; e += CH_THREADS) {
        const int r = e >> 6, c = e & 63;
        if (r < w && c < r) sD[r][c] *= rdiag[c];
        else if (r < w && c == r) sD[r][c] = cs->d[c];
    }
    __syncthreads();
}

#ifdef VEL_CHOL_TIMING
__device__ unsigned long long g_chol_t[8];
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#define CH_T(k)  do { if (blockIdx.x == 0 && threadIdx.x == 0) { unsigned long long now_ = gtime(); g_chol_t[k] += now_ - t_last; t_last = now_; } } while (0)
#else
#define CH_T(k)
#endif

// 64 x kw slab S[row0 .. row0+64)[col0 .. col0+kw) (kw <= 64) -> dst[64][CH_LDT], rows >= nrows and columns >= kw zero-filled
__device__ __forceinline__ void chol_load_tile(const double* S, long long lds, int row0, int nrows, int col0, int kw, double* dst, bool vec)
{
    const int tid = threadIdx.x;
    if (vec) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int q = tid + CH_THREADS * u, r = q >> 5, c2 = (q & 31) * 2;
            double2 v = make_double2(0.0, 0.0);
            if (r < nrows && c2 < kw) {
                const double* src = S + (long long)(row0 + r) * lds + col0 + c2;
                if (c2 + 1 < kw) v = __ldcg(reinterpret_cast<const double2*>(src));
                else v.x = __ldcg(src);
            }
            *reinterpret_cast<double2*>(dst + r * CH_LDT + c2) = v;
        }
    } else {
        for (int e = tid; e < CH_NB * CH_NB; e += CH_THREADS) {
            const int r = e >> 6, c2 = e & 63;
            dst[r * CH_LDT + c2] = (r < nrows && c2 < kw) ? __ldcg(S + (long long)(row0 + r) * lds + col0 + c2) : 0.0;
        }
    }
}

// One 64x64 tile of the trailing update:  A[bi][bj] -= sum over the K columns [kcol0, kcol0 + kw) of  P_bi P_bj^T
// (bi < 0: the right-hand-side strip b[bj block] -= y P_bj^T with y = b[kcol0 .. kcol0+kw)).  All threads of the CTA.
__device__ void chol_update_tile(double* S, long long lds, int n, double* b, int bi, int bj, int kcol0, int kw, double* bufA, double* bufB,
                                 bool vec)
{
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t4 = lane & 3;
    const int c0 = bj * CH_NB, wj = min(CH_NB, n - c0);
    const int r0 = bi * CH_NB, wi = bi < 0 ? 0 : min(CH_NB, n - r0);
    const int sr = (warp >> 2) * 16, sc = (warp & 3) * 16;
    double acc[2][2][2] = {};
    double strip = 0.0;
    for (int kc = 0; kc < kw; kc += CH_NB) {
        const int kww = min(CH_NB, kw - kc);
        __syncthreads();                                           // the previous user of the buffers is done
        chol_load_tile(S, lds, c0, wj, kcol0 + kc, kww, bufB, vec);
        if (bi < 0) {
            if (tid < kww) bufA[tid] = __ldcg(b + kcol0 + kc + tid);
            __syncthreads();
            if (tid < wj) {
                double s2 = 0.0;
                for (int c2 = 0; c2 < kww; ++c2) s2 += bufA[c2] * bufB[tid * CH_LDT + c2];
                strip += s2;
            }
            continue;
        }
        chol_load_tile(S, lds, r0, wi, kcol0 + kc, kww, bufA, vec);
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < CH_NB / 4; ++kk) {
            double a[2], bb[2];
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                a[i] = bufA[(sr + i * 8 + g) * CH_LDT + kk * 4 + t4];
                bb[i] = bufB[(sc + i * 8 + g) * CH_LDT + kk * 4 + t4];
            }
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < 2; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], bb[j]);
        }
    }
    if (bi < 0) {
        if (tid < wj) b[c0 + tid] = __ldcg(b + c0 + tid) - strip;
        return;
    }
    // read-modify-write of the tile: all loads first, then the stores
    double2 cur[2][2];
    bool m0[2][2], m1[2][2];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int r = sr + i * 8 + g, c2 = sc + j * 8 + 2 * t4;
            m0[i][j] = r < wi && c2 < wj && (bi > bj || c2 <= r);
            m1[i][j] = r < wi && c2 + 1 < wj && (bi > bj || c2 + 1 <= r);
            const double* p = S + (long long)(r0 + r) * lds + c0 + c2;
            cur[i][j] = make_double2(0.0, 0.0);
            if (vec && m0[i][j] && m1[i][j]) cur[i][j] = __ldcg(reinterpret_cast<const double2*>(p));
            else {
                if (m0[i][j]) cur[i][j].x = __ldcg(p);
                if (m1[i][j]) cur[i][j].y = __ldcg(p + 1);
            }
        }
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int r = sr + i * 8 + g, c2 = sc + j * 8 + 2 * t4;
            double* p = S + (long long)(r0 + r) * lds + c0 + c2;
            const double2 v = make_double2(cur[i][j].x - acc[i][j][0], cur[i][j].y - acc[i][j][1]);
            if (vec && m0[i][j] && m1[i][j]) *reinterpret_cast<double2*>(p) = v;
            else {
                if (m0[i][j]) p[0] = v.x;
                if (m1[i][j]) p[1] = v.y;
            }
        }
}

// trailing update of the block columns [jlo, jhi) (rows from the diagonal down, plus the right-hand-side strip) with the K
// columns [kcol0, kcol0 + kw): the tiles are dealt round-robin to the CTAs of the grid
__device__ void chol_update_columns(double* S, long long lds, int n, double* b, int nblk, int jlo, int jhi, int kcol0, int kw,
                                    double* bufA, double* bufB, bool vec)
{
    int t = blockIdx.x;
    for (int j = jlo; j < jhi; ++j) {
        const int cnt = nblk - j + 1;                  // tiles (j..nblk-1, j) and the strip
        while (t < cnt) {
            chol_update_tile(S, lds, n, b, t == cnt - 1 ? -1 : j + t, j, kcol0, kw, bufA, bufB, vec);
            t += gridDim.x;
        }
        t -= cnt;
    }
}

// A is the (n+1) x n "tall" matrix: rows 0..n-1 = S (lower triangle used), row n = b (stored separately).
// On exit: S holds L (lower), b holds x = S^-1 b.  info[0] = 0 on success, 1 when a pivot was not positive.
//
// Two-level blocking: 64-column panels inside 256-column outer blocks.  After a panel is factored only the REST OF ITS OUTER
// BLOCK is updated with it (at most 3 block columns: one tile per CTA); the matrix to the right of the outer block is updated
// once per outer block with all of its (up to 256) columns -- 4x fewer read-modify-write sweeps over S, 4x the work per sweep.
__global__ void __launch_bounds__(CH_THREADS, 1)
chol_solve_kernel(double* S, long long lds, int n, double* b, int* __restrict__ info)
{
    cg::grid_group grid = cg::this_grid();
    extern __shared__ __align__(16) double ch_smem[];
    double (*sD)[CH_LD] = reinterpret_cast<double (*)[CH_LD]>(ch_smem);        // diagonal block
    double* bufA = ch_smem + CH_NB * CH_LD + (CH_NB * CH_LD & 1);              // 16-byte aligned operand tiles
    double* bufB = bufA + CH_NB * CH_LDT;
    __shared__ int s_ok;
    __shared__ CholScratch cs;
    __shared__ double rdiag[CH_NB];              // reciprocals of the diagonal of the block in sD
    __shared__ double sx[2][CH_NB];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nblk = (n + CH_NB - 1) / CH_NB;
    const bool vec = (lds & 1) == 0 && ((size_t)S & 15) == 0;
    if (tid == 0) s_ok = 1;
    __syncthreads();
#ifdef VEL_CHOL_TIMING
    unsigned long long t_last = gtime();
#endif

    for (int kb = 0; kb < nblk; ++kb) {
        const int k0 = kb * CH_NB, w = min(CH_NB, n - k0);
        const int ob_end = min(nblk, (kb / (CH_NBO / CH_NB) + 1) * (CH_NBO / CH_NB));     // first block column after this outer block
        // ---- phase A: every CTA factors the diagonal block (redundantly: cheaper than a broadcast + barrier) ----------------
        for (int e = tid; e < CH_NB * CH_NB; e += CH_THREADS) {
            const int r = e >> 6, c2 = e & 63;
            if (r < w && c2 < w) sD[r][c2] = c2 <= r ? __ldcg(S + (long long)(k0 + r) * lds + k0 + c2) : 0.0;
        }
        __syncthreads();
        CH_T(0);
        chol_block(sD, w, &s_ok, rdiag, &cs);
        CH_T(1);
        if (blockIdx.x == 0) {
            for (int e = tid; e < CH_NB * CH_NB; e += CH_THREADS) {
                const int r = e >> 6, c2 = e & 63;
                if (r < w && c2 <= r) S[(long long)(k0 + r) * lds + k0 + c2] = sD[r][c2];
            }
        }
        // ---- phase B: rows below the block (and the right-hand side as row n): x L_kk^T = a, one warp per row ----------------
        {
            const int r_first = k0 + w, nrows = n - r_first + 1;       // + 1: the right-hand side
            const int warps_total = gridDim.x * (CH_THREADS / 32);
            for (int ri = blockIdx.x * (CH_THREADS / 32) + warp; ri < nrows; ri += warps_total) {
                const int r = r_first + ri;
                double* rowp = r < n ? S + (long long)r * lds + k0 : b + k0;
                // lane holds columns lane and lane + 32
                double a0 = lane < w ? __ldcg(rowp + lane) : 0.0, a1 = lane + 32 < w ? __ldcg(rowp + lane + 32) : 0.0;
                for (int cg8 = 0; cg8 < w; cg8 += 8) {
                    double l0[8], l1[8], rd[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {            // the L entries of this 8-column group, off the dependent chain
                        const int c2 = cg8 + u;
                        l0[u] = (c2 < w && lane > c2 && lane < w) ? sD[lane][c2] : 0.0;
                        l1[u] = (c2 < w && lane + 32 > c2 && lane + 32 < w) ? sD[lane + 32][c2] : 0.0;
                        rd[u] = c2 < w ? rdiag[c2] : 0.0;
                    }
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int c2 = cg8 + u;
                        if (c2 < w) {
                            const double src = c2 < 32 ? a0 : a1;
                            const double xc = __shfl_sync(0xffffffffu, src, c2 & 31) * rd[u];
                            if (lane == (c2 & 31)) { if (c2 < 32) a0 = xc; else a1 = xc; }
                            a0 -= xc * l0[u];
                            a1 -= xc * l1[u];
                        }
                    }
                }
                if (lane < w) rowp[lane] = a0;
                if (lane + 32 < w) rowp[lane + 32] = a1;
            }
        }
        CH_T(2);
        grid.sync();
        CH_T(3);
        // ---- phase C1: the rest of this outer block, with this panel ---------------------------------------------------------
        if (kb + 1 < ob_end) chol_update_columns(S, lds, n, b, nblk, kb + 1, ob_end, k0, w, bufA, bufB, vec);
        // ---- phase C2 (last panel of an outer block): everything to the right, with all columns of the outer block -----------
        if (kb + 1 == ob_end && ob_end < nblk) {
            const int ko0 = (kb / (CH_NBO / CH_NB)) * CH_NBO;
            chol_update_columns(S, lds, n, b, nblk, ob_end, nblk, ko0, k0 + w - ko0, bufA, bufB, vec);
        }
        CH_T(4);
        if (kb + 1 < nblk) grid.sync();
        CH_T(5);
    }
    // the diagonal blocks in shared memory were factored by every CTA: any CTA knows whether a pivot failed
    if (blockIdx.x == 0 && tid == 0) info[0] = s_ok ? 0 : 1;
    grid.sync();

    // ---- backward substitution  L^T x = y  (y is in b), block rows from the bottom -----------------------------------------
    for (int kb = nblk - 1; kb >= 0; --kb) {
        const int k0 = kb * CH_NB, w = min(CH_NB, n - k0);
        // every CTA solves the diagonal block redundantly (L_kk^T x_k = y_k) in shared memory
        for (int e = tid; e < CH_NB * CH_NB; e += CH_THREADS) {
            const int r = e >> 6, c2 = e & 63;
            if (r < w && c2 < w) sD[r][c2] = c2 <= r ? __ldcg(S + (long long)(k0 + r) * lds + k0 + c2) : 0.0;
        }
        if (tid < w) sx[0][tid] = __ldcg(b + k0 + tid);
        __syncthreads();
        if (tid < w) rdiag[tid] = 1.0 / sD[tid][tid];
        __syncthreads();
        if (warp == 0) {
            double y0 = lane < w ? sx[0][lane] : 0.0, y1 = lane + 32 < w ? sx[0][lane + 32] : 0.0;
            for (int cg8 = ((w - 1) >> 3) << 3; cg8 >= 0; cg8 -= 8) {
                double l0[8], l1[8], rd[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int c2 = cg8 + u;
                    l0[u] = (c2 < w && lane < c2) ? sD[c2][lane] : 0.0;              // column-oriented: y[m] -= L[c2][m] x[c2], m < c2
                    l1[u] = (c2 < w && lane + 32 < c2) ? sD[c2][lane + 32] : 0.0;
                    rd[u] = c2 < w ? rdiag[c2] : 0.0;
                }
#pragma unroll
                for (int u = 7; u >= 0; --u) {
                    const int c2 = cg8 + u;
                    if (c2 < w) {
                        const double src = c2 < 32 ? y0 : y1;
                        const double xc = __shfl_sync(0xffffffffu, src, c2 & 31) * rd[u];
                        if (lane == (c2 & 31)) { if (c2 < 32) y0 = xc; else y1 = xc; }
                        y0 -= xc * l0[u];
                        y1 -= xc * l1[u];
                    }
                }
            }
            if (lane < w) sx[1][lane] = y0;
            if (lane + 32 < w) sx[1][lane + 32] = y1;
        }
        __syncthreads();
        if (blockIdx.x == 0 && tid < w) b[k0 + tid] = sx[1][tid];
        // y[c] -= sum_r L[k0 + r][c] x[r]  for the columns c < k0, shared by all threads of the grid
        const int gtid = blockIdx.x * CH_THREADS + tid, gthreads = gridDim.x * CH_THREADS;
        for (int c2 = gtid; c2 < k0; c2 += gthreads) {
            double s2 = 0.0;
#pragma unroll 8
            for (int r = 0; r < w; ++r) s2 += __ldcg(S + (long long)(k0 + r) * lds + c2) * sx[1][r];
            b[c2] = __ldcg(b + c2) - s2;
        }
        CH_T(6);
        if (kb > 0) grid.sync();
        CH_T(7);
    }
}


// blocked form, right-hand side: b[r] -= sum_{c in [p0, p1)} L[r][c] y[c] for the rows r >= p1 below a factored panel group (one warp
// per row, fixed-order reduction)
__global__ void __launch_bounds__(256)
chol_rhs_gemv_kernel(const double* __restrict__ S, long long lds, double* __restrict__ b, int p0, int p1, int n, const int* __restrict__ gate)
{
    if (gate && *gate) return;
    const int r = p1 + blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (r >= n) return;
    const double* row = S + (long long)r * lds;
    double a = 0.0;
    for (int c = p0 + lane; c < p1; c += 32) a += row[c] * b[c];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == 0) b[r] -= a;
}

// ---- Cholesky as a task graph (default): no grid barrier in the factorisation -------------------------------------------------
// Tasks on 64x64 tiles:  D(k) factor the diagonal tile (+ L_kk^-1),  T(i,k) L_ik = A_ik L_kk^-T,  U(i,j,k) A_ij -= L_ik L_jk^T;
// the right-hand side is tile row nb.  Every tile has ONE owner CTA that applies all of its updates in panel order (fixed
// floating-point order, no atomics); readiness travels through release/acquire flags in global memory (dflag[k], tflag[i][k]).
// A CTA walks the panels k = 0, 1, ... and at each one runs its D, then its T, then its U tasks, spinning on the flags of the
// tiles it needs.  All waits point to tasks of the same or an earlier panel and the grid is co-resident (cooperative launch),
// so the schedule cannot deadlock; the look-ahead is implicit -- while the owner of the next diagonal tile factors it, the
// other CTAs apply the current panel to the rest of the matrix.  The two tiles on the critical path of panel k+1, (k+1,k) and
// (k+1,k+1), belong to the same CTA (k+1 mod G), so the chain D(k) -> T(k+1,k) -> U(k+1,k+1,k) -> D(k+1) crosses CTAs once.
#ifdef VEL_CHOL_TIMING
__device__ unsigned long long g_dag_t[2][8];
#define DG_T(k) do { if (threadIdx.x == 0) { unsigned long long now_ = gtime(); atomicAdd(&g_dag_t[blockIdx.x < nb ? 0 : 1][k], now_ - dg_last); dg_last = now_; } } while (0)
#else
#define DG_T(k)
#endif

struct DagOwn {
    int nb, G, W0, NW;
    __device__ __forceinline__ int owner(int i, int j) const      // i in j..nb (nb = right-hand-side strip)
    {
        if (i == j || i == j + 1) return i % G;                 // the two tiles on panel j+1's critical path share a CTA
        const long long e = (long long)j * (nb - 1) - (long long)j * (j - 1) / 2 + (i - (j + 2));
        return W0 + (int)(e % NW);
    }
};

__device__ __forceinline__ void dag_wait(const int* flag)
{
    if (threadIdx.x == 0) {
        while (ld_acquire(flag) == 0) __nanosleep(32);
    }
    __syncthreads();
}
__device__ __forceinline__ void dag_signal(int* flag)
{
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) st_release(flag, 1);
}

// Linv = L^-1 for the w x w lower-triangular L in sD (reciprocal diagonal in rdiag) -> sI [64][CH_LDT] (row-major, zero above the
// diagonal).  Recursive doubling, so that almost nothing sits on an FP64 latency chain: the eight 8x8 diagonal blocks are
// inverted by 64 threads in parallel (one 8-step chain), then three levels h = 8, 16, 32 fill the off-diagonal blocks with
//     Inv[R][C] = -Inv[R][R] * (L[R][C] * Inv[C][C])        R = lower half, C = upper half of a 2h x 2h diagonal block
// as two small tensor-core (DMMA) products per level.  tmp = [64][CH_LDT] scratch.  Rows/columns >= w behave as identity.
// stage 1 of tri_inverse_block: sI = 0, then the inverses of the eight 8x8 diagonal blocks (block-diagonal sI)
__device__ void tri_inverse_diag(double (*sD)[CH_LD], const double* rdiag, int w, double* sI)
{
    const int tid = threadIdx.x;
    for (int e = tid; e < CH_NB * CH_LDT; e += CH_THREADS) sI[e] = 0.0;
    __syncthreads();
    if (tid < 64) {
        const int m0 = tid & ~7, c = tid & 7;                // block m0/8, column c of the block
        double a[8], x[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) a[u] = u == c ? 1.0 : 0.0;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const bool in = m0 + q < w;
            x[q] = q >= c ? (in ? a[q] * rdiag[m0 + q] : a[q]) : 0.0;
#pragma unroll
            for (int u = q + 1; u < 8; ++u)
                if (m0 + u < w && in) a[u] -= sD[m0 + u][m0 + q] * x[q];
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) sI[(m0 + u) * CH_LDT + m0 + c] = x[u];
    }
    __syncthreads();
}

// stages 2-4: the off-diagonal blocks by recursive doubling (see tri_inverse_block)
__device__ void tri_inverse_levels(double (*sD)[CH_LD], int w, double* sI, double* tmp)
{
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t4 = lane & 3;
    for (int h = 8; h < CH_NB; h *= 2) {
        const int tpd = h >> 3, tpp = tpd * tpd, npair = CH_NB / (2 * h), ntile = npair * tpp;
        // tmp[R][C] = L[R][C] * Inv[C][C]
        for (int tt = warp; tt < ntile; tt += CH_THREADS / 32) {
            const int pr = tt / tpp, ta = (tt % tpp) / tpd, tb = tt % tpd;
            const int R0 = (2 * pr + 1) * h, C0 = 2 * pr * h;
            double d0 = 0.0, d1 = 0.0;
            for (int k4 = 0; k4 < h / 4; ++k4) {
                const int rr = R0 + ta * 8 + g, kk = C0 + k4 * 4 + t4;
                const double a = (rr < w && kk < w) ? sD[rr][kk] : 0.0;
                const double bv = sI[kk * CH_LDT + C0 + tb * 8 + g];
                dmma884(d0, d1, a, bv);
            }
            tmp[(R0 + ta * 8 + g) * CH_LDT + C0 + tb * 8 + 2 * t4] = d0;
            tmp[(R0 + ta * 8 + g) * CH_LDT + C0 + tb * 8 + 2 * t4 + 1] = d1;
        }
        __syncthreads();
        // Inv[R][C] = -Inv[R][R] * tmp[R][C]
        for (int tt = warp; tt < ntile; tt += CH_THREADS / 32) {
            const int pr = tt / tpp, ta = (tt % tpp) / tpd, tb = tt % tpd;
            const int R0 = (2 * pr + 1) * h, C0 = 2 * pr * h;
            double d0 = 0.0, d1 = 0.0;
            for (int k4 = 0; k4 < h / 4; ++k4) {
                const double a = sI[(R0 + ta * 8 + g) * CH_LDT + R0 + k4 * 4 + t4];
                const double bv = tmp[(R0 + k4 * 4 + t4) * CH_LDT + C0 + tb * 8 + g];
                dmma884(d0, d1, a, bv);
            }
            sI[(R0 + ta * 8 + g) * CH_LDT + C0 + tb * 8 + 2 * t4] = -d0;
            sI[(R0 + ta * 8 + g) * CH_LDT + C0 + tb * 8 + 2 * t4 + 1] = -d1;
        }
        __syncthreads();
    }
}

__device__ void tri_inverse_block(double (*sD)[CH_LD], const double* rdiag, int w, double* sI, double* tmp)
{
    tri_inverse_diag(sD, rdiag, w, sI);
    tri_inverse_levels(sD, w, sI, tmp);
}

// X = A L^-T for the 64 x 64 tile A (bufTile, row-major [64][CH_LDT]) and the lower-triangular 64 x 64 L (bufL, same layout; entries
// above the diagonal are never read), by blocked forward substitution over the eight column blocks of X
//     X_s = (A_s - sum_{m<s} X_m L_sm^T) L_ss^-T,         dinv[s] = L_ss^-1 (8 x 8, row-major)
// -> X (row-major [64][CH_LDT]).  Rows are independent: warp w < 8 owns rows 8w .. 8w+7 and walks its eight steps alone (two DMMAs
// per earlier block + two for the diagonal block: a chain of 72 DMMAs, ~1.3 us) -- the critical T task of the next panel runs this
// on L_kk as soon as the panel is factored, instead of waiting for the full inverse of L_kk (tri_inverse_levels, ~2.8 us).
__device__ void trsm_tile_64(const double* bufTile, const double* bufL, const double* dinv, double* X)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t4 = lane & 3;
    if (warp >= 8) return;
    const int r = warp * 8 + g;
    for (int s2 = 0; s2 < 8; ++s2) {
        double d0 = 0.0, d1 = 0.0;
        for (int m = 0; m < s2; ++m)
#pragma unroll
            for (int kk = 0; kk < 2; ++kk)
                dmma884(d0, d1, X[r * CH_LDT + 8 * m + 4 * kk + t4], bufL[(8 * s2 + g) * CH_LDT + 8 * m + 4 * kk + t4]);
        // tmp = A_s - acc, parked in X's own block s (only this warp touches rows 8w .. 8w+7)
        X[r * CH_LDT + 8 * s2 + 2 * t4] = bufTile[r * CH_LDT + 8 * s2 + 2 * t4] - d0;
        X[r * CH_LDT + 8 * s2 + 2 * t4 + 1] = bufTile[r * CH_LDT + 8 * s2 + 2 * t4 + 1] - d1;
        __syncwarp();
        double e0 = 0.0, e1 = 0.0;
#pragma unroll
        for (int kk = 0; kk < 2; ++kk)
            dmma884(e0, e1, X[r * CH_LDT + 8 * s2 + 4 * kk + t4], dinv[s2 * 64 + g * 8 + 4 * kk + t4]);
        __syncwarp();
        X[r * CH_LDT + 8 * s2 + 2 * t4] = e0;
        X[r * CH_LDT + 8 * s2 + 2 * t4 + 1] = e1;
        __syncwarp();
    }
}

// C(64x64 accumulators of the calling thread layout) = A(64 x kw) * B(64 x kw)^T from two staged tiles
__device__ __forceinline__ void tile_mma_64(const double* bufA, const double* bufB, double (&acc)[2][2][2])
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t4 = lane & 3;
    const int sr = (warp >> 2) * 16, sc = (warp & 3) * 16;
#pragma unroll
    for (int kk = 0; kk < CH_NB / 4; ++kk) {
        double a[2], bb[2];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            a[i] = bufA[(sr + i * 8 + g) * CH_LDT + kk * 4 + t4];
            bb[i] = bufB[(sc + i * 8 + g) * CH_LDT + kk * 4 + t4];
        }
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 2; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], bb[j]);
    }
}

__global__ void __launch_bounds__(CH_THREADS, 1)
chol_dag_kernel(double* S, long long lds, int n, double* b, int* __restrict__ info, int* dflag, int* tflag, int* xflag, double* Linv_g,
                const int* __restrict__ gate, int k_lo, int k_hi, int phase, int* eflag, double* Dinv_g)
{
    // phase 0: the whole factorisation (k_lo = 0, k_hi = nb) and the substitutions; phase 1: the panels [k_lo, k_hi) only -- tiles of
    // those COLUMNS, all rows -- for the blocked form of large systems (the trailing matrix is updated by the SYRK kernel between
    // launches); phase 2: the backward substitution only
    if (gate && *gate) return;
    cg::grid_group grid = cg::this_grid();
    extern __shared__ __align__(16) double ch_smem[];
    double (*sD)[CH_LD] = reinterpret_cast<double (*)[CH_LD]>(ch_smem);
    double* bufA = ch_smem + CH_NB * CH_LD + (CH_NB * CH_LD & 1);
    double* bufB = bufA + CH_NB * CH_LDT;
    double* sI = bufB + CH_NB * CH_LDT;                            // L_kk^-1 of the panel this CTA is working with
    double* bufT = sI + CH_NB * CH_LDT;                            // the resident sub-diagonal tile (me, me-1)
    __shared__ int s_ok, own_n;
    __shared__ CholScratch cs;
    __shared__ unsigned short own_i[CH_MAXOWN], own_j[CH_MAXOWN];
    __shared__ double rdiag[CH_NB];
    __shared__ double sx[2][CH_NB];
    __shared__ double sDinv[8 * 64];                               // the eight L_ss^-1 of the panel above this CTA's resident tile
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t4 = lane & 3;
    const int nb = (n + CH_NB - 1) / CH_NB;
    const bool vec = (lds & 1) == 0 && ((size_t)S & 15) == 0;
    const int me = blockIdx.x, G = gridDim.x;
    DagOwn own;
    own.nb = nb; own.G = G;
    if (nb <= G / 2) { own.W0 = nb; own.NW = G - nb; } else { own.W0 = 0; own.NW = G; }
    if (tid == 0) { s_ok = 1; own_n = 0; }
    __syncthreads();
    const int sr = (warp >> 2) * 16, sc = (warp & 3) * 16;

    // ---- this CTA's tiles, in (column, row) order -----------------------------------------------------------------------
    {
        const long long ntile = (long long)nb * (nb + 1) / 2 + nb;              // lower tiles + one strip tile per column
        const double hb = nb + 1.5;
        for (long long t = tid; t < ntile; t += CH_THREADS) {
            int j = (int)(hb - sqrt(fmax(hb * hb - 2.0 * (double)t, 0.0)));
            j = max(0, min(j, nb - 1));
            while (j > 0 && (long long)j * (nb + 1) - (long long)j * (j - 1) / 2 > t) --j;
            while (j < nb - 1 && (long long)(j + 1) * (nb + 1) - (long long)(j + 1) * j / 2 <= t) ++j;
            const int i = j + (int)(t - ((long long)j * (nb + 1) - (long long)j * (j - 1) / 2));
            if (j >= k_lo && j < k_hi && own.owner(i, j) == me) {
                const int idx = atomicAdd(&own_n, 1);
                if (idx < CH_MAXOWN) { own_i[idx] = (unsigned short)i; own_j[idx] = (unsigned short)j; }
            }
        }
        __syncthreads();
        if (own_n > CH_MAXOWN) {                      // cannot happen for n <= 64 * 65535 / ... with the sizes this library serves; report, do not hang
            if (me == 0 && tid == 0) info[0] = 2;
            own_n = 0;                                 // (every CTA takes the same decision only if all overflow; guarded on the host by the size check)
        }
        if (tid == 0) {                               // insertion sort by (j, i): a handful of entries
            for (int a = 1; a < own_n; ++a) {
                const unsigned short ti = own_i[a], tj = own_j[a];
                int bpos = a - 1;
                while (bpos >= 0 && (own_j[bpos] > tj || (own_j[bpos] == tj && own_i[bpos] > ti))) {
                    own_i[bpos + 1] = own_i[bpos]; own_j[bpos + 1] = own_j[bpos];
                    --bpos;
                }
                own_i[bpos + 1] = ti; own_j[bpos + 1] = tj;
            }
        }
        __syncthreads();
    }
    int p_lo = 0;                                      // first own tile with column >= k
    // The two tiles on the critical path of panel `me` -- the diagonal tile (me, me) and the sub-diagonal tile (me, me-1), both owned
    // by this CTA -- stay RESIDENT in shared memory from the start: their updates U(me, me, j), U(me, me-1, j) are applied there, the
    // triangular solve T(me, me-1) and the factorisation D(me) read them there, and the chain D(k) -> T(k+1,k) -> U(k+1,k+1,k) ->
    // D(k+1) crosses global memory only for what other CTAs need (L_kk^-1 in, L_(k+1,k) out).
    const bool resident = phase == 0 && nb <= G && me < nb;
    if (resident) {
        const int r0 = me * CH_NB, wme = min(CH_NB, n - r0);
        for (int e = tid; e < CH_NB * CH_NB; e += CH_THREADS) {
            const int r = e >> 6, c2 = e & 63;
            if (r < wme && c2 < wme) sD[r][c2] = c2 <= r ? __ldcg(S + (long long)(r0 + r) * lds + r0 + c2) : 0.0;
        }
        if (me > 0) chol_load_tile(S, lds, r0, wme, r0 - CH_NB, CH_NB, bufT, vec);
        __syncthreads();
    }
#ifdef VEL_CHOL_TIMING
    unsigned long long dg_last = gtime();
#endif

    for (int k = k_lo; k < k_hi && phase != 2; ++k) {
        const int k0 = k * CH_NB, w = min(CH_NB, n - k0);
        int have_linv = 0;
        int bufB_j = -1;                               // bufB holds L_(bufB_j, k)
        bool early_u = false;                          // U(me, me, k) was applied right behind the critical solve
        while (p_lo < own_n && own_j[p_lo] < k) ++p_lo;
        int p = p_lo;
        // ---------------- D(k) and T(i,k): own tiles of column k ----------------
        for (; p < own_n && own_j[p] == k; ++p) {
            const int i = own_i[p];
            if (i == k) {
                DG_T(7);
                if (!resident) {
                    for (int e = tid; e < CH_NB * CH_NB; e += CH_THREADS) {
                        const int r = e >> 6, c2 = e & 63;
                        if (r < w && c2 < w) sD[r][c2] = c2 <= r ? __ldcg(S + (long long)(k0 + r) * lds + k0 + c2) : 0.0;
                    }
                    __syncthreads();
                }
                DG_T(0);
                chol_block(sD, w, &s_ok, rdiag, &cs);
                DG_T(1);
                tri_inverse_diag(sD, rdiag, w, sI);
                // The next panel's critical tile (k+1, k) does not wait for the whole inverse: its owner solves against L_kk itself
                // (trsm_tile_64), which needs the factor and the inverses of its eight diagonal blocks -- published now, early.
                const bool early = resident && k + 1 < nb && w == CH_NB;
                if (early) {
                    for (int e = tid; e < CH_NB * CH_NB; e += CH_THREADS) {
                        const int r = e >> 6, c2 = e & 63;
                        if (c2 <= r) S[(long long)(k0 + r) * lds + k0 + c2] = sD[r][c2];
                    }
                    {
                        const int blk = tid >> 6, rr = (tid >> 3) & 7, cc = tid & 7;          // 512 threads = 8 blocks x 8 x 8
                        Dinv_g[(long long)k * 512 + tid] = sI[(8 * blk + rr) * CH_LDT + 8 * blk + cc];
                    }
                    dag_signal(eflag + k);
                }
                tri_inverse_levels(sD, w, sI, bufA);
                DG_T(2);
                // the inverse is what the other waiting T tasks need: publish it first, the factor itself (read by nobody else before
                // the kernel's final barrier) afterwards
                for (int e = tid; e < CH_NB * CH_NB; e += CH_THREADS) {
                    const int r = e >> 6, c2 = e & 63;
                    Linv_g[(long long)k * CH_NB * CH_NB + e] = (r < w && c2 < w) ? sI[r * CH_LDT + c2] : 0.0;
                }
                if (tid == 0 && !s_ok) info[0] = 1;
                dag_signal(dflag + k);
                if (!early) {
                    for (int e = tid; e < CH_NB * CH_NB; e += CH_THREADS) {
                        const int r = e >> 6, c2 = e & 63;
                        if (r < w && c2 <= r) S[(long long)(k0 + r) * lds + k0 + c2] = sD[r][c2];
                    }
                }
                DG_T(3);
                have_linv = 1;
                continue;
            }
            if (resident && i == me && k == me - 1) {
                // the critical tile (me, me-1), resident in bufT: solved against L_kk as soon as panel k is factored (early flag)
                DG_T(7);
                dag_wait(eflag + k);
                DG_T(4);
                chol_load_tile(S, lds, k0, CH_NB, k0, CH_NB, bufA, vec);       // L_kk (what lies above its diagonal is never read)
                sDinv[tid] = __ldcg(Dinv_g + (long long)k * 512 + tid);
                __syncthreads();
                trsm_tile_64(bufT, bufA, sDinv, bufB);                         // L_(me,k) -> bufB, where U(me, me, k) wants it
                __syncthreads();
                const int r0 = i * CH_NB, wi = min(CH_NB, n - r0);
                {
                    // U(me, me, k) at once: it is what D(me) waits for -- the copy of L_(me,k) to global memory and its flag (below) are
                    // for the other CTAs
                    double acc[2][2][2] = {};
                    if (!(sc > sr)) tile_mma_64(bufB, bufB, acc);
#pragma unroll
                    for (int ii = 0; ii < 2; ++ii)
#pragma unroll
                        for (int jj = 0; jj < 2; ++jj) {
                            const int r = sr + ii * 8 + g, c2 = sc + jj * 8 + 2 * t4;
                            if (r < wi && c2 < wi && c2 <= r) sD[r][c2] -= acc[ii][jj][0];
                            if (r < wi && c2 + 1 < wi && c2 + 1 <= r) sD[r][c2 + 1] -= acc[ii][jj][1];
                        }
                    early_u = true;
                }
                for (int e = tid; e < CH_NB * CH_NB; e += CH_THREADS) {
                    const int r = e >> 6, c2 = e & 63;
                    if (r < wi) S[(long long)(r0 + r) * lds + k0 + c2] = bufB[r * CH_LDT + c2];
                }
                bufB_j = i;
                dag_signal(tflag + (long long)i * nb + k);
                DG_T(5);
                continue;
            }
            if (!have_linv) {
                DG_T(7);
                dag_wait(dflag + k);
                DG_T(4);
                for (int e = tid; e < CH_NB * CH_NB; e += CH_THREADS) sI[(e >> 6) * CH_LDT + (e & 63)] = __ldcg(Linv_g + (long long)k * CH_NB * CH_NB + e);
                have_linv = 1;
            }
            __syncthreads();
            if (i == nb) {
                // right-hand-side strip: y[c] = sum_{m <= c} b[k0+m] Linv[c][m]
                if (tid < w) bufA[tid] = __ldcg(b + k0 + tid);
                __syncthreads();
                if (tid < w) {
                    double a0 = 0.0, a1 = 0.0;
                    for (int m = 0; m + 1 <= tid; m += 2) { a0 += bufA[m] * sI[tid * CH_LDT + m]; a1 += bufA[m + 1] * sI[tid * CH_LDT + m + 1]; }
                    if ((tid & 1) == 0) a0 += bufA[tid] * sI[tid * CH_LDT + tid];
                    b[k0 + tid] = a0 + a1;
                }
            } else {
                const int r0 = i * CH_NB, wi = min(CH_NB, n - r0);
                // the resident sub-diagonal tile; on a grid with fewer than 2 nb CTAs the panel CTAs share in the other tiles and may own
                // more of row `me` (found with VEL_CHOL_GRID=32: every tile of that row was then read from bufT)
                const bool res_t = resident && i == me && k == me - 1;
                if (!res_t) chol_load_tile(S, lds, r0, wi, k0, w, bufA, vec);
                __syncthreads();
                double acc[2][2][2] = {};
                tile_mma_64(res_t ? bufT : bufA, sI, acc);      // X = A Linv^T
#pragma unroll
                for (int ii = 0; ii < 2; ++ii)
#pragma unroll
                    for (int jj = 0; jj < 2; ++jj) {
                        const int r = sr + ii * 8 + g, c2 = sc + jj * 8 + 2 * t4;
                        if (r < wi) {
                            double* pp = S + (long long)(r0 + r) * lds + k0 + c2;
                            if (c2 < w) pp[0] = acc[ii][jj][0];
                            if (c2 + 1 < w) pp[1] = acc[ii][jj][1];
                        }
                        if (res_t) {                             // L_(me,k) stays in bufB for U(me, me, k): rows >= wi and columns >= w are zero
                            bufB[r * CH_LDT + c2] = acc[ii][jj][0];
                            bufB[r * CH_LDT + c2 + 1] = acc[ii][jj][1];
                        }
                    }
                if (res_t) bufB_j = i;
            }
            dag_signal(tflag + (long long)i * nb + k);
            DG_T(5);
        }
        // ---------------- U(i,j,k): own tiles of the columns j > k ----------------
        for (; p < own_n; ++p) {
            const int i = own_i[p], j = own_j[p];
            if (early_u && i == me && j == me) continue;
            const int c0 = j * CH_NB, wj = min(CH_NB, n - c0);
            if (j != bufB_j) {
                DG_T(7);
                dag_wait(tflag + (long long)j * nb + k);
                DG_T(4);
                chol_load_tile(S, lds, c0, wj, k0, w, bufB, vec);           // L_jk
                bufB_j = j;
            }
            if (i == nb) {
                dag_wait(tflag + (long long)nb * nb + k);
                if (tid < w) bufA[tid] = __ldcg(b + k0 + tid);
                __syncthreads();
                if (tid < wj) {
                    double s2 = 0.0;
                    for (int c2 = 0; c2 < w; ++c2) s2 += bufA[c2] * bufB[tid * CH_LDT + c2];
                    b[c0 + tid] = __ldcg(b + c0 + tid) - s2;
                }
                __syncthreads();
                continue;
            }
            const int r0 = i * CH_NB, wi = min(CH_NB, n - r0);
            if (i != j) {
                DG_T(6);
                dag_wait(tflag + (long long)i * nb + k);
                DG_T(4);
                chol_load_tile(S, lds, r0, wi, k0, w, bufA, vec);           // L_ik
            }
            __syncthreads();
            double acc[2][2][2] = {};
            if (!(i == j && sc > sr)) tile_mma_64(i == j ? bufB : bufA, bufB, acc);   // a diagonal tile needs its lower sub-tiles only
            if (resident && i == me && (j == me || j == me - 1)) {
                // the resident tiles take their update in shared memory
#pragma unroll
                for (int ii = 0; ii < 2; ++ii)
#pragma unroll
                    for (int jj = 0; jj < 2; ++jj) {
                        const int r = sr + ii * 8 + g, c2 = sc + jj * 8 + 2 * t4;
                        if (j == me) {
                            if (r < wi && c2 < wj && c2 <= r) sD[r][c2] -= acc[ii][jj][0];
                            if (r < wi && c2 + 1 < wj && c2 + 1 <= r) sD[r][c2 + 1] -= acc[ii][jj][1];
                        } else {
                            if (r < wi && c2 < wj) bufT[r * CH_LDT + c2] -= acc[ii][jj][0];
                            if (r < wi && c2 + 1 < wj) bufT[r * CH_LDT + c2 + 1] -= acc[ii][jj][1];
                        }
                    }
                __syncthreads();
                DG_T(6);
                continue;
            }
            double2 cur[2][2];
            bool m0[2][2], m1[2][2];
#pragma unroll
            for (int ii = 0; ii < 2; ++ii)
#pragma unroll
                for (int jj = 0; jj < 2; ++jj) {
                    const int r = sr + ii * 8 + g, c2 = sc + jj * 8 + 2 * t4;
                    m0[ii][jj] = r < wi && c2 < wj && (i > j || c2 <= r);
                    m1[ii][jj] = r < wi && c2 + 1 < wj && (i > j || c2 + 1 <= r);
                    const double* pp = S + (long long)(r0 + r) * lds + c0 + c2;
                    cur[ii][jj] = make_double2(0.0, 0.0);
                    if (vec && m0[ii][jj] && m1[ii][jj]) cur[ii][jj] = __ldcg(reinterpret_cast<const double2*>(pp));
                    else {
                        if (m0[ii][jj]) cur[ii][jj].x = __ldcg(pp);
                        if (m1[ii][jj]) cur[ii][jj].y = __ldcg(pp + 1);
                    }
                }
#pragma unroll
            for (int ii = 0; ii < 2; ++ii)
#pragma unroll
                for (int jj = 0; jj < 2; ++jj) {
                    const int r = sr + ii * 8 + g, c2 = sc + jj * 8 + 2 * t4;
                    double* pp = S + (long long)(r0 + r) * lds + c0 + c2;
                    const double2 v = make_double2(cur[ii][jj].x - acc[ii][jj][0], cur[ii][jj].y - acc[ii][jj][1]);
                    if (vec && m0[ii][jj] && m1[ii][jj]) *reinterpret_cast<double2*>(pp) = v;
                    else {
                        if (m0[ii][jj]) pp[0] = v.x;
                        if (m1[ii][jj]) pp[1] = v.y;
                    }
                }
            __syncthreads();                                        // the operand tiles are reloaded by the next task
            DG_T(6);
        }
    }
    if (phase == 1) return;                            // the launch boundary orders the panel against the trailing update that follows
    __threadfence();
    grid.sync();
    if (me == 0 && tid == 0 && info[0] != 1) info[0] = 0;

    // ---- backward substitution  L^T x = y ---------------------------------------------------------------------------------------
    if (nb <= G) {
        // Task-graph form, no grid barrier: CTA c owns block c of the solution.  It subtracts the contributions L[j][c]^T x_j of the
        // blocks below in FIXED order j = nb-1 ... c+1 as their x_j are published (xflag[j]), then x_c = Linv_c^T y_c and publishes.
        // Only the last contribution (j = c+1) and the product with Linv_c sit on the chain x_(c+1) -> x_c; their operands are staged
        // in shared memory beforehand.  out[col] = sum_r M[r][col] v[r] runs as 8 row classes x 64 columns (coalesced rows, 8-term
        // chains) whose partial sums are added in a fixed order.
        if (me < nb) {
            const int c = me, c0 = c * CH_NB, wc = min(CH_NB, n - c0);
            double* part = bufB;                                   // [8][CH_NB] partial sums
            const int col = tid & 63, cls = tid >> 6;
            if (c + 1 < nb) chol_load_tile(S, lds, c0 + CH_NB, min(CH_NB, n - c0 - CH_NB), c0, wc, bufA, vec);     // L[c+1][c]
            for (int e = tid; e < CH_NB * CH_NB; e += CH_THREADS) sI[(e >> 6) * CH_LDT + (e & 63)] = __ldcg(Linv_g + (long long)c * CH_NB * CH_NB + e);
            if (tid < CH_NB) sx[0][tid] = tid < wc ? __ldcg(b + c0 + tid) : 0.0;                                    // y_c
            __syncthreads();
            for (int j = nb - 1; j > c; --j) {
                const int j0 = j * CH_NB, wj = min(CH_NB, n - j0);
                dag_wait(xflag + j);
                if (tid < CH_NB) sx[1][tid] = tid < wj ? __ldcg(b + j0 + tid) : 0.0;                                // x_j
                __syncthreads();
                double a = 0.0;
                if (j == c + 1) {
                    for (int r = cls; r < wj; r += 8) a += bufA[r * CH_LDT + col] * sx[1][r];
                } else if (col < wc) {
                    for (int r = cls; r < wj; r += 8) a += __ldcg(S + (long long)(j0 + r) * lds + c0 + col) * sx[1][r];
                }
                part[cls * CH_NB + col] = a;
                __syncthreads();
                if (tid < CH_NB) {
                    double s2 = 0.0;
#pragma unroll
                    for (int q = 0; q < 8; ++q) s2 += part[q * CH_NB + tid];
                    sx[0][tid] -= s2;
                }
                __syncthreads();
            }
            {   // x_c[col] = sum_{m >= col} Linv_c[m][col] y_c[m]   (Linv is lower triangular: zero above the diagonal)
                double a = 0.0;
                for (int r = cls; r < wc; r += 8) a += sI[r * CH_LDT + col] * sx[0][r];
                part[cls * CH_NB + col] = a;
                __syncthreads();
                if (tid < wc) {
                    double s2 = 0.0;
#pragma unroll
                    for (int q = 0; q < 8; ++q) s2 += part[q * CH_NB + tid];
                    b[c0 + tid] = s2;
                }
            }
            dag_signal(xflag + c);
        }
        return;
    }
    // (larger systems than CTAs: one grid barrier per panel) with the stored inverses of the diagonal tiles: x_k = Linv_k^T y_k (a 64x64
    // product every CTA computes for itself), then y[c] -= sum_r L[k0+r][c] x[r] over the columns to the left
    for (int kb = nb - 1; kb >= 0; --kb) {
        const int k0 = kb * CH_NB, w = min(CH_NB, n - k0);
        for (int e = tid; e < CH_NB * CH_NB; e += CH_THREADS) sI[(e >> 6) * CH_LDT + (e & 63)] = __ldcg(Linv_g + (long long)kb * CH_NB * CH_NB + e);
        if (tid < CH_NB) sx[0][tid] = tid < w ? __ldcg(b + k0 + tid) : 0.0;
        __syncthreads();
        {
            // x[c] = sum_{m >= c} Linv[m][c] y[m]: 8 threads per output, fixed-order tree
            const int c = tid >> 3, part = tid & 7;
            double a = 0.0;
            for (int m = c + part; m < w; m += 8) a += sI[m * CH_LDT + c] * sx[0][m];
            a += __shfl_down_sync(0xffffffffu, a, 4, 8);
            a += __shfl_down_sync(0xffffffffu, a, 2, 8);
            a += __shfl_down_sync(0xffffffffu, a, 1, 8);
            if (part == 0) sx[1][c] = c < w ? a : 0.0;
        }
        __syncthreads();
        if (me == 0 && tid < w) b[k0 + tid] = sx[1][tid];
        const int gtid = me * CH_THREADS + tid, gthreads = G * CH_THREADS;
        for (int c2 = gtid; c2 < k0; c2 += gthreads) {
            double s2 = 0.0;
#pragma unroll 8
            for (int r = 0; r < w; ++r) s2 += __ldcg(S + (long long)(k0 + r) * lds + c2) * sx[1][r];
            b[c2] = __ldcg(b + c2) - s2;
        }
        if (kb > 0) grid.sync();
    }
}

inline size_t align256(size_t b) { return (b + 255) & ~(size_t)255; }

}  // namespace

// internal entry points (ba.cu) -------------------------------------------------------------------------------------------
size_t vel_dense_syrk_workspace(int m, int k)
{
    const SyrkPlan p = syrk_plan(m, k);
    return align256(sizeof(int) * ((size_t)p.ntiles + 1));      // turnstile flags + the work-queue counter
}

VEL_API size_t vel_syrk_lower_sub_workspace(int32_t m, int32_t k)
{
    if (m <= 0 || k <= 0) return 0;
    return vel_dense_syrk_workspace(m, k);
}

// the tile-row geometry the SYRK uses for an m x m product (the distributed form assigns whole tile rows to ranks)
VEL_API int vel_syrk_tile_rows(int32_t m, int32_t* nb, int32_t* rows_per_block)
{
    VEL_CHECK_ARG(m > 0 && nb && rows_per_block, "vel_syrk_tile_rows: bad argument");
    const SyrkPlan p = syrk_plan(m, SY_BK);
    *nb = p.nb;
    *rows_per_block = p.bme;
    return VEL_OK;
}

// gate (DEVICE, may be NULL): when *gate != 0 at launch the kernel returns at once -- the device-side loop control of vel_ba_iterate
int vel_dense_syrk_rows_gated(const double* E, int64_t ld, int32_t m, int32_t k, double* S, int64_t lds, void* work, size_t work_bytes,
                              int32_t blk_lo, int32_t blk_hi, const int* gate, vel_stream_t stream)
{
    VEL_CHECK_ARG(E && S && work, "vel_syrk_lower_sub: NULL argument");
    VEL_CHECK_ARG(m > 0 && k > 0 && lds >= m, "vel_syrk_lower_sub: bad sizes m=%d k=%d lds=%lld", m, k, (long long)lds);
    const long long kpad = ((long long)k + SY_BK - 1) / SY_BK * SY_BK;
    VEL_CHECK_ARG(ld >= kpad && ld % 2 == 0 && ((size_t)E & 15) == 0,
                  "vel_syrk_lower_sub: E needs 16-byte aligned rows (ld even) padded with zeros to a multiple of %d columns (ld %lld < %lld)",
                  SY_BK, (long long)ld, kpad);
    const SyrkPlan p = syrk_plan(m, k, blk_lo, blk_hi);
    if (p.ntiles == 0) return VEL_OK;
    VEL_CHECK_ARG(work_bytes >= sizeof(int) * ((size_t)p.ntiles + 1), "vel_syrk_lower_sub: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    int* flags = (int*)work;
    VEL_CUDA(cudaMemsetAsync(flags, 0, sizeof(int) * ((size_t)p.ntiles + 1), st));
    const char* qenv = getenv("VEL_SYRK_QUEUE");
    int* queue = (qenv && qenv[0] == '0') ? nullptr : flags + p.ntiles;
    static bool attr_set = false;
    if (!attr_set) {
        VEL_CUDA(cudaFuncSetAttribute(dsyrk_lower_sub_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SY_SMEM));
        attr_set = true;
    }
    long long ld_ = ld, lds_ = lds;
    int m_ = m, nb = p.nb, bme = p.bme, ntiles = p.ntiles, sk = p.sk, ktiles = p.ktiles, tile0 = p.tile0;
    void* args[] = {(void*)&E, &ld_, &m_, &nb, &bme, &ntiles, &sk, &ktiles, (void*)&S, &lds_, &flags, &tile0, (void*)&gate, (void*)&queue};
    // co-residency of the whole grid is what makes the turnstile wait safe: cooperative launch guarantees it (or fails)
    VEL_CUDA(cudaLaunchCooperativeKernel((void*)dsyrk_lower_sub_kernel, dim3(p.grid), dim3(SY_THREADS), args, SY_SMEM, st));
    return VEL_OK;
}

VEL_API int vel_syrk_lower_sub_rows(const double* E, int64_t ld, int32_t m, int32_t k, double* S, int64_t lds, void* work, size_t work_bytes,
                                    int32_t blk_lo, int32_t blk_hi, vel_stream_t stream)
{
    return vel_dense_syrk_rows_gated(E, ld, m, k, S, lds, work, work_bytes, blk_lo, blk_hi, nullptr, stream);
}

VEL_API int vel_syrk_lower_sub(const double* E, int64_t ld, int32_t m, int32_t k, double* S, int64_t lds, void* work, size_t work_bytes,
                               vel_stream_t stream)
{
    return vel_syrk_lower_sub_rows(E, ld, m, k, S, lds, work, work_bytes, 0, -1, stream);
}

#ifdef VEL_CHOL_TIMING
VEL_API void vel_chol_timing(unsigned long long* out8, int reset)
{
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out8, g_chol_t, sizeof(unsigned long long) * 8);
    cudaMemcpyFromSymbol(out8 + 8, g_cb_t, sizeof(unsigned long long) * 4);
    cudaMemcpyFromSymbol(out8 + 12, g_dag_t, sizeof(unsigned long long) * 16);
    if (reset) { unsigned long long z[8] = {0}; cudaMemcpyToSymbol(g_chol_t, z, sizeof(z)); cudaMemcpyToSymbol(g_cb_t, z, sizeof(unsigned long long) * 4); unsigned long long z2[16] = {0}; cudaMemcpyToSymbol(g_dag_t, z2, sizeof(z2)); }
}
#endif

constexpr int kBlockedMinPanels = 48;       // panels above which the blocked form takes over: 57 panels 2.6 vs 2.8 ms, 113: 10.4 vs 13.5, 225: 47 vs 91; 29 (C3) stays a pure task graph

int vel_dense_spd_solve_gated(double* S, int64_t lds, int32_t n, double* b, int32_t* info, const int* gate, vel_stream_t stream)
{
    VEL_CHECK_ARG(S && b && info, "vel_spd_solve: NULL argument");
    VEL_CHECK_ARG(n > 0 && lds >= n, "vel_spd_solve: bad sizes n=%d lds=%lld", n, (long long)lds);
    cudaStream_t st = (cudaStream_t)stream;
    const char* mode = getenv("VEL_CHOL");
    bool use_dag = !(mode && strcmp(mode, "sync") == 0);
    {   // the task-graph form keeps each CTA's tile list in shared memory: very large systems take the barrier form
        const long long nb_ = (n + CH_NB - 1) / CH_NB, ntile = nb_ * (nb_ + 1) / 2 + nb_;
        if (ntile > (long long)sm_count() * (CH_MAXOWN / 2)) use_dag = false;
    }
    static int max_grid = 0;
    if (max_grid == 0) {
        int per_sm = 0;
        VEL_CUDA(cudaFuncSetAttribute(chol_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CH_SMEM));
        VEL_CUDA(cudaFuncSetAttribute(chol_dag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CH_DAG_SMEM));
        VEL_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, chol_dag_kernel, CH_THREADS, CH_DAG_SMEM));
        VEL_CHECK_ARG(per_sm >= 1, "vel_spd_solve: kernel does not fit an SM");
        max_grid = sm_count();
    }
    const int nblk = (n + CH_NB - 1) / CH_NB;
    long long lds_ = lds;
    int n_ = n;
    if (gate && !use_dag) {
        vel_set_error("vel_spd_solve: the barrier form (VEL_CHOL=sync or a very large system) has no gated variant");
        return VEL_ERR_INVALID;
    }
    if (!use_dag) {
        int grid = max_grid;
        const int want = nblk * (nblk + 1) / 2 + nblk;          // tiles of the first trailing update
        if (want < grid) grid = want < 1 ? 1 : want;
        void* args[] = {(void*)&S, &lds_, &n_, (void*)&b, (void*)&info};
        VEL_CUDA(cudaLaunchCooperativeKernel((void*)chol_solve_kernel, dim3(grid), dim3(CH_THREADS), args, CH_SMEM, st));
        return VEL_OK;
    }
    // task-graph form: flags (dflag[nb], tflag[(nb+1)*nb]) and the inverses of the diagonal tiles live in stream-ordered scratch
    vel_keep_async_pool_cached();
    const size_t nflags = (size_t)nblk + (size_t)(nblk + 1) * nblk + (size_t)nblk + (size_t)nblk;     // dflag | tflag | xflag | eflag
    const size_t flag_bytes = align256(sizeof(int) * nflags);
    const size_t linv_bytes = sizeof(double) * (size_t)nblk * CH_NB * CH_NB;
    const size_t bytes = flag_bytes + linv_bytes + sizeof(double) * (size_t)nblk * 512;              // + the diagonal-block inverses per panel
    char* scratch = nullptr;
    VEL_CUDA(cudaMallocAsync((void**)&scratch, bytes, st));
    cudaError_t e = cudaMemsetAsync(scratch, 0, flag_bytes, st);
    if (e == cudaSuccess) e = cudaMemsetAsync(info, 0, sizeof(int), st);
    int* dflag = (int*)scratch;
    int* tflag = dflag + nblk;
    int* xflag = tflag + (size_t)(nblk + 1) * nblk;
    int* eflag = xflag + nblk;
    double* Linv_g = (double*)(scratch + flag_bytes);
    double* Dinv_g = (double*)(scratch + flag_bytes + linv_bytes);
    const int want = nblk * (nblk + 1) / 2 + nblk;
    int grid = max_grid;
    if (want < grid) grid = want < 1 ? 1 : want;
    if (const char* ge = getenv("VEL_CHOL_GRID")) { const int v = atoi(ge); if (v >= 1 && v < grid) grid = v; }      // tests: the forms a device with fewer SMs takes
    // Many panels (the multi-GPU global BA factors n = 3,594 ... 14,394 on its owner): the 64-tile task graph then spends its time
    // in latency-bound 64x64x64 update tasks (11 TFLOP/s at n = 14,394).  BLOCKED form: groups of CH_GROUP panels are factored by the task graph
    // restricted to their columns (phase 1), and everything to the right of a group is updated by the SYRK kernel (33 TFLOP/s at
    // this order) plus one GEMV for the right-hand side; the backward substitution is a last launch (phase 2).
    const char* benv = getenv("VEL_CHOL_BLOCKED");          // "0": never; a number > 1: use the blocked form above that many panels (A/B)
    int blocked_min = kBlockedMinPanels;
    if (benv && atoi(benv) > 1) blocked_min = atoi(benv);
    const bool blocked = nblk > blocked_min && (lds & 1) == 0 && ((size_t)S & 15) == 0 && !(benv && benv[0] == '0' && benv[1] == 0);
    int k_lo = 0, k_hi = nblk, phase = 0;
    void* args[] = {(void*)&S, &lds_, &n_, (void*)&b, (void*)&info, (void*)&dflag, (void*)&tflag, (void*)&xflag, (void*)&Linv_g, (void*)&gate,
                    &k_lo, &k_hi, &phase, (void*)&eflag, (void*)&Dinv_g};
    if (e == cudaSuccess && !blocked) {
        e = cudaLaunchCooperativeKernel((void*)chol_dag_kernel, dim3(grid), dim3(CH_THREADS), args, CH_DAG_SMEM, st);
    } else if (e == cudaSuccess) {
        int CH_GROUP = 16;                                                   // panels per group: K = 1024 for the trailing SYRK
        if (const char* ge = getenv("VEL_CHOL_GROUP")) { const int v = atoi(ge); if (v >= 2 && v <= 64) CH_GROUP = v; }
        const size_t sy_bytes = vel_dense_syrk_workspace(n, CH_GROUP * CH_NB);
        void* sy_work = nullptr;
        e = cudaMallocAsync(&sy_work, sy_bytes, st);
        for (int g0 = 0; g0 < nblk && e == cudaSuccess; g0 += CH_GROUP) {
            k_lo = g0; k_hi = g0 + CH_GROUP < nblk ? g0 + CH_GROUP : nblk; phase = 1;
            e = cudaLaunchCooperativeKernel((void*)chol_dag_kernel, dim3(grid), dim3(CH_THREADS), args, CH_DAG_SMEM, st);
            if (e != cudaSuccess || k_hi >= nblk) break;
            const int p0 = k_lo * CH_NB, p1 = k_hi * CH_NB, m2 = n - p1, kk = p1 - p0;
            chol_rhs_gemv_kernel<<<(m2 + 7) / 8, 256, 0, st>>>(S, lds, b, p0, p1, n, gate);
            e = cudaGetLastError();
            if (e != cudaSuccess) break;
            const int rc = vel_dense_syrk_rows_gated(S + (long long)p1 * lds + p0, lds, m2, kk, S + (long long)p1 * lds + p1, lds, sy_work, sy_bytes, 0,
                                                     -1, gate, stream);
            if (rc != VEL_OK) { cudaFreeAsync(sy_work, st); cudaFreeAsync(scratch, st); return rc; }
        }
        if (e == cudaSuccess) {
            k_lo = 0; k_hi = nblk; phase = 2;
            e = cudaLaunchCooperativeKernel((void*)chol_dag_kernel, dim3(grid), dim3(CH_THREADS), args, CH_DAG_SMEM, st);
        }
        if (sy_work) cudaFreeAsync(sy_work, st);
    }
    cudaFreeAsync(scratch, st);
    if (e != cudaSuccess) {
        vel_set_error("vel_spd_solve: %s", cudaGetErrorString(e));
        return VEL_ERR_CUDA;
    }
    return VEL_OK;
}

VEL_API int vel_spd_solve(double* S, int64_t lds, int32_t n, double* b, int32_t* info, vel_stream_t stream)
{
    return vel_dense_spd_solve_gated(S, lds, n, b, info, nullptr, stream);
}

// ---- measurement aid: the FP64 tensor-core (DMMA m8n8k4) rate of this device with operands in registers -----------------------
// MEASURED_PEAKS.json carries HBM and BF16 peaks only; bench.py reports the K8 SYRK against this number (TFLOP/s, best of 3).
// Synchronises.  Returns 0 on failure.
namespace {
__global__ void dmma_peak_kernel(double* out, int iters, double a0, double b0)
{
    double acc[16][2];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i][0] = acc[i][1] = 0.0;
    const double a = a0 + threadIdx.x, b = b0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) dmma884(acc[i][0], acc[i][1], a, b);
    }
    double s2 = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s2 += acc[i][0] + acc[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s2;
}
}  // namespace

VEL_API double vel_fp64_mma_peak_tflops(void)
{
    const int sms = sm_count(), threads = 512, iters = 4000;
    double* out = nullptr;
    if (cudaMalloc((void**)&out, sizeof(double) * sms * threads) != cudaSuccess) return 0.0;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    double best = 0.0;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        dmma_peak_kernel<<<sms, threads>>>(out, iters, 1.0, 1e-3);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) { best = 0.0; break; }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double tf = 2.0 * 256 * 16 * (double)iters * (threads / 32) * sms / ms / 1e9;
        if (rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    return best;
}
