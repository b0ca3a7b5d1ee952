// K4 (tensor-core form): all-pairs Hamming distance of 256-bit descriptors as an int8 GEMM on the
// 5th-generation tensor cores, with the 2-nearest-neighbour selection fused into the epilogue.
//
//   bit b -> (2b - 1) in {-1,+1}  =>  <q', t'> = 256 - 2 * hamming(q, t)   (exact in int32)
//
// so D = Q' T'^T (M x N x 256, int8 x int8 -> int32) carries every distance, and the largest dot
// product is the nearest neighbour.  The 8192 x 8192 distance matrix is never written: each
// 128 x 256 accumulator tile lives in TMEM, four epilogue warps read it with tcgen05.ld (one query
// row per thread) and keep a running top-2 in registers.  Ties resolve to the lower train index
// (columns are visited in ascending order with strict >), as cv2.BFMatcher does.
//
// The epilogue is what bounds this kernel (an M128 N256 K256 int8 tile is 0.55 us of tensor time, its 32768 accumulators must be
// scanned in less), so the MMA itself produces the sort keys: the operands are scaled (query bits -> +-2, train bits -> +-64,
// product +-128) and a NINTH K32 step multiplies a constant query column of ones with a constant train column holding
// 127 - (column mod 128), so the accumulator IS  128 * dot + tie-break  -- a running top-2 costs min / max / max per element and no
// key construction; dot = acc >> 7, column = 127 - (acc & 127) within its 128-column block.  Keys are comparable inside the 64
// columns a thread scans of one tile; per tile the thread's local top-2 is merged into its running global pair (20 instructions
// per 64 elements).  The constant operands are written into shared memory once per CTA (hand-swizzled, 48 KB).
//
// What bounds it (measured on B200, 8192 x 8192, 128 CTAs x 16 tiles): 23.6 us.  With the scan removed (loads, barriers and MMAs only)
// the same pipeline takes 20.5 us, with the operand loads removed as well it does not change -- the floor is reading the int32
// accumulators out of TMEM: 128 KB per tile at 64 B per cycle and SM = 2048 cycles against 1218 cycles of MMA (K = 256 is short:
// an accumulator costs 1.9x more to read than to compute), i.e. the tensor pipe cannot exceed ~50 % on this shape.  The scan
// itself (3 alu-pipe instructions per element, 2 issue cycles each) hides behind the reads with 16 epilogue warps.
//
// Kernel anatomy (one CTA per 128 queries x one split of the train set, 576 threads):
//   warp 0      TMA producer: Q' tile once (2 x 16 KB, SWIZZLE_128B), T' tiles double-buffered (2 x 64 KB)
//   warp 1      TMEM allocator + single-thread tcgen05.mma issuer: 9 x (M128 N256 K32, kind::i8) per tile,
//               accumulators double-buffered in 2 x 256 TMEM columns; tcgen05.commit frees smem / publishes D
//   warps 2..17 epilogue (four warps per TMEM lane quarter, 64 columns each): tcgen05.ld 32x32b.x16 double-buffered in registers
//               ACROSS tiles, running top-2 with four independent accumulator pairs, overlapped with the next tile's MMAs
// Reference call site: cv2.BFMatcher(NORM_HAMMING).knnMatch(k=2) (the ORB variant of utils/KLT.py:16-26;
// BASELINE config 4).  Roofline class: tensor (34.36 G int8-op per 8192^2 pair of frames).
#include <cuda.h>
#include <limits.h>
#include <stdlib.h>

#include "common.cuh"

namespace tc {

constexpr int TC_M = 128, TC_N = 256;
constexpr int TC_EPW = 4;         // epilogue warps per TMEM lane quarter
constexpr int TC_CW = TC_N / TC_EPW;   // columns of a tile one epilogue thread scans
constexpr int TC_THREADS = 64 + 32 * 4 * TC_EPW;   // TMA warp + MMA warp + 16 epilogue warps
constexpr int Q_POS = 2, T_POS = 64;   // operand scaling: product of two set / two clear bits = +128
constexpr uint32_t A_CHUNK = TC_M * 128;          // one 128-byte K chunk of the query tile
constexpr uint32_t B_CHUNK = TC_N * 128;
constexpr uint32_t SMEM_A = 2 * A_CHUNK;          // K = 256 bytes = 2 chunks
constexpr uint32_t SMEM_B_STAGE = 2 * B_CHUNK;
constexpr uint32_t SMEM_AE = A_CHUNK, SMEM_BE = B_CHUNK;   // the constant operands of the tie-break step (one 128-byte chunk each)
constexpr uint32_t SMEM_BARS = 256;
constexpr uint32_t SMEM_TOTAL = SMEM_A + 2 * SMEM_B_STAGE + SMEM_AE + SMEM_BE + SMEM_BARS + 1024;   // + alignment slack

// instruction descriptor: D = S32 (2<<4), A = S8 (1<<7), B = S8 (1<<10), K-major A and B, N>>3 at bit 17, M>>4 at bit 24
constexpr uint32_t IDESC = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TC_N >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);
// shared-memory matrix descriptor, high word: SBO = 1024 B (>>4 = 64), version 1 (bit 46), SWIZZLE_128B (2 at bits 61..63)
constexpr uint32_t DESC_HI = 64u | (1u << 14) | (2u << 29);

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr)
{
    const uint32_t lo = ((saddr & 0x3FFFFu) >> 4) | (1u << 16);   // start address, LBO = 1 (unused for swizzled K-major)
    return ((uint64_t)DESC_HI << 32) | lo;
}
__device__ __forceinline__ void mma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(IDESC), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, int (&v)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, int (&v)[16])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait16(int (&v)[16])
{
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]), "+r"(v[8]),
                   "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15])
                 :
                 : "memory");
}

// wait for the outstanding tcgen05.ld; the loaded registers are tied to the asm ("+r") so the compiler cannot
// schedule their first use above the wait
__device__ __forceinline__ void tmem_ld_wait(int (&v)[32])
{
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]), "+r"(v[8]),
                   "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]), "+r"(v[16]),
                   "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]), "+r"(v[24]),
                   "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31])
                 :
                 : "memory");
}

// Branch-free running top-2 on the accumulators themselves (acc = 128 * dot + 127 - (column mod 128): a larger value is a smaller
// Hamming distance and, among equal distances, the LOWER column; comparable inside the 64 columns a thread scans of one tile).  Four independent (k0, k1) pairs break the loop-carried
// dependency; they are merged once per tile.
struct Top2Keys {
    int k0[4], k1[4];
    __device__ __forceinline__ void init()
    {
#pragma unroll
        for (int a = 0; a < 4; ++a) { k0[a] = INT_MIN; k1[a] = INT_MIN; }
    }
    __device__ __forceinline__ void push(int a, int key)
    {
        k1[a] = max(k1[a], min(k0[a], key));
        k0[a] = max(k0[a], key);
    }
};

template <int N>
__device__ __forceinline__ void consume_chunk(const int (&v)[N], Top2Keys& T)
{
#pragma unroll
    for (int c = 0; c < N; ++c) T.push(c & 3, v[c]);
}

// (distance, index) lexicographic order, as cv2.BFMatcher ranks
__device__ __forceinline__ void push_lex(int& d0, int& i0, int& d1, int& i1, int d, int j)
{
    if (d < d0 || (d == d0 && j < i0)) { d1 = d0; i1 = i0; d0 = d; i0 = j; }
    else if (d < d1 || (d == d1 && j < i1)) { d1 = d; i1 = j; }
}

// bits -> scaled +-1 int8 (bit b of byte k -> element 8k + b); the first nq_bytes bytes are queries (+-Q_POS), the rest train (+-T_POS)
__global__ void expand_pm_kernel(const uint8_t* __restrict__ q, long long nq_bytes, const uint8_t* __restrict__ t, long long nt_bytes,
                                 uint2* __restrict__ qe, uint2* __restrict__ te, int* __restrict__ counters, int ncounters)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < ncounters) counters[i] = 0;            // the per-query-block arrival counters of the main kernel
    if (i >= nq_bytes + nt_bytes) return;
    const bool isq = i < nq_bytes;
    const unsigned b = isq ? q[i] : t[i - nq_bytes];
    const unsigned pos = isq ? (unsigned)Q_POS : (unsigned)T_POS, neg = (0u - pos) & 0xFFu;
    unsigned lo = 0, hi = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        lo |= (((b >> k) & 1u) ? pos : neg) << (8 * k);
        hi |= (((b >> (k + 4)) & 1u) ? pos : neg) << (8 * k);
    }
    if (isq) qe[i] = make_uint2(lo, hi);
    else te[i - nq_bytes] = make_uint2(lo, hi);
}

// End of a CTA's scan: merge the TC_EPW column ranges of each row through shared memory (xch [TC_EPW][TC_M]: the idle train stages),
// then the splits of a query block by whichever of its CTAs finishes last (a counter per query block) -- no separate merge launch.
// Called by the epilogue warps only (named barrier 1).  Not inlined: its temporaries must not raise the register count of the scan.
__device__ __noinline__ void epilogue_merge(int2* xch, int g0, int g1, int sub, int row, int m0, int col_base, int nq, int4* __restrict__ part,
                                            int* __restrict__ counters, int* __restrict__ idx, int* __restrict__ dist, int* s_last_p)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    xch[sub * TC_M + row] = make_int2(g0, g1);
    asm volatile("bar.sync 1, %0;" ::"n"(32 * 4 * TC_EPW) : "memory");                  // epilogue warps only
    const bool writer = sub == 0 && m0 + row < nq;
    if (writer) {
        int k0 = g0, k1 = g1;
#pragma unroll
        for (int a = 1; a < TC_EPW; ++a) {
            const int2 o = xch[a * TC_M + row];
            const int n0 = max(k0, o.x);
            const int n1 = max(min(k0, o.x), max(k1, o.y));
            k0 = n0; k1 = n1;
        }
        // key -> (dot, column): dot = key >> 16 (arithmetic), column = 0xFFFF - (key & 0xFFFF)
        const int dot0 = k0 >> 16, dot1 = k1 >> 16;
        const int idx0 = k0 != INT_MIN ? col_base + (0xFFFF - (k0 & 0xFFFF)) : -1;
        const int idx1 = k1 != INT_MIN ? col_base + (0xFFFF - (k1 & 0xFFFF)) : -1;
        const int d0 = idx0 >= 0 ? (256 - dot0) >> 1 : 0x7fffffff, d1 = idx1 >= 0 ? (256 - dot1) >> 1 : 0x7fffffff;
        part[(long long)blockIdx.y * nq + m0 + row] = make_int4(d0, idx0, d1, idx1);
        __threadfence();
    }
    asm volatile("bar.sync 1, %0;" ::"n"(32 * 4 * TC_EPW) : "memory");
    if (warp == 2 && lane == 0) *s_last_p = atomicAdd(&counters[blockIdx.x], 1) == (int)gridDim.y - 1 ? 1 : 0;
    asm volatile("bar.sync 1, %0;" ::"n"(32 * 4 * TC_EPW) : "memory");
    if (*s_last_p && writer) {
        __threadfence();
        int d0 = INT_MAX, i0 = -1, d1 = INT_MAX, i1 = -1;
        for (int sp = 0; sp < (int)gridDim.y; ++sp) {
            const int4 p = __ldcg(&part[(long long)sp * nq + m0 + row]);
            if (p.y >= 0) push_lex(d0, i0, d1, i1, p.x, p.y);
            if (p.w >= 0) push_lex(d0, i0, d1, i1, p.z, p.w);
        }
        const int qi = m0 + row;
        idx[2 * qi] = i0; idx[2 * qi + 1] = i1;
        dist[2 * qi] = i0 < 0 ? -1 : d0; dist[2 * qi + 1] = i1 < 0 ? -1 : d1;
    }
}

__global__ void __launch_bounds__(TC_THREADS, 1)  // 18 warps are allocated as 20: 96 registers per thread is the ceiling
knn2_hamming_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmT, int nq, int nt,
                       int nblk_per_split, int4* __restrict__ part, int* __restrict__ counters, int* __restrict__ idx, int* __restrict__ dist)
{
    extern __shared__ uint8_t smem_raw[];
    __shared__ int s_last;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;   // SWIZZLE_128B tiles need 1024-byte alignment
    const uint32_t sA = base, sB = base + SMEM_A, sAe = sB + 2 * SMEM_B_STAGE, sBe = sAe + SMEM_AE, sBar = sBe + SMEM_BE;
    // barriers (8 bytes each): 0 a_full | 1,2 b_full | 3,4 b_empty | 5,6 tmem_full | 7,8 tmem_empty ; +80: TMEM base slot
    const uint32_t bar_a = sBar, bar_bfull = sBar + 8, bar_bempty = sBar + 24, bar_tfull = sBar + 40, bar_tempty = sBar + 56;
    const uint32_t tmem_slot = sBar + 80;
    uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * TC_M;
    const int nblk_total = (nt + TC_N - 1) / TC_N;
    const int blk0 = blockIdx.y * nblk_per_split;
    const int nblk = min(nblk_per_split, nblk_total - blk0);

    if (threadIdx.x == 0) {
        mbar_init(bar_a, 1);
        for (int s = 0; s < 2; ++s) {
            mbar_init(bar_bfull + 8 * s, 1);
            mbar_init(bar_bempty + 8 * s, 1);
            mbar_init(bar_tfull + 8 * s, 1);
            mbar_init(bar_tempty + 8 * s, 4 * TC_EPW);   // one arrive per epilogue warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(tmem_slot) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    {   // constant operands of the tie-break step, K-major SWIZZLE_128B by hand: element (row, k = 0) sits in 16-byte chunk (row & 7)
        uint8_t* ce = smem_raw + (sAe - smem_u32(smem_raw));
        for (uint32_t o = threadIdx.x * 16u; o < SMEM_AE + SMEM_BE; o += TC_THREADS * 16u) {
            const uint32_t r = (o >> 7) & 0x1FFu, chunk = (o >> 4) & 7u;           // row within A_e (o < SMEM_AE) or A_e rows + B_e row
            uint4 val = make_uint4(0u, 0u, 0u, 0u);
            if (chunk == (r & 7u)) val.x = o < SMEM_AE ? 1u : (uint32_t)(127 - (int)(((o - SMEM_AE) >> 7) & 127u));
            *reinterpret_cast<uint4*>(ce + o) = val;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");                // generic-proxy writes -> visible to the tensor core
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    if (warp == 0) {
        if (lane == 0) {
            mbar_expect_tx(bar_a, SMEM_A);
            tma_load_2d(sA, &tmQ, 0, m0, bar_a);
            tma_load_2d(sA + A_CHUNK, &tmQ, 128, m0, bar_a);
            for (int i = 0; i < nblk; ++i) {
                const int s = i & 1;
                if (i >= 2) mbar_wait(bar_bempty + 8 * s, ((i >> 1) - 1) & 1);
                const uint32_t dst = sB + s * SMEM_B_STAGE;
                const int row = (blk0 + i) * TC_N;
                mbar_expect_tx(bar_bfull + 8 * s, SMEM_B_STAGE);
                tma_load_2d(dst, &tmT, 0, row, bar_bfull + 8 * s);
                tma_load_2d(dst + B_CHUNK, &tmT, 128, row, bar_bfull + 8 * s);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            mbar_wait(bar_a, 0);
            for (int i = 0; i < nblk; ++i) {
                const int s = i & 1;
                mbar_wait(bar_bfull + 8 * s, (i >> 1) & 1);
                if (i >= 2) mbar_wait(bar_tempty + 8 * s, ((i >> 1) - 1) & 1);
                tc_fence_after();
                const uint32_t d = tmem_base + (uint32_t)s * TC_N;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const uint32_t koff = (uint32_t)(k & 3) * 32u;
                    const uint64_t ad = make_desc(sA + (k >> 2) * A_CHUNK + koff);
                    const uint64_t bd = make_desc(sB + s * SMEM_B_STAGE + (k >> 2) * B_CHUNK + koff);
                    mma_i8(d, ad, bd, k > 0 ? 1u : 0u);
                }
                mma_i8(d, make_desc(sAe), make_desc(sBe), 1u);     // + (127 - column mod 128): the accumulator becomes the sort key
                mma_commit(bar_bempty + 8 * s);   // smem stage reusable once these MMAs have read it
                mma_commit(bar_tfull + 8 * s);    // accumulator tile complete
            }
        }
    } else {
        const int q = warp & 3;                   // TMEM lane quarter this warp may access
        const int sub = (warp - 2) >> 2;          // TC_EPW warps share a quarter: each scans TC_CW of the 256 columns
        const int row = q * 32 + lane;
        const int col_base = blk0 * TC_N;
        const int blk128 = (sub * TC_CW) & ~127;  // the 128-column block of the tile this warp's columns lie in
        int g0 = INT_MIN, g1 = INT_MIN;           // running global pair: (dot << 16) + (0xFFFF - column local to this split)
        // TMEM reads are the floor of this kernel (64 B per cycle and SM: 2048 cycles for the 128 KB of a tile against 1218 of MMA), so
        // they must never pause: chunks of 32 columns are double-buffered in registers, the next chunk -- of this tile or the first of
        // the NEXT tile -- is requested before the current one is scanned, and the accumulator buffer is handed back to the MMA warp
        // as soon as its last chunk has landed.
        constexpr int CHW = 16;                   // columns per tcgen05.ld: two 16-register buffers (32-column chunks would need 64 data
                                                  // registers; 18 warps are allocated as 20, so the ceiling is 96 per thread)
        constexpr int NCH = TC_CW / CHW;          // chunks per tile and thread (even: the two register buffers alternate across tiles)
        static_assert(NCH % 2 == 0, "register double buffering assumes an even chunk count per tile");
        const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(sub * TC_CW);
        int va[CHW], vb[CHW];
        if (nblk > 0) {
            mbar_wait(bar_tfull, 0);
            tc_fence_after();
            tmem_ld16(tlane, va);
        }
        for (int i = 0; i < nblk; ++i) {
            const int s = i & 1;
            const int jloc = i * TC_N;            // local column of the tile's first column
            const int valid = min(TC_N, nt - (col_base + jloc));   // columns of this tile that exist
            const uint32_t taddr = tlane + (uint32_t)s * TC_N;
            Top2Keys T;
            T.init();
#pragma unroll
            for (int ch = 0; ch < NCH; ++ch) {
                int (&cur)[CHW] = (ch & 1) ? vb : va;
                int (&nxt)[CHW] = (ch & 1) ? va : vb;
                tmem_ld_wait16(cur);
                if (ch + 1 < NCH) {
                    tmem_ld16(taddr + (ch + 1) * CHW, nxt);
                } else {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_tempty + 8 * s);          // every value of tile i is in registers
                    if (i + 1 < nblk) {
                        const int s1 = (i + 1) & 1;
                        mbar_wait(bar_tfull + 8 * s1, ((i + 1) >> 1) & 1);
                        tc_fence_after();
                        tmem_ld16(tlane + (uint32_t)s1 * TC_N, nxt);
                    }
                }
                if (valid < sub * TC_CW + (ch + 1) * CHW) {
#pragma unroll
                    for (int c = 0; c < CHW; ++c) if (sub * TC_CW + ch * CHW + c >= valid) cur[c] = INT_MIN;   // padded train rows never win
                }
                consume_chunk(cur, T);
            }
            // the tile's pair -> global keys -> merge (keys are unique: plain max / min merging is exact)
            int k0 = T.k0[0], k1 = T.k1[0];
#pragma unroll
            for (int a = 1; a < 4; ++a) {
                const int x0 = T.k0[a], x1 = T.k1[a];
                const int n0 = max(k0, x0);
                const int n1 = max(min(k0, x0), max(k1, x1));
                k0 = n0; k1 = n1;
            }
            const int inv0 = 0xFFFF - (jloc + blk128 + 127);      // + (acc & 127) = 0xFFFF - column
            const int t0 = k0 == INT_MIN ? INT_MIN : ((k0 >> 7) << 16) + inv0 + (k0 & 127);
            const int t1 = k1 == INT_MIN ? INT_MIN : ((k1 >> 7) << 16) + inv0 + (k1 & 127);
            const int n0 = max(g0, t0);
            const int n1 = max(min(g0, t0), max(g1, t1));
            g0 = n0; g1 = n1;
        }
        epilogue_merge(reinterpret_cast<int2*>(smem_raw + (sB - smem_u32(smem_raw))), g0, g1, sub, row, m0, col_base, nq, part, counters, idx, dist,
                       &s_last);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn()
{
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
            qres != cudaDriverEntryPointSuccess)
            return nullptr;
        fn = (EncodeTiledFn)p;
    }
    return fn;
}

bool make_map(CUtensorMap* tm, void* ptr, int rows, int box_rows)
{
    EncodeTiledFn enc = encode_fn();
    if (!enc) return false;
    cuuint64_t dims[2] = {256, (cuuint64_t)rows};
    cuuint64_t strides[1] = {256};
    cuuint32_t box[2] = {128, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    return enc(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace tc

// defined in match.cu
int vel_match_merge_hamming(const int4* part, int nq, int nsplit, int32_t* idx, int32_t* dist, cudaStream_t st);

// Tensor-core path of vel_match_knn2_hamming256 (called from match.cu for problems large enough to fill tiles).
int vel_match_knn2_hamming256_tc(const uint8_t* q, int32_t nq, const uint8_t* t, int32_t nt, int32_t* idx, int32_t* dist,
                                 cudaStream_t st)
{
    using namespace tc;
    const int mblocks = (nq + TC_M - 1) / TC_M;
    const int nblk_total = (nt + TC_N - 1) / TC_N;
    int nsplit = kNumSMs / mblocks;                            // one CTA per SM, a single wave
    if (nsplit > nblk_total) nsplit = nblk_total;
    if (nsplit < 1) nsplit = 1;
    int nblk_per_split = (nblk_total + nsplit - 1) / nsplit;
    if (nblk_per_split > 255) {                                // keys hold 16-bit local columns: <= 255 tiles of 256 per split
        nblk_per_split = 255;
    }
    nsplit = (nblk_total + nblk_per_split - 1) / nblk_per_split;
    // one stream-ordered scratch block: expanded operands | per-split partial results | per-query-block counters
    const size_t qe_bytes = ((size_t)nq * 256 + 255) & ~(size_t)255, te_bytes = ((size_t)nt * 256 + 255) & ~(size_t)255;
    const size_t part_bytes = (sizeof(int4) * (size_t)nq * nsplit + 255) & ~(size_t)255;
    char* scratch = nullptr;
    vel_keep_async_pool_cached();
    VEL_CUDA(cudaMallocAsync((void**)&scratch, qe_bytes + te_bytes + part_bytes + sizeof(int) * (size_t)mblocks, st));
    int8_t* qe = (int8_t*)scratch;
    int8_t* te = (int8_t*)(scratch + qe_bytes);
    int4* part = (int4*)(scratch + qe_bytes + te_bytes);
    int* counters = (int*)(scratch + qe_bytes + te_bytes + part_bytes);
    long long nexp = (long long)(nq + nt) * 32;
    if (nexp < mblocks) nexp = mblocks;
    expand_pm_kernel<<<(unsigned)((nexp + 255) / 256), 256, 0, st>>>(q, (long long)nq * 32, t, (long long)nt * 32, (uint2*)qe, (uint2*)te, counters,
                                                                     mblocks);
    VEL_LAUNCH_CHECK("expand_pm_kernel");

    CUtensorMap tmQ, tmT;
    if (!make_map(&tmQ, qe, nq, TC_M) || !make_map(&tmT, te, nt, TC_N)) {
        cudaFreeAsync(scratch, st);
        vel_set_error("vel_match_knn2_hamming256: cuTensorMapEncodeTiled failed");
        return VEL_ERR_CUDA;
    }
    static bool attr_set = false;
    if (!attr_set) {
        VEL_CUDA(cudaFuncSetAttribute(knn2_hamming_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_TOTAL));
        attr_set = true;
    }
    knn2_hamming_tc_kernel<<<dim3(mblocks, nsplit), TC_THREADS, SMEM_TOTAL, st>>>(tmQ, tmT, nq, nt, nblk_per_split, part, counters, idx, dist);
    VEL_LAUNCH_CHECK("knn2_hamming_tc_kernel");
    VEL_CUDA(cudaFreeAsync(scratch, st));
    return VEL_OK;
}
