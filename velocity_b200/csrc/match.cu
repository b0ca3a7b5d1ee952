// K4: all-pairs descriptor matching, two nearest neighbours per query.
//
// Replaces cv2.BFMatcher(norm).knnMatch(des1, des2, k=2) of the reference's descriptor fallback
// (utils/KLT.py:16,25): for every query row the two closest train rows, ascending distance, ties
// resolved to the LOWER train index (verified against cv2 4.13 in tests/golden/match_knn2.npz).
//   - Hamming, 256-bit descriptors (ORB; BASELINE config 4): exact integer popcount(xor).
//   - L2, float32 descriptors (what utils/KLT.py literally runs on SURF): sqrt(sum (a-b)^2), float32,
//     accumulated in index order exactly like the oracle.
// Decomposition: grid = (query tiles, train splits).  A CTA stages its train chunk tile by tile in
// shared memory (every thread reads the same train row -> broadcast), each thread keeps the running
// top-2 of one query in registers; partial results per split are merged in split order (strict <
// keeps the lowest index on ties), so the result does not depend on the split count.
#include "common.cuh"

#include <float.h>
#include <stdlib.h>

namespace {

constexpr int MQ_THREADS = 128;   // queries per CTA
constexpr int MT_TILE = 128;      // train rows per shared-memory tile

struct Top2i { int d0, i0, d1, i1; };

__device__ __forceinline__ void top2_push(int& d0, int& i0, int& d1, int& i1, int d, int j)
{
    if (d < d0) { d1 = d0; i1 = i0; d0 = d; i0 = j; }
    else if (d < d1) { d1 = d; i1 = j; }
}

__global__ void __launch_bounds__(MQ_THREADS)
knn2_hamming256_partial_kernel(const uint4* __restrict__ q, int nq, const uint4* __restrict__ t, int nt, int t_per_split,
                               int4* __restrict__ part)
{
    __shared__ uint4 st[MT_TILE][2];
    const int qi = blockIdx.x * MQ_THREADS + threadIdx.x;
    const int t0 = blockIdx.y * t_per_split, t1 = min(nt, t0 + t_per_split);
    uint4 qa = make_uint4(0, 0, 0, 0), qb = qa;
    if (qi < nq) { qa = __ldg(q + 2ll * qi); qb = __ldg(q + 2ll * qi + 1); }
    int d0 = INT_MAX, i0 = -1, d1 = INT_MAX, i1 = -1;
    for (int base = t0; base < t1; base += MT_TILE) {
        const int n = min(MT_TILE, t1 - base);
        __syncthreads();
        for (int k = threadIdx.x; k < 2 * n; k += MQ_THREADS) (&st[0][0])[k] = __ldg(t + 2ll * base + k);
        __syncthreads();
#pragma unroll 4
        for (int j = 0; j < n; ++j) {
            const uint4 a = st[j][0], b = st[j][1];
            const int d = __popc(qa.x ^ a.x) + __popc(qa.y ^ a.y) + __popc(qa.z ^ a.z) + __popc(qa.w ^ a.w) +
                          __popc(qb.x ^ b.x) + __popc(qb.y ^ b.y) + __popc(qb.z ^ b.z) + __popc(qb.w ^ b.w);
            top2_push(d0, i0, d1, i1, d, base + j);
        }
    }
    if (qi < nq) part[(long long)blockIdx.y * nq + qi] = make_int4(d0, i0, d1, i1);
}

// (distance, index) lexicographic order: partial results may cover interleaved column ranges
__device__ __forceinline__ void top2_push_lex(int& d0, int& i0, int& d1, int& i1, int d, int j)
{
    if (d < d0 || (d == d0 && j < i0)) { d1 = d0; i1 = i0; d0 = d; i0 = j; }
    else if (d < d1 || (d == d1 && j < i1)) { d1 = d; i1 = j; }
}

__global__ void knn2_hamming_merge_kernel(const int4* __restrict__ part, int nq, int nsplit, int* __restrict__ idx,
                                          int* __restrict__ dist)
{
    const int qi = blockIdx.x * blockDim.x + threadIdx.x;
    if (qi >= nq) return;
    int d0 = INT_MAX, i0 = -1, d1 = INT_MAX, i1 = -1;
    for (int s = 0; s < nsplit; ++s) {
        const int4 p = part[(long long)s * nq + qi];
        if (p.y >= 0) top2_push_lex(d0, i0, d1, i1, p.x, p.y);
        if (p.w >= 0) top2_push_lex(d0, i0, d1, i1, p.z, p.w);
    }
    idx[2 * qi] = i0; idx[2 * qi + 1] = i1;
    dist[2 * qi] = i0 < 0 ? -1 : d0; dist[2 * qi + 1] = i1 < 0 ? -1 : d1;
}

// ---- float32 L2 ------------------------------------------------------------------------------------
constexpr int L2_DIM_TILE = 64;   // descriptor dims staged per pass (SURF = 64, SIFT = 128 -> 2 passes not needed: see below)

__device__ __forceinline__ void top2_pushf(float& d0, int& i0, float& d1, int& i1, float d, int j)
{
    if (d < d0) { d1 = d0; i1 = i0; d0 = d; i0 = j; }
    else if (d < d1) { d1 = d; i1 = j; }
}

// one thread per query; the query row stays in registers for dim <= 128 via a templated bound
template <int DIM>
__global__ void __launch_bounds__(MQ_THREADS)
knn2_l2_partial_kernel(const float* __restrict__ q, int nq, const float* __restrict__ t, int nt, int t_per_split,
                       float4* __restrict__ part)
{
    constexpr int TT = 32;  // train rows per tile
    __shared__ float st[TT][DIM];
    const int qi = blockIdx.x * MQ_THREADS + threadIdx.x;
    const int t0 = blockIdx.y * t_per_split, t1 = min(nt, t0 + t_per_split);
    float qr[DIM];
#pragma unroll
    for (int k = 0; k < DIM; ++k) qr[k] = qi < nq ? __ldg(q + (long long)qi * DIM + k) : 0.f;
    float d0 = FLT_MAX, d1 = FLT_MAX;
    int i0 = -1, i1 = -1;
    for (int base = t0; base < t1; base += TT) {
        const int n = min(TT, t1 - base);
        __syncthreads();
        for (int k = threadIdx.x; k < n * DIM; k += MQ_THREADS) (&st[0][0])[k] = __ldg(t + (long long)base * DIM + k);
        __syncthreads();
        for (int j = 0; j < n; ++j) {
            float s = 0.f;
#pragma unroll
            for (int k = 0; k < DIM; ++k) {
                const float e = __fsub_rn(qr[k], st[j][k]);
                s = __fadd_rn(s, __fmul_rn(e, e));   // index-order accumulation, no FMA: equals the oracle bit for bit
            }
            top2_pushf(d0, i0, d1, i1, s, base + j);
        }
    }
    if (qi < nq) part[(long long)blockIdx.y * nq + qi] = make_float4(d0, __int_as_float(i0), d1, __int_as_float(i1));
}

__global__ void knn2_l2_merge_kernel(const float4* __restrict__ part, int nq, int nsplit, int* __restrict__ idx,
                                     float* __restrict__ dist)
{
    const int qi = blockIdx.x * blockDim.x + threadIdx.x;
    if (qi >= nq) return;
    float d0 = FLT_MAX, d1 = FLT_MAX;
    int i0 = -1, i1 = -1;
    for (int s = 0; s < nsplit; ++s) {
        const float4 p = part[(long long)s * nq + qi];
        const int a = __float_as_int(p.y), b = __float_as_int(p.w);
        if (a >= 0) top2_pushf(d0, i0, d1, i1, p.x, a);
        if (b >= 0) top2_pushf(d0, i0, d1, i1, p.z, b);
    }
    idx[2 * qi] = i0; idx[2 * qi + 1] = i1;
    dist[2 * qi] = i0 < 0 ? -1.f : __fsqrt_rn(d0);
    dist[2 * qi + 1] = i1 < 0 ? -1.f : __fsqrt_rn(d1);
}

int choose_splits(int nq, int nt, int tile)
{
    const int qblocks = (nq + MQ_THREADS - 1) / MQ_THREADS;
    int nsplit = (4 * kNumSMs + qblocks - 1) / qblocks;           // ~4 CTAs per SM in flight
    const int max_split = (nt + tile - 1) / tile;
    if (nsplit > max_split) nsplit = max_split;
    return nsplit < 1 ? 1 : nsplit;
}

}  // namespace

// shared with the tensor-core path (match_tc.cu): merge per-split partial top-2 in split order
int vel_match_merge_hamming(const int4* part, int nq, int nsplit, int32_t* idx, int32_t* dist, cudaStream_t st)
{
    knn2_hamming_merge_kernel<<<(nq + 255) / 256, 256, 0, st>>>(part, nq, nsplit, idx, dist);
    VEL_LAUNCH_CHECK("knn2_hamming_merge_kernel");
    return VEL_OK;
}

int vel_match_knn2_hamming256_tc(const uint8_t* q, int32_t nq, const uint8_t* t, int32_t nt, int32_t* idx, int32_t* dist,
                                 cudaStream_t st);

// exact CUDA-core (xor + popcount) form; used for problems too small to fill a 128 x 256 tensor tile
static int knn2_hamming256_popc(const uint8_t* q, int32_t nq, const uint8_t* t, int32_t nt, int32_t* idx, int32_t* dist,
                                cudaStream_t st)
{
    const int nsplit = choose_splits(nq, nt, MT_TILE);
    int t_per_split = (nt + nsplit - 1) / nsplit;
    t_per_split = ((t_per_split + MT_TILE - 1) / MT_TILE) * MT_TILE;
    if (t_per_split < MT_TILE) t_per_split = MT_TILE;
    const int nsp = nt > 0 ? (nt + t_per_split - 1) / t_per_split : 1;
    int4* part = nullptr;
    vel_keep_async_pool_cached();
    VEL_CUDA(cudaMallocAsync((void**)&part, sizeof(int4) * (size_t)nq * nsp, st));
    dim3 grid((nq + MQ_THREADS - 1) / MQ_THREADS, nsp);
    knn2_hamming256_partial_kernel<<<grid, MQ_THREADS, 0, st>>>((const uint4*)q, nq, (const uint4*)t, nt, t_per_split, part);
    VEL_LAUNCH_CHECK("knn2_hamming256_partial_kernel");
    const int rc = vel_match_merge_hamming(part, nq, nsp, idx, dist, st);
    if (rc != VEL_OK) return rc;
    VEL_CUDA(cudaFreeAsync(part, st));
    return VEL_OK;
}

// test hook: VEL_MATCH_FORCE=popc|tc selects one form regardless of size
VEL_API int vel_match_knn2_hamming256(const uint8_t* q, int32_t nq, const uint8_t* t, int32_t nt, int32_t* idx, int32_t* dist,
                                      vel_stream_t stream)
{
    VEL_CHECK_ARG(q && t && idx && dist, "vel_match_knn2_hamming256: NULL argument");
    VEL_CHECK_ARG(nq >= 0 && nt >= 0, "vel_match_knn2_hamming256: negative size");
    VEL_CHECK_ARG(((((uintptr_t)q) | ((uintptr_t)t)) & 15) == 0, "vel_match_knn2_hamming256: descriptors must be 16-byte aligned");
    if (nq == 0) return VEL_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const char* force = getenv("VEL_MATCH_FORCE");
    bool use_tc = nq >= 128 && nt >= 256;
    if (force && force[0] == 'p') use_tc = false;
    if (force && force[0] == 't' && nt >= 1) use_tc = true;
    if (use_tc) return vel_match_knn2_hamming256_tc(q, nq, t, nt, idx, dist, st);
    return knn2_hamming256_popc(q, nq, t, nt, idx, dist, st);
}

VEL_API int vel_match_knn2_l2(const float* q, int32_t nq, const float* t, int32_t nt, int32_t dim, int32_t* idx, float* dist,
                              vel_stream_t stream)
{
    VEL_CHECK_ARG(q && t && idx && dist, "vel_match_knn2_l2: NULL argument");
    VEL_CHECK_ARG(nq >= 0 && nt >= 0, "vel_match_knn2_l2: negative size");
    if (dim != 64 && dim != 128) {
        vel_set_error("vel_match_knn2_l2: descriptor dim %d not supported (64 = SURF, 128 = SURF-extended/SIFT)", dim);
        return VEL_ERR_UNSUPPORTED;
    }
    if (nq == 0) return VEL_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int TT = 32;
    const int nsplit = choose_splits(nq, nt, TT);
    int t_per_split = (nt + nsplit - 1) / nsplit;
    t_per_split = ((t_per_split + TT - 1) / TT) * TT;
    if (t_per_split < TT) t_per_split = TT;
    const int nsp = nt > 0 ? (nt + t_per_split - 1) / t_per_split : 1;
    float4* part = nullptr;
    vel_keep_async_pool_cached();
    VEL_CUDA(cudaMallocAsync((void**)&part, sizeof(float4) * (size_t)nq * nsp, st));
    dim3 grid((nq + MQ_THREADS - 1) / MQ_THREADS, nsp);
    if (dim == 64) knn2_l2_partial_kernel<64><<<grid, MQ_THREADS, 0, st>>>(q, nq, t, nt, t_per_split, part);
    else knn2_l2_partial_kernel<128><<<grid, MQ_THREADS, 0, st>>>(q, nq, t, nt, t_per_split, part);
    VEL_LAUNCH_CHECK("knn2_l2_partial_kernel");
    knn2_l2_merge_kernel<<<(nq + 255) / 256, 256, 0, st>>>(part, nq, nsp, idx, dist);
    VEL_LAUNCH_CHECK("knn2_l2_merge_kernel");
    VEL_CUDA(cudaFreeAsync(part, st));
    return VEL_OK;
}
