// K6: multi-view ray triangulation.
//
//   vel_triangulate_2v <- fcn2vintercept (utils/MSV.py:98-142): for every frame pair j<k the two
//       closest-approach points of rays (A_j,U_j), (A_k,U_k); the tie point is the mean of all
//       2*C(nf,2) of them:  C0 = (sum_{j<k} (t1*v + s1*u) + (nf-1)*sum_f A_f) / (2*C(nf,2)).
//   vel_triangulate_nv <- fcnNvintercept (utils/MSV.py:146-175): C0 = S1^-1 S2 with
//       S1 = sum_f (I - u u^T), S2 = sum_f (I - u u^T) A_f.
// A [nf][3], U [3][nf][nv] (component-major, so consecutive threads = consecutive points read
// consecutive addresses), C0 [nv][3]; float64 throughout.
#include "common.cuh"

namespace {

constexpr int TRI_THREADS = 128;

// ---- N-view least squares -------------------------------------------------------------------------
// S1_i = sum_f (I - u u^T), S2_i = sum_f (I - u u^T) A_f, C0_i = S1_i^-1 S2_i (utils/MSV.py:146-175).  grid = (point blocks, frame
// chunks): a thread sums its point's rays over one chunk of frames (one thread per point over all 300 frames left the machine to
// 128 warps on a chain of dependent loads: 190 us at C3); the chunk sums are added in chunk order by the solve kernel --
// deterministic, no atomics.
constexpr int TRI_NACC = 9;
__global__ void __launch_bounds__(TRI_THREADS)
tri_nv_partial_kernel(const double* __restrict__ A, const double* __restrict__ U, int nf, int nv, int f_per_chunk, double* __restrict__ part)
{
    const int i = blockIdx.x * TRI_THREADS + threadIdx.x;
    if (i >= nv) return;
    const int f0 = blockIdx.y * f_per_chunk, f1 = min(nf, f0 + f_per_chunk);
    const long long plane = (long long)nf * nv;
    double m00 = 0, m01 = 0, m02 = 0, m11 = 0, m12 = 0, m22 = 0, b0 = 0, b1 = 0, b2 = 0;
    for (int f = f0; f < f1; ++f) {
        const double ux = U[(long long)f * nv + i], uy = U[plane + (long long)f * nv + i], uz = U[2 * plane + (long long)f * nv + i];
        const double ax = A[3 * f], ay = A[3 * f + 1], az = A[3 * f + 2];
        const double v00 = 1 - ux * ux, v01 = -ux * uy, v02 = -ux * uz, v11 = 1 - uy * uy, v12 = -uy * uz, v22 = 1 - uz * uz;
        m00 += v00; m01 += v01; m02 += v02; m11 += v11; m12 += v12; m22 += v22;
        b0 += ax * v00 + ay * v01 + az * v02;
        b1 += ax * v01 + ay * v11 + az * v12;
        b2 += ax * v02 + ay * v12 + az * v22;
    }
    double* o = part + (long long)blockIdx.y * TRI_NACC * nv + i;        // [chunk][quantity][point]: coalesced
    o[0] = m00; o[(long long)nv] = m01; o[2ll * nv] = m02; o[3ll * nv] = m11; o[4ll * nv] = m12; o[5ll * nv] = m22;
    o[6ll * nv] = b0; o[7ll * nv] = b1; o[8ll * nv] = b2;
}

__global__ void __launch_bounds__(TRI_THREADS)
tri_nv_solve_kernel(const double* __restrict__ part, int nv, int nchunks, double* __restrict__ C0)
{
    const int i = blockIdx.x * TRI_THREADS + threadIdx.x;
    if (i >= nv) return;
    double q[TRI_NACC];
#pragma unroll
    for (int k = 0; k < TRI_NACC; ++k) q[k] = 0.0;
    for (int c = 0; c < nchunks; ++c) {
        const double* o = part + (long long)c * TRI_NACC * nv + i;
#pragma unroll
        for (int k = 0; k < TRI_NACC; ++k) q[k] += o[(long long)k * nv];
    }
    const double m00 = q[0], m01 = q[1], m02 = q[2], m11 = q[3], m12 = q[4], m22 = q[5], b0 = q[6], b1 = q[7], b2 = q[8];
    // C0 = inv(S1) @ S2 for the symmetric 3x3 S1 (adjugate form)
    const double c00 = m11 * m22 - m12 * m12, c01 = m02 * m12 - m01 * m22, c02 = m01 * m12 - m02 * m11;
    const double c11 = m00 * m22 - m02 * m02, c12 = m01 * m02 - m00 * m12, c22 = m00 * m11 - m01 * m01;
    const double det = m00 * c00 + m01 * c01 + m02 * c02;
    const double inv = 1.0 / det;
    C0[3ll * i + 0] = (c00 * b0 + c01 * b1 + c02 * b2) * inv;
    C0[3ll * i + 1] = (c01 * b0 + c11 * b1 + c12 * b2) * inv;
    C0[3ll * i + 2] = (c02 * b0 + c12 * b1 + c22 * b2) * inv;
}

// ---- pairwise closest approach --------------------------------------------------------------------
// grid = (point blocks, J chunks): chunk c sums the pairs whose first frame j is in its range;
// partial sums go to part[c][nv][3] and are added in chunk order by the finalize kernel
// (deterministic, no atomics).
__global__ void __launch_bounds__(TRI_THREADS)
tri_2v_partial_kernel(const double* __restrict__ A, const double* __restrict__ U, int nf, int nv, int j_per_chunk,
                      double* __restrict__ part)
{
    const int i = blockIdx.x * TRI_THREADS + threadIdx.x;
    if (i >= nv) return;
    const int j0 = blockIdx.y * j_per_chunk, j1 = min(nf - 1, j0 + j_per_chunk);
    const long long plane = (long long)nf * nv;
    double sx = 0, sy = 0, sz = 0;
    for (int j = j0; j < j1; ++j) {
        const double ux = U[(long long)j * nv + i], uy = U[plane + (long long)j * nv + i], uz = U[2 * plane + (long long)j * nv + i];
        const double ajx = A[3 * j], ajy = A[3 * j + 1], ajz = A[3 * j + 2];
        for (int k = j + 1; k < nf; ++k) {
            const double vx = U[(long long)k * nv + i], vy = U[plane + (long long)k * nv + i], vz = U[2 * plane + (long long)k * nv + i];
            const double bx = ajx - A[3 * k], by = ajy - A[3 * k + 1], bz = ajz - A[3 * k + 2];
            const double d = ux * vx + uy * vy + uz * vz;
            const double e = ux * bx + uy * by + uz * bz;
            const double f = vx * bx + vy * by + vz * bz;
            const double g = 1 - d * d;
            const double s1 = (d * f - e) / g;
            const double t1 = (f - d * e) / g;
            sx += t1 * vx + s1 * ux;
            sy += t1 * vy + s1 * uy;
            sz += t1 * vz + s1 * uz;
        }
    }
    double* o = part + ((long long)blockIdx.y * nv + i) * 3;
    o[0] = sx; o[1] = sy; o[2] = sz;
}

__global__ void __launch_bounds__(TRI_THREADS)
tri_2v_finalize_kernel(const double* __restrict__ A, const double* __restrict__ part, int nf, int nv, int nchunks,
                       double* __restrict__ C0)
{
    const int i = blockIdx.x * TRI_THREADS + threadIdx.x;
    if (i >= nv) return;
    double sx = 0, sy = 0, sz = 0;
    for (int c = 0; c < nchunks; ++c) {
        const double* o = part + ((long long)c * nv + i) * 3;
        sx += o[0]; sy += o[1]; sz += o[2];
    }
    double ax = 0, ay = 0, az = 0;
    for (int f = 0; f < nf; ++f) { ax += A[3 * f]; ay += A[3 * f + 1]; az += A[3 * f + 2]; }
    const double den = (double)((long long)nf * (nf - 1));  // number of pairs times 2
    const double m = (double)(nf - 1);
    C0[3ll * i + 0] = (sx + ax * m) / den;
    C0[3ll * i + 1] = (sy + ay * m) / den;
    C0[3ll * i + 2] = (sz + az * m) / den;
}


// ---- fcnMSV1_t (utils/MSV.py:8-49): LM on the last camera's translation with pairwise ---------------
// re-triangulation inside every iteration.  The whole loop (<= max_iter iterations) runs in ONE
// persistent CTA: a host-driven loop would need a device->host sync per iteration for the
// rms(delta) < 1e-8 test and be slower than the reference's numpy loop.
constexpr int MSV_THREADS = 256;

__device__ __forceinline__ double msv_warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ void msv_project(const double* K, double ax, double ay, double az, double& u, double& v)
{
    const double q0 = ax * K[0] + ay * K[3] + az * K[6];
    const double q1 = ax * K[1] + ay * K[4] + az * K[7];
    const double q2 = ax * K[2] + ay * K[5] + az * K[8];
    u = q0 / q2;
    v = q1 / q2;
}

__global__ void __launch_bounds__(MSV_THREADS)
msv1_t_kernel(const double* __restrict__ Kg, const double* __restrict__ Afix, const double* __restrict__ U, int nf, int ng,
              const double* __restrict__ z, const double* __restrict__ x0, int max_iter, double* __restrict__ xout,
              double* __restrict__ b0out, int* __restrict__ iters)
{
    constexpr int NW = MSV_THREADS / 32;
    __shared__ double sK[9], sx[3], sAsum[3];
    __shared__ double sred[NW][9];
    __shared__ int s_done;
    const int tid = threadIdx.x;
    const long long plane = (long long)nf * ng;
    if (tid < 9) sK[tid] = Kg[tid];
    if (tid < 3) {
        sx[tid] = x0[tid];
        double a = 0.0;
        for (int f = 0; f < nf - 1; ++f) a += Afix[3 * f + tid];
        sAsum[tid] = a;   // sum of the fixed origins, added in frame order like A.sum(0)
    }
    if (tid == 0) s_done = 0;
    __syncthreads();
    const double dx = 1e-6;
    const double den = (double)((long long)nf * (nf - 1));
    int it = 0;
    for (; it < max_iter; ++it) {
        const double lx = -sx[0], ly = -sx[1], lz = -sx[2];   // origin of the last camera: -x
        double acc[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) acc[k] = 0.0;
        for (int i = tid; i < ng; i += MSV_THREADS) {
            double sx_ = 0, sy_ = 0, sz_ = 0;
            for (int j = 0; j < nf - 1; ++j) {
                const double ux = U[(long long)j * ng + i], uy = U[plane + (long long)j * ng + i], uz = U[2 * plane + (long long)j * ng + i];
                const double ajx = Afix[3 * j], ajy = Afix[3 * j + 1], ajz = Afix[3 * j + 2];
                for (int k = j + 1; k < nf; ++k) {
                    const double vx = U[(long long)k * ng + i], vy = U[plane + (long long)k * ng + i], vz = U[2 * plane + (long long)k * ng + i];
                    const bool last = (k == nf - 1);
                    const double bx = ajx - (last ? lx : Afix[3 * k]), by = ajy - (last ? ly : Afix[3 * k + 1]),
                                 bz = ajz - (last ? lz : Afix[3 * k + 2]);
                    const double d = ux * vx + uy * vy + uz * vz;
                    const double e = ux * bx + uy * by + uz * bz;
                    const double f = vx * bx + vy * by + vz * bz;
                    const double g = 1 - d * d;
                    const double s1 = (d * f - e) / g, t1 = (f - d * e) / g;
                    sx_ += t1 * vx + s1 * ux; sy_ += t1 * vy + s1 * uy; sz_ += t1 * vz + s1 * uz;
                }
            }
            const double m = (double)(nf - 1);
            const double cx = (sx_ + (sAsum[0] + lx) * m) / den, cy = (sy_ + (sAsum[1] + ly) * m) / den,
                         cz = (sz_ + (sAsum[2] + lz) * m) / den;
            const double bx0 = cx + sx[0], by0 = cy + sx[1], bz0 = cz + sx[2];   // b0 = C0 + x
            b0out[3ll * i] = bx0; b0out[3ll * i + 1] = by0; b0out[3ll * i + 2] = bz0;
            double u0, v0, u, v, ju[3], jv[3];
            msv_project(sK, bx0, by0, bz0, u0, v0);
            msv_project(sK, bx0 + dx, by0, bz0, u, v); ju[0] = (u - u0) / dx; jv[0] = (v - v0) / dx;
            msv_project(sK, bx0, by0 + dx, bz0, u, v); ju[1] = (u - u0) / dx; jv[1] = (v - v0) / dx;
            msv_project(sK, bx0, by0, bz0 + dx, u, v); ju[2] = (u - u0) / dx; jv[2] = (v - v0) / dx;
            const double ru = z[2ll * i] - u0, rv = z[2ll * i + 1] - v0;
            acc[0] += ju[0] * ju[0] + jv[0] * jv[0]; acc[1] += ju[0] * ju[1] + jv[0] * jv[1]; acc[2] += ju[0] * ju[2] + jv[0] * jv[2];
            acc[3] += ju[1] * ju[1] + jv[1] * jv[1]; acc[4] += ju[1] * ju[2] + jv[1] * jv[2]; acc[5] += ju[2] * ju[2] + jv[2] * jv[2];
            acc[6] += ju[0] * ru + jv[0] * rv; acc[7] += ju[1] * ru + jv[1] * rv; acc[8] += ju[2] * ru + jv[2] * rv;
        }
#pragma unroll
        for (int k = 0; k < 9; ++k) acc[k] = msv_warp_sum(acc[k]);
        if ((tid & 31) == 0) {
#pragma unroll
            for (int k = 0; k < 9; ++k) sred[tid >> 5][k] = acc[k];
        }
        __syncthreads();
        if (tid == 0) {
            double t[9];
            for (int k = 0; k < 9; ++k) {
                double s = 0.0;
                for (int w = 0; w < NW; ++w) s += sred[w][k];
                t[k] = s;
            }
            // (JtJ + I) delta = Jt r, symmetric 3x3 by the adjugate
            const double m00 = t[0] + 1.0, m01 = t[1], m02 = t[2], m11 = t[3] + 1.0, m12 = t[4], m22 = t[5] + 1.0;
            const double c00 = m11 * m22 - m12 * m12, c01 = m02 * m12 - m01 * m22, c02 = m01 * m12 - m02 * m11;
            const double c11 = m00 * m22 - m02 * m02, c12 = m01 * m02 - m00 * m12, c22 = m00 * m11 - m01 * m01;
            const double inv = 1.0 / (m00 * c00 + m01 * c01 + m02 * c02);
            const double d0 = (c00 * t[6] + c01 * t[7] + c02 * t[8]) * inv;
            const double d1 = (c01 * t[6] + c11 * t[7] + c12 * t[8]) * inv;
            const double d2 = (c02 * t[6] + c12 * t[7] + c22 * t[8]) * inv;
            sx[0] += d0; sx[1] += d1; sx[2] += d2;
            if (sqrt((d0 * d0 + d1 * d1 + d2 * d2) / 3.0) < 1e-8) s_done = 1;
        }
        __syncthreads();
        if (s_done) break;
    }
    if (tid < 3) xout[tid] = sx[tid];
    if (tid == 0) *iters = s_done ? it + 1 : -max_iter;
}

}  // namespace

VEL_API int vel_triangulate_nv(const double* A, const double* U, int32_t nf, int32_t nv, double* C0, vel_stream_t stream)
{
    VEL_CHECK_ARG(A && U && C0, "vel_triangulate_nv: NULL argument");
    VEL_CHECK_ARG(nf >= 1 && nv >= 0, "vel_triangulate_nv: bad sizes nf=%d nv=%d", nf, nv);
    if (nv == 0) return VEL_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int pblocks = (nv + TRI_THREADS - 1) / TRI_THREADS;
    // enough frame chunks to fill the machine (~4 CTAs per SM), at least 8 frames each
    int nchunks = (4 * kNumSMs + pblocks - 1) / pblocks;
    if (nchunks > (nf + 7) / 8) nchunks = (nf + 7) / 8;
    if (nchunks < 1) nchunks = 1;
    const int f_per_chunk = (nf + nchunks - 1) / nchunks;
    nchunks = (nf + f_per_chunk - 1) / f_per_chunk;
    double* part = nullptr;
    vel_keep_async_pool_cached();
    VEL_CUDA(cudaMallocAsync((void**)&part, sizeof(double) * TRI_NACC * (size_t)nv * nchunks, st));
    tri_nv_partial_kernel<<<dim3(pblocks, nchunks), TRI_THREADS, 0, st>>>(A, U, nf, nv, f_per_chunk, part);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) {
        tri_nv_solve_kernel<<<pblocks, TRI_THREADS, 0, st>>>(part, nv, nchunks, C0);
        e = cudaGetLastError();
    }
    cudaFreeAsync(part, st);
    if (e != cudaSuccess) {
        vel_set_error("vel_triangulate_nv: %s", cudaGetErrorString(e));
        return VEL_ERR_CUDA;
    }
    return VEL_OK;
}

VEL_API int vel_triangulate_2v(const double* A, const double* U, int32_t nf, int32_t nv, double* C0, vel_stream_t stream)
{
    VEL_CHECK_ARG(A && U && C0, "vel_triangulate_2v: NULL argument");
    VEL_CHECK_ARG(nf >= 2 && nv >= 0, "vel_triangulate_2v: need nf >= 2 (got %d), nv >= 0 (got %d)", nf, nv);
    if (nv == 0) return VEL_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int pblocks = (nv + TRI_THREADS - 1) / TRI_THREADS;
    // enough chunks to fill the machine (~4 CTAs per SM) but never more than nf-1
    int nchunks = (4 * kNumSMs + pblocks - 1) / pblocks;
    if (nchunks > nf - 1) nchunks = nf - 1;
    if (nchunks < 1) nchunks = 1;
    const int j_per_chunk = (nf - 1 + nchunks - 1) / nchunks;
    nchunks = (nf - 1 + j_per_chunk - 1) / j_per_chunk;
    double* part = nullptr;
    vel_keep_async_pool_cached();
    VEL_CUDA(cudaMallocAsync((void**)&part, sizeof(double) * 3ull * nv * nchunks, st));
    tri_2v_partial_kernel<<<dim3(pblocks, nchunks), TRI_THREADS, 0, st>>>(A, U, nf, nv, j_per_chunk, part);
    VEL_LAUNCH_CHECK("tri_2v_partial_kernel");
    tri_2v_finalize_kernel<<<pblocks, TRI_THREADS, 0, st>>>(A, part, nf, nv, nchunks, C0);
    VEL_LAUNCH_CHECK("tri_2v_finalize_kernel");
    VEL_CUDA(cudaFreeAsync(part, st));
    return VEL_OK;
}

VEL_API int vel_msv1_t(const double* K, const double* A_fixed, const double* U, int32_t nf, int32_t ng, const double* z,
                       const double* x0, int32_t max_iter, double* x, double* b0, int32_t* iters, vel_stream_t stream)
{
    VEL_CHECK_ARG(K && A_fixed && U && z && x0 && x && b0 && iters, "vel_msv1_t: NULL argument");
    VEL_CHECK_ARG(nf >= 2 && ng >= 1 && max_iter >= 1, "vel_msv1_t: bad sizes nf=%d ng=%d max_iter=%d", nf, ng, max_iter);
    msv1_t_kernel<<<1, MSV_THREADS, 0, (cudaStream_t)stream>>>(K, A_fixed, U, nf, ng, z, x0, max_iter, x, b0, iters);
    VEL_LAUNCH_CHECK("msv1_t_kernel");
    return VEL_OK;
}
