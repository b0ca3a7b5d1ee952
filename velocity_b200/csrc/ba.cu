// K7 / K8: bundle adjustment of fcnNLS_batch (utils/NLS.py:186-250) in block-sparse form.
//
// The reference builds a DENSE forward-difference Jacobian JT [nx][nz] by re-evaluating the whole
// projection once per parameter (nx evaluations of nz measurements, 277 GB at the BASELINE size) and
// inverts the dense nx x nx system.  Each measurement depends on one point (3 params) and one camera
// (6 params: position, roll/pitch/yaw; camera 0 is fixed at identity), so JtJ has the block layout
//      [ V  W^T ]   V = blockdiag(V_i 3x3),  U = blockdiag(U_j 6x6),  W_ji 6x3
//      [ W  U   ]
// K7 accumulates exactly the non-zero blocks, with the SAME forward-difference Jacobian entries
// (perturb one parameter by 1e-6, re-project, subtract, divide) the reference computes, so the two
// linear systems are identical up to float64 summation order.  K8 solves (JtJ + I) delta = Jt r by
// Schur complement on the point blocks and applies x += 0.9 * delta (utils/NLS.py:234-235).
//
// Layouts:  x = [points nt*3 | camera positions nc*3 | camera rpy nc*3]  (utils/NLS.py:203)
//           z = [2][nc+1][nt]  (all u then all v, camera-major, track fastest, :198-199)
//           V [nt][6] upper triangles, U [nc][21] upper triangles (pos,rpy order),
//           W [nc*6][nt*3] row-major (row 6*(c-1)+a, column 3*i+b), g [nt*3 + nc*6] in x order.
// Cameras are indexed 0..nc (0 = the fixed one); a rank owning cameras [cam_first, cam_first +
// cam_count) fills its rows of U/W/g_c and PARTIAL V/g_p/cost over its cameras (all-reduce across
// ranks).
#ifdef VEL_WITH_VENDOR_SOLVER   // optional A/B build (python -m velocity_b200.build --vendor-solver): cuBLAS / cuSOLVER behind VEL_BA_SOLVER=vendor
#include <cublas_v2.h>
#include <cusolverDn.h>
#endif
#include <stdlib.h>
#include <string.h>

#include "ba_math.cuh"

namespace {

// per-camera constants for the point kernel: R0 (9) + pos (3); camera 0 = identity / zero
__global__ void ba_cam_setup_kernel(const double* __restrict__ x, int nt, int nc, double* __restrict__ cams)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c > nc) return;
    double* o = cams + 12ll * c;
    if (c == 0) {
        o[0] = 1; o[1] = 0; o[2] = 0; o[3] = 0; o[4] = 1; o[5] = 0; o[6] = 0; o[7] = 0; o[8] = 1; o[9] = 0; o[10] = 0; o[11] = 0;
        return;
    }
    const double* pos = x + 3ll * nt + 3ll * (c - 1);
    const double* rpy = x + 3ll * nt + 3ll * nc + 3ll * (c - 1);
    rpy2dcm(rpy, o);
    o[9] = pos[0]; o[10] = pos[1]; o[11] = pos[2];
}

// Jacobian of one observation wrt its point: forward differences exactly as the reference forms them
// (camera 0 projects the point itself: alist[0] = pw, utils/NLS.py:208)
__device__ __forceinline__ void point_jacobian(const double* K, const double* R, const double* pos, bool fixed_cam, double X,
                                               double Y, double Z, double& u0, double& v0, double (&ju)[3], double (&jv)[3])
{
    double ax, ay, az, u, v;
    if (fixed_cam) {
        project(K, X, Y, Z, u0, v0);
        project(K, X + JDX, Y, Z, u, v); ju[0] = (u - u0) / JDX; jv[0] = (v - v0) / JDX;
        project(K, X, Y + JDX, Z, u, v); ju[1] = (u - u0) / JDX; jv[1] = (v - v0) / JDX;
        project(K, X, Y, Z + JDX, u, v); ju[2] = (u - u0) / JDX; jv[2] = (v - v0) / JDX;
        return;
    }
    rot(R, X, Y, Z, ax, ay, az);
    project(K, ax + pos[0], ay + pos[1], az + pos[2], u0, v0);
    rot(R, X + JDX, Y, Z, ax, ay, az);
    project(K, ax + pos[0], ay + pos[1], az + pos[2], u, v); ju[0] = (u - u0) / JDX; jv[0] = (v - v0) / JDX;
    rot(R, X, Y + JDX, Z, ax, ay, az);
    project(K, ax + pos[0], ay + pos[1], az + pos[2], u, v); ju[1] = (u - u0) / JDX; jv[1] = (v - v0) / JDX;
    rot(R, X, Y, Z + JDX, ax, ay, az);
    project(K, ax + pos[0], ay + pos[1], az + pos[2], u, v); ju[2] = (u - u0) / JDX; jv[2] = (v - v0) / JDX;
}

// ---- camera kernel: one CTA per parameterised camera ------------------------------------------------
__global__ void __launch_bounds__(CAM_THREADS)
ba_camera_kernel(const double* __restrict__ Kg, const double* __restrict__ x, const double* __restrict__ z, int nt, int nc,
                 int cam_first, double* __restrict__ U, double* __restrict__ W, double* __restrict__ g)
{
    constexpr int NACC = 21 + 6;
    constexpr int NWARP = CAM_THREADS / 32;
    __shared__ double sK[9], sR[4][9], sP[3];
    __shared__ double sred[NWARP][NACC];
    const int c = cam_first + blockIdx.x;  // camera index, >= 1
    const int tid = threadIdx.x;
    const double* pos = x + 3ll * nt + 3ll * (c - 1);
    const double* rpy = x + 3ll * nt + 3ll * nc + 3ll * (c - 1);
    if (tid < 9) sK[tid] = Kg[tid];
    if (tid < 3) sP[tid] = pos[tid];
    if (tid < 4) {
        double r[3] = {rpy[0], rpy[1], rpy[2]};
        if (tid > 0) r[tid - 1] = r[tid - 1] + JDX;
        rpy2dcm(r, sR[tid]);
    }
    __syncthreads();
    const double* zu = z + (long long)c * nt;
    const double* zv = z + (long long)(nc + 1) * nt + (long long)c * nt;

    double acc[NACC];
#pragma unroll
    for (int k = 0; k < NACC; ++k) acc[k] = 0.0;
    for (int i = tid; i < nt; i += CAM_THREADS) {
        const double X = x[3ll * i], Y = x[3ll * i + 1], Z = x[3ll * i + 2];
        double u0, v0, pu[3], pv[3];
        point_jacobian(sK, sR[0], sP, false, X, Y, Z, u0, v0, pu, pv);
        double cu[6], cv[6], ax, ay, az, u, v;
        rot(sR[0], X, Y, Z, ax, ay, az);
        project(sK, ax + (sP[0] + JDX), ay + sP[1], az + sP[2], u, v); cu[0] = (u - u0) / JDX; cv[0] = (v - v0) / JDX;
        project(sK, ax + sP[0], ay + (sP[1] + JDX), az + sP[2], u, v); cu[1] = (u - u0) / JDX; cv[1] = (v - v0) / JDX;
        project(sK, ax + sP[0], ay + sP[1], az + (sP[2] + JDX), u, v); cu[2] = (u - u0) / JDX; cv[2] = (v - v0) / JDX;
#pragma unroll
        for (int m = 1; m < 4; ++m) {
            rot(sR[m], X, Y, Z, ax, ay, az);
            project(sK, ax + sP[0], ay + sP[1], az + sP[2], u, v);
            cu[2 + m] = (u - u0) / JDX; cv[2 + m] = (v - v0) / JDX;
        }
        const double ru = zu[i] - u0, rv = zv[i] - v0;
        int k = 0;
#pragma unroll
        for (int r = 0; r < 6; ++r)
#pragma unroll
            for (int q = r; q < 6; ++q) acc[k++] += cu[r] * cu[q] + cv[r] * cv[q];
#pragma unroll
        for (int r = 0; r < 6; ++r) acc[21 + r] += cu[r] * ru + cv[r] * rv;
        if (W) {
#pragma unroll
            for (int a = 0; a < 6; ++a) {
                double* w = W + ((long long)(6 * (c - 1) + a) * nt + i) * 3;
                w[0] = cu[a] * pu[0] + cv[a] * pv[0];
                w[1] = cu[a] * pu[1] + cv[a] * pv[1];
                w[2] = cu[a] * pu[2] + cv[a] * pv[2];
            }
        }
    }
#pragma unroll
    for (int k = 0; k < NACC; ++k) acc[k] = warp_sum(acc[k]);
    if ((tid & 31) == 0) {
#pragma unroll
        for (int k = 0; k < NACC; ++k) sred[tid >> 5][k] = acc[k];
    }
    __syncthreads();
    if (tid < NACC) {
        double s = 0.0;
        for (int w = 0; w < NWARP; ++w) s += sred[w][tid];
        if (tid < 21) U[21ll * (c - 1) + tid] = s;
        else {
            const int r = tid - 21;  // g order: cam positions block, then cam rpy block
            const long long idx = r < 3 ? 3ll * nt + 3ll * (c - 1) + r : 3ll * nt + 3ll * nc + 3ll * (c - 1) + (r - 3);
            g[idx] = s;
        }
    }
}

// ---- point kernel: thread = (point, camera chunk); a second kernel adds the chunks in a fixed order ----------------
// (one thread per point looping over all cameras is a single 300-trip dependent FP64 chain on 32 CTAs: latency-bound)
constexpr int PT_MAX_CHUNKS = 32;

__global__ void __launch_bounds__(PT_THREADS)
ba_point_kernel(const double* __restrict__ Kg, const double* __restrict__ x, const double* __restrict__ z,
                const double* __restrict__ cams, int nt, int nc, int cam_first, int cam_count, int nchunks, double* __restrict__ part)
{
    __shared__ double sK[9];
    const int tid = threadIdx.x, i = blockIdx.x * PT_THREADS + tid, chunk = blockIdx.y;
    if (tid < 9) sK[tid] = Kg[tid];
    __syncthreads();
    if (i >= nt) return;
    const int per = (cam_count + nchunks - 1) / nchunks;
    const int c0 = cam_first + chunk * per, c1 = min(cam_first + cam_count, c0 + per);
    const double X = x[3ll * i], Y = x[3ll * i + 1], Z = x[3ll * i + 2];
    double v00 = 0, v01 = 0, v02 = 0, v11 = 0, v12 = 0, v22 = 0, g0 = 0, g1 = 0, g2 = 0, cost = 0;
    for (int c = c0; c < c1; ++c) {
        const double* cm = cams + 12ll * c;
        double u0, w0, ju[3], jv[3];
        point_jacobian(sK, cm, cm + 9, c == 0, X, Y, Z, u0, w0, ju, jv);
        const double ru = z[(long long)c * nt + i] - u0;
        const double rv = z[(long long)(nc + 1) * nt + (long long)c * nt + i] - w0;
        v00 += ju[0] * ju[0] + jv[0] * jv[0]; v01 += ju[0] * ju[1] + jv[0] * jv[1]; v02 += ju[0] * ju[2] + jv[0] * jv[2];
        v11 += ju[1] * ju[1] + jv[1] * jv[1]; v12 += ju[1] * ju[2] + jv[1] * jv[2]; v22 += ju[2] * ju[2] + jv[2] * jv[2];
        g0 += ju[0] * ru + jv[0] * rv; g1 += ju[1] * ru + jv[1] * rv; g2 += ju[2] * ru + jv[2] * rv;
        cost += ru * ru + rv * rv;
    }
    double* o = part + (long long)chunk * 10 * nt + i;     // [chunk][10][nt]: coalesced across the points of a CTA
    o[0] = v00; o[(long long)nt] = v01; o[2ll * nt] = v02; o[3ll * nt] = v11; o[4ll * nt] = v12; o[5ll * nt] = v22;
    o[6ll * nt] = g0; o[7ll * nt] = g1; o[8ll * nt] = g2; o[9ll * nt] = cost;
}

__global__ void __launch_bounds__(PT_THREADS)
ba_point_reduce_kernel(const double* __restrict__ part, int nt, int nchunks, double* __restrict__ V, double* __restrict__ g,
                       double* __restrict__ cost_part)
{
    __shared__ double sred[PT_THREADS / 32];
    const int tid = threadIdx.x, i = blockIdx.x * PT_THREADS + tid;
    double cost = 0.0;
    if (i < nt) {
        double a[10];
#pragma unroll
        for (int k = 0; k < 10; ++k) a[k] = 0.0;
        for (int ch = 0; ch < nchunks; ++ch) {
            const double* o = part + (long long)ch * 10 * nt + i;
#pragma unroll
            for (int k = 0; k < 10; ++k) a[k] += o[(long long)k * nt];
        }
        double* v = V + 6ll * i;
#pragma unroll
        for (int k = 0; k < 6; ++k) v[k] = a[k];
        g[3ll * i] = a[6]; g[3ll * i + 1] = a[7]; g[3ll * i + 2] = a[8];
        cost = a[9];
    }
    cost = warp_sum(cost);
    if ((tid & 31) == 0) sred[tid >> 5] = cost;
    __syncthreads();
    if (tid == 0) {
        double s = 0.0;
        for (int w = 0; w < PT_THREADS / 32; ++w) s += sred[w];
        cost_part[blockIdx.x] = s;
    }
}

__global__ void ba_cost_finalize_kernel(const double* __restrict__ part, int n, double* __restrict__ cost)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double s = 0.0;
        for (int k = 0; k < n; ++k) s += part[k];
        *cost = s;
    }
}

// ---- solve ----------------------------------------------------------------------------------------------
// per point: L_i = chol((V_i + I)^-1) so that W' = W * blockdiag(L_i) gives W' W'^T = W (V+I)^-1 W^T;
// also y_i = (V_i + I)^-1 g_p,i
__global__ void __launch_bounds__(PT_THREADS)
ba_point_prep_kernel(const double* __restrict__ V, const double* __restrict__ g, int nt, double* __restrict__ Vinv,
                     double* __restrict__ Lf, double* __restrict__ y)
{
    const int i = blockIdx.x * PT_THREADS + threadIdx.x;
    if (i >= nt) return;
    const double* v = V + 6ll * i;
    const double m00 = v[0] + 1.0, m01 = v[1], m02 = v[2], m11 = v[3] + 1.0, m12 = v[4], m22 = v[5] + 1.0;
    const double c00 = m11 * m22 - m12 * m12, c01 = m02 * m12 - m01 * m22, c02 = m01 * m12 - m02 * m11;
    const double c11 = m00 * m22 - m02 * m02, c12 = m01 * m02 - m00 * m12, c22 = m00 * m11 - m01 * m01;
    const double inv = 1.0 / (m00 * c00 + m01 * c01 + m02 * c02);
    const double i00 = c00 * inv, i01 = c01 * inv, i02 = c02 * inv, i11 = c11 * inv, i12 = c12 * inv, i22 = c22 * inv;
    double* o = Vinv + 6ll * i;
    o[0] = i00; o[1] = i01; o[2] = i02; o[3] = i11; o[4] = i12; o[5] = i22;
    // lower Cholesky factor of the (SPD) inverse
    const double l00 = sqrt(i00), l10 = i01 / l00, l20 = i02 / l00;
    const double l11 = sqrt(i11 - l10 * l10), l21 = (i12 - l20 * l10) / l11;
    const double l22 = sqrt(i22 - l20 * l20 - l21 * l21);
    double* l = Lf + 6ll * i;
    l[0] = l00; l[1] = l10; l[2] = l11; l[3] = l20; l[4] = l21; l[5] = l22;
    const double g0 = g[3ll * i], g1 = g[3ll * i + 1], g2 = g[3ll * i + 2];
    y[3ll * i] = i00 * g0 + i01 * g1 + i02 * g2;
    y[3ll * i + 1] = i01 * g0 + i11 * g1 + i12 * g2;
    y[3ll * i + 2] = i02 * g0 + i12 * g1 + i22 * g2;
}

// W'[row][3i..3i+2] = W[row][3i..3i+2] * L_i   (row vector times lower-triangular 3x3)
__global__ void __launch_bounds__(256)
ba_scale_w_kernel(const double* __restrict__ W, const double* __restrict__ Lf, int nrows, int nt, double* __restrict__ Wp, long long ldo)
{
    const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
    if (idx >= (long long)nrows * nt) return;
    const int i = (int)(idx % nt);
    const double* w = W + idx * 3;
    const double* l = Lf + 6ll * i;
    const double w0 = w[0], w1 = w[1], w2 = w[2];
    double* o = Wp + (idx / nt) * ldo + 3ll * i;
    o[0] = w0 * l[0] + w1 * l[1] + w2 * l[3];
    o[1] = w1 * l[2] + w2 * l[4];
    o[2] = w2 * l[5];
}

// S (col-major n6 x n6, full) = blockdiag(U_j + I); rhs = g_c
__global__ void ba_init_s_kernel(const double* __restrict__ U, const double* __restrict__ g, int nt, int nc,
                                 double* __restrict__ S, double* __restrict__ rhs)
{
    const int n6 = 6 * nc;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)n6 * n6) return;
    const int r = (int)(idx % n6), c = (int)(idx / n6);
    double v = 0.0;
    if (r / 6 == c / 6) {
        // S ordering is camera-major: row 6*j + a, a in (pos xyz, rpy xyz)
        const int j = r / 6, a = r % 6, b = c % 6;
        const int lo = a < b ? a : b, hi = a < b ? b : a;
        const int k = lo * 6 - lo * (lo - 1) / 2 + (hi - lo);
        v = U[21ll * j + k] + (a == b ? 1.0 : 0.0);
    }
    S[idx] = v;
    if (c == 0) {
        const int j = r / 6, a = r % 6;
        rhs[r] = a < 3 ? g[3ll * nt + 3ll * j + a] : g[3ll * nt + 3ll * nc + 3ll * j + (a - 3)];
    }
}

// delta_p,i = y_i - (V_i+I)^-1 t_i with t = W^T delta_c;  x += 0.9*delta;  partial sum of delta^2
__global__ void __launch_bounds__(PT_THREADS)
ba_update_kernel(const double* __restrict__ Vinv, const double* __restrict__ y, const double* __restrict__ t,
                 const double* __restrict__ dc, int nt, int nc, double* __restrict__ x, double* __restrict__ ss_part)
{
    __shared__ double sred[PT_THREADS / 32];
    const int tid = threadIdx.x;
    const long long i = (long long)blockIdx.x * PT_THREADS + tid;
    double ss = 0.0;
    if (i < nt) {
        const double* v = Vinv + 6 * i;
        const double t0 = t[3 * i], t1 = t[3 * i + 1], t2 = t[3 * i + 2];
        const double d0 = y[3 * i] - (v[0] * t0 + v[1] * t1 + v[2] * t2);
        const double d1 = y[3 * i + 1] - (v[1] * t0 + v[3] * t1 + v[4] * t2);
        const double d2 = y[3 * i + 2] - (v[2] * t0 + v[4] * t1 + v[5] * t2);
        x[3 * i] += 0.9 * d0; x[3 * i + 1] += 0.9 * d1; x[3 * i + 2] += 0.9 * d2;
        ss = 0.81 * (d0 * d0 + d1 * d1 + d2 * d2);
    } else if (i < nt + 6ll * nc) {
        // camera part: dc is camera-major (6 per camera), x is [pos block | rpy block]
        const int r = (int)(i - nt), j = r / 6, a = r % 6;
        const double d = dc[r];
        const long long xi = a < 3 ? 3ll * nt + 3ll * j + a : 3ll * nt + 3ll * nc + 3ll * j + (a - 3);
        x[xi] += 0.9 * d;
        ss = 0.81 * d * d;
    }
    ss = warp_sum(ss);
    if ((tid & 31) == 0) sred[tid >> 5] = ss;
    __syncthreads();
    if (tid == 0) {
        double s = 0.0;
        for (int w = 0; w < PT_THREADS / 32; ++w) s += sred[w];
        ss_part[blockIdx.x] = s;
    }
}

// info != 0 (the Cholesky of the reduced system failed: not positive definite to working precision) turns rms(delta)
// into NaN, so that the caller's convergence test cannot mistake a garbage update for progress
__global__ void ba_rms_finalize_kernel(const double* __restrict__ part, int n, long long nx, double* __restrict__ rms,
                                       const int* __restrict__ info)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double s = 0.0;
        for (int k = 0; k < n; ++k) s += part[k];
        *rms = (info && *info != 0) ? __longlong_as_double(0x7ff8000000000000ll) : sqrt(s / (double)nx);
    }
}

// The dense part of the solve (K8) runs on this library's own kernels (csrc/dense_f64.cu: FP64 tensor-core SYRK, task-graph
// Cholesky, hand-written GEMVs); the shipped library links no vendor BLAS / solver.  A build with -DVEL_WITH_VENDOR_SOLVER
// (python -m velocity_b200.build --vendor-solver) additionally compiles the cuBLAS DSYRK / cuSOLVER DPOTRF path of round 1,
// selected at run time by VEL_BA_SOLVER=vendor, for A/B measurements (DESIGN.md has the numbers).
bool native_solver()
{
#ifdef VEL_WITH_VENDOR_SOLVER
    const char* e = getenv("VEL_BA_SOLVER");
    return !(e && strcmp(e, "vendor") == 0);
#else
    return true;
#endif
}

constexpr int GEMV_ROW_CHUNKS = 8;

// rhs[r] -= sum_k W[r][k] y[k]   (one CTA per row, fixed-order reduction)
__global__ void __launch_bounds__(256)
ba_gemv_rows_sub_kernel(const double* __restrict__ W, long long ld, int n3, const double* __restrict__ y, double* __restrict__ rhs)
{
    __shared__ double sred[8];
    const double* w = W + (long long)blockIdx.x * ld;
    double s = 0.0;
    for (int k = threadIdx.x; k < n3; k += 256) s += w[k] * y[k];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int q = 0; q < 8; ++q) t += sred[q];
        rhs[blockIdx.x] -= t;
    }
}

// tpart[c][k] = sum over the rows of chunk c of W[r][k] dc[r]   (thread per column, coalesced across k)
__global__ void __launch_bounds__(256)
ba_gemv_cols_kernel(const double* __restrict__ W, long long ld, int nrows, int n3, const double* __restrict__ dc, double* __restrict__ tpart)
{
    const int k = blockIdx.x * 256 + threadIdx.x;
    if (k >= n3) return;
    const int per = (nrows + GEMV_ROW_CHUNKS - 1) / GEMV_ROW_CHUNKS;
    const int r0 = blockIdx.y * per, r1 = min(nrows, r0 + per);
    double s = 0.0;
#pragma unroll 4
    for (int r = r0; r < r1; ++r) s += W[(long long)r * ld + k] * dc[r];
    tpart[(long long)blockIdx.y * n3 + k] = s;
}

__global__ void ba_gemv_cols_reduce_kernel(const double* __restrict__ tpart, int n3, double* __restrict__ t)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n3) return;
    double s = 0.0;
    for (int c = 0; c < GEMV_ROW_CHUNKS; ++c) s += tpart[(long long)c * n3 + k];
    t[k] = s;
}

// zero the padding columns [n3, ld) of W'
__global__ void ba_zero_pad_kernel(double* __restrict__ Wp, long long ld, int nrows, int n3)
{
    const int pad = (int)(ld - n3);
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (pad <= 0 || idx >= (long long)nrows * pad) return;
    Wp[(idx / pad) * ld + n3 + idx % pad] = 0.0;
}

#ifdef VEL_WITH_VENDOR_SOLVER
struct Handles {
    cublasHandle_t blas = nullptr;
    cusolverDnHandle_t solver = nullptr;
    int device = -1;
};

// one pair of handles per host thread AND device (a handle is bound to the device it was created on)
Handles* handles()
{
    static thread_local Handles h[16];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return nullptr;
    Handles& hd = h[dev];
    if (!hd.blas || !hd.solver) {
        if (!hd.blas && cublasCreate(&hd.blas) != CUBLAS_STATUS_SUCCESS) { hd.blas = nullptr; return nullptr; }
        if (!hd.solver && cusolverDnCreate(&hd.solver) != CUSOLVER_STATUS_SUCCESS) { hd.solver = nullptr; return nullptr; }
        hd.device = dev;
    }
    return &hd;
}
#endif

inline size_t align256(size_t b) { return (b + 255) & ~(size_t)255; }

struct SolveLayout {
    size_t off_wp, off_s, off_vinv, off_l, off_y, off_rhs, off_t, off_part, off_info, off_flags, off_tpart, off_potrf, total, ldw;
    int lwork;
};

bool solve_layout(int nt, int nc, SolveLayout* L, bool query_potrf)
{
    const size_t n6 = 6ull * nc, n3 = 3ull * nt;
    size_t o = 0;
    L->ldw = (n3 + 31) / 32 * 32;                          // the native SYRK wants rows zero-padded to its k-tile
    L->off_wp = o; o += align256(sizeof(double) * n6 * L->ldw);
    L->off_s = o; o += align256(sizeof(double) * n6 * n6);
    L->off_vinv = o; o += align256(sizeof(double) * 6 * nt);
    L->off_l = o; o += align256(sizeof(double) * 6 * nt);
    L->off_y = o; o += align256(sizeof(double) * n3);
    L->off_rhs = o; o += align256(sizeof(double) * (n6 + 8));
    L->off_t = o; o += align256(sizeof(double) * n3);
    L->off_part = o; o += align256(sizeof(double) * ((nt + n6) / PT_THREADS + 2));
    L->off_info = o; o += 256;
    L->off_flags = o; o += vel_dense_syrk_workspace((int)(n6 > 0 ? n6 : 1), (int)(n3 > 0 ? n3 : 1));
    L->off_tpart = o; o += align256(sizeof(double) * GEMV_ROW_CHUNKS * n3);
    L->lwork = 0;
#ifdef VEL_WITH_VENDOR_SOLVER
    if (query_potrf && nc > 0 && !native_solver()) {
        Handles* h = handles();
        if (!h) return false;
        int lwork = 0;
        if (cusolverDnDpotrf_bufferSize(h->solver, CUBLAS_FILL_MODE_LOWER, (int)n6, nullptr, (int)n6, &lwork) != CUSOLVER_STATUS_SUCCESS)
            return false;
        L->lwork = lwork;
    }
#else
    (void)query_potrf;
#endif
    L->off_potrf = o; o += align256(sizeof(double) * (size_t)(L->lwork > 0 ? L->lwork : 1));
    L->total = o;
    return true;
}

}  // namespace

VEL_API int vel_ba_accumulate(const double* K, const double* x, const double* z, int32_t nt, int32_t nc, int32_t cam_first,
                              int32_t cam_count, double* V, double* U, double* W, double* g, double* cost, vel_stream_t stream)
{
    VEL_CHECK_ARG(K && x && z && V && U && g && cost, "vel_ba_accumulate: NULL argument");
    VEL_CHECK_ARG(nt > 0 && nc >= 0, "vel_ba_accumulate: bad sizes nt=%d nc=%d", nt, nc);
    VEL_CHECK_ARG(cam_first >= 0 && cam_count >= 0 && cam_first + cam_count <= nc + 1,
                  "vel_ba_accumulate: camera slice [%d,%d) outside [0,%d]", cam_first, cam_first + cam_count, nc);
    cudaStream_t st = (cudaStream_t)stream;
    const int pblocks = (nt + PT_THREADS - 1) / PT_THREADS;
    vel_keep_async_pool_cached();
    const int nchunks = cam_count >= 16 ? (cam_count / 8 < PT_MAX_CHUNKS ? cam_count / 8 : PT_MAX_CHUNKS) : 1;
    double* tmp = nullptr;
    VEL_CUDA(cudaMallocAsync((void**)&tmp, sizeof(double) * (12ull * (nc + 1) + pblocks + 10ull * nchunks * nt), st));
    double* cams = tmp;
    double* cost_part = tmp + 12ull * (nc + 1);
    double* part = cost_part + pblocks;
    ba_cam_setup_kernel<<<(nc + 1 + 127) / 128, 128, 0, st>>>(x, nt, nc, cams);
    VEL_LAUNCH_CHECK("ba_cam_setup_kernel");
    const int first_param = cam_first < 1 ? 1 : cam_first;
    const int n_param = cam_first + cam_count - first_param;
    if (n_param > 0) {
        ba_camera_kernel<<<n_param, CAM_THREADS, 0, st>>>(K, x, z, nt, nc, first_param, U, W, g);
        VEL_LAUNCH_CHECK("ba_camera_kernel");
    }
    ba_point_kernel<<<dim3(pblocks, nchunks), PT_THREADS, 0, st>>>(K, x, z, cams, nt, nc, cam_first, cam_count, nchunks, part);
    VEL_LAUNCH_CHECK("ba_point_kernel");
    ba_point_reduce_kernel<<<pblocks, PT_THREADS, 0, st>>>(part, nt, nchunks, V, g, cost_part);
    VEL_LAUNCH_CHECK("ba_point_reduce_kernel");
    ba_cost_finalize_kernel<<<1, 32, 0, st>>>(cost_part, pblocks, cost);
    VEL_LAUNCH_CHECK("ba_cost_finalize_kernel");
    VEL_CUDA(cudaFreeAsync(tmp, st));
    return VEL_OK;
}

VEL_API size_t vel_ba_solve_workspace(int32_t nt, int32_t nc)
{
    if (nt <= 0 || nc < 0) return 0;
    SolveLayout L;
    if (!solve_layout(nt, nc, &L, true)) { vel_set_error("vel_ba_solve_workspace: cuSOLVER/cuBLAS handle creation failed"); return 0; }
    return L.total;
}

// internal entry points of dense_f64.cu (declared in the public header as well)
// The dense part of the solve with this library's own kernels:  S(lower) -= W' W'^T;  rhs -= W y;  S delta_c = rhs;  t = W^T delta_c.
// W' [nrows][ldw] zero-padded, W [nrows][n3] the unscaled cross blocks.  info (device) is raised by a failed Cholesky.
static int dense_solve_native(const double* W, const double* Wp, long long ldw, int nrows, int n3, double* S, const double* y, double* rhs,
                              double* t, double* tpart, void* flags, size_t flags_bytes, int* info, cudaStream_t st)
{
    int rc = vel_syrk_lower_sub(Wp, ldw, nrows, n3, S, nrows, flags, flags_bytes, (vel_stream_t)st);
    if (rc != VEL_OK) return rc;
    ba_gemv_rows_sub_kernel<<<nrows, 256, 0, st>>>(W, n3, n3, y, rhs);
    VEL_LAUNCH_CHECK("ba_gemv_rows_sub_kernel");
    rc = vel_spd_solve(S, nrows, nrows, rhs, info, (vel_stream_t)st);
    if (rc != VEL_OK) return rc;
    ba_gemv_cols_kernel<<<dim3((n3 + 255) / 256, GEMV_ROW_CHUNKS), 256, 0, st>>>(W, n3, nrows, n3, rhs, tpart);
    VEL_LAUNCH_CHECK("ba_gemv_cols_kernel");
    ba_gemv_cols_reduce_kernel<<<(n3 + 255) / 256, 256, 0, st>>>(tpart, n3, t);
    VEL_LAUNCH_CHECK("ba_gemv_cols_reduce_kernel");
    return VEL_OK;
}

VEL_API int vel_ba_solve(const double* V, const double* U, const double* W, const double* g, int32_t nt, int32_t nc, double* x,
                         double* rms_delta, void* work, size_t work_bytes, vel_stream_t stream)
{
    VEL_CHECK_ARG(V && g && x && rms_delta && work, "vel_ba_solve: NULL argument");
    VEL_CHECK_ARG(nt > 0 && nc >= 0, "vel_ba_solve: bad sizes nt=%d nc=%d", nt, nc);
    VEL_CHECK_ARG(nc == 0 || (U && W), "vel_ba_solve: U and W are required when nc > 0");
    SolveLayout L;
    VEL_CHECK_ARG(solve_layout(nt, nc, &L, true), "vel_ba_solve: cuSOLVER/cuBLAS handle creation failed");
    VEL_CHECK_ARG(work_bytes >= L.total, "vel_ba_solve: workspace %zu B < required %zu B", work_bytes, L.total);
    cudaStream_t st = (cudaStream_t)stream;
    char* wb = (char*)work;
    double* Wp = (double*)(wb + L.off_wp);
    double* S = (double*)(wb + L.off_s);
    double* Vinv = (double*)(wb + L.off_vinv);
    double* Lf = (double*)(wb + L.off_l);
    double* y = (double*)(wb + L.off_y);
    double* rhs = (double*)(wb + L.off_rhs);
    double* t = (double*)(wb + L.off_t);
    double* part = (double*)(wb + L.off_part);
    int* info = (int*)(wb + L.off_info);
    double* potrf_work = (double*)(wb + L.off_potrf);
    const int n6 = 6 * nc, n3 = 3 * nt;
    const int pblocks = (nt + PT_THREADS - 1) / PT_THREADS;

    const bool native = native_solver();
    VEL_CUDA(cudaMemsetAsync(info, 0, sizeof(int), st));
    ba_point_prep_kernel<<<pblocks, PT_THREADS, 0, st>>>(V, g, nt, Vinv, Lf, y);
    VEL_LAUNCH_CHECK("ba_point_prep_kernel");
    if (nc > 0 && native) {
        const long long nel = (long long)n6 * nt;
        ba_scale_w_kernel<<<(unsigned)((nel + 255) / 256), 256, 0, st>>>(W, Lf, n6, nt, Wp, (long long)L.ldw);
        VEL_LAUNCH_CHECK("ba_scale_w_kernel");
        if ((int)L.ldw > n3) {
            const long long npad = (long long)n6 * (L.ldw - n3);
            ba_zero_pad_kernel<<<(unsigned)((npad + 255) / 256), 256, 0, st>>>(Wp, (long long)L.ldw, n6, n3);
            VEL_LAUNCH_CHECK("ba_zero_pad_kernel");
        }
        const long long ns = (long long)n6 * n6;
        ba_init_s_kernel<<<(unsigned)((ns + 255) / 256), 256, 0, st>>>(U, g, nt, nc, S, rhs);
        VEL_LAUNCH_CHECK("ba_init_s_kernel");
        const int rc = dense_solve_native(W, Wp, (long long)L.ldw, n6, n3, S, y, rhs, t, (double*)(wb + L.off_tpart), wb + L.off_flags,
                                          L.off_tpart - L.off_flags, info, st);
        if (rc != VEL_OK) return rc;
#ifdef VEL_WITH_VENDOR_SOLVER
    } else if (nc > 0) {
        Handles* h = handles();
        VEL_CHECK_ARG(h != nullptr, "vel_ba_solve: cuBLAS/cuSOLVER handles unavailable");
        cublasSetStream(h->blas, st);
        cusolverDnSetStream(h->solver, st);
        cublasSetPointerMode(h->blas, CUBLAS_POINTER_MODE_HOST);
        const long long nel = (long long)n6 * nt;
        ba_scale_w_kernel<<<(unsigned)((nel + 255) / 256), 256, 0, st>>>(W, Lf, n6, nt, Wp, (long long)n3);
        VEL_LAUNCH_CHECK("ba_scale_w_kernel");
        const long long ns = (long long)n6 * n6;
        ba_init_s_kernel<<<(unsigned)((ns + 255) / 256), 256, 0, st>>>(U, g, nt, nc, S, rhs);
        VEL_LAUNCH_CHECK("ba_init_s_kernel");
        const double one = 1.0, neg = -1.0, zero = 0.0;
        // row-major W [n6][n3] is the column-major matrix Wc [n3][n6] (lda = n3)
        // S(lower) -= Wc'^T Wc'
        if (cublasDsyrk(h->blas, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_T, n6, n3, &neg, Wp, n3, &one, S, n6) != CUBLAS_STATUS_SUCCESS) {
            vel_set_error("vel_ba_solve: cublasDsyrk failed");
            return VEL_ERR_CUDA;
        }
        // rhs = g_c - W y   (W y = Wc^T y)
        if (cublasDgemv(h->blas, CUBLAS_OP_T, n3, n6, &neg, W, n3, y, 1, &one, rhs, 1) != CUBLAS_STATUS_SUCCESS) {
            vel_set_error("vel_ba_solve: cublasDgemv failed");
            return VEL_ERR_CUDA;
        }
        if (cusolverDnDpotrf(h->solver, CUBLAS_FILL_MODE_LOWER, n6, S, n6, potrf_work, L.lwork, info) != CUSOLVER_STATUS_SUCCESS ||
            cusolverDnDpotrs(h->solver, CUBLAS_FILL_MODE_LOWER, n6, 1, S, n6, rhs, n6, info) != CUSOLVER_STATUS_SUCCESS) {
            vel_set_error("vel_ba_solve: cuSOLVER Cholesky failed");
            return VEL_ERR_CUDA;
        }
        // t = W^T delta_c = Wc delta_c
        if (cublasDgemv(h->blas, CUBLAS_OP_N, n3, n6, &one, W, n3, rhs, 1, &zero, t, 1) != CUBLAS_STATUS_SUCCESS) {
            vel_set_error("vel_ba_solve: cublasDgemv failed");
            return VEL_ERR_CUDA;
        }
#endif
    } else {
        VEL_CUDA(cudaMemsetAsync(t, 0, sizeof(double) * n3, st));
    }
    const int ublocks = (nt + n6 + PT_THREADS - 1) / PT_THREADS;
    ba_update_kernel<<<ublocks, PT_THREADS, 0, st>>>(Vinv, y, t, rhs, nt, nc, x, part);
    VEL_LAUNCH_CHECK("ba_update_kernel");
    ba_rms_finalize_kernel<<<1, 32, 0, st>>>(part, ublocks, (long long)n3 + n6, rms_delta, info);
    VEL_LAUNCH_CHECK("ba_rms_finalize_kernel");
    return VEL_OK;
}

// ---- the solve in three stages (the sharded form runs collectives between them) ---------------------------
// vel_ba_reduce : (V+I)^-1 factors, W' = W blockdiag(L), S[rows of tile rows blk_lo..blk_hi) = U + I - W' W'^T, rhs = g_c - W y
// vel_ba_factor : Cholesky of S and delta_c = S^-1 rhs (in place in rhs)           -- the owner rank only
// vel_ba_update : t = W^T delta_c, delta_p = y - (V+I)^-1 t, x += 0.9 delta, rms(delta)
// S is row-major, lower triangle, [6nc][6nc]; byte offsets of S and rhs inside `work` come from vel_ba_solve_layout.
VEL_API int vel_ba_solve_layout(int32_t nt, int32_t nc, int64_t* off_S, int64_t* off_rhs)
{
    VEL_CHECK_ARG(nt > 0 && nc >= 0 && off_S && off_rhs, "vel_ba_solve_layout: bad argument");
    SolveLayout L;
    VEL_CHECK_ARG(solve_layout(nt, nc, &L, false), "vel_ba_solve_layout: layout failed");
    *off_S = (int64_t)L.off_s;
    *off_rhs = (int64_t)L.off_rhs;
    return VEL_OK;
}

VEL_API int vel_ba_reduce(const double* V, const double* U, const double* W, const double* g, int32_t nt, int32_t nc, int32_t blk_lo,
                          int32_t blk_hi, void* work, size_t work_bytes, vel_stream_t stream)
{
    VEL_CHECK_ARG(V && U && W && g && work, "vel_ba_reduce: NULL argument");
    VEL_CHECK_ARG(nt > 0 && nc > 0, "vel_ba_reduce: bad sizes nt=%d nc=%d", nt, nc);
    SolveLayout L;
    VEL_CHECK_ARG(solve_layout(nt, nc, &L, true), "vel_ba_reduce: layout failed");
    VEL_CHECK_ARG(work_bytes >= L.total, "vel_ba_reduce: workspace %zu B < required %zu B", work_bytes, L.total);
    cudaStream_t st = (cudaStream_t)stream;
    char* wb = (char*)work;
    double* Wp = (double*)(wb + L.off_wp);
    double* S = (double*)(wb + L.off_s);
    double* Vinv = (double*)(wb + L.off_vinv);
    double* Lf = (double*)(wb + L.off_l);
    double* y = (double*)(wb + L.off_y);
    double* rhs = (double*)(wb + L.off_rhs);
    int* info = (int*)(wb + L.off_info);
    const int n6 = 6 * nc, n3 = 3 * nt;
    const int pblocks = (nt + PT_THREADS - 1) / PT_THREADS;
    VEL_CUDA(cudaMemsetAsync(info, 0, sizeof(int), st));
    ba_point_prep_kernel<<<pblocks, PT_THREADS, 0, st>>>(V, g, nt, Vinv, Lf, y);
    VEL_LAUNCH_CHECK("ba_point_prep_kernel");
    const long long nel = (long long)n6 * nt;
    ba_scale_w_kernel<<<(unsigned)((nel + 255) / 256), 256, 0, st>>>(W, Lf, n6, nt, Wp, (long long)L.ldw);
    VEL_LAUNCH_CHECK("ba_scale_w_kernel");
    if ((int)L.ldw > n3) {
        const long long npad = (long long)n6 * (L.ldw - n3);
        ba_zero_pad_kernel<<<(unsigned)((npad + 255) / 256), 256, 0, st>>>(Wp, (long long)L.ldw, n6, n3);
        VEL_LAUNCH_CHECK("ba_zero_pad_kernel");
    }
    const long long ns = (long long)n6 * n6;
    ba_init_s_kernel<<<(unsigned)((ns + 255) / 256), 256, 0, st>>>(U, g, nt, nc, S, rhs);
    VEL_LAUNCH_CHECK("ba_init_s_kernel");
    int rc = vel_syrk_lower_sub_rows(Wp, (long long)L.ldw, n6, n3, S, n6, wb + L.off_flags, L.off_tpart - L.off_flags, blk_lo, blk_hi, stream);
    if (rc != VEL_OK) return rc;
    ba_gemv_rows_sub_kernel<<<n6, 256, 0, st>>>(W, n3, n3, y, rhs);
    VEL_LAUNCH_CHECK("ba_gemv_rows_sub_kernel");
    return VEL_OK;
}

VEL_API int vel_ba_factor(int32_t nt, int32_t nc, void* work, size_t work_bytes, vel_stream_t stream)
{
    VEL_CHECK_ARG(work && nt > 0 && nc > 0, "vel_ba_factor: bad argument");
    SolveLayout L;
    VEL_CHECK_ARG(solve_layout(nt, nc, &L, true), "vel_ba_factor: layout failed");
    VEL_CHECK_ARG(work_bytes >= L.total, "vel_ba_factor: workspace %zu B < required %zu B", work_bytes, L.total);
    cudaStream_t st = (cudaStream_t)stream;
    char* wb = (char*)work;
    double* S = (double*)(wb + L.off_s);
    double* rhs = (double*)(wb + L.off_rhs);
    int* info = (int*)(wb + L.off_info);
    const int n6 = 6 * nc;
    if (native_solver()) return vel_spd_solve(S, n6, n6, rhs, info, stream);
#ifdef VEL_WITH_VENDOR_SOLVER
    Handles* h = handles();
    VEL_CHECK_ARG(h != nullptr, "vel_ba_factor: cuSOLVER handle unavailable");
    cusolverDnSetStream(h->solver, st);
    // row-major lower == column-major upper
    if (cusolverDnDpotrf(h->solver, CUBLAS_FILL_MODE_UPPER, n6, S, n6, (double*)(wb + L.off_potrf), L.lwork, info) != CUSOLVER_STATUS_SUCCESS ||
        cusolverDnDpotrs(h->solver, CUBLAS_FILL_MODE_UPPER, n6, 1, S, n6, rhs, n6, info) != CUSOLVER_STATUS_SUCCESS) {
        vel_set_error("vel_ba_factor: cuSOLVER Cholesky failed");
        return VEL_ERR_CUDA;
    }
#else
    (void)st;
#endif
    return VEL_OK;
}

VEL_API int vel_ba_update(const double* W, int32_t nt, int32_t nc, double* x, double* rms_delta, void* work, size_t work_bytes,
                          vel_stream_t stream)
{
    VEL_CHECK_ARG(W && x && rms_delta && work && nt > 0 && nc > 0, "vel_ba_update: bad argument");
    SolveLayout L;
    VEL_CHECK_ARG(solve_layout(nt, nc, &L, false), "vel_ba_update: layout failed");
    cudaStream_t st = (cudaStream_t)stream;
    char* wb = (char*)work;
    double* Vinv = (double*)(wb + L.off_vinv);
    double* y = (double*)(wb + L.off_y);
    double* rhs = (double*)(wb + L.off_rhs);
    double* t = (double*)(wb + L.off_t);
    double* part = (double*)(wb + L.off_part);
    int* info = (int*)(wb + L.off_info);
    const int n6 = 6 * nc, n3 = 3 * nt;
    ba_gemv_cols_kernel<<<dim3((n3 + 255) / 256, GEMV_ROW_CHUNKS), 256, 0, st>>>(W, n3, n6, n3, rhs, (double*)(wb + L.off_tpart));
    VEL_LAUNCH_CHECK("ba_gemv_cols_kernel");
    ba_gemv_cols_reduce_kernel<<<(n3 + 255) / 256, 256, 0, st>>>((const double*)(wb + L.off_tpart), n3, t);
    VEL_LAUNCH_CHECK("ba_gemv_cols_reduce_kernel");
    const int ublocks = (nt + n6 + PT_THREADS - 1) / PT_THREADS;
    ba_update_kernel<<<ublocks, PT_THREADS, 0, st>>>(Vinv, y, t, rhs, nt, nc, x, part);
    VEL_LAUNCH_CHECK("ba_update_kernel");
    ba_rms_finalize_kernel<<<1, 32, 0, st>>>(part, ublocks, (long long)n3 + n6, rms_delta, info);
    VEL_LAUNCH_CHECK("ba_rms_finalize_kernel");
    return VEL_OK;
}

// =====================================================================================================
// fcnNLS_batch2 (utils/NLS.py:253-328): same damped Gauss-Newton, different camera model.
// Parameters x = [points nt*3 | q], q = [joint roll,pitch,yaw (3) | el | az | range_1..range_nc]:
//   pc = pw @ rpy2dcm(crpy);  camera 0 sees pc;  camera c >= 1 sees pc + sc2cc([range_c, el, az]) @ cam2ned()
// i.e. offset_c = [a sin(az), -range_c sin(el), a cos(az)], a = range_c cos(el).
// Five of the camera-side parameters are shared by every observation, so the camera block of JtJ is
// a dense (5+nc)^2 matrix G; the point blocks stay 3x3.  Forward differences (1e-6) as in the reference.
namespace {

constexpr int B2_SHARED = 5;

// per camera: offsets for the base parameters and for el+h, az+h, range+h  (4 x 3 doubles); camera 0 = zeros
__global__ void ba2_setup_kernel(const double* __restrict__ x, int nt, int nc, double* __restrict__ offs, double* __restrict__ dcm)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const double* q = x + 3ll * nt;
    if (c == 0) {
        // the four DCMs: base and roll/pitch/yaw + h
        for (int m = 0; m < 4; ++m) {
            double r[3] = {q[0], q[1], q[2]};
            if (m > 0) r[m - 1] = r[m - 1] + JDX;
            rpy2dcm(r, dcm + 9 * m);
        }
    }
    if (c > nc) return;
    double* o = offs + 12ll * c;
    if (c == 0) {
        for (int k = 0; k < 12; ++k) o[k] = 0.0;
        return;
    }
    const double el = q[3], az = q[4], rg = q[5 + (c - 1)];
    for (int m = 0; m < 4; ++m) {
        const double e = m == 1 ? el + JDX : el, a_ = m == 2 ? az + JDX : az, r = m == 3 ? rg + JDX : rg;
        const double a = r * cos(e);
        o[3 * m + 0] = a * sin(a_);
        o[3 * m + 1] = -r * sin(e);
        o[3 * m + 2] = a * cos(a_);
    }
}

// residual + forward-difference Jacobian rows of one observation: columns 0..2 point, 3..5 crpy, 6 el, 7 az, 8 range
__device__ __forceinline__ void ba2_jacobian(const double* K, const double* dcm, const double* off, bool cam0, double X, double Y,
                                             double Z, double& u0, double& v0, double (&ju)[9], double (&jv)[9])
{
    double ax, ay, az, u, v;
    rot(dcm, X, Y, Z, ax, ay, az);
    project(K, ax + off[0], ay + off[1], az + off[2], u0, v0);
    const double bx = ax, by = ay, bz = az;
    rot(dcm, X + JDX, Y, Z, ax, ay, az); project(K, ax + off[0], ay + off[1], az + off[2], u, v); ju[0] = (u - u0) / JDX; jv[0] = (v - v0) / JDX;
    rot(dcm, X, Y + JDX, Z, ax, ay, az); project(K, ax + off[0], ay + off[1], az + off[2], u, v); ju[1] = (u - u0) / JDX; jv[1] = (v - v0) / JDX;
    rot(dcm, X, Y, Z + JDX, ax, ay, az); project(K, ax + off[0], ay + off[1], az + off[2], u, v); ju[2] = (u - u0) / JDX; jv[2] = (v - v0) / JDX;
#pragma unroll
    for (int m = 1; m < 4; ++m) {
        rot(dcm + 9 * m, X, Y, Z, ax, ay, az);
        project(K, ax + off[0], ay + off[1], az + off[2], u, v);
        ju[2 + m] = (u - u0) / JDX; jv[2 + m] = (v - v0) / JDX;
    }
    if (cam0) {
#pragma unroll
        for (int m = 6; m < 9; ++m) { ju[m] = 0.0; jv[m] = 0.0; }
    } else {
#pragma unroll
        for (int m = 1; m < 4; ++m) {
            project(K, bx + off[3 * m], by + off[3 * m + 1], bz + off[3 * m + 2], u, v);
            ju[5 + m] = (u - u0) / JDX; jv[5 + m] = (v - v0) / JDX;
        }
    }
}

// one CTA per camera c in [0, nc]: H66 (21) and g6 (6) over the camera-side columns 3..8, and the range row of W
__global__ void __launch_bounds__(CAM_THREADS)
ba2_camera_kernel(const double* __restrict__ Kg, const double* __restrict__ x, const double* __restrict__ z,
                  const double* __restrict__ offs, const double* __restrict__ dcm, int nt, int nc, double* __restrict__ Hc,
                  double* __restrict__ gc, double* __restrict__ W)
{
    constexpr int NACC = 21 + 6;
    constexpr int NWARP = CAM_THREADS / 32;
    __shared__ double sK[9], sD[36], sO[12];
    __shared__ double sred[NWARP][NACC];
    const int c = blockIdx.x, tid = threadIdx.x;
    if (tid < 9) sK[tid] = Kg[tid];
    if (tid < 36) sD[tid] = dcm[tid];
    if (tid < 12) sO[tid] = offs[12ll * c + tid];
    __syncthreads();
    const double* zu = z + (long long)c * nt;
    const double* zv = z + (long long)(nc + 1) * nt + (long long)c * nt;
    double acc[NACC];
#pragma unroll
    for (int k = 0; k < NACC; ++k) acc[k] = 0.0;
    for (int i = tid; i < nt; i += CAM_THREADS) {
        double u0, v0, ju[9], jv[9];
        ba2_jacobian(sK, sD, sO, c == 0, x[3ll * i], x[3ll * i + 1], x[3ll * i + 2], u0, v0, ju, jv);
        const double ru = zu[i] - u0, rv = zv[i] - v0;
        int k = 0;
#pragma unroll
        for (int r = 0; r < 6; ++r)
#pragma unroll
            for (int q = r; q < 6; ++q) acc[k++] += ju[3 + r] * ju[3 + q] + jv[3 + r] * jv[3 + q];
#pragma unroll
        for (int r = 0; r < 6; ++r) acc[21 + r] += ju[3 + r] * ru + jv[3 + r] * rv;
        if (c > 0) {
            double* w = W + ((long long)(B2_SHARED + c - 1) * nt + i) * 3;   // range_c row
            w[0] = ju[8] * ju[0] + jv[8] * jv[0];
            w[1] = ju[8] * ju[1] + jv[8] * jv[1];
            w[2] = ju[8] * ju[2] + jv[8] * jv[2];
        }
    }
#pragma unroll
    for (int k = 0; k < NACC; ++k) acc[k] = warp_sum(acc[k]);
    if ((tid & 31) == 0) {
#pragma unroll
        for (int k = 0; k < NACC; ++k) sred[tid >> 5][k] = acc[k];
    }
    __syncthreads();
    if (tid < NACC) {
        double s = 0.0;
        for (int w = 0; w < NWARP; ++w) s += sred[w][tid];
        if (tid < 21) Hc[21ll * c + tid] = s;
        else gc[6ll * c + (tid - 21)] = s;
    }
}

// one thread per point: V_i, g_p,i, the five shared rows of W (summed over cameras) and the cost
__global__ void __launch_bounds__(PT_THREADS)
ba2_point_kernel(const double* __restrict__ Kg, const double* __restrict__ x, const double* __restrict__ z,
                 const double* __restrict__ offs, const double* __restrict__ dcm, int nt, int nc, double* __restrict__ V,
                 double* __restrict__ g, double* __restrict__ W, double* __restrict__ cost_part)
{
    __shared__ double sK[9], sD[36];
    __shared__ double sred[PT_THREADS / 32];
    const int tid = threadIdx.x, i = blockIdx.x * PT_THREADS + tid;
    if (tid < 9) sK[tid] = Kg[tid];
    if (tid < 36) sD[tid] = dcm[tid];
    __syncthreads();
    double cost = 0.0;
    if (i < nt) {
        const double X = x[3ll * i], Y = x[3ll * i + 1], Z = x[3ll * i + 2];
        double v[6] = {0, 0, 0, 0, 0, 0}, gp[3] = {0, 0, 0}, w5[B2_SHARED][3];
#pragma unroll
        for (int a = 0; a < B2_SHARED; ++a) { w5[a][0] = 0; w5[a][1] = 0; w5[a][2] = 0; }
        for (int c = 0; c <= nc; ++c) {
            double u0, v0, ju[9], jv[9];
            ba2_jacobian(sK, sD, offs + 12ll * c, c == 0, X, Y, Z, u0, v0, ju, jv);
            const double ru = z[(long long)c * nt + i] - u0, rv = z[(long long)(nc + 1) * nt + (long long)c * nt + i] - v0;
            v[0] += ju[0] * ju[0] + jv[0] * jv[0]; v[1] += ju[0] * ju[1] + jv[0] * jv[1]; v[2] += ju[0] * ju[2] + jv[0] * jv[2];
            v[3] += ju[1] * ju[1] + jv[1] * jv[1]; v[4] += ju[1] * ju[2] + jv[1] * jv[2]; v[5] += ju[2] * ju[2] + jv[2] * jv[2];
#pragma unroll
            for (int b = 0; b < 3; ++b) gp[b] += ju[b] * ru + jv[b] * rv;
#pragma unroll
            for (int a = 0; a < B2_SHARED; ++a)
#pragma unroll
                for (int b = 0; b < 3; ++b) w5[a][b] += ju[3 + a] * ju[b] + jv[3 + a] * jv[b];
            cost += ru * ru + rv * rv;
        }
#pragma unroll
        for (int k = 0; k < 6; ++k) V[6ll * i + k] = v[k];
#pragma unroll
        for (int b = 0; b < 3; ++b) g[3ll * i + b] = gp[b];
#pragma unroll
        for (int a = 0; a < B2_SHARED; ++a)
#pragma unroll
            for (int b = 0; b < 3; ++b) W[((long long)a * nt + i) * 3 + b] = w5[a][b];
    }
    cost = warp_sum(cost);
    if ((tid & 31) == 0) sred[tid >> 5] = cost;
    __syncthreads();
    if (tid == 0) {
        double s = 0.0;
        for (int w = 0; w < PT_THREADS / 32; ++w) s += sred[w];
        cost_part[blockIdx.x] = s;
    }
}

// G (nq x nq, dense, symmetric) and the camera-side gradient from the per-camera 6x6 blocks
__global__ void ba2_assemble_kernel(const double* __restrict__ Hc, const double* __restrict__ gc, int nt, int nc,
                                    double* __restrict__ G, double* __restrict__ g)
{
    const int nq = B2_SHARED + nc;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)nq * nq) return;
    const int r = (int)(idx % nq), c = (int)(idx / nq);
    auto h = [&](int cam, int a, int b) {   // upper-triangle lookup of camera `cam`'s 6x6 block
        const int lo = a < b ? a : b, hi = a < b ? b : a;
        return Hc[21ll * cam + lo * 6 - lo * (lo - 1) / 2 + (hi - lo)];
    };
    double v = 0.0;
    if (r < B2_SHARED && c < B2_SHARED) {
        for (int cam = 0; cam <= nc; ++cam) v += h(cam, r, c);
    } else if (r < B2_SHARED) {
        v = h(c - B2_SHARED + 1, r, 5);
    } else if (c < B2_SHARED) {
        v = h(r - B2_SHARED + 1, c, 5);
    } else if (r == c) {
        v = h(r - B2_SHARED + 1, 5, 5);
    }
    G[idx] = v;
    if (c == 0) {
        double s = 0.0;
        if (r < B2_SHARED) { for (int cam = 0; cam <= nc; ++cam) s += gc[6ll * cam + r]; }
        else s = gc[6ll * (r - B2_SHARED + 1) + 5];
        g[3ll * nt + r] = s;
    }
}

// S = G + I, rhs = g_q   (column-major nq x nq)
__global__ void ba2_init_s_kernel(const double* __restrict__ G, const double* __restrict__ g, int nt, int nq, double* __restrict__ S,
                                  double* __restrict__ rhs)
{
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)nq * nq) return;
    const int r = (int)(idx % nq), c = (int)(idx / nq);
    S[idx] = G[idx] + (r == c ? 1.0 : 0.0);
    if (c == 0) rhs[r] = g[3ll * nt + r];
}

// delta_p from the point blocks, x += 0.9 * delta for x = [points | q], partial sums of delta^2
__global__ void __launch_bounds__(PT_THREADS)
ba2_update_kernel(const double* __restrict__ Vinv, const double* __restrict__ y, const double* __restrict__ t,
                  const double* __restrict__ dq, int nt, int nq, double* __restrict__ x, double* __restrict__ ss_part)
{
    __shared__ double sred[PT_THREADS / 32];
    const int tid = threadIdx.x;
    const long long i = (long long)blockIdx.x * PT_THREADS + tid;
    double ss = 0.0;
    if (i < nt) {
        const double* v = Vinv + 6 * i;
        const double t0 = t[3 * i], t1 = t[3 * i + 1], t2 = t[3 * i + 2];
        const double d0 = y[3 * i] - (v[0] * t0 + v[1] * t1 + v[2] * t2);
        const double d1 = y[3 * i + 1] - (v[1] * t0 + v[3] * t1 + v[4] * t2);
        const double d2 = y[3 * i + 2] - (v[2] * t0 + v[4] * t1 + v[5] * t2);
        x[3 * i] += 0.9 * d0; x[3 * i + 1] += 0.9 * d1; x[3 * i + 2] += 0.9 * d2;
        ss = 0.81 * (d0 * d0 + d1 * d1 + d2 * d2);
    } else if (i < nt + nq) {
        const int r = (int)(i - nt);
        const double d = dq[r];
        x[3ll * nt + r] += 0.9 * d;
        ss = 0.81 * d * d;
    }
    ss = warp_sum(ss);
    if ((tid & 31) == 0) sred[tid >> 5] = ss;
    __syncthreads();
    if (tid == 0) {
        double s = 0.0;
        for (int w = 0; w < PT_THREADS / 32; ++w) s += sred[w];
        ss_part[blockIdx.x] = s;
    }
}

}  // namespace

VEL_API int vel_ba2_accumulate(const double* K, const double* x, const double* z, int32_t nt, int32_t nc, double* V, double* G,
                               double* W, double* g, double* cost, vel_stream_t stream)
{
    VEL_CHECK_ARG(K && x && z && V && G && W && g && cost, "vel_ba2_accumulate: NULL argument");
    VEL_CHECK_ARG(nt > 0 && nc >= 1, "vel_ba2_accumulate: bad sizes nt=%d nc=%d", nt, nc);
    cudaStream_t st = (cudaStream_t)stream;
    const int pblocks = (nt + PT_THREADS - 1) / PT_THREADS;
    const int nq = B2_SHARED + nc;
    double* tmp = nullptr;
    const size_t n_tmp = 12ull * (nc + 1) + 36 + 27ull * (nc + 1) + pblocks;
    vel_keep_async_pool_cached();
    VEL_CUDA(cudaMallocAsync((void**)&tmp, sizeof(double) * n_tmp, st));
    double* offs = tmp;
    double* dcm = offs + 12ull * (nc + 1);
    double* Hc = dcm + 36;
    double* gc = Hc + 21ull * (nc + 1);
    double* cost_part = gc + 6ull * (nc + 1);
    ba2_setup_kernel<<<(nc + 1 + 127) / 128, 128, 0, st>>>(x, nt, nc, offs, dcm);
    VEL_LAUNCH_CHECK("ba2_setup_kernel");
    ba2_camera_kernel<<<nc + 1, CAM_THREADS, 0, st>>>(K, x, z, offs, dcm, nt, nc, Hc, gc, W);
    VEL_LAUNCH_CHECK("ba2_camera_kernel");
    ba2_point_kernel<<<pblocks, PT_THREADS, 0, st>>>(K, x, z, offs, dcm, nt, nc, V, g, W, cost_part);
    VEL_LAUNCH_CHECK("ba2_point_kernel");
    const long long nel = (long long)nq * nq;
    ba2_assemble_kernel<<<(unsigned)((nel + 255) / 256), 256, 0, st>>>(Hc, gc, nt, nc, G, g);
    VEL_LAUNCH_CHECK("ba2_assemble_kernel");
    ba_cost_finalize_kernel<<<1, 32, 0, st>>>(cost_part, pblocks, cost);
    VEL_LAUNCH_CHECK("ba_cost_finalize_kernel");
    VEL_CUDA(cudaFreeAsync(tmp, st));
    return VEL_OK;
}

// Solve (JtJ + I) delta = g for x = [points | q] with a dense nq x nq camera-side block G; x += 0.9 delta.
// Workspace: vel_ba_solve_workspace(nt, ceil(nq / 6)) bytes are sufficient.
VEL_API int vel_ba2_solve(const double* V, const double* G, const double* W, const double* g, int32_t nt, int32_t nq, double* x,
                          double* rms_delta, void* work, size_t work_bytes, vel_stream_t stream)
{
    VEL_CHECK_ARG(V && G && W && g && x && rms_delta && work, "vel_ba2_solve: NULL argument");
    VEL_CHECK_ARG(nt > 0 && nq > 0, "vel_ba2_solve: bad sizes nt=%d nq=%d", nt, nq);
    const int nc_equiv = (nq + 5) / 6;
    SolveLayout L;
    VEL_CHECK_ARG(solve_layout(nt, nc_equiv, &L, false), "vel_ba2_solve: layout failed");
    const bool native = native_solver();
    int lwork = 0;
#ifdef VEL_WITH_VENDOR_SOLVER
    Handles* h = native ? nullptr : handles();
    VEL_CHECK_ARG(native || h != nullptr, "vel_ba2_solve: cuBLAS/cuSOLVER handles unavailable");
    if (!native && cusolverDnDpotrf_bufferSize(h->solver, CUBLAS_FILL_MODE_LOWER, nq, nullptr, nq, &lwork) != CUSOLVER_STATUS_SUCCESS) {
        vel_set_error("vel_ba2_solve: cusolverDnDpotrf_bufferSize failed");
        return VEL_ERR_CUDA;
    }
#endif
    const size_t need = L.off_potrf + align256(sizeof(double) * (size_t)(lwork > 0 ? lwork : 1));
    VEL_CHECK_ARG(work_bytes >= need, "vel_ba2_solve: workspace %zu B < required %zu B", work_bytes, need);
    cudaStream_t st = (cudaStream_t)stream;
    char* wb = (char*)work;
    double* Wp = (double*)(wb + L.off_wp);
    double* S = (double*)(wb + L.off_s);
    double* Vinv = (double*)(wb + L.off_vinv);
    double* Lf = (double*)(wb + L.off_l);
    double* y = (double*)(wb + L.off_y);
    double* rhs = (double*)(wb + L.off_rhs);
    double* t = (double*)(wb + L.off_t);
    double* part = (double*)(wb + L.off_part);
    int* info = (int*)(wb + L.off_info);
    double* potrf_work = (double*)(wb + L.off_potrf);
    const int n3 = 3 * nt;
    const int pblocks = (nt + PT_THREADS - 1) / PT_THREADS;
    VEL_CUDA(cudaMemsetAsync(info, 0, sizeof(int), st));
    ba_point_prep_kernel<<<pblocks, PT_THREADS, 0, st>>>(V, g, nt, Vinv, Lf, y);
    VEL_LAUNCH_CHECK("ba_point_prep_kernel");
    if (native) {
        // the layout was sized for ceil(nq/6) cameras: nq <= 6*nc_equiv rows of the same padded pitch
        const long long nel_n = (long long)nq * nt;
        ba_scale_w_kernel<<<(unsigned)((nel_n + 255) / 256), 256, 0, st>>>(W, Lf, nq, nt, Wp, (long long)L.ldw);
        VEL_LAUNCH_CHECK("ba_scale_w_kernel");
        if ((int)L.ldw > n3) {
            const long long npad = (long long)nq * (L.ldw - n3);
            ba_zero_pad_kernel<<<(unsigned)((npad + 255) / 256), 256, 0, st>>>(Wp, (long long)L.ldw, nq, n3);
            VEL_LAUNCH_CHECK("ba_zero_pad_kernel");
        }
        const long long ns_n = (long long)nq * nq;
        ba2_init_s_kernel<<<(unsigned)((ns_n + 255) / 256), 256, 0, st>>>(G, g, nt, nq, S, rhs);
        VEL_LAUNCH_CHECK("ba2_init_s_kernel");
        const int rc = dense_solve_native(W, Wp, (long long)L.ldw, nq, n3, S, y, rhs, t, (double*)(wb + L.off_tpart), wb + L.off_flags,
                                          L.off_tpart - L.off_flags, info, st);
        if (rc != VEL_OK) return rc;
        const int ublocks_n = (nt + nq + PT_THREADS - 1) / PT_THREADS;
        ba2_update_kernel<<<ublocks_n, PT_THREADS, 0, st>>>(Vinv, y, t, rhs, nt, nq, x, part);
        VEL_LAUNCH_CHECK("ba2_update_kernel");
        ba_rms_finalize_kernel<<<1, 32, 0, st>>>(part, ublocks_n, (long long)n3 + nq, rms_delta, info);
        VEL_LAUNCH_CHECK("ba_rms_finalize_kernel");
        return VEL_OK;
    }
#ifdef VEL_WITH_VENDOR_SOLVER
    cublasSetStream(h->blas, st);
    cusolverDnSetStream(h->solver, st);
    cublasSetPointerMode(h->blas, CUBLAS_POINTER_MODE_HOST);
    const long long nel = (long long)nq * nt;
    ba_scale_w_kernel<<<(unsigned)((nel + 255) / 256), 256, 0, st>>>(W, Lf, nq, nt, Wp, (long long)n3);
    VEL_LAUNCH_CHECK("ba_scale_w_kernel");
    const long long ns = (long long)nq * nq;
    ba2_init_s_kernel<<<(unsigned)((ns + 255) / 256), 256, 0, st>>>(G, g, nt, nq, S, rhs);
    VEL_LAUNCH_CHECK("ba2_init_s_kernel");
    const double one = 1.0, neg = -1.0, zero = 0.0;
    if (cublasDsyrk(h->blas, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_T, nq, n3, &neg, Wp, n3, &one, S, nq) != CUBLAS_STATUS_SUCCESS ||
        cublasDgemv(h->blas, CUBLAS_OP_T, n3, nq, &neg, W, n3, y, 1, &one, rhs, 1) != CUBLAS_STATUS_SUCCESS) {
        vel_set_error("vel_ba2_solve: cuBLAS failed");
        return VEL_ERR_CUDA;
    }
    if (cusolverDnDpotrf(h->solver, CUBLAS_FILL_MODE_LOWER, nq, S, nq, potrf_work, lwork, info) != CUSOLVER_STATUS_SUCCESS ||
        cusolverDnDpotrs(h->solver, CUBLAS_FILL_MODE_LOWER, nq, 1, S, nq, rhs, nq, info) != CUSOLVER_STATUS_SUCCESS) {
        vel_set_error("vel_ba2_solve: cuSOLVER Cholesky failed");
        return VEL_ERR_CUDA;
    }
    if (cublasDgemv(h->blas, CUBLAS_OP_N, n3, nq, &one, W, n3, rhs, 1, &zero, t, 1) != CUBLAS_STATUS_SUCCESS) {
        vel_set_error("vel_ba2_solve: cublasDgemv failed");
        return VEL_ERR_CUDA;
    }
    const int ublocks = (nt + nq + PT_THREADS - 1) / PT_THREADS;
    ba2_update_kernel<<<ublocks, PT_THREADS, 0, st>>>(Vinv, y, t, rhs, nt, nq, x, part);
    VEL_LAUNCH_CHECK("ba2_update_kernel");
    ba_rms_finalize_kernel<<<1, 32, 0, st>>>(part, ublocks, (long long)n3 + nq, rms_delta, info);
    VEL_LAUNCH_CHECK("ba_rms_finalize_kernel");
#endif
    return VEL_OK;
}
