// K5: small-dof Levenberg-Marquardt pose solvers, batched over independent frames.
//
//   vel_nls_t   <- fcnNLS_t   (utils/NLS.py:102-129)  3-dof translation
//   vel_nls_rt  <- fcnNLS_Rt  (utils/NLS.py:133-183)  6-dof roll/pitch/yaw + translation
//
// Same algorithm as the reference, step for step, in float64:
//   zhat = fzK(a, K);  J by FORWARD DIFFERENCES with dx = 1e-6 (the reference's numeric Jacobian is
//   reproduced rather than replaced by an analytic one, so iterates agree to ~1e-12, not ~1e-6);
//   delta = inv(JtJ + I) Jt (z - zhat) * min(((i+1)*0.2)^2, 1);  x += delta;
//   stop when rms(delta) < 1e-8, at most 30 iterations.
// One CTA per problem; per-observation 2xDOF Jacobian rows are formed in registers, the normal
// equations are reduced with warp shuffles + a fixed-order cross-warp sum (deterministic), thread 0
// solves the DOFxDOF system.  Reduction-bound (HBM/L2 streaming of p, pw once per iteration).
#include "common.cuh"

namespace {

constexpr int NLS_THREADS = 256;
constexpr int NLS_MAX_ITER = 30;

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Z-Y-X Euler angles -> DCM, exactly the element formulas of utils/transforms.py:7-23
__device__ void rpy2dcm(const double* rpy, double* C)
{
    double sr, cr, sp, cp, sy, cy;
    sincos(rpy[0], &sr, &cr);
    sincos(rpy[1], &sp, &cp);
    sincos(rpy[2], &sy, &cy);
    C[0] = cp * cy; C[1] = sr * sp * cy - cr * sy; C[2] = cr * sp * cy + sr * sy;
    C[3] = cp * sy; C[4] = sr * sp * sy + cr * cy; C[5] = cr * sp * sy - sr * cy;
    C[6] = -sp;     C[7] = sr * cp;                C[8] = cr * cp;
}

// (u, v) = pscale(a @ K), K row-major 3x3 in the reference's row-vector convention
__device__ __forceinline__ void project(const double* K, double ax, double ay, double az, double& u, double& v)
{
    const double q0 = ax * K[0] + ay * K[3] + az * K[6];
    const double q1 = ax * K[1] + ay * K[4] + az * K[7];
    const double q2 = ax * K[2] + ay * K[5] + az * K[8];
    u = q0 / q2;
    v = q1 / q2;
}

// solve (H + I) d = g for a DOF x DOF symmetric H given as a full matrix; partial pivoting
template <int DOF>
__device__ void solve_damped(double* H, const double* g, double* d)
{
    double M[DOF][DOF + 1];
    for (int r = 0; r < DOF; ++r) {
        for (int c = 0; c < DOF; ++c) M[r][c] = H[r * DOF + c] + (r == c ? 1.0 : 0.0);
        M[r][DOF] = g[r];
    }
    for (int k = 0; k < DOF; ++k) {
        int piv = k;
        double best = fabs(M[k][k]);
        for (int r = k + 1; r < DOF; ++r)
            if (fabs(M[r][k]) > best) { best = fabs(M[r][k]); piv = r; }
        if (piv != k)
            for (int c = k; c <= DOF; ++c) { const double t = M[k][c]; M[k][c] = M[piv][c]; M[piv][c] = t; }
        for (int r = k + 1; r < DOF; ++r) {
            const double f = M[r][k] / M[k][k];
            for (int c = k; c <= DOF; ++c) M[r][c] -= f * M[k][c];
        }
    }
    for (int r = DOF - 1; r >= 0; --r) {
        double s = M[r][DOF];
        for (int c = r + 1; c < DOF; ++c) s -= M[r][c] * d[c];
        d[r] = s / M[r][r];
    }
}

template <int DOF>
__global__ void __launch_bounds__(NLS_THREADS)
nls_kernel(const double* __restrict__ Kg, const double* __restrict__ p, const double* __restrict__ pw,
           const int* __restrict__ first, const int* __restrict__ count, const double* __restrict__ x0, double* __restrict__ xout,
           int* __restrict__ iters)
{
    constexpr int NH = DOF * (DOF + 1) / 2;  // upper triangle of JtJ
    constexpr int NACC = NH + DOF;           // + Jt r
    constexpr int NW = NLS_THREADS / 32;
    __shared__ double sK[9];
    __shared__ double sx[DOF];
    __shared__ double sR[4][9];      // DCMs: base and the three perturbed (6-dof only)
    __shared__ double sred[NW][NACC];
    __shared__ int s_done;

    const int prob = blockIdx.x, tid = threadIdx.x;
    const int base = first[prob], n = count[prob];
    const double* P = p + 2ll * base;
    const double* PW = pw + 3ll * base;
    if (tid < 9) sK[tid] = Kg[tid];
    if (tid < DOF) sx[tid] = x0[prob * DOF + tid];
    if (tid == 0) s_done = 0;
    __syncthreads();

    const double dx = 1e-6;
    int it = 0;
    for (; it < NLS_MAX_ITER; ++it) {
        if (DOF == 6 && tid < 4) {
            double r[3] = {sx[0], sx[1], sx[2]};
            if (tid > 0) r[tid - 1] = r[tid - 1] + dx;   // rpy2dcm(x03 + dx_k)
            rpy2dcm(r, sR[tid]);
        }
        __syncthreads();
        double acc[NACC];
#pragma unroll
        for (int k = 0; k < NACC; ++k) acc[k] = 0.0;

        for (int i = tid; i < n; i += NLS_THREADS) {
            const double X = PW[3 * i], Y = PW[3 * i + 1], Z = PW[3 * i + 2];
            const double zu = P[2 * i], zv = P[2 * i + 1];
            double ju[DOF], jv[DOF], u0, v0;
            if (DOF == 3) {
                const double bx = X + sx[0], by = Y + sx[1], bz = Z + sx[2];   // b0 = pw + x
                project(sK, bx, by, bz, u0, v0);
                double u, v;
                project(sK, bx + dx, by, bz, u, v); ju[0] = (u - u0) / dx; jv[0] = (v - v0) / dx;
                project(sK, bx, by + dx, bz, u, v); ju[1] = (u - u0) / dx; jv[1] = (v - v0) / dx;
                project(sK, bx, by, bz + dx, u, v); ju[2] = (u - u0) / dx; jv[2] = (v - v0) / dx;
            } else {
                double a[4][3];
#pragma unroll
                for (int m = 0; m < 4; ++m) {   // a_m = pw @ R_m
                    const double* R = sR[m];
                    a[m][0] = X * R[0] + Y * R[3] + Z * R[6];
                    a[m][1] = X * R[1] + Y * R[4] + Z * R[7];
                    a[m][2] = X * R[2] + Y * R[5] + Z * R[8];
                }
                const double t0 = sx[3], t1 = sx[4], t2 = sx[5];
                project(sK, a[0][0] + t0, a[0][1] + t1, a[0][2] + t2, u0, v0);
                double u, v;
#pragma unroll
                for (int m = 1; m < 4; ++m) {   // rotation columns: a_m + b0
                    project(sK, a[m][0] + t0, a[m][1] + t1, a[m][2] + t2, u, v);
                    ju[m - 1] = (u - u0) / dx; jv[m - 1] = (v - v0) / dx;
                }
                // translation columns: a0 + (x36 + dx_k)
                project(sK, a[0][0] + (t0 + dx), a[0][1] + t1, a[0][2] + t2, u, v); ju[3] = (u - u0) / dx; jv[3] = (v - v0) / dx;
                project(sK, a[0][0] + t0, a[0][1] + (t1 + dx), a[0][2] + t2, u, v); ju[4] = (u - u0) / dx; jv[4] = (v - v0) / dx;
                project(sK, a[0][0] + t0, a[0][1] + t1, a[0][2] + (t2 + dx), u, v); ju[5] = (u - u0) / dx; jv[5] = (v - v0) / dx;
            }
            const double ru = zu - u0, rv = zv - v0;
            int k = 0;
#pragma unroll
            for (int r = 0; r < DOF; ++r)
#pragma unroll
                for (int c = r; c < DOF; ++c) acc[k++] += ju[r] * ju[c] + jv[r] * jv[c];
#pragma unroll
            for (int r = 0; r < DOF; ++r) acc[NH + r] += ju[r] * ru + jv[r] * rv;
        }
#pragma unroll
        for (int k = 0; k < NACC; ++k) acc[k] = warp_sum(acc[k]);
        if ((tid & 31) == 0) {
#pragma unroll
            for (int k = 0; k < NACC; ++k) sred[tid >> 5][k] = acc[k];
        }
        __syncthreads();
        if (tid == 0) {
            double tot[NACC];
            for (int k = 0; k < NACC; ++k) {
                double s = 0.0;
                for (int w = 0; w < NW; ++w) s += sred[w][k];
                tot[k] = s;
            }
            double H[DOF * DOF], d[DOF];
            int k = 0;
            for (int r = 0; r < DOF; ++r)
                for (int c = r; c < DOF; ++c) { H[r * DOF + c] = tot[k]; H[c * DOF + r] = tot[k]; ++k; }
            solve_damped<DOF>(H, tot + NH, d);
            const double ramp = (double)(it + 1) * 0.2;
            const double sc = fmin(ramp * ramp, 1.0);
            double ss = 0.0;
            for (int r = 0; r < DOF; ++r) {
                d[r] *= sc;
                sx[r] = sx[r] + d[r];
                ss += d[r] * d[r];
            }
            if (sqrt(ss / DOF) < 1e-8) s_done = 1;
        }
        __syncthreads();
        if (s_done) break;
    }
    if (tid < DOF) xout[prob * DOF + tid] = sx[tid];
    if (tid == 0) iters[prob] = s_done ? it + 1 : -NLS_MAX_ITER;
}

template <int DOF>
int launch_nls(const double* K, const double* p, const double* pw, const int32_t* first, const int32_t* n, int32_t nprob,
               const double* x0, double* x, int32_t* iters, vel_stream_t stream, const char* name)
{
    VEL_CHECK_ARG(K && p && pw && first && n && x0 && x && iters, "%s: NULL argument", name);
    VEL_CHECK_ARG(nprob >= 0, "%s: nprob < 0", name);
    if (nprob == 0) return VEL_OK;
    nls_kernel<DOF><<<nprob, NLS_THREADS, 0, (cudaStream_t)stream>>>(K, p, pw, first, n, x0, x, iters);
    VEL_LAUNCH_CHECK(name);
    return VEL_OK;
}

}  // namespace

VEL_API int vel_nls_t(const double* K, const double* p, const double* pw, const int32_t* first, const int32_t* n, int32_t nprob,
                      const double* x0, double* x, int32_t* iters, vel_stream_t stream)
{
    return launch_nls<3>(K, p, pw, first, n, nprob, x0, x, iters, stream, "vel_nls_t");
}

VEL_API int vel_nls_rt(const double* K, const double* p, const double* pw, const int32_t* first, const int32_t* n, int32_t nprob,
                       const double* x0, double* x, int32_t* iters, vel_stream_t stream)
{
    return launch_nls<6>(K, p, pw, first, n, nprob, x0, x, iters, stream, "vel_nls_rt");
}
