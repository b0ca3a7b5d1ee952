// K7 + K8 as ONE device-resident loop: vel_ba_iterate = the whole `for i in range(10)` of fcnNLS_batch (utils/NLS.py:222-242).
//
// vel_ba_accumulate / vel_ba_solve (ba.cu) mirror the reference's loop BODY, so a caller pays one host synchronisation per
// iteration for the `if xr < 1e-7: break` test (:238) -- 0.4 ms of a 3.6 ms iteration at the BASELINE size.  Here the test
// runs on the device: every kernel of an iteration takes a gate word and returns at once when it is set, the last kernel of
// an iteration evaluates rms(delta) < tol and closes the gate, and the host enqueues max_iter iterations back to back and
// reads the (cost, rms(delta)) history once.  The iterations that run are the ones the reference would run.
//
// The arithmetic is the block-sparse system of ba.cu (same forward-difference Jacobian entries, same damping, same Schur
// complement) with the cross blocks held ONLY in their scaled form  W' = W blockdiag(L_i),  L_i L_i^T = (V_i + I)^-1:
//     S     = U + I - W' W'^T                       (DMMA SYRK, dense_f64.cu)
//     rhs   = g_c - W y,   y = (V+I)^-1 g_p         (accumulated inside the camera kernel: W_ji y_i is a 6x3 by 3 product)
//     dc    = S^-1 rhs                              (task-graph Cholesky, dense_f64.cu)
//     dp_i  = y_i - (V_i+I)^-1 W_i^T dc = y_i - L_i t'_i,   t' = W'^T dc
// so the point blocks are reduced FIRST (V needs every camera), the camera kernel then writes W' directly, and neither the
// unscaled W (177 MB at nt=4096, nc=299) nor the separate scaling / W y passes over it exist.  Per iteration: 10 launches.
//
// Two deliberate departures from the reference's floating-point sequence, both below its own forward-difference noise
// (eps * |u| / 1e-6 ~ 1e-7 in a Jacobian entry): a projection divides once (reciprocal, then two multiplies) instead of
// twice, and the difference quotient multiplies by 1e6 instead of dividing by 1e-6.  Parity: tests/test_sfm_gpu.py
// (the reference's own fcnNLS_batch outputs at nt = 24 ... 512, the sparse oracle at the BASELINE size).
#include <math.h>

#include "ba_math.cuh"

namespace {

constexpr double JDX_INV = 1e6;
constexpr int CAMREC = 39;                 // per camera: R0 (9) | pos (3) | R(roll+h) (9) | R(pitch+h) (9) | R(yaw+h) (9)
constexpr int PTL_MAX_CHUNKS = 32;
constexpr int TCHUNKS = 24;                // row chunks of t' = W'^T dc

// reciprocal to full double precision: 20-bit hardware seed + two Newton steps (no slow path: q is a depth, never 0 / inf / denormal)
__device__ __forceinline__ double rcp_nr(double x)
{
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double e = fma(-x, y, 1.0);
    y = fma(y, e, y);
    e = fma(-x, y, 1.0);
    return fma(y, e, y);
}

__device__ __forceinline__ void project_r(const double* K, double ax, double ay, double az, double& u, double& v)
{
    const double q0 = ax * K[0] + ay * K[3] + az * K[6];
    const double q1 = ax * K[1] + ay * K[4] + az * K[7];
    const double q2 = ax * K[2] + ay * K[5] + az * K[8];
    const double r = rcp_nr(q2);
    u = q0 * r;
    v = q1 * r;
}

// forward-difference Jacobian of one observation wrt its point (camera 0 projects the point itself, utils/NLS.py:208)
__device__ __forceinline__ void point_jac_r(const double* K, const double* R, const double* pos, bool fixed_cam, double X, double Y, double Z,
                                            double& u0, double& v0, double (&ju)[3], double (&jv)[3])
{
    double ax, ay, az, u, v;
    if (fixed_cam) {
        project_r(K, X, Y, Z, u0, v0);
        project_r(K, X + JDX, Y, Z, u, v); ju[0] = (u - u0) * JDX_INV; jv[0] = (v - v0) * JDX_INV;
        project_r(K, X, Y + JDX, Z, u, v); ju[1] = (u - u0) * JDX_INV; jv[1] = (v - v0) * JDX_INV;
        project_r(K, X, Y, Z + JDX, u, v); ju[2] = (u - u0) * JDX_INV; jv[2] = (v - v0) * JDX_INV;
        return;
    }
    rot(R, X, Y, Z, ax, ay, az);
    project_r(K, ax + pos[0], ay + pos[1], az + pos[2], u0, v0);
    rot(R, X + JDX, Y, Z, ax, ay, az);
    project_r(K, ax + pos[0], ay + pos[1], az + pos[2], u, v); ju[0] = (u - u0) * JDX_INV; jv[0] = (v - v0) * JDX_INV;
    rot(R, X, Y + JDX, Z, ax, ay, az);
    project_r(K, ax + pos[0], ay + pos[1], az + pos[2], u, v); ju[1] = (u - u0) * JDX_INV; jv[1] = (v - v0) * JDX_INV;
    rot(R, X, Y, Z + JDX, ax, ay, az);
    project_r(K, ax + pos[0], ay + pos[1], az + pos[2], u, v); ju[2] = (u - u0) * JDX_INV; jv[2] = (v - v0) * JDX_INV;
}

// ---- loop state -----------------------------------------------------------------------------------------------------------------
struct LoopState {
    int gate;      // 1 = converged (or stopped): every later kernel returns at once
    int iter;      // iterations completed
    int info;      // raised by a failed Cholesky
};

__global__ void bal_init_kernel(LoopState* st, double* __restrict__ hist, int max_iter, double* __restrict__ Wp, long long ldw, int nrows, int n3)
{
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx == 0) { st->gate = 0; st->iter = 0; st->info = 0; }
    if (idx < 2ll * max_iter) hist[idx] = __longlong_as_double(0x7ff8000000000000ll);
    const int pad = (int)(ldw - n3);                         // zero columns the SYRK's last k-tile reads
    if (pad > 0 && idx < (long long)nrows * pad) Wp[(idx / pad) * ldw + n3 + idx % pad] = 0.0;
}

// per-camera constants: camera 0 = identity / zero
__global__ void bal_cam_setup_kernel(const double* __restrict__ x, int nt, int nc, double* __restrict__ cams, const LoopState* __restrict__ st)
{
    if (st->gate) return;
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c > nc) return;
    double* o = cams + (long long)CAMREC * c;
    if (c == 0) {
        for (int k = 0; k < CAMREC; ++k) o[k] = 0.0;
        o[0] = o[4] = o[8] = 1.0;
        return;
    }
    const double* pos = x + 3ll * nt + 3ll * (c - 1);
    const double* rpy = x + 3ll * nt + 3ll * nc + 3ll * (c - 1);
    for (int m = 0; m < 4; ++m) {
        double r[3] = {rpy[0], rpy[1], rpy[2]};
        if (m > 0) r[m - 1] = r[m - 1] + JDX;                // utils/NLS.py:226-228: x[k] + dx, then the whole chain
        rpy2dcm(r, o + (m == 0 ? 0 : 3 + 9 * m));
    }
    o[9] = pos[0]; o[10] = pos[1]; o[11] = pos[2];
}

// ---- point side: thread = (point, camera chunk) -> partial V_i, g_p,i, cost ---------------------------------------------------------
__global__ void __launch_bounds__(PT_THREADS)
bal_point_kernel(const double* __restrict__ Kg, const double* __restrict__ x, const double* __restrict__ z, const double* __restrict__ cams,
                 int nt, int nc, int nchunks, double* __restrict__ part, const LoopState* __restrict__ st)
{
    if (st->gate) return;
    __shared__ double sK[9];
    const int tid = threadIdx.x, i = blockIdx.x * PT_THREADS + tid, chunk = blockIdx.y;
    if (tid < 9) sK[tid] = Kg[tid];
    __syncthreads();
    if (i >= nt) return;
    const int per = (nc + 1 + nchunks - 1) / nchunks;
    const int c0 = chunk * per, c1 = min(nc + 1, c0 + per);
    const double X = x[3ll * i], Y = x[3ll * i + 1], Z = x[3ll * i + 2];
    double v00 = 0, v01 = 0, v02 = 0, v11 = 0, v12 = 0, v22 = 0, g0 = 0, g1 = 0, g2 = 0, cost = 0;
    for (int c = c0; c < c1; ++c) {
        const double* cm = cams + (long long)CAMREC * c;
        double u0, w0, ju[3], jv[3];
        point_jac_r(sK, cm, cm + 9, c == 0, X, Y, Z, u0, w0, ju, jv);
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

// chunks added in a fixed order; then per point: (V_i + I)^-1 = L_i L_i^T, y_i = (V_i + I)^-1 g_p,i
// CTA = 32 points x 10 quantities (thread (k, p) adds the chunks of quantity k of point p: 128 CTAs of independent coalesced
// streams instead of 32 CTAs of threads walking 320 values each), then one warp finishes the 32 points.
constexpr int RP_PTS = 32, RP_THREADS = RP_PTS * 10;

__global__ void __launch_bounds__(RP_THREADS)
bal_point_reduce_prep_kernel(const double* __restrict__ part, int nt, int nchunks, double* __restrict__ Lf, double* __restrict__ y,
                             double* __restrict__ cost_part, const LoopState* __restrict__ st)
{
    if (st->gate) return;
    __shared__ double sa[10][RP_PTS];
    const int tid = threadIdx.x, k = tid / RP_PTS, p = tid % RP_PTS, i = blockIdx.x * RP_PTS + p;
    double acc = 0.0;
    if (i < nt) {
        const double* o = part + (long long)k * nt + i;
        const long long stride = 10ll * nt;
        int ch = 0;
        for (; ch + 4 <= nchunks; ch += 4) {          // four loads in flight, added in chunk order
            const double v0 = o[ch * stride], v1 = o[(ch + 1) * stride], v2 = o[(ch + 2) * stride], v3 = o[(ch + 3) * stride];
            acc += v0; acc += v1; acc += v2; acc += v3;
        }
        for (; ch < nchunks; ++ch) acc += o[ch * stride];
    }
    sa[k][p] = acc;
    __syncthreads();
    if (tid >= RP_PTS) return;
    double cost = 0.0;
    if (i < nt) {
        double a[10];
#pragma unroll
        for (int q = 0; q < 10; ++q) a[q] = sa[q][p];
        cost = a[9];
        const double m00 = a[0] + 1.0, m01 = a[1], m02 = a[2], m11 = a[3] + 1.0, m12 = a[4], m22 = a[5] + 1.0;
        const double c00 = m11 * m22 - m12 * m12, c01 = m02 * m12 - m01 * m22, c02 = m01 * m12 - m02 * m11;
        const double c11 = m00 * m22 - m02 * m02, c12 = m01 * m02 - m00 * m12, c22 = m00 * m11 - m01 * m01;
        const double inv = 1.0 / (m00 * c00 + m01 * c01 + m02 * c02);
        const double i00 = c00 * inv, i01 = c01 * inv, i02 = c02 * inv, i11 = c11 * inv, i12 = c12 * inv, i22 = c22 * inv;
        const double l00 = sqrt(i00), l10 = i01 / l00, l20 = i02 / l00;
        const double l11 = sqrt(i11 - l10 * l10), l21 = (i12 - l20 * l10) / l11;
        const double l22 = sqrt(i22 - l20 * l20 - l21 * l21);
        double* l = Lf + 6ll * i;
        l[0] = l00; l[1] = l10; l[2] = l11; l[3] = l20; l[4] = l21; l[5] = l22;
        y[3ll * i] = i00 * a[6] + i01 * a[7] + i02 * a[8];
        y[3ll * i + 1] = i01 * a[6] + i11 * a[7] + i12 * a[8];
        y[3ll * i + 2] = i02 * a[6] + i12 * a[7] + i22 * a[8];
    }
    cost = warp_sum(cost);
    if (tid == 0) cost_part[blockIdx.x] = cost;
}

// ---- camera side: work items (camera j, chunk of CAM_CHUNK points) over a grid that fills the machine exactly ------------------------
//   partial sums of U_j, g_c,j and sum_i W_ji y_i -> camp [j][chunk][33] (added in fixed order by bal_camera_reduce_kernel),
//   W'_ji = W_ji L_i -> rows 6j..6j+5 of W'
// One CTA per camera would be 299 CTAs of which 148 x (resident CTAs per SM) run at a time: a last round for three cameras.  Items
// are 8x finer and walked with a grid stride, so the tail is one item.  A warp handles 32 consecutive points per trip; its 6 x 96
// block of W' goes through shared memory so that the global stores are whole 256-byte segments (a thread's own 3 values per row sit
// 24 bytes apart).
constexpr int CAM_T = 128;                  // threads per CTA
constexpr int CAM_CHUNK = 512;              // points per item
constexpr int CAM_NACC = 21 + 6 + 6;

__global__ void __launch_bounds__(CAM_T)
bal_camera_kernel(const double* __restrict__ Kg, const double* __restrict__ x, const double* __restrict__ z, const double* __restrict__ cams,
                  const double* __restrict__ Lf, const double* __restrict__ y, int nt, int nc, int nchunk, double* __restrict__ Wp, long long ldw,
                  double* __restrict__ camp, const LoopState* __restrict__ st)
{
    if (st->gate) return;
    constexpr int NACC = CAM_NACC;
    constexpr int NWARP = CAM_T / 32;
    __shared__ double sK[9], sC[CAMREC];
    __shared__ double sred[NWARP][NACC];
    __shared__ double sW[NWARP][6][97];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid < 9) sK[tid] = Kg[tid];
    const int n3 = 3 * nt;
  for (int item = blockIdx.x; item < nc * nchunk; item += gridDim.x) {
    const int c = 1 + item / nchunk, chunk = item % nchunk;  // camera index >= 1, point chunk
    const int i_lo = chunk * CAM_CHUNK, i_hi = min(nt, i_lo + CAM_CHUNK);
    __syncthreads();                                         // the previous item's readers of sC / sred are done
    if (tid < CAMREC) sC[tid] = cams[(long long)CAMREC * c + tid];
    __syncthreads();
    const double* R0 = sC;
    const double* sP = sC + 9;
    const double* zu = z + (long long)c * nt;
    const double* zv = z + (long long)(nc + 1) * nt + (long long)c * nt;

    double acc[NACC];
#pragma unroll
    for (int k = 0; k < NACC; ++k) acc[k] = 0.0;
    for (int i0 = i_lo + warp * 32; i0 < i_hi; i0 += CAM_T) {
        const int i = i0 + lane;
        double wv[6][3];
        if (i < i_hi) {
            const double X = x[3ll * i], Y = x[3ll * i + 1], Z = x[3ll * i + 2];
            double u0, v0, pu[3], pv[3];
            point_jac_r(sK, R0, sP, false, X, Y, Z, u0, v0, pu, pv);
            double cu[6], cv[6], ax, ay, az, u, v;
            rot(R0, X, Y, Z, ax, ay, az);
            project_r(sK, ax + (sP[0] + JDX), ay + sP[1], az + sP[2], u, v); cu[0] = (u - u0) * JDX_INV; cv[0] = (v - v0) * JDX_INV;
            project_r(sK, ax + sP[0], ay + (sP[1] + JDX), az + sP[2], u, v); cu[1] = (u - u0) * JDX_INV; cv[1] = (v - v0) * JDX_INV;
            project_r(sK, ax + sP[0], ay + sP[1], az + (sP[2] + JDX), u, v); cu[2] = (u - u0) * JDX_INV; cv[2] = (v - v0) * JDX_INV;
#pragma unroll
            for (int m = 1; m < 4; ++m) {
                rot(sC + 3 + 9 * m, X, Y, Z, ax, ay, az);
                project_r(sK, ax + sP[0], ay + sP[1], az + sP[2], u, v);
                cu[2 + m] = (u - u0) * JDX_INV; cv[2 + m] = (v - v0) * JDX_INV;
            }
            const double ru = zu[i] - u0, rv = zv[i] - v0;
            int k = 0;
#pragma unroll
            for (int r = 0; r < 6; ++r)
#pragma unroll
                for (int q = r; q < 6; ++q) acc[k++] += cu[r] * cu[q] + cv[r] * cv[q];
#pragma unroll
            for (int r = 0; r < 6; ++r) acc[21 + r] += cu[r] * ru + cv[r] * rv;
            const double* l = Lf + 6ll * i;
            const double l00 = l[0], l10 = l[1], l11 = l[2], l20 = l[3], l21 = l[4], l22 = l[5];
            const double y0 = y[3ll * i], y1 = y[3ll * i + 1], y2 = y[3ll * i + 2];
#pragma unroll
            for (int a = 0; a < 6; ++a) {
                const double w0 = cu[a] * pu[0] + cv[a] * pv[0];
                const double w1 = cu[a] * pu[1] + cv[a] * pv[1];
                const double w2 = cu[a] * pu[2] + cv[a] * pv[2];
                acc[27 + a] += w0 * y0 + w1 * y1 + w2 * y2;
                wv[a][0] = w0 * l00 + w1 * l10 + w2 * l20;      // row vector times lower-triangular L_i
                wv[a][1] = w1 * l11 + w2 * l21;
                wv[a][2] = w2 * l22;
            }
        } else {
#pragma unroll
            for (int a = 0; a < 6; ++a) wv[a][0] = wv[a][1] = wv[a][2] = 0.0;
        }
#pragma unroll
        for (int a = 0; a < 6; ++a) {
            sW[warp][a][3 * lane] = wv[a][0];
            sW[warp][a][3 * lane + 1] = wv[a][1];
            sW[warp][a][3 * lane + 2] = wv[a][2];
        }
        __syncwarp();
        const int col0 = 3 * i0;
#pragma unroll
        for (int a = 0; a < 6; ++a) {
            double* row = Wp + (long long)(6 * (c - 1) + a) * ldw + col0;
#pragma unroll
            for (int q = 0; q < 3; ++q)
                if (col0 + 32 * q + lane < n3) row[32 * q + lane] = sW[warp][a][32 * q + lane];
        }
        __syncwarp();
    }
#pragma unroll
    for (int k = 0; k < NACC; ++k) acc[k] = warp_sum(acc[k]);
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < NACC; ++k) sred[warp][k] = acc[k];
    }
    __syncthreads();
    if (tid < NACC) {
        double s = 0.0;
        for (int w = 0; w < NWARP; ++w) s += sred[w][tid];
        camp[((long long)(c - 1) * nchunk + chunk) * NACC + tid] = s;
    }
  }
}

// chunk partials added in a fixed order:  U_j + I -> the diagonal block of S,  rhs_j = g_c,j - sum_i W_ji y_i.  The CTA of camera j
// also clears rows 6j .. 6j+5 of S first (the SYRK subtracts from S), so no separate pass over the 26 MB of S is needed.
__global__ void __launch_bounds__(256)
bal_camera_reduce_kernel(const double* __restrict__ camp, int nc, int nchunk, double* __restrict__ S, double* __restrict__ rhs,
                         const LoopState* __restrict__ st)
{
    if (st->gate) return;
    __shared__ double sacc[CAM_NACC];
    const int c = 1 + blockIdx.x, tid = threadIdx.x;
    const int n6 = 6 * nc, r0 = 6 * (c - 1);
    {
        double* rows = S + (long long)r0 * n6;                 // 6 * n6 doubles, 16-byte aligned (n6 is even, S is 256-byte aligned)
        const int n2 = 3 * n6;                                  // double2 count
        for (int e = tid; e < n2; e += 256) reinterpret_cast<double2*>(rows)[e] = make_double2(0.0, 0.0);
    }
    if (tid < CAM_NACC) {
        double s = 0.0;
        for (int ch = 0; ch < nchunk; ++ch) s += camp[((long long)(c - 1) * nchunk + ch) * CAM_NACC + tid];
        sacc[tid] = s;
    }
    __syncthreads();
    if (tid < 36) {
        const int a = tid / 6, b = tid % 6, lo = a < b ? a : b, hi = a < b ? b : a;
        const int k = lo * 6 - lo * (lo - 1) / 2 + (hi - lo);      // upper-triangle index of (lo, hi)
        S[(long long)(r0 + a) * n6 + r0 + b] = sacc[k] + (a == b ? 1.0 : 0.0);
    } else if (tid < 42) {
        const int a = tid - 36;
        rhs[r0 + a] = sacc[21 + a] - sacc[27 + a];
    }
}

// tpart[ch][k] = sum over the rows of chunk ch of W'[r][k] dc[r]   (thread per column, coalesced across k)
__global__ void __launch_bounds__(256)
bal_gemv_cols_kernel(const double* __restrict__ Wp, long long ldw, int nrows, int n3, const double* __restrict__ dc, double* __restrict__ tpart,
                     const LoopState* __restrict__ st)
{
    if (st->gate) return;
    const int k = blockIdx.x * 256 + threadIdx.x;
    if (k >= n3) return;
    const int per = (nrows + TCHUNKS - 1) / TCHUNKS;
    const int r0 = blockIdx.y * per, r1 = min(nrows, r0 + per);
    double s0 = 0.0, s1 = 0.0;
    int r = r0;
    for (; r + 8 <= r1; r += 8) {
        double w[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) w[q] = Wp[(long long)(r + q) * ldw + k];
#pragma unroll
        for (int q = 0; q < 8; q += 2) { s0 += w[q] * dc[r + q]; s1 += w[q + 1] * dc[r + q + 1]; }
    }
    for (; r < r1; ++r) s0 += Wp[(long long)r * ldw + k] * dc[r];
    tpart[(long long)blockIdx.y * n3 + k] = s0 + s1;
}

// delta_p,i = y_i - L_i t'_i;  x += 0.9 delta;  partial sums of delta^2
__global__ void __launch_bounds__(PT_THREADS)
bal_update_kernel(const double* __restrict__ Lf, const double* __restrict__ y, const double* __restrict__ tpart, const double* __restrict__ dc,
                  int nt, int nc, double* __restrict__ x, double* __restrict__ ss_part, const LoopState* __restrict__ st)
{
    if (st->gate) return;
    __shared__ double sred[PT_THREADS / 32];
    const int tid = threadIdx.x;
    const long long i = (long long)blockIdx.x * PT_THREADS + tid;
    const long long n3 = 3ll * nt;
    double ss = 0.0;
    if (i < nt) {
        double t0 = 0.0, t1 = 0.0, t2 = 0.0;
        if (nc > 0) {
            for (int ch = 0; ch < TCHUNKS; ++ch) {
                const double* tp = tpart + ch * n3 + 3 * i;
                t0 += tp[0]; t1 += tp[1]; t2 += tp[2];
            }
        }
        const double* l = Lf + 6 * i;
        const double d0 = y[3 * i] - l[0] * t0;
        const double d1 = y[3 * i + 1] - (l[1] * t0 + l[2] * t1);
        const double d2 = y[3 * i + 2] - (l[3] * t0 + l[4] * t1 + l[5] * t2);
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

// end of an iteration: hist[it] = (sum of squared residuals BEFORE the update, rms(delta)); the gate closes on rms(delta) < tol
// (utils/NLS.py:237-238).  A failed Cholesky turns rms(delta) into NaN: it never passes the test and the caller sees it.
__global__ void bal_finalize_kernel(const double* __restrict__ cost_part, int ncost, const double* __restrict__ ss_part, int nss, long long nx,
                                    double tol, double* __restrict__ hist, LoopState* st, int* __restrict__ iters_run)
{
    if (st->gate) return;
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double cs = 0.0, ss = 0.0;
        for (int k = 0; k < ncost; ++k) cs += cost_part[k];
        for (int k = 0; k < nss; ++k) ss += ss_part[k];
        const double rms = st->info != 0 ? __longlong_as_double(0x7ff8000000000000ll) : sqrt(ss / (double)nx);
        const int it = st->iter;
        hist[2 * it] = cs;
        hist[2 * it + 1] = rms;
        st->iter = it + 1;
        *iters_run = it + 1;
        if (rms < tol) st->gate = 1;
    }
}

inline size_t align256(size_t b) { return (b + 255) & ~(size_t)255; }

struct LoopLayout {
    size_t off_wp, off_s, off_l, off_y, off_rhs, off_tpart, off_part, off_cams, off_camp, off_cost, off_ss, off_state, off_flags, flags_bytes, total, ldw;
    int nchunks, pblocks, rblocks, ublocks, cam_chunks;
};

LoopLayout loop_layout(int nt, int nc)
{
    LoopLayout L;
    const size_t n6 = 6ull * nc, n3 = 3ull * nt;
    L.ldw = (n3 + 31) / 32 * 32;
    L.nchunks = nc + 1 >= 16 ? ((nc + 1) / 8 < PTL_MAX_CHUNKS ? (nc + 1) / 8 : PTL_MAX_CHUNKS) : 1;
    L.pblocks = (nt + PT_THREADS - 1) / PT_THREADS;
    L.rblocks = (nt + RP_PTS - 1) / RP_PTS;
    L.ublocks = (int)((nt + n6 + PT_THREADS - 1) / PT_THREADS);
    size_t o = 0;
    L.off_wp = o; o += align256(sizeof(double) * (n6 ? n6 : 1) * L.ldw);
    L.off_s = o; o += align256(sizeof(double) * (n6 * n6 + 2));
    L.off_l = o; o += align256(sizeof(double) * 6 * nt);
    L.off_y = o; o += align256(sizeof(double) * n3);
    L.off_rhs = o; o += align256(sizeof(double) * (n6 + 8));
    L.off_tpart = o; o += align256(sizeof(double) * TCHUNKS * n3);
    L.off_part = o; o += align256(sizeof(double) * 10ull * L.nchunks * nt);
    L.off_cams = o; o += align256(sizeof(double) * CAMREC * (nc + 1));
    L.cam_chunks = (nt + CAM_CHUNK - 1) / CAM_CHUNK;
    L.off_camp = o; o += align256(sizeof(double) * CAM_NACC * (size_t)(nc > 0 ? nc : 1) * L.cam_chunks);
    L.off_cost = o; o += align256(sizeof(double) * L.rblocks);
    L.off_ss = o; o += align256(sizeof(double) * L.ublocks);
    L.off_state = o; o += 256;
    L.flags_bytes = vel_dense_syrk_workspace((int)(n6 > 0 ? n6 : 1), (int)n3);
    L.off_flags = o; o += align256(L.flags_bytes);
    L.total = o;
    return L;
}

}  // namespace

VEL_API size_t vel_ba_iterate_workspace(int32_t nt, int32_t nc)
{
    if (nt <= 0 || nc < 0) return 0;
    return loop_layout(nt, nc).total;
}

VEL_API int vel_ba_iterate(const double* K, const double* z, int32_t nt, int32_t nc, double* x, int32_t max_iter, double tol, double* hist,
                           int32_t* iters_run, void* work, size_t work_bytes, vel_stream_t stream)
{
    VEL_CHECK_ARG(K && z && x && hist && iters_run && work, "vel_ba_iterate: NULL argument");
    VEL_CHECK_ARG(nt > 0 && nc >= 0 && max_iter > 0 && max_iter <= 1000, "vel_ba_iterate: bad sizes nt=%d nc=%d max_iter=%d", nt, nc, max_iter);
    VEL_CHECK_ARG(((size_t)work & 255) == 0, "vel_ba_iterate: workspace must be 256-byte aligned");
    const LoopLayout L = loop_layout(nt, nc);
    VEL_CHECK_ARG(work_bytes >= L.total, "vel_ba_iterate: workspace %zu B < required %zu B", work_bytes, L.total);
    cudaStream_t st = (cudaStream_t)stream;
    char* wb = (char*)work;
    double* Wp = (double*)(wb + L.off_wp);
    double* S = (double*)(wb + L.off_s);
    double* Lf = (double*)(wb + L.off_l);
    double* y = (double*)(wb + L.off_y);
    double* rhs = (double*)(wb + L.off_rhs);
    double* tpart = (double*)(wb + L.off_tpart);
    double* part = (double*)(wb + L.off_part);
    double* cams = (double*)(wb + L.off_cams);
    double* camp = (double*)(wb + L.off_camp);
    double* cost_part = (double*)(wb + L.off_cost);
    double* ss_part = (double*)(wb + L.off_ss);
    LoopState* state = (LoopState*)(wb + L.off_state);
    const int n6 = 6 * nc, n3 = 3 * nt;
    {
        const long long npad = (long long)n6 * (long long)(L.ldw - n3);
        const long long n = npad > 2ll * max_iter ? npad : 2ll * max_iter;
        bal_init_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(state, hist, max_iter, Wp, (long long)L.ldw, n6, n3);
        VEL_LAUNCH_CHECK("bal_init_kernel");
        VEL_CUDA(cudaMemsetAsync(iters_run, 0, sizeof(int32_t), st));
    }
    for (int it = 0; it < max_iter; ++it) {
        bal_cam_setup_kernel<<<(nc + 1 + 127) / 128, 128, 0, st>>>(x, nt, nc, cams, state);
        VEL_LAUNCH_CHECK("bal_cam_setup_kernel");
        bal_point_kernel<<<dim3(L.pblocks, L.nchunks), PT_THREADS, 0, st>>>(K, x, z, cams, nt, nc, L.nchunks, part, state);
        VEL_LAUNCH_CHECK("bal_point_kernel");
        bal_point_reduce_prep_kernel<<<L.rblocks, RP_THREADS, 0, st>>>(part, nt, L.nchunks, Lf, y, cost_part, state);
        VEL_LAUNCH_CHECK("bal_point_reduce_prep_kernel");
        if (nc > 0) {
            const int cam_items = nc * L.cam_chunks, cam_grid = cam_items < 3 * kNumSMs ? cam_items : 3 * kNumSMs;   // 3 CTAs of 168 registers x 128 threads per SM
            bal_camera_kernel<<<cam_grid, CAM_T, 0, st>>>(K, x, z, cams, Lf, y, nt, nc, L.cam_chunks, Wp, (long long)L.ldw, camp, state);
            VEL_LAUNCH_CHECK("bal_camera_kernel");
            bal_camera_reduce_kernel<<<nc, 256, 0, st>>>(camp, nc, L.cam_chunks, S, rhs, state);
            VEL_LAUNCH_CHECK("bal_camera_reduce_kernel");
            int rc = vel_dense_syrk_rows_gated(Wp, (int64_t)L.ldw, n6, n3, S, n6, wb + L.off_flags, L.flags_bytes, 0, -1, &state->gate, stream);
            if (rc != VEL_OK) return rc;
            rc = vel_dense_spd_solve_gated(S, n6, n6, rhs, &state->info, &state->gate, stream);
            if (rc != VEL_OK) return rc;
            bal_gemv_cols_kernel<<<dim3((n3 + 255) / 256, TCHUNKS), 256, 0, st>>>(Wp, (long long)L.ldw, n6, n3, rhs, tpart, state);
            VEL_LAUNCH_CHECK("bal_gemv_cols_kernel");
        }
        bal_update_kernel<<<L.ublocks, PT_THREADS, 0, st>>>(Lf, y, tpart, rhs, nt, nc, x, ss_part, state);
        VEL_LAUNCH_CHECK("bal_update_kernel");
        bal_finalize_kernel<<<1, 32, 0, st>>>(cost_part, L.rblocks, ss_part, L.ublocks, (long long)n3 + n6, tol, hist, state, iters_run);
        VEL_LAUNCH_CHECK("bal_finalize_kernel");
    }
    return VEL_OK;
}
