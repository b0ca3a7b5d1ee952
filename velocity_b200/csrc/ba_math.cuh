// Device helpers shared by the bundle-adjustment translation units (ba.cu: the stage-by-stage entry points, ba_loop.cu: the
// device-resident iteration loop): the reference's projection chain, utils/NLS.py:71-78 (fzK), utils/transforms.py:7-23 (rpy2dcm).
#pragma once
#include "common.cuh"

namespace {

constexpr double JDX = 1e-6;
constexpr int CAM_THREADS = 256;
constexpr int PT_THREADS = 128;

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ void rpy2dcm(const double* rpy, double* C)  // utils/transforms.py:7-23
{
    double sr, cr, sp, cp, sy, cy;
    sincos(rpy[0], &sr, &cr);
    sincos(rpy[1], &sp, &cp);
    sincos(rpy[2], &sy, &cy);
    C[0] = cp * cy; C[1] = sr * sp * cy - cr * sy; C[2] = cr * sp * cy + sr * sy;
    C[3] = cp * sy; C[4] = sr * sp * sy + cr * cy; C[5] = cr * sp * sy - sr * cy;
    C[6] = -sp;     C[7] = sr * cp;                C[8] = cr * cp;
}

__device__ __forceinline__ void project(const double* K, double ax, double ay, double az, double& u, double& v)
{
    const double q0 = ax * K[0] + ay * K[3] + az * K[6];
    const double q1 = ax * K[1] + ay * K[4] + az * K[7];
    const double q2 = ax * K[2] + ay * K[5] + az * K[8];
    u = q0 / q2;
    v = q1 / q2;
}

__device__ __forceinline__ void rot(const double* R, double X, double Y, double Z, double& ax, double& ay, double& az)
{
    ax = X * R[0] + Y * R[3] + Z * R[6];
    ay = X * R[1] + Y * R[4] + Z * R[7];
    az = X * R[2] + Y * R[5] + Z * R[8];
}

}  // namespace
