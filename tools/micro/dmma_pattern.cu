// What DMMA rate is reachable with a REAL register pattern?  16 accumulators = 4 A fragments x 4 B fragments per k4-step, fragments
// (a) loop-invariant registers, (b) re-loaded from shared memory every step (the GEMM inner loop).  nvcc -O3 -arch sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
template <int MODE>
__global__ void k(double* out, int iters)
{
    __shared__ double sm[2][128 * 20];
    for (int i = threadIdx.x; i < 2 * 128 * 20; i += blockDim.x) (&sm[0][0])[i] = 1e-3 * (i % 7);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t4 = lane & 3;
    const double* pa = &sm[0][((warp >> 2) * 32 + g) * 20 + t4];
    const double* pb = &sm[1][((warp & 3) * 32 + g) * 20 + t4];
    double acc[4][4][2] = {};
    double a[4], b[4];
    for (int i = 0; i < 4; ++i) { a[i] = pa[i * 8 * 20]; b[i] = pb[i * 8 * 20]; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            if (MODE == 1) {
#pragma unroll
                for (int i = 0; i < 4; ++i) { a[i] = pa[i * 8 * 20 + kk * 4]; b[i] = pb[i * 8 * 20 + kk * 4]; }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
    }
    double s = 0;
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) s += acc[i][j][0] + acc[i][j][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main()
{
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double* out; cudaMalloc(&out, 8 * sms * 1024);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 2000;
    for (int mode = 0; mode < 2; ++mode)
        for (int threads : {128, 256, 512}) {
            float ms = 0;
            for (int rep = 0; rep < 2; ++rep) {
                cudaEventRecord(e0);
                if (mode) k<1><<<sms, threads>>>(out, iters); else k<0><<<sms, threads>>>(out, iters);
                cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
            }
            double flops = 2.0 * 256 * 16 * 4 * (double)iters * (threads / 32) * sms;
            printf("%s fragments, %3d thr/SM: %.3f ms  %.2f TFLOP/s\n", mode ? "LDS-loaded " : "register   ", threads, ms, flops / ms / 1e9);
        }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
