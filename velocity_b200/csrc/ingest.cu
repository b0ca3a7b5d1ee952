// Frame ingest: cv2.cvtColor(imbgr, cv2.COLOR_BGR2GRAY) of vidExample.py:91.
//
// OpenCV 4.13's 8-bit path is 15-bit fixed point (verified exhaustively against cv2 by the oracle tests):
//     gray = (3735*B + 19235*G + 9798*R + 16384) >> 15
// Streaming, HBM-bound: 3 bytes read + 1 byte written per pixel.  Each thread converts 4 pixels:
// three aligned 32-bit loads (12 bytes of BGR), one 32-bit store.
#include "common.cuh"

namespace {

__device__ __forceinline__ unsigned gray_of(unsigned b, unsigned g, unsigned r)
{
    return (3735u * b + 19235u * g + 9798u * r + 16384u) >> 15;
}

__global__ void __launch_bounds__(256)
bgr2gray_kernel(const uint8_t* __restrict__ bgr_base, long long bgr_stride, int bpitch, int width, int height,
                uint8_t* __restrict__ gray_base, long long gray_stride, int gpitch, bool vec_ok)
{
    const int x4 = blockIdx.x * blockDim.x + threadIdx.x;   // group of 4 pixels
    const int y = blockIdx.y;
    const uint8_t* __restrict__ src = bgr_base + (long long)blockIdx.z * bgr_stride + (long long)y * bpitch;
    uint8_t* __restrict__ dst = gray_base + (long long)blockIdx.z * gray_stride + (long long)y * gpitch;
    const int x = 4 * x4;
    if (x >= width) return;
    if (vec_ok && x + 3 < width) {
        const unsigned* p = reinterpret_cast<const unsigned*>(src + 3 * x);
        const unsigned w0 = __ldg(p), w1 = __ldg(p + 1), w2 = __ldg(p + 2);
        // bytes: w0 = B0 G0 R0 B1 | w1 = G1 R1 B2 G2 | w2 = R2 B3 G3 R3
        const unsigned g0 = gray_of(w0 & 0xff, (w0 >> 8) & 0xff, (w0 >> 16) & 0xff);
        const unsigned g1 = gray_of(w0 >> 24, w1 & 0xff, (w1 >> 8) & 0xff);
        const unsigned g2 = gray_of((w1 >> 16) & 0xff, w1 >> 24, w2 & 0xff);
        const unsigned g3 = gray_of((w2 >> 8) & 0xff, (w2 >> 16) & 0xff, w2 >> 24);
        *reinterpret_cast<unsigned*>(dst + x) = g0 | (g1 << 8) | (g2 << 16) | (g3 << 24);
    } else {
        for (int k = 0; k < 4 && x + k < width; ++k) {
            const uint8_t* q = src + 3 * (x + k);
            dst[x + k] = (uint8_t)gray_of(q[0], q[1], q[2]);
        }
    }
}

}  // namespace

VEL_API int vel_bgr2gray_u8(const uint8_t* bgr, int64_t bgr_stride, int32_t bgr_pitch, int32_t nframes, int32_t width,
                            int32_t height, uint8_t* gray, int64_t gray_stride, int32_t gray_pitch, vel_stream_t stream)
{
    VEL_CHECK_ARG(bgr && gray, "vel_bgr2gray_u8: NULL argument");
    VEL_CHECK_ARG(width > 0 && height > 0 && height <= 65535 && nframes > 0 && nframes <= 65535, "vel_bgr2gray_u8: bad geometry");
    VEL_CHECK_ARG(bgr_pitch >= 3 * width && gray_pitch >= width, "vel_bgr2gray_u8: pitch too small");
    const bool vec_ok = ((((uintptr_t)bgr) | (uintptr_t)bgr_stride | (uintptr_t)bgr_pitch | ((uintptr_t)gray) | (uintptr_t)gray_stride |
                          (uintptr_t)gray_pitch) & 3) == 0;
    dim3 grid(((width + 3) / 4 + 255) / 256, height, nframes);
    bgr2gray_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(bgr, bgr_stride, bgr_pitch, width, height, gray, gray_stride, gray_pitch,
                                                          vec_ok);
    VEL_LAUNCH_CHECK("bgr2gray_kernel");
    return VEL_OK;
}
