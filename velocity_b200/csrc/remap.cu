// K3: affine ROI warp, utils/KLT.py:70-73.
//
//   x, y   = meshgrid(arange(x0, x1, float32), arange(y0, y1, float32))
//   mapx   = x*T[0,0] + y*T[1,0] + T[2,0]          (float32, evaluated left to right by numpy)
//   mapy   = x*T[0,1] + y*T[1,1] + T[2,1]
//   out    = cv2.remap(im, mapx, mapy, INTER_LINEAR)           (BORDER_CONSTANT, value 0)
//
// cv2.remap's bilinear path is fixed point: coordinates rounded (half-even) to 1/32 px, the four
// weights taken from a 32x32 table of 15-bit integers ((32-fx)(32-fy)*32 etc., exact for the
// bilinear kernel), result (sum + 2^14) >> 15.  Taps outside the source contribute 0.
// Streaming kernel: algorithmic bytes = ROI bytes read (once, via L2) + ROI bytes written.
#include "common.cuh"

namespace {

struct Affine { float t00, t01, t10, t11, t20, t21; };

__global__ void __launch_bounds__(256)
remap_affine_kernel(const uint8_t* __restrict__ src, int sw, int sh, int spitch, Affine T, int x0, int y0, int dw, int dh,
                    uint8_t* __restrict__ dst, int dpitch)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y * blockDim.y + threadIdx.y;
    if (c >= dw || r >= dh) return;
    const float x = (float)(x0 + c), y = (float)(y0 + r);
    const float mx = __fadd_rn(__fadd_rn(__fmul_rn(x, T.t00), __fmul_rn(y, T.t10)), T.t20);
    const float my = __fadd_rn(__fadd_rn(__fmul_rn(x, T.t01), __fmul_rn(y, T.t11)), T.t21);
    const int sx = __float2int_rn(__fmul_rn(mx, 32.f)), sy = __float2int_rn(__fmul_rn(my, 32.f));
    int ix = sx >> 5, iy = sy >> 5;
    ix = max(-32768, min(32767, ix));  // OpenCV keeps the integer part as saturated int16
    iy = max(-32768, min(32767, iy));
    const int fx = sx & 31, fy = sy & 31;
    const int w00 = (32 - fx) * (32 - fy) * 32, w01 = fx * (32 - fy) * 32, w10 = (32 - fx) * fy * 32, w11 = fx * fy * 32;
    const bool in_x0 = ix >= 0 && ix < sw, in_x1 = ix + 1 >= 0 && ix + 1 < sw;
    const bool in_y0 = iy >= 0 && iy < sh, in_y1 = iy + 1 >= 0 && iy + 1 < sh;
    int acc = 0;
    if (in_y0 && in_x0) acc += (int)ldg_u8(src + (long long)iy * spitch + ix) * w00;
    if (in_y0 && in_x1) acc += (int)ldg_u8(src + (long long)iy * spitch + ix + 1) * w01;
    if (in_y1 && in_x0) acc += (int)ldg_u8(src + (long long)(iy + 1) * spitch + ix) * w10;
    if (in_y1 && in_x1) acc += (int)ldg_u8(src + (long long)(iy + 1) * spitch + ix + 1) * w11;
    dst[(long long)r * dpitch + c] = (uint8_t)((acc + (1 << 14)) >> 15);
}

}  // namespace

VEL_API int vel_remap_affine_u8(const uint8_t* src, int32_t width, int32_t height, int32_t pitch, const float* T_host, int32_t x0,
                                int32_t y0, int32_t dst_width, int32_t dst_height, uint8_t* dst, int32_t dst_pitch,
                                vel_stream_t stream)
{
    VEL_CHECK_ARG(src && dst && T_host, "vel_remap_affine_u8: NULL argument");
    VEL_CHECK_ARG(width > 0 && height > 0 && dst_width > 0 && dst_height > 0 && pitch >= width && dst_pitch >= dst_width,
                  "vel_remap_affine_u8: bad geometry");
    Affine T = {T_host[0], T_host[1], T_host[2], T_host[3], T_host[4], T_host[5]};
    dim3 block(64, 4), grid((dst_width + 63) / 64, (dst_height + 3) / 4);
    remap_affine_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(src, width, height, pitch, T, x0, y0, dst_width, dst_height,
                                                                   dst, dst_pitch);
    VEL_LAUNCH_CHECK("remap_affine_kernel");
    return VEL_OK;
}
