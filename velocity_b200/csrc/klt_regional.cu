// a2: KLTregional (utils/KLT.py:55-95) as ONE call -- the entry point SURVEY.md 8(b) suggests (`vel_klt_regional`).
//
// The reference cuts the region of interest out of both frames (:60-66: boundingRect(p0) + 50 px, clipped), brings the current
// frame into the previous frame's coordinates -- an integer shift of the ROI (:62-66, translateFlag) or an affine remap of the
// ROI grid (:68-73) --, runs cv2calcOpticalFlowPyrLK with the forward-backward gate on the two crops (:75, utils/KLT.py:37-51:
// cv2 builds the pyramids OF THE CROPS, with REFLECT_101 at the crop borders) and maps the result back (:88-93).
// Everything between the ROI decision and the map-back runs here without leaving the stream: point shift, K3 remap, K1 pyramids
// of both crops into caller-provided scratch, K2 with the fused backward pass.  The ROI rectangle and T are host values (the
// reference computes them on the host from its point list); the map-back stays with the caller (a 3x2 float32 product whose
// rounding is numpy's).  With the ROI set to the whole frame and a zero shift this is plain cv2calcOpticalFlowPyrLK.
#include "common.cuh"

namespace {

__global__ void shift_points_kernel(const float2* __restrict__ p, int n, float dx, float dy, float2* __restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const float2 v = p[i];
        out[i] = make_float2(__fsub_rn(v.x, dx), __fsub_rn(v.y, dy));
    }
}

inline size_t al256(size_t b) { return (b + 255) & ~(size_t)255; }

struct RegionalLayout {
    vel_pyr_layout pyr;
    size_t off_pts, off_remap, off_pyr0, off_pyr1, total;
    int remap_pitch;
};

int regional_layout(int roi_w, int roi_h, int win_w, int win_h, int max_level, int npts, RegionalLayout* L)
{
    const int rc = vel_pyr_layout_make(roi_w, roi_h, win_w, win_h, max_level, &L->pyr);
    if (rc != VEL_OK) return rc;
    L->remap_pitch = (roi_w + 15) & ~15;
    size_t o = 0;
    L->off_pts = o; o += al256(sizeof(float) * 2 * (size_t)(npts > 0 ? npts : 1));
    L->off_remap = o; o += al256((size_t)L->remap_pitch * roi_h);
    L->off_pyr0 = o; o += al256((size_t)(L->pyr.bytes > 0 ? L->pyr.bytes : 16));
    L->off_pyr1 = o; o += al256((size_t)(L->pyr.bytes > 0 ? L->pyr.bytes : 16));
    L->total = o;
    return VEL_OK;
}

}  // namespace

VEL_API size_t vel_klt_regional_workspace(int32_t roi_w, int32_t roi_h, int32_t win_w, int32_t win_h, int32_t max_level, int32_t npts)
{
    RegionalLayout L;
    if (roi_w <= 0 || roi_h <= 0 || npts < 0 || regional_layout(roi_w, roi_h, win_w, win_h, max_level, npts, &L) != VEL_OK) return 0;
    return L.total;
}

VEL_API int vel_klt_regional(const uint8_t* im0, const uint8_t* im, int32_t width, int32_t height, int32_t pitch0, int32_t pitch,
                             const float* p0, int32_t npts, int32_t x0, int32_t x1, int32_t y0, int32_t y1, const float* T_host,
                             int32_t flags, const vel_lk_params* params, void* work, size_t work_bytes, float* pa_roi, uint8_t* status,
                             float* err, vel_stream_t stream)
{
    VEL_CHECK_ARG(im0 && im && params && work && T_host, "vel_klt_regional: NULL argument");
    VEL_CHECK_ARG(npts >= 0 && (npts == 0 || (p0 && pa_roi && status && err)), "vel_klt_regional: NULL point arrays");
    VEL_CHECK_ARG(width > 0 && height > 0 && pitch0 >= width && pitch >= width, "vel_klt_regional: bad frame geometry");
    VEL_CHECK_ARG(0 <= x0 && x0 < x1 && x1 <= width && 0 <= y0 && y0 < y1 && y1 <= height, "vel_klt_regional: ROI [%d,%d) x [%d,%d) outside the %d x %d frame",
                  x0, x1, y0, y1, width, height);
    const int rw = x1 - x0, rh = y1 - y0;
    RegionalLayout L;
    int rc = regional_layout(rw, rh, params->win_w, params->win_h, params->max_level, npts, &L);
    if (rc != VEL_OK) return rc;
    VEL_CHECK_ARG(rw > params->win_w && rh > params->win_h, "vel_klt_regional: the ROI (%d x %d) must be larger than the window (%d x %d)", rw, rh,
                  params->win_w, params->win_h);
    VEL_CHECK_ARG(work_bytes >= L.total && ((size_t)work & 255) == 0, "vel_klt_regional: workspace %zu B < required %zu B (or not 256-byte aligned)",
                  work_bytes, L.total);
    if (npts == 0) return VEL_OK;
    cudaStream_t st = (cudaStream_t)stream;
    char* wb = (char*)work;
    float* pts_roi = (float*)(wb + L.off_pts);
    uint8_t* remap = (uint8_t*)(wb + L.off_remap);
    uint8_t* pyr0 = (uint8_t*)(wb + L.off_pyr0);
    uint8_t* pyr1 = (uint8_t*)(wb + L.off_pyr1);

    const bool translate = (flags & VEL_KLT_TRANSLATE) != 0;
    const float* pts_in = p0;
    if (!(flags & VEL_KLT_POINTS_IN_ROI)) {       // :76 p0 - xy0 (float32 - float32)
        shift_points_kernel<<<(npts + 255) / 256, 256, 0, st>>>(reinterpret_cast<const float2*>(p0), npts, (float)x0, (float)y0,
                                                                reinterpret_cast<float2*>(pts_roi));
        VEL_LAUNCH_CHECK("shift_points_kernel");
        pts_in = pts_roi;
    }
    const uint8_t* roi_prev = im0 + (long long)y0 * pitch0 + x0;
    const uint8_t* roi_next;
    int next_pitch;
    if (translate) {
        const int dx = (int)T_host[4], dy = (int)T_host[5];           // :62 int(T[2, 0]), int(T[2, 1]): truncation toward zero
        VEL_CHECK_ARG(y0 + dy >= 0 && x0 + dx >= 0 && y1 + dy <= height && x1 + dx <= width,
                      "vel_klt_regional: the shifted ROI leaves the frame (cv2 asserts on mismatched pyramid sizes here)");
        roi_next = im + (long long)(y0 + dy) * pitch + (x0 + dx);
        next_pitch = pitch;
    } else {
        rc = vel_remap_affine_u8(im, width, height, pitch, T_host, x0, y0, rw, rh, remap, L.remap_pitch, stream);
        if (rc != VEL_OK) return rc;
        roi_next = remap;
        next_pitch = L.remap_pitch;
    }
    if (L.pyr.max_level > 0) {
        rc = vel_pyramid_u8(roi_prev, 0, pitch0, 1, &L.pyr, pyr0, L.pyr.bytes, stream);
        if (rc != VEL_OK) return rc;
        rc = vel_pyramid_u8(roi_next, 0, next_pitch, 1, &L.pyr, pyr1, L.pyr.bytes, stream);
        if (rc != VEL_OK) return rc;
    }
    return vel_lk_track(roi_prev, 0, pitch0, pyr0, 0, roi_next, 0, next_pitch, pyr1, 0, &L.pyr, 1, pts_in, 0, npts, params, pa_roi, status, err,
                        nullptr, stream);
}
