/*
 * velocity_b200.h -- C ABI of libvelocity_b200.so (B200 / sm_100a).
 *
 * This is the drop-in boundary for the SFM speed-estimation hot path of ultralytics/velocity.
 * The reference has no FFI layer of its own (pure Python calling cv2/numpy); each
 * entry point below replaces the arithmetic behind one reference call site, cited per function.
 * The Python package velocity_b200/ binds these with ctypes (see INTEGRATION.md for the stub a
 * reference maintainer would add to utils/KLT.py etc.).
 *
 * Conventions
 *   - plain pointers and sizes only; all data pointers are DEVICE pointers unless the name ends
 *     in _host; the caller owns every buffer, the library never frees or retains them.
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*; NULL = default
 *     stream), performs no hidden synchronisation and is re-entrant across streams.
 *   - return value: VEL_OK or a negative VEL_ERR_*; vel_last_error() returns a thread-local
 *     description of the last failure.  No exceptions cross this boundary.
 *   - points are float32 (x, y) pairs, 0-based pixel centres, exactly as cv2 takes them.
 *   - images are uint8, row-major, `pitch` bytes between rows.
 */
#ifndef VELOCITY_B200_H
#define VELOCITY_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VEL_OK 0
#define VEL_ERR_INVALID (-1)     /* bad argument */
#define VEL_ERR_CUDA (-2)        /* CUDA runtime error (message in vel_last_error) */
#define VEL_ERR_UNSUPPORTED (-3) /* valid request outside the implemented envelope */

#define VEL_MAX_LEVELS 8

typedef void* vel_stream_t; /* cudaStream_t */

/* Geometry of one frame's image pyramid.  Level 0 is the caller's frame itself (never copied);
 * levels >= 1 live in a separate pyramid buffer at byte offset `offset[l]`, row pitch `pitch[l]`. */
typedef struct vel_pyr_layout {
    int32_t max_level; /* effective top level after OpenCV's truncation rule (levels 0..max_level) */
    int32_t width[VEL_MAX_LEVELS];
    int32_t height[VEL_MAX_LEVELS];
    int32_t pitch[VEL_MAX_LEVELS]; /* pitch[0] is ignored: level-0 pitch is passed per call */
    int64_t offset[VEL_MAX_LEVELS];
    int64_t bytes; /* pyramid buffer bytes per frame (levels >= 1), 256-byte aligned */
} vel_pyr_layout;

/* cv2.calcOpticalFlowPyrLK parameters as the reference passes them (utils/KLT.py:106-107). */
typedef struct vel_lk_params {
    int32_t win_w, win_h;
    int32_t max_level;       /* requested; the layout carries the effective value */
    int32_t max_count;       /* TERM_CRITERIA_COUNT, clamped to [0,100] like OpenCV */
    double eps;              /* TERM_CRITERIA_EPS (squared internally), clamped to [0,10] */
    float min_eig_threshold; /* 1e-4 = OpenCV default, which is what the reference uses */
    float fb_threshold;      /* >= 0: fused backward pass + forward-backward gate (utils/KLT.py:47-50); < 0: forward only */
} vel_lk_params;

int vel_version(void);
const char* vel_last_error(void);

/* Pyramid geometry for a w x h frame tracked with a win_w x win_h window up to max_level
 * (OpenCV buildOpticalFlowPyramid rule: stop before the first level with width <= win_w or
 * height <= win_h).  Host-only helper, no CUDA call. */
int vel_pyr_layout_make(int32_t width, int32_t height, int32_t win_w, int32_t win_h, int32_t max_level,
                        vel_pyr_layout* out);

/* K1.  cv2.pyrDown chain (5-tap [1 4 6 4 1]/16 separable, BORDER_REFLECT_101, (s+128)>>8) for a
 * batch of frames: the pyramid that cv2.calcOpticalFlowPyrLK builds internally (utils/KLT.py:45,48).
 * frames + i*frame_stride is frame i (level 0, row pitch `pitch`); levels 1.. are written to
 * pyr + i*pyr_stride + layout->offset[l]. */
int vel_pyramid_u8(const uint8_t* frames, int64_t frame_stride, int32_t pitch, int32_t nframes,
                   const vel_pyr_layout* layout, uint8_t* pyr, int64_t pyr_stride, vel_stream_t stream);

/* cv2.resize(im, (0,0), fx=1/4, fy=1/4, INTER_NEAREST) == im[::4, ::4] (utils/KLT.py:111,113). */
int vel_decimate4_u8(const uint8_t* src, int32_t width, int32_t height, int32_t pitch, uint8_t* dst, int32_t dst_width,
                     int32_t dst_height, int32_t dst_pitch, vel_stream_t stream);

/* Frame ingest: cv2.cvtColor(imbgr, cv2.COLOR_BGR2GRAY) of vidExample.py:91 for a batch of
 * interleaved 8-bit BGR frames, in OpenCV's 15-bit fixed point:
 * gray = (3735*B + 19235*G + 9798*R + 16384) >> 15 (bit-exact against cv2 4.13). */
int vel_bgr2gray_u8(const uint8_t* bgr, int64_t bgr_stride, int32_t bgr_pitch, int32_t nframes, int32_t width, int32_t height,
                    uint8_t* gray, int64_t gray_stride, int32_t gray_pitch, vel_stream_t stream);

/* K9.  Feature initialisation: cv2.goodFeaturesToTrack(roi, maxCorners, quality, 0,
 * blockSize=5, useHarrisDetector=True) of vidExample.py:110 (Harris k = 0.04 there), in OpenCV 4.13's arithmetic
 * (fused Sobel taps, float64 running box sums, unfused response; restated in oracle/gftt_oracle.py): the same
 * corners in the same order as cv2.  img is one uint8 image; out_xy receives up to max_corners (x, y) float32
 * pairs, out_count[1] their number (both DEVICE); response, if not NULL, receives the float32 [height][width]
 * Harris response.  work must hold vel_good_features_workspace() bytes.  Only block_size 5 and minDistance 0
 * (what the reference passes) are implemented. */
size_t vel_good_features_workspace(int32_t width, int32_t height, int32_t max_corners);
int vel_good_features_harris_u8(const uint8_t* img, int32_t width, int32_t height, int32_t pitch, int32_t max_corners,
                                double quality, int32_t block_size, double k, void* work, size_t work_bytes, float* response,
                                float* out_xy, int32_t* out_count, vel_stream_t stream);

/* cv2.cornerSubPix(im, p, (win_w, win_h), (-1,-1), criteria) of vidExample.py:113-115 (CV_8UC1 image): refines npts
 * (x, y) float32 corners IN PLACE (DEVICE), in OpenCV 4.13's arithmetic (restated in oracle/velocity_oracle.c,
 * orc_corner_subpix_u8): bit-identical to cv2, including corners whose sampling window hangs over the frame border.
 * max_iters / eps are the criteria's COUNT / EPS members (clamped like cv2: 1..100, eps >= 0; pass 100 / 0 for an
 * absent member). */
int vel_corner_subpix_u8(const uint8_t* img, int32_t width, int32_t height, int32_t pitch, float* pts, int32_t npts, int32_t win_w,
                         int32_t win_h, int32_t max_iters, double eps, vel_stream_t stream);

/* K10.  cv2.estimateAffine2D(from, to, method=cv2.RANSAC) (utils/KLT.py:116,127,33) with caller-supplied threshold /
 * confidence / iteration budget (cv2's defaults: 3.0, 0.99, 2000) on npts (3..8192) float32 correspondences (DEVICE,
 * interleaved x,y).  OpenCV's RANSAC loop with its fixed-seed RNG (restated in oracle/ransac_oracle.py): the inlier
 * mask (uint8 [npts], DEVICE) is bit-identical to cv2's; T (double [6] = 2x3 row-major, DEVICE) is the model after
 * the refinement on the inliers (refine != 0; cv2's 10 LM steps converge to this least-squares fit: agreement
 * ~1e-12).  info (int32 [3], DEVICE) = {found (0/1), inlier count, RANSAC iterations run}. */
int vel_estimate_affine2d_ransac(const float* from_xy, const float* to_xy, int32_t npts, double threshold, double confidence,
                                 int32_t max_iters, int32_t refine, uint8_t* inliers, double* T, int32_t* info, vel_stream_t stream);

/* The tracker's call shape of the same fit (utils/KLT.py:116-117 and :127): `T, inl = cv2.estimateAffine2D(p0[v], p[v]); v[v] = inl`
 * with v the status mask of the LK stage before it, as ONE launch on device data: the rows of from_xy / to_xy [npts][2] with
 * mask[i] != 0 are compacted in order, the `to` points first mapped in float32 as to * to_scale + (to_off_x, to_off_y) -- the
 * `p /= scale` of :114 (to_scale 4) or the map-back of the translated ROI, :89 (the offset) --, the fit runs on the compacted
 * rows exactly as vel_estimate_affine2d_ransac, and mask_out[i] = mask[i] & inlier (mask_out may be mask itself).
 * to_mapped (may be NULL) receives the mapped `to` of ALL rows.  info (int32 [4], DEVICE) = {found, inliers, iterations, rows
 * kept}; fewer than 3 rows kept = not found (cv2 returns None).  work: vel_estimate_affine2d_ransac_masked_workspace(npts)
 * bytes, 8-byte aligned.  npts 1..8192. */
size_t vel_estimate_affine2d_ransac_masked_workspace(int32_t npts);
int vel_estimate_affine2d_ransac_masked(const float* from_xy, const float* to_xy, const uint8_t* mask, int32_t npts, float to_scale,
                                        float to_off_x, float to_off_y, double threshold, double confidence, int32_t max_iters,
                                        int32_t refine, void* work, size_t work_bytes, uint8_t* mask_out, float* to_mapped, double* T,
                                        int32_t* info, vel_stream_t stream);

/* K2.  cv2calcOpticalFlowPyrLK (utils/KLT.py:37-51) for a batch of frame pairs: pyramidal
 * Lucas-Kanade forward pass and, when params->fb_threshold >= 0, the backward pass from the
 * forward result fused in the same kernel with
 *     status = st_fwd & st_bwd & (||p1 - p1'||_2 < fb_threshold).
 * Pair k tracks prev frame k -> next frame k.  prev_pts + k*pts_stride floats holds npts (x,y)
 * pairs (pts_stride 0 = the same points for every pair).  Outputs are [npairs][npts]:
 * next_pts (x,y) float32, status uint8 (0/1), err float32 (cv2's L1 patch error of the forward
 * pass; 0 where the forward status is 0).  back_pts may be NULL. */
int vel_lk_track(const uint8_t* prev_frames, int64_t prev_frame_stride, int32_t prev_pitch, const uint8_t* prev_pyr,
                 int64_t prev_pyr_stride, const uint8_t* next_frames, int64_t next_frame_stride, int32_t next_pitch,
                 const uint8_t* next_pyr, int64_t next_pyr_stride, const vel_pyr_layout* layout, int32_t npairs,
                 const float* prev_pts, int64_t pts_stride, int32_t npts, const vel_lk_params* params, float* next_pts,
                 uint8_t* status, float* err, float* back_pts, vel_stream_t stream);

/* a2.  KLTregional (utils/KLT.py:55-95) as ONE call: crop both frames to the ROI [x0,x1) x [y0,y1) (HOST values: boundingRect(p0) + 50,
 * clipped, :60 / utils/images.py:6-19), bring the current frame into the previous frame's coordinates -- flags & VEL_KLT_TRANSLATE: the integer
 * shift int(T[2,0]), int(T[2,1]) of the ROI (:62-66); else the affine remap of the ROI grid (K3, :68-73) --, build the pyramids of the
 * two CROPS (what cv2 does inside calcOpticalFlowPyrLK: REFLECT_101 at the crop borders) and track p0 - (x0, y0) with the fused
 * forward-backward gate (utils/KLT.py:37-51).  T_host: the 3x2 row-vector affine as float32 (T00,T01,T10,T11,T20,T21), HOST pointer.
 * p0 [npts][2] DEVICE, full-frame coordinates.  Outputs (DEVICE): pa_roi [npts][2] the tracked points in ROI coordinates (the map-back
 * :88-93 stays with the caller), status [npts] (gated), err [npts].  With the ROI = the whole frame and a zero shift this is
 * cv2calcOpticalFlowPyrLK on the pair.  work: vel_klt_regional_workspace(...) bytes, 256-byte aligned.  Stream-ordered, no sync. */
#define VEL_KLT_TRANSLATE 1      /* flags: integer-shift form (translateFlag); otherwise the affine remap */
#define VEL_KLT_POINTS_IN_ROI 2  /* flags: p0 is already p0 - (x0, y0) (a caller that subtracts in float64 like numpy does for float64 points) */
size_t vel_klt_regional_workspace(int32_t roi_w, int32_t roi_h, int32_t win_w, int32_t win_h, int32_t max_level, int32_t npts);
int vel_klt_regional(const uint8_t* im0, const uint8_t* im, int32_t width, int32_t height, int32_t pitch0, int32_t pitch,
                     const float* p0, int32_t npts, int32_t x0, int32_t x1, int32_t y0, int32_t y1, const float* T_host,
                     int32_t flags, const vel_lk_params* params, void* work, size_t work_bytes, float* pa_roi, uint8_t* status,
                     float* err, vel_stream_t stream);

/* K3.  utils/KLT.py:70-73: float32 affine map of the ROI grid x in [x0,x0+dw), y in [y0,y0+dh)
 * through the 3x2 row-vector affine T (T[0..5] = T00,T01,T10,T11,T20,T21, HOST pointer), then
 * cv2.remap(INTER_LINEAR, BORDER_CONSTANT 0) in OpenCV's fixed point (1/32 px, 15-bit weights). */
int vel_remap_affine_u8(const uint8_t* src, int32_t width, int32_t height, int32_t pitch, const float* T_host, int32_t x0,
                        int32_t y0, int32_t dst_width, int32_t dst_height, uint8_t* dst, int32_t dst_pitch,
                        vel_stream_t stream);

/* K5.  fcnNLS_t (utils/NLS.py:102-129) for a batch of independent frames: 3-dof translation
 * Levenberg-Marquardt with the reference's forward-difference Jacobian (dx = 1e-6), fixed damping
 * JtJ + I, step schedule min(((i+1)*0.2)^2, 1), <= 30 iterations, stop at rms(delta) < 1e-8.
 * K is the 3x3 row-vector intrinsic matrix (float64, row-major, DEVICE).  Problem b uses
 * n[b] points at offset first[b] of p (float64 [.,2]) and pw (float64 [.,3]); x0/x are [nprob][3].
 * iters[b] receives the number of iterations run, or -(30) if the iteration cap was hit (the
 * reference prints a WARNING in that case).  All float64. */
int vel_nls_t(const double* K, const double* p, const double* pw, const int32_t* first, const int32_t* n, int32_t nprob,
              const double* x0, double* x, int32_t* iters, vel_stream_t stream);

/* fcnNLS_Rt (utils/NLS.py:133-183): 6-dof (roll,pitch,yaw,t) variant; x0/x are [nprob][6]. */
int vel_nls_rt(const double* K, const double* p, const double* pw, const int32_t* first, const int32_t* n, int32_t nprob,
               const double* x0, double* x, int32_t* iters, vel_stream_t stream);

/* K6.  fcn2vintercept (utils/MSV.py:98-142): mean of the closest-approach points over all
 * C(nf,2) ray pairs.  A [nf][3] ray origins, U [3][nf][nv] unit directions, C0 [nv][3]; float64. */
int vel_triangulate_2v(const double* A, const double* U, int32_t nf, int32_t nv, double* C0, vel_stream_t stream);

/* fcnNvintercept (utils/MSV.py:146-175): least-squares intersection of nf rays per point. */
int vel_triangulate_nv(const double* A, const double* U, int32_t nf, int32_t nv, double* C0, vel_stream_t stream);

/* fcnMSV1_t (utils/MSV.py:8-49): Levenberg-Marquardt on the LAST camera's translation x with the
 * pairwise triangulation (vel_triangulate_2v) re-run inside every iteration; the whole loop runs
 * in one persistent kernel.  A_fixed [nf-1][3] are the origins of cameras 0..nf-2, the last origin
 * is -x; U [3][nf][ng]; z [ng][2] are the last frame's pixels.  Outputs x [3], b0 [ng][3] (the
 * triangulated points relative to the last camera at the last evaluated iterate), iters (number of
 * iterations, or -max_iter if the cap was hit). */
int vel_msv1_t(const double* K, const double* A_fixed, const double* U, int32_t nf, int32_t ng, const double* z,
               const double* x0, int32_t max_iter, double* x, double* b0, int32_t* iters, vel_stream_t stream);

/* K7.  One Gauss-Newton linearisation of fcnNLS_batch (utils/NLS.py:186-250) in block form.
 * Parameters x = [points nt*3 | camera positions nc*3 | camera rpy nc*3] (camera 0 is fixed at
 * identity and not a parameter); z = observations [2][(nc+1)][nt] (all x then all y, track
 * fastest, utils/NLS.py:198-199).  Writes, for J = forward-difference Jacobian (1e-6):
 *   V  [nt][6]      upper triangle of the 3x3 point blocks of JtJ
 *   U  [nc][21]     upper triangle of the 6x6 camera blocks of JtJ (pos, rpy)
 *   W  [nc*6][nt*3] camera-point cross blocks as one row-major matrix: row 6*(c-1)+a (a = pos xyz,
 *                   rpy xyz), column 3*i+b                       (may be NULL)
 *   g  [nt*3+nc*6]  Jt (z - zhat), in parameter order
 *   cost[1]         sum of squared residuals
 * Cameras are indexed 0..nc (0 = the fixed identity camera, which has measurements but no
 * parameters).  cam_first/cam_count select the cameras this call (this rank) owns: it fills their
 * rows of U/W/g and PARTIAL sums over them in V, the point part of g and cost, to be all-reduced
 * across ranks.  U rows are indexed by camera-1. */
int vel_ba_accumulate(const double* K, const double* x, const double* z, int32_t nt, int32_t nc, int32_t cam_first,
                      int32_t cam_count, double* V, double* U, double* W, double* g, double* cost, vel_stream_t stream);

/* K8.  Solve (JtJ + I) delta = g by Schur complement on the point blocks and apply
 * x += 0.9 * delta (utils/NLS.py:234-235).  work must hold vel_ba_solve_workspace(nt, nc) bytes.
 * rms_delta[1] receives rms(delta) (the reference's convergence measure, :238). */
size_t vel_ba_solve_workspace(int32_t nt, int32_t nc);
int vel_ba_solve(const double* V, const double* U, const double* W, const double* g, int32_t nt, int32_t nc, double* x,
                 double* rms_delta, void* work, size_t work_bytes, vel_stream_t stream);

/* K7 + K8 as one device-resident loop: the whole `for i in range(10)` of fcnNLS_batch (utils/NLS.py:222-242) -- residuals, the
 * block-sparse forward-difference normal equations, the Schur solve, x += 0.9 delta and the `rms(delta) < tol: break` test (:238),
 * the test evaluated ON THE DEVICE: max_iter iterations are enqueued back to back, every kernel of an iteration returns at once
 * after the test has passed, and nothing synchronises with the host.  x [nt*3 + nc*6] is updated in place (layout as above).
 *   hist [max_iter][2] (DEVICE)  row i = (sum of squared residuals before update i, rms(delta_i)); NaN rows = not run
 *   iters_run[1]       (DEVICE)  number of iterations that ran
 * A failed Cholesky of the reduced system makes rms(delta) NaN for that iteration.  The cross blocks exist only in the scaled form
 * W' = W blockdiag(chol((V_i+I)^-1)) inside `work` (vel_ba_iterate_workspace(nt, nc) bytes, 256-byte aligned). */
size_t vel_ba_iterate_workspace(int32_t nt, int32_t nc);
int vel_ba_iterate(const double* K, const double* z, int32_t nt, int32_t nc, double* x, int32_t max_iter, double tol, double* hist,
                   int32_t* iters_run, void* work, size_t work_bytes, vel_stream_t stream);

/* K8 in three stages, for the camera-sharded form: every rank accumulates its cameras (vel_ba_accumulate with a
 * camera slice), the point blocks are all-reduced and the camera rows all-gathered; then
 *   vel_ba_reduce   per-point (V+I)^-1 factors, W' = W blockdiag(L), the rank's TILE ROWS [blk_lo, blk_hi) of the reduced camera
 *                   system S = U + I - W' W'^T (row-major, lower; tile-row geometry from vel_syrk_tile_rows(6*nc)), rhs = g_c - W y
 *   (the ranks send their rows of S to the owner)
 *   vel_ba_factor   owner only: Cholesky of S, delta_c = S^-1 rhs in place in rhs
 *   (the owner broadcasts rhs)
 *   vel_ba_update   t = W^T delta_c, delta_p, x += 0.9 delta, rms(delta)
 * blk_lo = 0, blk_hi = -1 computes all of S; then the three calls equal vel_ba_solve.  vel_ba_solve_layout returns the byte offsets of
 * S [6nc][6nc] and rhs [6nc] inside `work` (vel_ba_solve_workspace bytes). */
int vel_ba_solve_layout(int32_t nt, int32_t nc, int64_t* off_S, int64_t* off_rhs);
int vel_ba_reduce(const double* V, const double* U, const double* W, const double* g, int32_t nt, int32_t nc, int32_t blk_lo, int32_t blk_hi,
                  void* work, size_t work_bytes, vel_stream_t stream);
int vel_ba_factor(int32_t nt, int32_t nc, void* work, size_t work_bytes, vel_stream_t stream);
int vel_ba_update(const double* W, int32_t nt, int32_t nc, double* x, double* rms_delta, void* work, size_t work_bytes, vel_stream_t stream);

/* K8 building blocks -- the dense float64 algebra of `inv(JJ^T + I) @ ...` (utils/NLS.py:236) after the point blocks are
 * eliminated, hand-written (FP64 tensor-core MMA; no cuBLAS / cuSOLVER anywhere in this library):
 *   vel_syrk_lower_sub: S(lower triangle, row-major [m][lds]) -= E E^T for E row-major [m][ld] with k used columns;
 *     rows must be 16-byte aligned (ld even) and zero-padded to a multiple of 16 columns.  Deterministic (split-K partial
 *     products are applied in a fixed order).  work: vel_syrk_lower_sub_workspace(m, k) bytes.
 *   vel_spd_solve: Cholesky S = L L^T in place (lower, row-major) and b <- S^-1 b; info[0] (DEVICE) = 0, or 1 when S is not
 *     positive definite to working precision (b is then meaningless). */
size_t vel_syrk_lower_sub_workspace(int32_t m, int32_t k);
/* the tile-row geometry of the product (nb blocks of rows_per_block rows) and the product restricted to tile rows [blk_lo, blk_hi) */
int vel_syrk_tile_rows(int32_t m, int32_t* nb, int32_t* rows_per_block);
int vel_syrk_lower_sub_rows(const double* E, int64_t ld, int32_t m, int32_t k, double* S, int64_t lds, void* work, size_t work_bytes,
                            int32_t blk_lo, int32_t blk_hi, vel_stream_t stream);
int vel_syrk_lower_sub(const double* E, int64_t ld, int32_t m, int32_t k, double* S, int64_t lds, void* work, size_t work_bytes,
                       vel_stream_t stream);
int vel_spd_solve(double* S, int64_t lds, int32_t n, double* b, int32_t* info, vel_stream_t stream);

/* Measurement aid (bench.py's roofline denominator for the K8 SYRK; MEASURED_PEAKS.json has no FP64 entry): the FP64 tensor-core
 * (mma.sync m8n8k4 f64) rate of the current device with operands in registers, TFLOP/s, best of 3.  Synchronises; 0 on failure. */
double vel_fp64_mma_peak_tflops(void);

/* fcnNLS_batch2 (utils/NLS.py:253-328): the range/elevation/azimuth-parametrised bundle adjustment.
 * x = [points nt*3 | q], q = [joint roll,pitch,yaw | el | az | range_1..range_nc] (nq = 5 + nc).
 * vel_ba2_accumulate writes V [nt][6], the dense symmetric camera-side block G [nq][nq], the cross
 * block W [nq][nt*3] (row-major), g [nt*3 + nq] and cost[1]; vel_ba2_solve solves
 * (JtJ + I) delta = g through the Schur complement of the point blocks and applies x += 0.9 delta.
 * Workspace for vel_ba2_solve: vel_ba_solve_workspace(nt, (nq + 5) / 6) bytes. */
int vel_ba2_accumulate(const double* K, const double* x, const double* z, int32_t nt, int32_t nc, double* V, double* G,
                       double* W, double* g, double* cost, vel_stream_t stream);
int vel_ba2_solve(const double* V, const double* G, const double* W, const double* g, int32_t nt, int32_t nq, double* x,
                  double* rms_delta, void* work, size_t work_bytes, vel_stream_t stream);

/* K4.  cv2.BFMatcher(NORM_HAMMING).knnMatch(q, t, k=2) for 256-bit descriptors (the ORB variant
 * of the reference's descriptor fallback, utils/KLT.py:16-26; BASELINE config 4): the two nearest
 * train rows per query, ascending distance, ties to the lower train index.  q [nq][32], t [nt][32]
 * uint8; idx/dist [nq][2] int32 (-1 where nt < 2). */
int vel_match_knn2_hamming256(const uint8_t* q, int32_t nq, const uint8_t* t, int32_t nt, int32_t* idx, int32_t* dist,
                              vel_stream_t stream);

/* cv2.BFMatcher().knnMatch(q, t, k=2), NORM_L2 on float32 descriptors (what utils/KLT.py:16,25
 * literally runs on SURF descriptors).  dist = sqrt(sum (a-b)^2) in float32. */
int vel_match_knn2_l2(const float* q, int32_t nq, const float* t, int32_t nt, int32_t dim, int32_t* idx, float* dist,
                      vel_stream_t stream);

/* ---- Device-resident sequence bookkeeping (the connected C3 pipeline) ------------------------
 * The reference's frame loop (vidExample.py:75-166) keeps p / vg / P / B / S in numpy and touches them between
 * every cv2 and solver call.  These entry points keep the same state in HBM, FRAME-MAJOR:
 *   tracks [n][npts][2] float32  == P[0:2] transposed (row k = the tracked points in frame k, row 0 = the seeds)
 *   alive  [n][npts]    uint8    == vg after frame k
 *   proj   [n][npts][2] float32  == P[2:4] transposed (reprojections; NaN where a point is not in the fit)
 *   B [n][14], S [n][9] float32  exactly the reference's arrays (vidExample.py:44-45)
 * A failed track stays in place at an out-of-frame sentinel (K2 drops it without touching memory); the reference
 * compacts instead (p[v]) -- the surviving tracks are identical either way. */

/* vidExample.py:134-135 for a run of nframes consecutive frames: for k = 0..nframes-2
 *     tracks[k+1], status = cv2calcOpticalFlowPyrLK(frame k, frame k+1, tracks[k], fbt)   (K2, vel_lk_track)
 *     alive[k+1] = alive[k] & status
 * tracks row 0 / alive row 0 are inputs.  frames / pyr as in vel_lk_track (pyramids built by vel_pyramid_u8);
 * err [nframes-1][npts]; status is scratch of npts bytes.  The loop runs inside the library: no host work per frame. */
int vel_klt_sequence(const uint8_t* frames, int64_t frame_stride, int32_t pitch, const uint8_t* pyr, int64_t pyr_stride,
                     const vel_pyr_layout* layout, int32_t nframes, int32_t npts, const vel_lk_params* params, float* tracks,
                     uint8_t* alive, float* err, uint8_t* status, vel_stream_t stream);

/* vidExample.py:139-146 for frames 1..nframes-1 in one launch (one CTA per frame): t = fcnNLS_t(K, p[vp], p3[vp], x0)
 * (utils/NLS.py:102-129, float64, result rounded to float32 as the reference does), p_proj = world2image(K, I, t, p3),
 * residual = rms(p - p_proj).  A point enters frame f's fit when alive[f][i] != 0 and (subset == NULL or subset[i] != 0)
 * (subset = the reference's vp mask, vidExample.py:124,136).  Writes B[f,3:6] = t, B[f,0:3] = B[0,0:3] + t, S[f,2] =
 * number of live tracks, S[f,3] = residual, proj[f] (may be NULL), iters[f] (iterations, or -30 at the cap).
 * K, p3 float64 (DEVICE); x0_host = the start value (HOST double[3]; the reference always starts from (0,0,1)). */
int vel_seq_pose_t(const double* K, const float* tracks, const uint8_t* alive, const uint8_t* subset, const double* p3,
                   int32_t nframes, int32_t npts, const double* x0_host, float* B, float* S, float* proj, int32_t* iters,
                   vel_stream_t stream);

/* vidExample.py:142-146,164: S[i,0] = i, S[i,4] = dt = B[i,12]-B[i-1,12], S[i,5] = B[i,12]-B[0,12], S[i,6] = dr =
 * norm(t + B[0,0:3] - B[i-1,0:3]), S[i,7] = cumulative distance, S[i,8] = dr/dt*3.6 (km/h), S[0,2] = live tracks of
 * frame 0 -- all in the reference's float32 arithmetic.  B[:,12] (frame times) and B[:,0:6] must be filled. */
int vel_seq_stats(const float* B, const uint8_t* alive, int32_t nframes, int32_t npts, float* S, vel_stream_t stream);

/* utils/NLS.py:190-191: the full-length tracks (alive in the last frame, and in subset if given), as an
 * order-preserving index list idx[0..count) (DEVICE int32; count[1] DEVICE). */
int vel_seq_select(const uint8_t* alive_last, const uint8_t* subset, int32_t npts, int32_t* idx, int32_t* count, vel_stream_t stream);

/* utils/MSV.py:13-16: U [3][nframes][nsel] = pixel2uvec(K, tracks[f][idx[j]]) (float64) and, when A != NULL, the ray
 * origins A [nframes][3] = B[0,0:3] - B[f,0:3].  idx may be NULL (identity, nsel == npts). */
int vel_seq_rays(const double* K, const float* tracks, const int32_t* idx, int32_t nframes, int32_t npts, int32_t nsel,
                 const float* B, double* U, double* A, vel_stream_t stream);

/* utils/NLS.py:198-203: the bundle-adjustment inputs of fcnNLS_batch(K, P, pw, cw = B[:,3:6]) from the device arrays:
 * z [2][nframes][nsel] float64 and x = [pw (nsel*3) | B[1:,3:6] | zeros] (the layout vel_ba_accumulate takes). */
int vel_seq_pack_ba(const float* tracks, const int32_t* idx, int32_t nframes, int32_t npts, int32_t nsel, const double* pw,
                    const float* B, double* z, double* x, vel_stream_t stream);

/* The commented call site vidExample.py:157, `B[0:i, 3:6], p3[vg] = fcnNLS_batch(K, P, p3, B[0:i, 3:6])`: B_ba = B with the
 * camera positions replaced by the bundle-adjusted ones (x = the parameter vector of vel_ba_solve, nsel points) and
 * S_ba = S (follow with vel_seq_stats(B_ba, ..., S_ba) for the speed table after the adjustment). */
int vel_seq_ba_cameras(const double* x, int32_t nsel, int32_t nframes, const float* B, const float* S, float* B_ba, float* S_ba,
                       vel_stream_t stream);

/* vidExample.py:128,151-153: the reference's P [5][npts][nframes] float32 (x, y, xproj, yproj, frame index; NaN =
 * invalid) from the frame-major device arrays (proj may be NULL). */
int vel_seq_export_P(const float* tracks, const float* proj, const uint8_t* alive, int32_t nframes, int32_t npts, float* P,
                     vel_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* VELOCITY_B200_H */
