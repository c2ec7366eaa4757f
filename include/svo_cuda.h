/* svo_cuda.h — C-ABI of the B200 (sm_100a) tracking hot path of stereo-svo-slam.
 *
 * The reference (eichenberger/stereo-svo-slam) has no FFI seam of its own: its seam is the C++ class
 * StereoSlam (src/include/stereo_slam.hpp:27-79).  This header is the boundary a maintainer binds
 * instead of the reference's CPU stages; every entry point cites the reference code it replaces.
 * Two layers are exported from one shared library (libstereosvo_b200.so):
 *
 *   svo_*        device layer: context, device-resident image sets ("slots"), one entry point per
 *                hot-path stage, and the fused per-frame tracking call.
 *   svo_slam_*   host facade with the semantics of StereoSlam (new_image / get_frame / keyframes /
 *                trajectory / update_pose); written in C++ on top of svo_* only.
 *
 * Conventions: plain pointers and sizes, caller-owned host buffers, no exceptions; every function
 * returns an int status (SVO_OK == 0).  Images are 8-bit single channel with an explicit row stride
 * (cv::Mat::step — inputs may be ROIs, src/app/video_input.cpp:35-36).  One context per sequence; calls on one
 * context must be serialised by the caller (the reference is not re-entrant either);
 * any number of contexts may live in one process.  There is NO CPU fallback: without a CUDA device
 * every call fails with SVO_ERR_NO_DEVICE.
 */
#ifndef SVO_CUDA_H
#define SVO_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SVO_OK 0
#define SVO_ERR_INVALID 1   /* bad argument */
#define SVO_ERR_CUDA 2      /* CUDA runtime error, see svo_last_error */
#define SVO_ERR_NO_DEVICE 3 /* no CUDA device / extension built without one */
#define SVO_ERR_CAPACITY 4  /* more keypoints / keyframes than the context was created for */
#define SVO_ERR_STATE 5     /* call order violated (e.g. tracking before the first frame) */

/* Field-for-field the reference's CameraSettings (src/include/stereo_slam_types.hpp:16-36). */
typedef struct svo_camera_settings {
    float baseline, fx, fy, cx, cy, k1, k2, k3, p1, p2;
    int grid_height, grid_width, search_x, search_y;
    int window_size_pose_estimator, window_size_opt_flow, window_size_depth_calculator;
    int max_pyramid_levels, min_pyramid_level_pose_estimation;
} svo_camera_settings;

typedef struct svo_ctx svo_ctx;

/* keypoint flag bits == KeyPointInformation::{ignore_during_refinement, ignore_completely, ignore_temporary}
 * (src/include/stereo_slam_types.hpp:94-98) */
#define SVO_FLAG_IGNORE_REFINEMENT 1
#define SVO_FLAG_IGNORE_COMPLETELY 2
#define SVO_FLAG_IGNORE_TEMPORARY 4
/* KeyPointType (stereo_slam_types.hpp:52-55) */
#define SVO_KP_FAST 0
#define SVO_KP_EDGELET 1

/* ------------------------------------------------------------------ context ------------------- */
/* Replaces StereoSlam::StereoSlam's per-instance state (stereo_slam.cpp:29-41) on the device side. */
int svo_ctx_create(const svo_camera_settings *settings, int device, int width, int height, int max_keypoints,
                   svo_ctx **out);
int svo_ctx_destroy(svo_ctx *ctx);
/* back to an empty sequence: all image sets and keyframes are given back, keyframe tables and template cache rewound; device
 * memory, streams and captured frame graphs are kept (a new StereoSlam on the same camera without a single allocation) */
int svo_ctx_reset(svo_ctx *ctx);
const char *svo_last_error(svo_ctx *ctx); /* ctx may be NULL: last error of a failed svo_ctx_create */
int svo_device_count(void);

/* ------------------------------------------------------------------ image sets ---------------- */
/* Upload one stereo pair and build, on the device, everything new_image builds per frame
 * (stereo_slam.cpp:135-139): the halfSample pyramid of `left` (max_pyramid_levels levels, :93-121),
 * level 0 of `right`, and the 3-level LK pyramid of `left` (cv::buildOpticalFlowPyramid(left, win, 2):
 * Gaussian pyrDown, every level stored with its REFLECT_101 frame).  The Scharr derivative levels of that pyramid are
 * materialised only for image sets that become keyframes (svo_keyframe_commit) — OpenCV builds them for every frame, but
 * only a keyframe's are ever read (optical_flow.cpp:41-44) — or on demand (svo_download_level kind 3, svo_klt_slots).
 * Returns a slot id; slots are reference counted.  A keyframe owns a COPY of its frame's image set (the reference
 * shares the cv::Mat buffers, keyframe_manager.cpp:27-29; a copy lets frames alternate between two fixed slots). */
int svo_upload_stereo(svo_ctx *ctx, const uint8_t *left, size_t left_stride, const uint8_t *right, size_t right_stride,
                      int *slot_out);
/* Same, for images that already live in device memory of the context's GPU (e.g. a camera/decoder pipeline
 * that lands frames in HBM): device-to-device copy, then the same pyramid build. */
int svo_upload_stereo_device(svo_ctx *ctx, const uint8_t *d_left, size_t left_stride, const uint8_t *d_right,
                             size_t right_stride, int *slot_out);
/* EuRoC front end (src/app/euroc_input.cpp:48-49, :69-73): cv::initUndistortRectifyMap(K, D, R, P[:3,:3], size, CV_32F)
 * is evaluated once on the device; from then on every uploaded image `which` (0 = the `left` argument, 1 = `right`)
 * is a raw camera image and goes through cv::remap(INTER_LINEAR, BORDER_CONSTANT 0) fused in front of the pyramid
 * build (bit-exact with OpenCV's fixed-point remap).  K, R, P: row-major 3x3 (P = first three columns of the
 * 3x4 projection), D = (k1, k2, p1, p2, k3).  Source and rectified size are the context's width x height.
 * Note the reference feeds cam0 rectified with LEFT.* as `right` and cam1 with RIGHT.* as `left` (:69-70, :100-101). */
int svo_set_rectification(svo_ctx *ctx, int which, const double K[9], const double D[5], const double R[9],
                          const double P[9]);
int svo_clear_rectification(svo_ctx *ctx);
/* the float maps of input `which` (width*height each) — parity with cv::initUndistortRectifyMap */
int svo_rectification_maps(svo_ctx *ctx, int which, float *map1, float *map2);
/* number of keypoints a tracking frame / a keyframe may hold (max_keypoints of svo_ctx_create rounded up to 32; default 4 per
 * grid cell) */
int svo_keypoint_capacity(svo_ctx *ctx, int *max_keypoints);
int svo_slot_retain(svo_ctx *ctx, int slot);
int svo_slot_release(svo_ctx *ctx, int slot);
/* kind: 0 = left halfSample pyramid level, 1 = right level 0, 2 = LK pyramid level (unpadded view),
 *       3 = Scharr derivative level of the LK pyramid: (Ix, Iy) int16 pairs, 4 bytes per pixel (built on demand — the
 *           tracking path only materialises them for keyframes, the only image sets ever used as LK reference),
 *       4 / 5 = the LK image level / the Scharr level TOGETHER WITH the 32-pixel frame kept around them: BORDER_REFLECT_101
 *           pixels around the image, BORDER_CONSTANT zeros around the derivatives (cv::buildOpticalFlowPyramid's pyrBorder /
 *           derivBorder defaults, stereo_slam.cpp:139).  Sizes are in pixels; out_stride in bytes. */
int svo_slot_level_size(svo_ctx *ctx, int kind, int level, int *width, int *height);
int svo_download_level(svo_ctx *ctx, int slot, int kind, int level, uint8_t *out, size_t out_stride);

/* ------------------------------------------------------------------ stages -------------------- */
/* CornerDetector::detect_keypoints (corner_detector.cpp:13-79) on left pyramid level `level` with the
 * given grid: FAST-9/16 (threshold 6, NMS) best response per cell, else Sobel-x edgelet arg-max;
 * exactly one keypoint per cell, row-major cell order.  xy: n*2 (level coordinates), score: n, type: n. */
int svo_detect_keypoints(svo_ctx *ctx, int slot, int level, int grid_width, int grid_height, int max_out, float *xy,
                         float *score, int *type, int *n_out);
/* test/diagnostic: the raw FAST list (x, y, score) in raster order == cv::FastFeatureDetector(6) */
int svo_fast_corners(svo_ctx *ctx, int slot, int level, int max_out, int *xys, int *n_out);

/* Stereo SSD match: DepthCalculator::calculate_depth's template match (depth_calculator.cpp:200-239,
 * mode 0) and DepthFilter::calculate_disparities (depth_filter.cpp:259-327, mode 1: clamps to >= 0.5 and
 * returns -1 for windows outside the image).  Exact integer SSD (the specification behind
 * cv::matchTemplate(TM_SQDIFF)), first minimum in raster order, tie rule of :226-237. */
int svo_stereo_match(svo_ctx *ctx, int slot, const float *kps2d, int n, int mode, float *disparity);

/* PoseEstimator::estimate_pose (pose_estimator.cpp:115-130): coarse-to-fine sparse image alignment of the
 * previous frame's keypoints into the current frame, Gauss-Newton with step halving solved in-kernel.
 * flags may be NULL (all keypoints used); keypoints with SVO_FLAG_IGNORE_TEMPORARY are skipped
 * (pose_estimator.cpp:238-245).  evals16: per level l, [2l] = cost evaluations, [2l+1] = gradient calls. */
int svo_align(svo_ctx *ctx, int prev_slot, int cur_slot, const float *kps2d, const float *kps3d, const uint8_t *flags,
              int n, const float pose_in[6], float pose_out[6], float *cost, int *evals16);
/* one cost evaluation + one gradient at `pose` on pyramid level `level` (PoseEstimatorCallback::do_calc /
 * get_gradient, pose_estimator.cpp:275-300, :418-539) — parity probe */
int svo_align_probe(svo_ctx *ctx, int prev_slot, int cur_slot, const float *kps2d, const float *kps3d, int n, int level,
                    const float pose[6], float *cost, float grad[6]);

/* OpticalFlow::calculate_optical_flow -> cv::calcOpticalFlowPyrLK (optical_flow.cpp:14-56): win x win,
 * 3 levels, 30 iterations / eps 0.01, OPTFLOW_USE_INITIAL_FLOW, minEig 1e-4.  prev points live in the
 * keyframe whose id is keyframe_ids[i] (registered with svo_keyframe_commit); status 0 => err = +inf. */
int svo_klt(svo_ctx *ctx, const int *keyframe_ids, int cur_slot, const float *prev_pts, const float *init_pts, int n,
            float *next_pts, uint8_t *status, float *err);
/* same, between two explicit slots (parity tests) */
int svo_klt_slots(svo_ctx *ctx, int prev_slot, int cur_slot, const float *prev_pts, const float *init_pts, int n,
                  float *next_pts, uint8_t *status, float *err);

/* PoseRefiner::update_pose (pose_refinement.cpp:236-290, :321-412): Gauss-Newton on the reprojection
 * error. evals2 = {cost evaluations, gradient calls}. */
int svo_reproj_refine(svo_ctx *ctx, const float *kps2d, const float *kps3d, const uint8_t *flags, int n,
                      const float pose_in[6], float pose_out[6], float *cost, int *evals2);

/* DepthFilter::update_depth (depth_filter.cpp:40-50) as one stage: stereo SSD disparities at kps2d on image set `slot`
 * (calculate_disparities :259-327), outlier vote (:52-128), triangulation against the origin keyframe + 1x1 Kalman update
 * on the inverse depth (:130-257), then the flag post-processing and the re-projection with `pose` that
 * StereoSlam::new_image applies right after (stereo_slam.cpp:205-229).  kps3d / flags / counters / kf_state are updated in
 * place; disparity and kps2d_out may be NULL.  keyframe_ids refer to keyframes registered with svo_keyframe_commit. */
int svo_depth_filter_update(svo_ctx *ctx, int slot, int n, const float *kps2d, const float *ref_kps2d, const int *keyframe_ids,
                            float *kps3d, uint8_t *flags, int *inlier_count, int *outlier_count, float *kf_state,
                            const float pose[6], float *disparity, float *kps2d_out);

/* project_keypoints (transform_keypoints.cpp:11-47) */
int svo_project(svo_ctx *ctx, const float pose[6], const float *kps3d, int n, float *kps2d);

/* Register the image set `slot` as keyframe (KeyFrameManager::create_keyframe, keyframe_manager.cpp:15-32):
 * the device keeps a copy of the image set (LK pyramid for svo_klt, images for the getters) and the pose resident. */
int svo_keyframe_commit(svo_ctx *ctx, int slot, const float pose[6], int *keyframe_id_out);
/* slot holding keyframe `keyframe_id`'s own copy of the image set (for getters) */
int svo_keyframe_slot(svo_ctx *ctx, int keyframe_id, int *slot_out);
/* Build and keep the 31x31 LK templates (window samples, Scharr derivatives, structure tensor; all three levels) of the
 * keypoints keyframe `keyframe_id` introduced: entries first .. first + count - 1 of its keypoint list, at kps2d (count*2,
 * keyframe coordinates).  cv::calcOpticalFlowPyrLK (optical_flow.cpp:41-44) rebuilds them from the keyframe's pyramid for
 * every frame; pyramid and positions never change after KeyFrameManager::create_keyframe, so svo_frame_begin /
 * svo_track_frame fetch them instead when svo_track_io::keypoint_index is given (same bits either way).  No-op unless
 * window_size_opt_flow == 31 (env SVO_NO_TEMPLATES=1 turns the cache off). */
int svo_keyframe_set_templates(svo_ctx *ctx, int keyframe_id, const float *kps2d, int first, int count);

/* ------------------------------------------------------------------ fused per-frame tracking --- */
/* Everything StereoSlam::new_image does on a tracking frame between the pyramid build and the keyframe
 * decision (stereo_slam.cpp:196-229), enqueued as one stream-ordered sequence with a single host
 * synchronisation at the end:
 *   sparse alignment (estimate_pose :58-90) -> projection -> KLT against the origin keyframes + gating
 *   (pose_refinement.cpp:62-150) -> reprojection Gauss-Newton (:236-290) -> depth filter: stereo SSD
 *   disparities, outlier vote, triangulation + per-point Kalman update (depth_filter.cpp:40-327) ->
 *   flag post-processing (:205-226) -> re-projection of the updated points (:228-229).
 * All arrays are host memory of length n (x2 / x3 where noted); in/out arrays are updated in place. */
typedef struct svo_track_io {
    int n;
    const float *prev_kps2d;  /* n*2 in : previous frame's keypoints (alignment reference patches) */
    float *kps3d;             /* n*3 in/out : world points; out = depth-filter-updated */
    const float *ref_kps2d;   /* n*2 in : keypoint position in its origin keyframe */
    const int *keyframe_id;   /* n   in : origin keyframe id */
    uint8_t *flags;           /* n   in/out : SVO_FLAG_* */
    int *inlier_count;        /* n   in/out */
    int *outlier_count;       /* n   in/out */
    float *kf_state;          /* n*2 in/out : (inverse depth, variance) of the per-point Kalman filter */
    float pose_prior[6];      /* in  */
    float *kps2d;             /* n*2 out : re-projected positions under pose_refined */
    float pose_aligned[6];    /* out : after sparse alignment */
    float pose_refined[6];    /* out : after reprojection refinement (== frame pose before the motion filter) */
    float align_cost, refine_cost;
    int align_evals[16];
    int refine_evals[2];
    /* optional traces (NULL to skip) */
    float *klt_pts;           /* n*2 : tracked positions */
    float *klt_err;           /* n */
    uint8_t *klt_status;      /* n */
    float *disparity;         /* n */
    float *kps2d_refine_in;   /* n*2 : positions fed to the refinement / depth filter */
    int *klt_iters;           /* n   : LK iterations summed over the three levels */
    /* optional input (NULL: the LK templates are rebuilt from the keyframe pyramids for every frame) */
    const int *keypoint_index; /* n  in : index of the keypoint in its origin keyframe's list (KeyPointInformation::keypoint_index,
                                *         stereo_slam_types.hpp:93) — the key of the template cache, svo_keyframe_set_templates */
} svo_track_io;

int svo_track_frame(svo_ctx *ctx, int prev_slot, int cur_slot, svo_track_io *io);
/* svo_upload_stereo[_device] + svo_track_frame_begin in one call — everything new_image enqueues for a tracking frame
 * (stereo_slam.cpp:135-139 and :196-229).  In steady state everything after the frame ingest (half-sample pyramid, the LK
 * pyramid's two kernels with the keypoint import, alignment, KLT, refinement, stereo SSD, depth filter with the result
 * export: 8 kernels on two graph branches) is replayed as ONE CUDA graph launch; finish with svo_track_frame_end.
 * on_device != 0: left/right are device pointers. */
int svo_frame_begin(svo_ctx *ctx, const uint8_t *left, size_t left_stride, const uint8_t *right, size_t right_stride,
                    int on_device, int prev_slot, svo_track_io *io, int *cur_slot_out);
/* Number of SMs (CTAs of one thread-block cluster: 1, 2, 4, 8 or 16) that share the alignment solve of a frame.  8
 * minimises the latency of one sequence with a few hundred keypoints; 16 (a non-portable cluster size) pays with
 * thousands of keypoints per frame; 1-4 hold fewer SMs per solve when many sequences share the GPU.  0 (default) picks
 * 8, or 16 for frames with more than 1024 keypoints.  Results agree within the pose tolerance (the cross-keypoint sums
 * are grouped differently), each setting is deterministic.  Env SVO_ALIGN_CLUSTER sets the default. */
int svo_set_align_cluster(svo_ctx *ctx, int ctas);
/* Latency mode of the tracking kernels.  Width of the line searches of the two Gauss-Newton solvers (and, with it, two warps per
 * keypoint in the KLT kernel for frames of up to 1024 keypoints).  Most trial steps of a running sequence are rejected and halved, and
 * every trial pose x0 + 2^-j * step is known as soon as the step is: wide = 1 evaluates several of them per round on otherwise
 * idle SMs (refinement: an 8-CTA cluster, CTA c takes trial j + c; alignment: two half-clusters take trials j and j + 1) and
 * replays the reference's accept / halve / stop rule over the costs in order — the same evaluations, decisions and bits as the
 * sequential search (wide = 0), fewer dependent rounds, more SMs held per frame.  -1 (default): wide while fewer than four
 * sequences share the device (latency of one stream), sequential otherwise (throughput of many).  Env SVO_SOLVER_WIDTH sets
 * the default. */
int svo_set_solver_width(svo_ctx *ctx, int wide);
/* CUDA-graph replay on/off (default on; env SVO_NO_GRAPHS=1 turns it off); counters for tests */
int svo_set_graphs(svo_ctx *ctx, int on);
int svo_graph_stats(svo_ctx *ctx, long long *graph_launches, long long *graph_captures);
/* asynchronous pair (multi-sequence throughput): enqueue on the context's stream / wait + unpack */
int svo_track_frame_begin(svo_ctx *ctx, int prev_slot, int cur_slot, svo_track_io *io);
int svo_track_frame_end(svo_ctx *ctx, svo_track_io *io);

/* GPU time (ms, CUDA events on the context's stream) of the last svo_track_frame, and the number of kernel
 * launches it issued */
int svo_last_track_timing(svo_ctx *ctx, float *gpu_ms, int *launches);
/* cumulative number of kernels this context has launched since creation */
int svo_launch_count(svo_ctx *ctx, long long *launches);
/* per-stage CUDA-event timing of the tracking frame (adds one event record per stage when on).
 * stage_ms8 = {upload+pyramids, alignment, klt, reprojection refinement, stereo ssd, depth filter, d2h, total} */
int svo_set_profiling(svo_ctx *ctx, int on);
int svo_last_stage_ms(svo_ctx *ctx, float *stage_ms8);
int svo_sync(svo_ctx *ctx);
/* developer aid (env SVO_DEBUG_MARKS=1 at context creation): progress stamps written by the device between the stages
 * of a frame: out16[s] = sequence number of the last stamp of stage s (1 ingest start, 2 ingest done, 3 pyramids,
 * 4 inputs on device, 5 alignment, 6 KLT, 7 refinement, 8 SSD, 9 depth filter, 10 D2H enqueued), out16[0] = latest stamp,
 * out16[15] = stamps enqueued by the host, [16..63] = record of a barrier wait of the alignment kernel that timed out;
 * all -1 when the aid is off */
int svo_debug_marks(svo_ctx *ctx, int *out64);
/* developer aid (env SVO_SOLVER_TRACE=1 at context creation): phase stamps of the last alignment (which = 0) or refinement
 * (which = 1) solve, or of the last depth-filter / export kernel (which = 2; tags 1 matrices ready, 2 keypoint updated, 3 all
 * keypoints done, 4 export table ready, 5 copies issued, 6 copies done in the CTA, 7 system fence passed): out[0] = entries n, out[1..n] = SM clock << 16 | extra << 8 | tag (0 start, 1 level start, 2 level images
 * staged, 3 reference terms cached, 4 cost round done (extra = trial poses in it), 5 gradient round done, 6 end) */
int svo_debug_solver_trace(svo_ctx *ctx, int which, unsigned long long *out, int cap);
/* developer probe: bandwidth (GB/s) at which `ctas` CTAs of 256 threads read page-locked host memory over PCIe with the
 * access pattern of the frame-ingest kernel (16-byte loads) — the e2e roofline of zero-copy ingest */
int svo_debug_zero_copy_bandwidth(svo_ctx *ctx, const void *pinned_host, size_t bytes, int ctas, int reps, float *gb_per_s);

/* ================================================================== host facade ================ */
/* StereoSlam (src/include/stereo_slam.hpp:27-79) with plain-C types. */
typedef struct svo_slam svo_slam;

typedef struct svo_pose {  /* == struct Pose, src/include/pose_manager.hpp:21-28 */
    float x, y, z, rx, ry, rz;
} svo_pose;

typedef struct svo_keypoint_info {  /* KeyPointInformation without the cv::KalmanFilter (stereo_slam_types.hpp:86-100) */
    float score;
    int level;
    int type;
    uint64_t keyframe_id;
    uint64_t keypoint_index;
    uint8_t color[3];
    uint8_t ignore_during_refinement, ignore_completely, ignore_temporary;
    int outlier_count, inlier_count;
    float kf_inv_depth, kf_variance; /* statePost / errorCovPost of the 1x1 filter */
} svo_keypoint_info;

int svo_slam_create(const svo_camera_settings *settings, int device, int width, int height, svo_slam **out);
/* same, with an explicit keypoint capacity (max_keypoints of svo_ctx_create; 0 = default) */
int svo_slam_create_with_capacity(const svo_camera_settings *settings, int device, int width, int height, int max_keypoints,
                                  svo_slam **out);
int svo_slam_destroy(svo_slam *s);
/* a fresh StereoSlam on the same camera settings (frame, keyframes, trajectory, motion filter, counters of the sequence all
 * start over; keyframe ids restart at 0) that keeps every device resource of the old one: the next image is a first image */
int svo_slam_reset(svo_slam *s);
const char *svo_slam_last_error(svo_slam *s);
/* StereoSlam::new_image (stereo_slam.cpp:123-271) */
int svo_slam_new_image(svo_slam *s, const uint8_t *left, size_t left_stride, const uint8_t *right, size_t right_stride,
                       float time_stamp);
/* pipelined variant for many independent sequences: begin enqueues the frame, end finishes the host part */
int svo_slam_new_image_begin(svo_slam *s, const uint8_t *left, size_t left_stride, const uint8_t *right,
                             size_t right_stride, float time_stamp);
int svo_slam_new_image_end(svo_slam *s);
/* same as _begin with the stereo pair already resident in device memory (svo_upload_stereo_device) */
int svo_slam_new_image_device_begin(svo_slam *s, const uint8_t *d_left, size_t left_stride, const uint8_t *d_right,
                                    size_t right_stride, float time_stamp);
/* EurocInput's rectification (euroc_input.cpp:48-49, :69-73) applied on the device to every image passed as
 * `left` (which = 0) / `right` (which = 1) from now on; see svo_set_rectification */
int svo_slam_set_rectification(svo_slam *s, int which, const double K[9], const double D[5], const double R[9],
                               const double P[9]);
/* StereoSlam::get_frame (stereo_slam.cpp:278-284): returns SVO_ERR_STATE before the first image */
int svo_slam_get_frame(svo_slam *s, uint64_t *id, svo_pose *pose, double *time_stamp, int *n_keypoints);
int svo_slam_get_frame_keypoints(svo_slam *s, int max, float *kps2d, float *kps3d, svo_keypoint_info *info);
int svo_slam_get_frame_image(svo_slam *s, int kind, int level, uint8_t *out, size_t out_stride);
/* StereoSlam::get_keyframe(s) (stereo_slam.cpp:273-276, :286-289); index -1 = last keyframe */
int svo_slam_keyframe_count(svo_slam *s);
int svo_slam_get_keyframe(svo_slam *s, int index, uint64_t *id, svo_pose *pose, double *time_stamp, int *n_keypoints);
int svo_slam_get_keyframe_keypoints(svo_slam *s, int index, int max, float *kps2d, float *kps3d, svo_keypoint_info *info);
int svo_slam_get_keyframe_image(svo_slam *s, int index, int kind, int level, uint8_t *out, size_t out_stride);
/* StereoSlam::get_trajectory (stereo_slam.cpp:291-294): returns the number of poses, copies up to max */
int svo_slam_get_trajectory(svo_slam *s, int max, svo_pose *out);
/* StereoSlam::update_pose (stereo_slam.cpp:296-359) */
int svo_slam_update_pose(svo_slam *s, const svo_pose *pose, const float speed[6], const float pose_variance[6],
                         const float speed_variance[6], double dt, svo_pose *filtered);
/* diagnostics of the last new_image: device ms, kernel launches, whether a keyframe was created */
int svo_slam_last_stats(svo_slam *s, float *gpu_ms, int *launches, int *keyframe_created);
/* work counters of the last tracking frame: {keypoints, keypoints used by the alignment, alignment cost
 * evaluations, alignment gradient evaluations, refinement cost evaluations, refinement gradient evaluations,
 * LK iterations summed over keypoints and levels, keypoints tracked by LK (status 1)} */
int svo_slam_last_counters(svo_slam *s, long long *out8);
/* cumulative since creation: {frames, tracking frames, keyframes created, alignment patches (keypoint x evaluation),
 * LK windows (keypoint x level x iteration), keypoints summed over tracking frames, alignment evaluations, refinement
 * evaluations} — the numerators of frames/s, Mpatches/s and Mwindows/s */
int svo_slam_total_counters(svo_slam *s, long long *out8);
/* n independent sequences advanced by n_frames frames each in ONE call (StereoSlam::new_image for every (sequence, frame)):
 * left[i * n_frames + f] / right[...] / time_stamps[...] are frame f of sequence i (host pointers, or device pointers with
 * on_device != 0; one row stride for all).  `workers` host threads share the sequences (worker t owns sequences t, t + workers,
 * ...) and each walks its sequences through their frames on its own — no barrier between frames, the per-frame work of the
 * host (packing, launches, bookkeeping, keyframe creation) is native code on as many cores as the caller grants.  Frames of
 * one sequence are processed in order; results are those of n_frames svo_slam_new_image calls per sequence.  On failure the
 * first error code is returned, *failed_sequence (may be NULL) names the sequence, its message is in svo_slam_last_error. */
int svo_slam_run_many(svo_slam *const *slams, int n, int n_frames, const uint8_t *const *left, const uint8_t *const *right,
                      size_t left_stride, size_t right_stride, const float *time_stamps, int on_device, int workers,
                      int *failed_sequence);
/* same, for streams made of many finite sequences: restart[i * n_frames + f] != 0 (restart may be NULL) means frame f of
 * stream i is the FIRST image of a new sequence — svo_slam_reset runs before it */
int svo_slam_run_many_restart(svo_slam *const *slams, int n, int n_frames, const uint8_t *const *left, const uint8_t *const *right,
                              size_t left_stride, size_t right_stride, const float *time_stamps, const uint8_t *restart,
                              int on_device, int workers, int *failed_sequence);
/* the device context behind the facade (for stage-level probes in tests) */
svo_ctx *svo_slam_ctx(svo_slam *s);
/* new keypoints that keyframes could not take because the device keypoint block (svo_keypoint_capacity) was full; the
 * reference's lists are unbounded, here a keyframe then keeps its best-scored new keypoints (0 in every tested sequence) */
int svo_slam_dropped_keypoints(svo_slam *s, long long *dropped);

/* ================================================================== host stages ================ */
/* The bookkeeping of DepthCalculator / KeyFrameManager / StereoSlam that stays on the CPU between the device stages, one
 * entry point each.  Pure host code: no context, no device — they run (and are tested against the reference) without a GPU.
 * The facade calls the same functions. */
/* select_best_keypoints (depth_calculator.cpp:37-65): the per-level lists of CornerDetector::detect_keypoints folded into
 * one choice per LIST INDEX (entry j of every level competes with entry j of level 0): FAST beats edgelet, within a type the
 * coarser level wins unless the finer score is strictly higher; coarser positions are scaled by 2^level. */
int svo_host_select_best_keypoints(int n_levels, const int *n_per_level, const float *const *xy, const float *const *score,
                                   const int *const *type, int max_out, float *kps2d, float *score_out, int *type_out,
                                   int *level_out, int *n_out);
/* find_bad_keypoints (depth_calculator.cpp:67-86): keep[i] = 0 for keypoints outside [0, width] x [0, height] or flagged
 * ignore_completely / ignore_during_refinement */
int svo_host_find_bad_keypoints(int width, int height, int n, const float *kps2d, const uint8_t *flags, uint8_t *keep);
/* merge_keypoints (depth_calculator.cpp:88-130) as DepthCalculator::calculate_depth calls it (:179-180, grid arguments
 * swapped): appended[] = indices into the new list of the keypoints that join the old ones, in the order they are appended */
int svo_host_merge_keypoints(int width, int height, int grid_width, int grid_height, int n_old, const float *old_kps2d,
                             int n_new, const float *new_kps2d, int max_out, int *appended, int *n_out);
/* KeyFrameManager::keyframe_needed (keyframe_manager.cpp:47-74) */
int svo_host_keyframe_needed(int width, int height, int grid_width, int grid_height, int n, const float *kps2d,
                             const uint8_t *flags, int *needed);
/* the 12-state motion filter of StereoSlam (cv::KalmanFilter(12, 12), stereo_slam.cpp:29-41) and one
 * StereoSlam::update_pose step (:296-359); state_pre12 (may be NULL) = kf.statePre, the pose prior of the next image (:184-189) */
typedef struct svo_motion_filter svo_motion_filter;
int svo_motion_filter_create(svo_motion_filter **out);
int svo_motion_filter_destroy(svo_motion_filter *m);
int svo_motion_filter_update(svo_motion_filter *m, const svo_pose *pose, const float speed[6], const float pose_variance[6],
                             const float speed_variance[6], double dt, svo_pose *filtered, float state_pre12[12]);

#ifdef __cplusplus
}
#endif
#endif /* SVO_CUDA_H */
