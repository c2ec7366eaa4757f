// ORACLE — TEST INFRASTRUCTURE ONLY.  Never linked, imported or executed by the product path.
//
// extern "C" handle API over the REFERENCE's own StereoSlam class (src/include/stereo_slam.hpp:27-79), compiled
// together with the unmodified reference sources /root/reference/src/lib/*.cpp against oracle/cvshim into
// oracle/_ref/libstereosvo_ref.so (oracle/Makefile, target _ref).  Nothing here restates the reference: it only
// calls its public API (new_image / get_frame / get_keyframes / get_trajectory / update_pose) and flattens the
// structs of src/include/stereo_slam_types.hpp into arrays for ctypes.
//
// The library keeps process-global state (depth_calculator.cpp:135 `static uint64_t keyframe_count`,
// keyframe_manager.cpp:8 `KeyFrameManager::keyframe_counter`): one StereoSlam per loaded copy of the .so — oracle.py
// loads a private copy of the file per RefSlam instance.
#include <cstdint>
#include <cstring>
#include <vector>
#include <iostream>

#include "stereo_slam.hpp"

namespace {
struct Handle {
    StereoSlam slam;
    int w, h;
    explicit Handle(const CameraSettings &cs, int w_, int h_) : slam(cs), w(w_), h(h_) {}
};

// per keypoint: info8 = level, type, keyframe_id, keypoint_index, flags (1 ignore_during_refinement | 2 ignore_completely |
// 4 ignore_temporary), inlier_count, outlier_count, colour (r | g<<8 | b<<16); finfo3 = score, depth-filter state
// (kf.statePost, kf.errorCovPost; 0 when the filter was never initialised)
int dump_kps(const KeyPoints &k, float *kps2d, float *kps3d, int32_t *info8, float *finfo3, int max)
{
    int n = (int)k.info.size();
    int c = n < max ? n : max;
    for (int i = 0; i < c; i++) {
        if (kps2d && (size_t)i < k.kps2d.size()) { kps2d[2 * i] = k.kps2d[i].x; kps2d[2 * i + 1] = k.kps2d[i].y; }
        if (kps3d && (size_t)i < k.kps3d.size()) { kps3d[3 * i] = k.kps3d[i].x; kps3d[3 * i + 1] = k.kps3d[i].y; kps3d[3 * i + 2] = k.kps3d[i].z; }
        const KeyPointInformation &in = k.info[i];
        if (info8) {
            int32_t *o = info8 + 8 * i;
            o[0] = in.level; o[1] = (int)in.type; o[2] = (int32_t)in.keyframe_id; o[3] = (int32_t)in.keypoint_index;
            o[4] = (in.ignore_during_refinement ? 1 : 0) | (in.ignore_completely ? 2 : 0) | (in.ignore_temporary ? 4 : 0);
            o[5] = in.inlier_count; o[6] = in.outlier_count;
            o[7] = in.color.r | (in.color.g << 8) | (in.color.b << 16);
        }
        if (finfo3) {
            float *o = finfo3 + 3 * i;
            o[0] = in.score;
            o[1] = in.kf.statePost.empty() ? 0.f : in.kf.statePost.at<float>(0);
            o[2] = in.kf.errorCovPost.empty() ? 0.f : in.kf.errorCovPost.at<float>(0);
        }
    }
    return n;
}
void pose6(const PoseManager &pm, float *p)
{
    Pose q = pm.get_pose();
    p[0] = q.x; p[1] = q.y; p[2] = q.z; p[3] = q.rx; p[4] = q.ry; p[5] = q.rz;
}
}  // namespace

extern "C" {

// the reference prints several lines per frame and per solver step to std::cout (e.g. stereo_slam.cpp:22, :68, :191);
// quiet = 1 parks the stream in a failed state so those insertions cost nothing and the terminal stays usable
void ref_set_quiet(int quiet)
{
    if (quiet) std::cout.setstate(std::ios_base::failbit);
    else std::cout.clear();
}

// digits used by the reference's own progress prints (debugging aid: its per-stage poses and costs at full precision)
void ref_cout_precision(int digits) { std::cout.precision(digits); }

// CameraSettings is passed with the reference's own layout (src/include/stereo_slam_types.hpp:16-36)
void *ref_slam_create(const CameraSettings *cs, int w, int h) { return new Handle(*cs, w, h); }
void ref_slam_destroy(void *hd) { delete (Handle *)hd; }
int ref_sizeof_camera_settings() { return (int)sizeof(CameraSettings); }

// new_image borrows its inputs (pyr[0] aliases them, stereo_slam.cpp:115); like the reference app
// (src/app/slam_app.cpp:172-173) hand it clones it may keep.
int ref_slam_new_image(void *hd, const uint8_t *left, const uint8_t *right, int stride, float ts)
{
    Handle *H = (Handle *)hd;
    try {
        cv::Mat l = cv::Mat(H->h, H->w, CV_8U, (void *)left, (size_t)stride).clone();
        cv::Mat r = cv::Mat(H->h, H->w, CV_8U, (void *)right, (size_t)stride).clone();
        H->slam.new_image(l, r, ts);
    } catch (const std::exception &e) {
        std::cerr << "ref_slam_new_image: " << e.what() << std::endl;
        return -1;
    }
    return 0;
}
void ref_slam_update_pose(void *hd, const float *pose, const float *speed, const float *pv, const float *sv, double dt, float *out)
{
    Handle *H = (Handle *)hd;
    Pose p = {pose[0], pose[1], pose[2], pose[3], pose[4], pose[5]};
    Pose r = H->slam.update_pose(p, cv::Vec6f(speed), cv::Vec6f(pv), cv::Vec6f(sv), dt);
    out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.rx; out[4] = r.ry; out[5] = r.rz;
}
int ref_slam_get_pose(void *hd, float *p6)
{
    Frame f;
    if (!((Handle *)hd)->slam.get_frame(f)) return 0;
    pose6(f.pose, p6);
    return 1;
}
// current frame: returns the keypoint count (or -1 before the first image); id/time stamp through the out pointers
int ref_slam_frame(void *hd, float *p6, uint64_t *id, double *ts, float *kps2d, float *kps3d, int32_t *info8, float *finfo3, int max)
{
    Frame f;
    if (!((Handle *)hd)->slam.get_frame(f)) return -1;
    if (p6) pose6(f.pose, p6);
    if (id) *id = f.id;
    if (ts) *ts = f.time_stamp;
    return dump_kps(f.kps, kps2d, kps3d, info8, finfo3, max);
}
int ref_slam_n_keyframes(void *hd)
{
    std::vector<KeyFrame> kfs;
    ((Handle *)hd)->slam.get_keyframes(kfs);
    return (int)kfs.size();
}
int ref_slam_keyframe(void *hd, int k, float *p6, uint64_t *id, float *kps2d, float *kps3d, int32_t *info8, float *finfo3, int max)
{
    std::vector<KeyFrame> kfs;
    ((Handle *)hd)->slam.get_keyframes(kfs);
    if (k < 0 || k >= (int)kfs.size()) return -1;
    if (p6) pose6(kfs[k].pose, p6);
    if (id) *id = kfs[k].id;
    return dump_kps(kfs[k].kps, kps2d, kps3d, info8, finfo3, max);
}
// level `lvl` of the latest keyframe's pyramids: which = 0 left (halfSample), 1 right, 2 LK image level, 3 LK Scharr level (int16 x2)
int ref_slam_keyframe_image(void *hd, int which, int lvl, uint8_t *out, int max_bytes, int *w, int *h)
{
    KeyFrame kf;
    ((Handle *)hd)->slam.get_keyframe(kf);
    const std::vector<cv::Mat> &v = which == 0 ? kf.stereo_image.left : which == 1 ? kf.stereo_image.right : kf.stereo_image.opt_flow;
    size_t idx = which <= 1 ? (size_t)lvl : (size_t)(2 * lvl + (which == 3 ? 1 : 0));
    if (idx >= v.size()) return -1;
    const cv::Mat &m = v[idx];
    if (w) *w = m.cols;
    if (h) *h = m.rows;
    size_t rb = (size_t)m.cols * m.elemSize();
    if (out && (size_t)max_bytes >= rb * m.rows)
        for (int y = 0; y < m.rows; y++) memcpy(out + rb * y, m.ptr(y), rb);
    return (int)(rb * m.rows);
}
int ref_slam_trajectory(void *hd, float *out, int max)
{
    std::vector<Pose> t;
    ((Handle *)hd)->slam.get_trajectory(t);
    int n = (int)t.size() < max ? (int)t.size() : max;
    for (int i = 0; i < n; i++) { out[6 * i] = t[i].x; out[6 * i + 1] = t[i].y; out[6 * i + 2] = t[i].z; out[6 * i + 3] = t[i].rx; out[6 * i + 4] = t[i].ry; out[6 * i + 5] = t[i].rz; }
    return (int)t.size();
}

// ---- reference-owned stages on their own (the classes behind new_image), for stage-level pins of the oracle ----
}  // extern "C"

#include "pose_manager.hpp"
#include "transform_keypoints.hpp"
#include "exponential_map.hpp"
#include "corner_detector.hpp"
#include "image_comparison.hpp"

extern "C" {
// exponential_map.hpp:12-37
void ref_expmap(const float *tw, float *out)
{
    cv::Mat twist(6, 1, CV_32F), pose(6, 1, CV_32F);
    memcpy(twist.ptr<float>(), tw, 6 * sizeof(float));
    exponential_map(twist, pose);
    memcpy(out, pose.ptr<float>(), 6 * sizeof(float));
}
// transform_keypoints.cpp:11-47
void ref_project(const CameraSettings *cs, const float *p6, const float *pts3, int n, float *out2)
{
    Pose p = {p6[0], p6[1], p6[2], p6[3], p6[4], p6[5]};
    PoseManager pm;
    pm.set_pose(p);
    std::vector<KeyPoint3d> in(n);
    for (int i = 0; i < n; i++) { in[i].x = pts3[3 * i]; in[i].y = pts3[3 * i + 1]; in[i].z = pts3[3 * i + 2]; }
    std::vector<KeyPoint2d> out;
    project_keypoints(pm, in, *cs, out);
    for (int i = 0; i < n; i++) { out2[2 * i] = out[i].x; out2[2 * i + 1] = out[i].y; }
}
// corner_detector.cpp:13-79
int ref_detect_keypoints(const uint8_t *img, int w, int h, int grid_w, int grid_h, int level, int max, float *xy, float *score, int *type)
{
    cv::Mat im(h, w, CV_8U, (void *)img);
    CornerDetector det;
    std::vector<KeyPoint2d> k;
    std::vector<KeyPointInformation> inf;
    det.detect_keypoints(im, grid_w, grid_h, k, inf, level);
    int n = (int)inf.size() < max ? (int)inf.size() : max;
    for (int i = 0; i < n; i++) { xy[2 * i] = k[i].x; xy[2 * i + 1] = k[i].y; score[i] = inf[i].score; type[i] = (int)inf[i].type; }
    return (int)inf.size();
}
// image_comparison.cpp:103-120
float ref_total_intensity_diff(const uint8_t *im1, const uint8_t *im2, int w, int h, const float *k1, const float *k2, int n, int patch)
{
    cv::Mat a(h, w, CV_8U, (void *)im1), b(h, w, CV_8U, (void *)im2);
    std::vector<KeyPoint2d> v1(n), v2(n);
    for (int i = 0; i < n; i++) { v1[i].x = k1[2 * i]; v1[i].y = k1[2 * i + 1]; v2[i].x = k2[2 * i]; v2[i].y = k2[2 * i + 1]; }
    return get_total_intensity_diff(a, b, v1, v2, (size_t)patch);
}
}  // extern "C"
