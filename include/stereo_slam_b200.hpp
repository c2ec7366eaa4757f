// stereo_slam_b200.hpp — drop-in `StereoSlam` for callers of the reference library.
//
// Same class name, constructor and member signatures as the reference's public header
// (eichenberger/stereo-svo-slam, src/include/stereo_slam.hpp:27-79), same public structs
// (src/include/stereo_slam_types.hpp:16-131, src/include/pose_manager.hpp:21-63), implemented on top of the
// C-ABI in svo_cuda.h (libstereosvo_b200.so).  src/app, src/ar-app and the Cython wrapper compile against this
// header unchanged: include it instead of "stereo_slam.hpp" and link -lstereosvo_b200 instead of -lstereosvo.
//
// Needs the OpenCV C++ headers (cv::Mat in the signatures), exactly like the reference header.  Where they are
// absent (this build container) the header compiles against the minimal stand-in types in
// tests/cpp/opencv_stub.hpp for a syntax/ABI check only (tests/test_facade_header.py).
#ifndef STEREO_SLAM_B200_HPP
#define STEREO_SLAM_B200_HPP

#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#if defined(SVO_USE_OPENCV_STUB)
#include "opencv_stub.hpp"
#else
#include <opencv2/opencv.hpp>
#endif

#include "svo_cuda.h"

// ---- public types: field-for-field the reference's (stereo_slam_types.hpp / pose_manager.hpp) ----------------
struct CameraSettings {
    float baseline, fx, fy, cx, cy, k1, k2, k3, p1, p2;
    int grid_height, grid_width, search_x, search_y;
    int window_size_pose_estimator, window_size_opt_flow, window_size_depth_calculator;
    int max_pyramid_levels, min_pyramid_level_pose_estimation;
};
static_assert(sizeof(CameraSettings) == sizeof(svo_camera_settings), "CameraSettings must match svo_camera_settings");

struct Pose { float x, y, z, rx, ry, rz; };
static_assert(sizeof(Pose) == sizeof(svo_pose), "Pose must match svo_pose");

class PoseManager {
public:
    PoseManager() : pose{0, 0, 0, 0, 0, 0} {}
    void set_pose(Pose &p)
    {
        pose = p;
        angles = cv::Vec3f(p.rx, p.ry, p.rz);
        translation = cv::Vec3f(p.x, p.y, p.z);
        cv::Rodrigues(angles, rot_mat);
        cv::Rodrigues(-angles, inv_rot_mat);
    }
    void set_vector(cv::Vec6f &v) { Pose p{v[0], v[1], v[2], v[3], v[4], v[5]}; set_pose(p); }
    cv::Matx33f get_rotation_matrix() const { return rot_mat; }
    cv::Matx33f get_inv_rotation_matrix() const { return inv_rot_mat; }
    cv::Vec3f get_translation() const { return translation; }
    cv::Vec3f get_angles() const { return angles; }
    Pose get_pose() const { return pose; }
    cv::Vec6f get_vector() const { return cv::Vec6f(pose.x, pose.y, pose.z, pose.rx, pose.ry, pose.rz); }

private:
    Pose pose;
    cv::Matx33f rot_mat, inv_rot_mat;
    cv::Vec3f angles, translation;
};

struct StereoImage {
    std::vector<cv::Mat> left, right, opt_flow;  // left: halfSample pyramid; right: level 0; opt_flow: LK image levels
};
enum KeyPointType { KP_FAST, KP_EDGELET };
struct KeyPoint2d { float x, y; };
struct KeyPoint3d { float x, y, z; };
struct Color { uint8_t r, g, b; };
struct KeyPointInformation {
    float score;
    int level;
    enum KeyPointType type;
    uint64_t keyframe_id;
    size_t keypoint_index;
    Color color;
    bool ignore_during_refinement, ignore_completely;
    int outlier_count, inlier_count;
    bool ignore_temporary;
    cv::KalmanFilter kf;  // statePost / errorCovPost are filled from the device-side filter state
};
struct KeyPoints {
    std::vector<KeyPoint2d> kps2d;
    std::vector<KeyPoint3d> kps3d;
    std::vector<KeyPointInformation> info;
};
struct Frame {
    uint64_t id;
    PoseManager pose;
    StereoImage stereo_image;
    KeyPoints kps;
    double time_stamp;
};
struct KeyFrame : Frame {};

class StereoSlam {
public:
    explicit StereoSlam(const CameraSettings &camera_settings) : camera_settings(camera_settings), slam(nullptr), width(0), height(0) {}
    ~StereoSlam() { if (slam) svo_slam_destroy(slam); }
    StereoSlam(const StereoSlam &) = delete;
    StereoSlam &operator=(const StereoSlam &) = delete;

    // StereoSlam::new_image (src/lib/stereo_slam.cpp:123): CV_8U single channel, ROIs (Mat::step) honoured
    void new_image(const cv::Mat &left, const cv::Mat &right, const float time_stamp)
    {
        if (left.type() != CV_8U || right.type() != CV_8U || left.size() != right.size())
            throw std::invalid_argument("StereoSlam::new_image: CV_8U single-channel images of equal size required");
        if (!slam) {
            width = left.cols; height = left.rows;
            svo_camera_settings s;
            std::memcpy(&s, &camera_settings, sizeof(s));
            check(svo_slam_create(&s, device_from_env(), width, height, &slam), nullptr);
        }
        if (left.cols != width || left.rows != height) throw std::invalid_argument("StereoSlam::new_image: image size changed");
        check(svo_slam_new_image(slam, left.data, left.step, right.data, right.step, time_stamp), slam);
    }
    void get_keyframe(KeyFrame &keyframe) { fetch(-1, keyframe, true); }
    void get_keyframes(std::vector<KeyFrame> &keyframes)
    {
        int n = slam ? svo_slam_keyframe_count(slam) : 0;
        keyframes.resize(n);
        for (int i = 0; i < n; i++) fetch(i, keyframes[i], true);
    }
    bool get_frame(Frame &frame)
    {
        if (!slam) return false;
        return fetch(0, frame, false);
    }
    void get_trajectory(std::vector<Pose> &trajectory)
    {
        int n = slam ? svo_slam_get_trajectory(slam, 0, nullptr) : 0;
        trajectory.resize(n);
        if (n) svo_slam_get_trajectory(slam, n, reinterpret_cast<svo_pose *>(trajectory.data()));
    }
    Pose update_pose(const Pose &pose, const cv::Vec6f &speed, const cv::Vec6f &pose_variance, const cv::Vec6f &speed_variance,
                     double dt)
    {
        Pose out = pose;
        if (!slam) return out;  // the motion filter lives with the device context, created by the first image
        float sp[6], pv[6], sv[6];
        for (int i = 0; i < 6; i++) { sp[i] = speed[i]; pv[i] = pose_variance[i]; sv[i] = speed_variance[i]; }
        check(svo_slam_update_pose(slam, reinterpret_cast<const svo_pose *>(&pose), sp, pv, sv, dt, reinterpret_cast<svo_pose *>(&out)), slam);
        return out;
    }

private:
    static int device_from_env()
    {
        const char *e = std::getenv("SVO_CUDA_DEVICE");
        return e ? std::atoi(e) : 0;
    }
    static void check(int rc, svo_slam *s)
    {
        // the reference has no error channel: device errors surface as exceptions
        if (rc != SVO_OK) throw std::runtime_error(std::string("stereosvo_b200: ") + svo_slam_last_error(s));
    }
    bool fetch(int index, Frame &f, bool keyframe)
    {
        uint64_t id = 0;
        svo_pose p;
        double ts = 0;
        int n = 0;
        int rc = keyframe ? svo_slam_get_keyframe(slam, index, &id, &p, &ts, &n) : svo_slam_get_frame(slam, &id, &p, &ts, &n);
        if (rc == SVO_ERR_STATE) return false;
        check(rc, slam);
        f.id = id; f.time_stamp = ts;
        Pose pp{p.x, p.y, p.z, p.rx, p.ry, p.rz};
        f.pose.set_pose(pp);
        std::vector<svo_keypoint_info> info(n);
        f.kps.kps2d.resize(n); f.kps.kps3d.resize(n); f.kps.info.resize(n);
        if (keyframe) check(svo_slam_get_keyframe_keypoints(slam, index, n, reinterpret_cast<float *>(f.kps.kps2d.data()),
                                                            reinterpret_cast<float *>(f.kps.kps3d.data()), info.data()), slam);
        else check(svo_slam_get_frame_keypoints(slam, n, reinterpret_cast<float *>(f.kps.kps2d.data()),
                                                reinterpret_cast<float *>(f.kps.kps3d.data()), info.data()), slam);
        for (int i = 0; i < n; i++) {
            KeyPointInformation &o = f.kps.info[i];
            const svo_keypoint_info &s = info[i];
            o.score = s.score; o.level = s.level; o.type = s.type == SVO_KP_FAST ? KP_FAST : KP_EDGELET;
            o.keyframe_id = s.keyframe_id; o.keypoint_index = (size_t)s.keypoint_index;
            o.color = Color{s.color[0], s.color[1], s.color[2]};
            o.ignore_during_refinement = s.ignore_during_refinement; o.ignore_completely = s.ignore_completely;
            o.ignore_temporary = s.ignore_temporary; o.outlier_count = s.outlier_count; o.inlier_count = s.inlier_count;
            o.kf.init(1, 1);
            o.kf.statePost.at<float>(0) = s.kf_inv_depth;
            o.kf.errorCovPost.at<float>(0, 0) = s.kf_variance;
        }
        auto pull = [&](int kind, int levels, std::vector<cv::Mat> &dst) {
            dst.resize(levels);
            svo_ctx *ctx = svo_slam_ctx(slam);
            for (int l = 0; l < levels; l++) {
                int w = 0, h = 0;
                svo_slot_level_size(ctx, kind, l, &w, &h);
                dst[l].create(h, w, CV_8U);
                int r = keyframe ? svo_slam_get_keyframe_image(slam, index, kind, l, dst[l].data, dst[l].step)
                                 : svo_slam_get_frame_image(slam, kind, l, dst[l].data, dst[l].step);
                check(r, slam);
            }
        };
        pull(0, camera_settings.max_pyramid_levels, f.stereo_image.left);
        pull(1, 1, f.stereo_image.right);
        pull(2, 3, f.stereo_image.opt_flow);  // image levels only (the Scharr levels live on the device: svo_download_level kind 3)
        return true;
    }

    const CameraSettings camera_settings;
    svo_slam *slam;
    int width, height;
};

#endif
