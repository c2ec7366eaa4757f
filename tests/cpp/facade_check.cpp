// Compile/link check of the OpenCV-typed facade header against the C-ABI library (no GPU needed to build;
// at run time without a GPU the first new_image throws, which is what this program verifies).
#define SVO_USE_OPENCV_STUB 1
#include "../../include/stereo_slam_b200.hpp"
#include <cstdio>
int main()
{
    CameraSettings cs = {47.9f, 435.2f, 435.2f, 367.4f, 252.2f, 0, 0, 0, 0, 0, 24, 30, 50, 6, 4, 31, 31, 4, 2};
    StereoSlam slam(cs);
    Frame f;
    if (slam.get_frame(f)) return 2;            // no frame before the first image (stereo_slam.cpp:278-284)
    std::vector<Pose> traj;
    slam.get_trajectory(traj);
    if (!traj.empty()) return 3;
    cv::Mat l(480, 752, CV_8U), r(480, 752, CV_8U);
    try {
        slam.new_image(l, r, 0.f);
        slam.get_trajectory(traj);
        std::printf("tracked: trajectory %zu\n", traj.size());
        return traj.size() == 1 ? 0 : 4;
    } catch (const std::exception &e) {
        std::printf("exception: %s\n", e.what());
        return 10;                               // expected without a CUDA device
    }
}
