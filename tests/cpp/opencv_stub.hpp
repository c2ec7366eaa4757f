// Minimal stand-ins for the handful of OpenCV types the public facade header uses — ONLY for compile-checking
// include/stereo_slam_b200.hpp in a container without the OpenCV C++ SDK.  Not a product file.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <vector>
#define CV_8U 0
#define CV_32F 5
namespace cv {
struct Size { int width = 0, height = 0; bool operator!=(const Size &o) const { return width != o.width || height != o.height; } };
struct Mat {
    int rows = 0, cols = 0, type_ = CV_8U;
    size_t step = 0;
    uint8_t *data = nullptr;
    std::vector<uint8_t> buf;
    Mat() {}
    Mat(int r, int c, int t) { create(r, c, t); }
    void create(int r, int c, int t)
    {
        rows = r; cols = c; type_ = t;
        size_t es = t == CV_32F ? 4 : 1;
        step = (size_t)c * es;
        buf.assign((size_t)r * step, 0);
        data = buf.data();
    }
    int type() const { return type_; }
    Size size() const { Size s; s.width = cols; s.height = rows; return s; }
    template <class T> T &at(int i) { return reinterpret_cast<T *>(data)[i]; }
    template <class T> T &at(int i, int j) { return reinterpret_cast<T *>(data + (size_t)i * step)[j]; }
};
template <class T, int N> struct Vec {
    T v[N];
    Vec() { for (int i = 0; i < N; i++) v[i] = 0; }
    Vec(T a, T b, T c) { static_assert(N == 3, ""); v[0] = a; v[1] = b; v[2] = c; }
    Vec(T a, T b, T c, T d, T e, T f) { static_assert(N == 6, ""); v[0] = a; v[1] = b; v[2] = c; v[3] = d; v[4] = e; v[5] = f; }
    T &operator[](int i) { return v[i]; }
    const T &operator[](int i) const { return v[i]; }
    Vec operator-() const { Vec r; for (int i = 0; i < N; i++) r.v[i] = -v[i]; return r; }
};
typedef Vec<float, 3> Vec3f;
typedef Vec<float, 6> Vec6f;
struct Matx33f { float val[9]; };
inline void Rodrigues(const Vec3f &r, Matx33f &R)
{
    double rx = r[0], ry = r[1], rz = r[2], th = std::sqrt(rx * rx + ry * ry + rz * rz);
    double c = std::cos(th), s = std::sin(th), c1 = 1 - c, it = th > 0 ? 1 / th : 0;
    rx *= it; ry *= it; rz *= it;
    double K[9] = {0, -rz, ry, rz, 0, -rx, -ry, rx, 0}, rr[9] = {rx * rx, rx * ry, rx * rz, rx * ry, ry * ry, ry * rz, rx * rz, ry * rz, rz * rz};
    for (int i = 0; i < 9; i++) R.val[i] = (float)(th > 0 ? c * (i % 4 == 0) + c1 * rr[i] + s * K[i] : (i % 4 == 0));
}
struct KalmanFilter {
    Mat statePost, errorCovPost;
    void init(int n, int) { statePost.create(n, 1, CV_32F); errorCovPost.create(n, n, CV_32F); }
};
}  // namespace cv
