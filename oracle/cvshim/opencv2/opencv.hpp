// ORACLE — TEST INFRASTRUCTURE ONLY. Never included, linked or executed by the product path.
//
// <opencv2/opencv.hpp> stand-in: just the slice of the OpenCV 4.x C++ API that the reference library
// (eichenberger/stereo-svo-slam, src/lib/*.cpp + src/include/*.hpp) uses, so that the UNMODIFIED reference
// sources compile in an image that has no OpenCV C++ SDK (oracle/Makefile, target _ref).  The result,
// oracle/_ref/libstereosvo_ref.so, is the reference's own control flow and arithmetic; what is ours is this file.
//
// What this header is:
//  * containers with OpenCV's semantics where the reference depends on them: cv::Mat is a reference-counted
//    header (copies share pixels; KeyPointInformation copies therefore share their KalmanFilter state, and a
//    keyframe shares its frame's images), Mat::create() keeps an existing buffer of the right size, ROIs alias;
//    cv::Matx/Vec arithmetic is element-wise in _Tp with `s = 0; for k: s += a(i,k)*b(k,j)` products and
//    saturate_cast<_Tp>(a*alpha) scaling in the scalar's own type (matx.hpp: Matx_MatMulOp / Matx_ScaleOp);
//  * the OpenCV-owned algorithms (FAST, Sobel, matchTemplate, minMaxLoc, buildOpticalFlowPyramid,
//    calcOpticalFlowPyrLK, projectPoints, Rodrigues, KalmanFilter, invert/solve with DECOMP_SVD) forwarded to
//    oracle/ocv_prims.hpp, whose restatements are pinned against the real library (cv2 4.13.0) by
//    tests/test_oracle_golden.py.  matchTemplate(TM_SQDIFF) returns the exact integer SSD as float — OpenCV's
//    own map carries ±~40 of DFT noise; the exact map is the specification.
//  * <math.h> is included on purpose: the real opencv.hpp pulls it in (opencv2/flann/lsh_table.h), and with
//    libstdc++ that puts the float overloads of cos/sin/sqrt/floor/fabs into the global namespace — which decides
//    whether `cos(_norm)` in exponential_map.hpp:33 is float or double arithmetic.
//
// Everything not used by src/lib is absent.  Compile with -ffp-contract=off (see ocv_prims.hpp).
#pragma once
#include <math.h>
#include <cstdint>
#include <cstring>
#include <cassert>
#include <cmath>
#include <cfloat>
#include <vector>
#include <map>
#include <memory>
#include <string>
#include <iostream>
#include <algorithm>
#include <chrono>
#include <stdexcept>
#include <limits>
#include <type_traits>
#include <initializer_list>

#include "../../ocv_prims.hpp"

typedef unsigned char uchar;

#define CV_8U 0
#define CV_8S 1
#define CV_16U 2
#define CV_16S 3
#define CV_32S 4
#define CV_32F 5
#define CV_64F 6
#define CV_CN_SHIFT 3
#define CV_MAT_DEPTH(t) ((t) & 7)
#define CV_MAT_CN(t) ((((t) >> CV_CN_SHIFT) & 63) + 1)
#define CV_MAKETYPE(depth, cn) (CV_MAT_DEPTH(depth) + (((cn) - 1) << CV_CN_SHIFT))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_16SC2 CV_MAKETYPE(CV_16S, 2)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_32FC2 CV_MAKETYPE(CV_32F, 2)
#define CV_64FC1 CV_MAKETYPE(CV_64F, 1)

namespace cv {

class Exception : public std::runtime_error {
public:
    explicit Exception(const std::string &m) : std::runtime_error(m) {}
};
#define CVSHIM_ASSERT(c) do { if (!(c)) throw cv::Exception(std::string("cvshim assertion failed: ") + #c); } while (0)

enum DecompTypes { DECOMP_LU = 0, DECOMP_SVD = 1, DECOMP_EIG = 2, DECOMP_CHOLESKY = 3, DECOMP_QR = 4, DECOMP_NORMAL = 16 };
enum TemplateMatchModes { TM_SQDIFF = 0, TM_SQDIFF_NORMED = 1, TM_CCORR = 2, TM_CCORR_NORMED = 3, TM_CCOEFF = 4, TM_CCOEFF_NORMED = 5 };
enum { OPTFLOW_USE_INITIAL_FLOW = 4, OPTFLOW_LK_GET_MIN_EIGENVALS = 8 };
enum BorderTypes { BORDER_CONSTANT = 0, BORDER_REPLICATE = 1, BORDER_REFLECT = 2, BORDER_WRAP = 3, BORDER_REFLECT_101 = 4, BORDER_DEFAULT = 4 };

template <typename T> static inline T saturate_cast(double v) { return (T)v; }
template <typename T> static inline T saturate_cast(float v) { return (T)v; }
template <typename T> static inline T saturate_cast(int v) { return (T)v; }

// ------------------------------------------------------------------------------------------- Ptr
// cvstd_wrapper.hpp: a std::shared_ptr with the OpenCV 3.x extras the reference uses
// (implicit construction from a raw pointer, implicit conversion to T*, empty(), release()).
template <typename T> struct Ptr : public std::shared_ptr<T> {
    Ptr() {}
    Ptr(std::nullptr_t) {}
    template <typename Y> Ptr(Y *p) : std::shared_ptr<T>(p) {}
    Ptr(const std::shared_ptr<T> &o) : std::shared_ptr<T>(o) {}
    template <typename Y> Ptr(const Ptr<Y> &o) : std::shared_ptr<T>(o) {}
    void release() { std::shared_ptr<T>::reset(); }
    operator T *() const { return std::shared_ptr<T>::get(); }
    bool empty() const { return std::shared_ptr<T>::get() == nullptr; }
};
template <typename T, typename... A> static inline Ptr<T> makePtr(A &&...a) { return Ptr<T>(std::make_shared<T>(std::forward<A>(a)...)); }

// ------------------------------------------------------------------------------------------- Matx / Vec
template <typename T, int m, int n> struct Matx {
    enum { rows = m, cols = n, channels = m * n };
    T val[m * n];
    Matx() { for (int i = 0; i < m * n; i++) val[i] = T(0); }
    explicit Matx(T v0) { for (int i = 0; i < m * n; i++) val[i] = T(0); val[0] = v0; }
    explicit Matx(const T *vals) { for (int i = 0; i < m * n; i++) val[i] = vals[i]; }
    Matx(std::initializer_list<T> l) { int i = 0; for (T v : l) { if (i < m * n) val[i++] = v; } for (; i < m * n; i++) val[i] = T(0); }
    // Matx(v0, v1, ...): OpenCV spells these out for 2..16 values; remaining entries are zero.
    template <typename A0, typename A1, typename... A,
              typename = typename std::enable_if<(2 + sizeof...(A) <= m * n) && std::is_arithmetic<A0>::value && std::is_arithmetic<A1>::value>::type>
    Matx(A0 a0, A1 a1, A... a)
    {
        const T tmp[] = {static_cast<T>(a0), static_cast<T>(a1), static_cast<T>(a)...};
        int k = (int)(sizeof(tmp) / sizeof(T));
        for (int i = 0; i < m * n; i++) val[i] = i < k ? tmp[i] : T(0);
    }
    static Matx zeros() { return Matx(); }
    static Matx all(T v) { Matx r; for (int i = 0; i < m * n; i++) r.val[i] = v; return r; }
    static Matx eye() { Matx r; for (int i = 0; i < (m < n ? m : n); i++) r.val[i * n + i] = T(1); return r; }
    const T &operator()(int i, int j) const { return val[i * n + j]; }
    T &operator()(int i, int j) { return val[i * n + j]; }
    const T &operator()(int i) const { return val[i]; }
    T &operator()(int i) { return val[i]; }
    Matx<T, n, m> t() const { Matx<T, n, m> r; for (int i = 0; i < m; i++) for (int j = 0; j < n; j++) r.val[j * m + i] = val[i * n + j]; return r; }
    Matx mul(const Matx &b) const { Matx r; for (int i = 0; i < m * n; i++) r.val[i] = saturate_cast<T>(val[i] * b.val[i]); return r; }
    // Matx::inv — matx.hpp: for m == n > 3 (or any SVD request) cv::invert(*this, b, method); zeros when not ok
    Matx<T, n, m> inv(int method = DECOMP_LU, bool *p_is_ok = nullptr) const
    {
        static_assert(m == n && std::is_same<T, float>::value, "cvshim: only square float Matx::inv");
        CVSHIM_ASSERT(method == DECOMP_SVD);
        Matx<T, n, m> b;
        bool ok = orc::invert_svd_f(val, n, b.val);
        if (p_is_ok) *p_is_ok = ok;
        return ok ? b : Matx<T, n, m>::zeros();
    }
};

template <typename T, int cn> struct Vec : public Matx<T, cn, 1> {
    typedef Matx<T, cn, 1> Base;
    Vec() {}
    explicit Vec(const T *vals) : Base(vals) {}
    Vec(std::initializer_list<T> l) : Base(l) {}
    Vec(const Base &a) : Base(a) {}
    template <typename A0, typename A1, typename... A,
              typename = typename std::enable_if<(2 + sizeof...(A) <= cn) && std::is_arithmetic<A0>::value && std::is_arithmetic<A1>::value>::type>
    Vec(A0 a0, A1 a1, A... a) : Base(a0, a1, a...) {}
    const T &operator[](int i) const { return this->val[i]; }
    T &operator[](int i) { return this->val[i]; }
    const T &operator()(int i) const { return this->val[i]; }
    T &operator()(int i) { return this->val[i]; }
};

// element-wise ops (matx.hpp Matx_AddOp / Matx_SubOp / Matx_ScaleOp / Matx_MatMulOp)
template <typename T, int m, int n> static inline Matx<T, m, n> operator+(const Matx<T, m, n> &a, const Matx<T, m, n> &b) { Matx<T, m, n> r; for (int i = 0; i < m * n; i++) r.val[i] = saturate_cast<T>(a.val[i] + b.val[i]); return r; }
template <typename T, int m, int n> static inline Matx<T, m, n> operator-(const Matx<T, m, n> &a, const Matx<T, m, n> &b) { Matx<T, m, n> r; for (int i = 0; i < m * n; i++) r.val[i] = saturate_cast<T>(a.val[i] - b.val[i]); return r; }
template <typename T, int m, int n> static inline Matx<T, m, n> operator-(const Matx<T, m, n> &a) { Matx<T, m, n> r; for (int i = 0; i < m * n; i++) r.val[i] = saturate_cast<T>(-a.val[i]); return r; }
template <typename T, int m, int n, typename T2> static inline Matx<T, m, n> &operator+=(Matx<T, m, n> &a, const Matx<T2, m, n> &b) { for (int i = 0; i < m * n; i++) a.val[i] = saturate_cast<T>(a.val[i] + b.val[i]); return a; }
template <typename T, int m, int n, typename T2> static inline Matx<T, m, n> &operator-=(Matx<T, m, n> &a, const Matx<T2, m, n> &b) { for (int i = 0; i < m * n; i++) a.val[i] = saturate_cast<T>(a.val[i] - b.val[i]); return a; }
#define CVSHIM_SCALE(S)                                                                                                                  \
    template <typename T, int m, int n> static inline Matx<T, m, n> operator*(const Matx<T, m, n> &a, S alpha) { Matx<T, m, n> r; for (int i = 0; i < m * n; i++) r.val[i] = saturate_cast<T>(a.val[i] * alpha); return r; } \
    template <typename T, int m, int n> static inline Matx<T, m, n> operator*(S alpha, const Matx<T, m, n> &a) { Matx<T, m, n> r; for (int i = 0; i < m * n; i++) r.val[i] = saturate_cast<T>(a.val[i] * alpha); return r; } \
    template <typename T, int m, int n> static inline Matx<T, m, n> &operator*=(Matx<T, m, n> &a, S alpha) { for (int i = 0; i < m * n; i++) a.val[i] = saturate_cast<T>(a.val[i] * alpha); return a; } \
    template <typename T, int cn> static inline Vec<T, cn> operator*(const Vec<T, cn> &a, S alpha) { Vec<T, cn> r; for (int i = 0; i < cn; i++) r.val[i] = saturate_cast<T>(a.val[i] * alpha); return r; } \
    template <typename T, int cn> static inline Vec<T, cn> operator*(S alpha, const Vec<T, cn> &a) { Vec<T, cn> r; for (int i = 0; i < cn; i++) r.val[i] = saturate_cast<T>(a.val[i] * alpha); return r; }
CVSHIM_SCALE(int)
CVSHIM_SCALE(float)
CVSHIM_SCALE(double)
#undef CVSHIM_SCALE
// matx.hpp: Vec / int and Vec / double scale by the DOUBLE 1./alpha, Vec / float by the float 1.f/alpha
template <typename T, int cn> static inline Vec<T, cn> operator/(const Vec<T, cn> &a, int alpha) { double s = 1. / alpha; Vec<T, cn> r; for (int i = 0; i < cn; i++) r.val[i] = saturate_cast<T>(a.val[i] * s); return r; }
template <typename T, int cn> static inline Vec<T, cn> operator/(const Vec<T, cn> &a, float alpha) { float s = 1.f / alpha; Vec<T, cn> r; for (int i = 0; i < cn; i++) r.val[i] = saturate_cast<T>(a.val[i] * s); return r; }
template <typename T, int cn> static inline Vec<T, cn> operator/(const Vec<T, cn> &a, double alpha) { double s = 1. / alpha; Vec<T, cn> r; for (int i = 0; i < cn; i++) r.val[i] = saturate_cast<T>(a.val[i] * s); return r; }
template <typename T, int m, int n, int l> static inline Matx<T, m, n> operator*(const Matx<T, m, l> &a, const Matx<T, l, n> &b)
{
    Matx<T, m, n> r;
    for (int i = 0; i < m; i++)
        for (int j = 0; j < n; j++) {
            T s = 0;
            for (int k = 0; k < l; k++) s += a(i, k) * b(k, j);
            r.val[i * n + j] = s;
        }
    return r;
}
template <typename T, int m, int n> static inline Vec<T, m> operator*(const Matx<T, m, n> &a, const Vec<T, n> &b)
{
    Vec<T, m> r;
    for (int i = 0; i < m; i++) {
        T s = 0;
        for (int k = 0; k < n; k++) s += a(i, k) * b.val[k];
        r.val[i] = s;
    }
    return r;
}
template <typename T, int cn> static inline Vec<T, cn> operator+(const Vec<T, cn> &a, const Vec<T, cn> &b) { Vec<T, cn> r; for (int i = 0; i < cn; i++) r.val[i] = saturate_cast<T>(a.val[i] + b.val[i]); return r; }
template <typename T, int cn> static inline Vec<T, cn> operator-(const Vec<T, cn> &a, const Vec<T, cn> &b) { Vec<T, cn> r; for (int i = 0; i < cn; i++) r.val[i] = saturate_cast<T>(a.val[i] - b.val[i]); return r; }
template <typename T, int cn> static inline Vec<T, cn> operator-(const Vec<T, cn> &a) { Vec<T, cn> r; for (int i = 0; i < cn; i++) r.val[i] = saturate_cast<T>(-a.val[i]); return r; }
template <typename T, int cn> static inline Vec<T, cn> &operator+=(Vec<T, cn> &a, const Vec<T, cn> &b) { for (int i = 0; i < cn; i++) a.val[i] = saturate_cast<T>(a.val[i] + b.val[i]); return a; }
template <typename T, int cn> static inline Vec<T, cn> &operator-=(Vec<T, cn> &a, const Vec<T, cn> &b) { for (int i = 0; i < cn; i++) a.val[i] = saturate_cast<T>(a.val[i] - b.val[i]); return a; }
template <typename T, int cn> static inline std::ostream &operator<<(std::ostream &os, const Vec<T, cn> &v)
{
    os << "[";
    for (int i = 0; i < cn; i++) os << v.val[i] << (i + 1 < cn ? ", " : "");
    return os << "]";
}

typedef Matx<float, 1, 2> Matx12f;
typedef Matx<float, 1, 6> Matx16f;
typedef Matx<float, 2, 1> Matx21f;
typedef Matx<float, 3, 3> Matx33f;
typedef Matx<double, 3, 3> Matx33d;
typedef Matx<float, 6, 1> Matx61f;
typedef Matx<float, 6, 6> Matx66f;
typedef Vec<float, 2> Vec2f;
typedef Vec<float, 3> Vec3f;
typedef Vec<float, 4> Vec4f;
typedef Vec<float, 6> Vec6f;
typedef Vec<double, 3> Vec3d;

// ------------------------------------------------------------------------------------------- small types
template <typename T> struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T x_, T y_) : x(x_), y(y_) {}
};
typedef Point_<int> Point2i;
typedef Point2i Point;
typedef Point_<float> Point2f;
template <typename T> struct Point3_ {
    T x, y, z;
    Point3_() : x(0), y(0), z(0) {}
    Point3_(T x_, T y_, T z_) : x(x_), y(y_), z(z_) {}
    Point3_(const Vec<T, 3> &v) : x(v.val[0]), y(v.val[1]), z(v.val[2]) {}
};
typedef Point3_<float> Point3f;
template <typename T> struct Size_ {
    T width, height;
    Size_() : width(0), height(0) {}
    Size_(T w, T h) : width(w), height(h) {}
};
typedef Size_<int> Size;
template <typename T> struct Rect_ {
    T x, y, width, height;
    Rect_() : x(0), y(0), width(0), height(0) {}
    Rect_(T x_, T y_, T w, T h) : x(x_), y(y_), width(w), height(h) {}
};
typedef Rect_<int> Rect;
struct Range {
    int start, end;
    Range() : start(0), end(0) {}
    Range(int s, int e) : start(s), end(e) {}
    static Range all() { return Range(INT32_MIN, INT32_MAX); }
};
struct Scalar {
    double val[4];
    Scalar() { val[0] = val[1] = val[2] = val[3] = 0; }
    Scalar(double v0, double v1 = 0, double v2 = 0, double v3 = 0) { val[0] = v0; val[1] = v1; val[2] = v2; val[3] = v3; }
    static Scalar all(double v) { return Scalar(v, v, v, v); }
    double operator[](int i) const { return val[i]; }
};
struct KeyPoint {
    Point2f pt;
    float size = 0, angle = -1, response = 0;
    int octave = 0, class_id = -1;
    KeyPoint() {}
    KeyPoint(float x, float y, float size_, float angle_ = -1, float response_ = 0, int octave_ = 0, int class_id_ = -1)
        : pt(x, y), size(size_), angle(angle_), response(response_), octave(octave_), class_id(class_id_) {}
};
struct TermCriteria {
    enum Type { COUNT = 1, MAX_ITER = COUNT, EPS = 2 };
    int type = 0, maxCount = 0;
    double epsilon = 0;
    TermCriteria() {}
    TermCriteria(int t, int c, double e) : type(t), maxCount(c), epsilon(e) {}
};
class TickMeter {
public:
    void start() { t0 = clock::now(); running = true; }
    void stop() { if (running) { acc += std::chrono::duration<double, std::milli>(clock::now() - t0).count(); cnt++; running = false; } }
    void reset() { acc = 0; cnt = 0; running = false; }
    double getTimeMilli() const { return acc; }
    double getTimeSec() const { return acc * 1e-3; }
    int64_t getCounter() const { return cnt; }
private:
    typedef std::chrono::steady_clock clock;
    clock::time_point t0;
    double acc = 0;
    int64_t cnt = 0;
    bool running = false;
};

// ------------------------------------------------------------------------------------------- Mat
template <typename T> class Mat_;
template <typename T> class MatCommaInitializer_;

class Mat {
public:
    int rows = 0, cols = 0;
    uchar *data = nullptr;
    size_t step = 0;
    std::shared_ptr<void> aux;  // cvshim: buildOpticalFlowPyramid hangs the padded level set of the pyramid here

    Mat() {}
    Mat(int r, int c, int type) { create(r, c, type); }
    Mat(int r, int c, int type, void *d, size_t st = 0) : rows(r), cols(c), data((uchar *)d), type_(type)
    {
        step = st ? st : (size_t)c * elemSize();
    }
    template <typename T, int m, int n> explicit Mat(const Matx<T, m, n> &mx, bool copy = true) { from_array(mx.val, m, n, copy); }
    template <typename T, int n> explicit Mat(const Vec<T, n> &v, bool copy = true) { from_array(v.val, n, 1, copy); }

    void create(int r, int c, int type)
    {
        if (data && rows == r && cols == c && type_ == type) return;  // OpenCV keeps a matching buffer
        type_ = type; rows = r; cols = c;
        step = (size_t)c * elemSize();
        size_t bytes = step * (size_t)r;
        buf.reset(new uchar[bytes ? bytes : 1], std::default_delete<uchar[]>());
        data = buf.get();
        aux.reset();
    }
    void release() { buf.reset(); aux.reset(); data = nullptr; rows = cols = 0; step = 0; }
    int type() const { return type_; }
    int depth() const { return CV_MAT_DEPTH(type_); }
    int channels() const { return CV_MAT_CN(type_); }
    size_t elemSize() const
    {
        static const int sz[] = {1, 1, 2, 2, 4, 4, 8, 2};
        return (size_t)sz[depth()] * channels();
    }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    bool isContinuous() const { return step == (size_t)cols * elemSize() || rows == 1; }
    size_t total() const { return (size_t)rows * cols; }
    Size size() const { return Size(cols, rows); }

    uchar *ptr(int y = 0) { return data + step * (size_t)y; }
    const uchar *ptr(int y = 0) const { return data + step * (size_t)y; }
    template <typename T> T *ptr(int y = 0) { return (T *)(data + step * (size_t)y); }
    template <typename T> const T *ptr(int y = 0) const { return (const T *)(data + step * (size_t)y); }
    template <typename T> T *ptr(int y, int x) { return (T *)(data + step * (size_t)y) + x; }
    template <typename T> const T *ptr(int y, int x) const { return (const T *)(data + step * (size_t)y) + x; }
    template <typename T> T &at(int y, int x) { return ((T *)(data + step * (size_t)y))[x]; }
    template <typename T> const T &at(int y, int x) const { return ((const T *)(data + step * (size_t)y))[x]; }
    // mat.inl.hpp Mat::at(int i0): linear index for continuous / single-row, row index for single-column matrices
    template <typename T> T &at(int i0) { return *(T *)at_addr(i0, sizeof(T)); }
    template <typename T> const T &at(int i0) const { return *(const T *)const_cast<Mat *>(this)->at_addr(i0, sizeof(T)); }

    void copyTo(Mat &dst) const
    {
        if (dst.data == data && dst.rows == rows && dst.cols == cols) return;
        dst.create(rows, cols, type_);
        size_t rb = (size_t)cols * elemSize();
        for (int y = 0; y < rows; y++) std::memcpy(dst.data + dst.step * y, data + step * y, rb);
    }
    Mat clone() const { Mat m; copyTo(m); return m; }
    Mat operator()(Range rr, Range cr) const
    {
        Mat m(*this);
        if (rr.start != INT32_MIN) { CVSHIM_ASSERT(0 <= rr.start && rr.start <= rr.end && rr.end <= rows); m.rows = rr.end - rr.start; m.data += step * (size_t)rr.start; }
        if (cr.start != INT32_MIN) { CVSHIM_ASSERT(0 <= cr.start && cr.start <= cr.end && cr.end <= cols); m.cols = cr.end - cr.start; m.data += elemSize() * (size_t)cr.start; }
        m.aux.reset();
        return m;
    }
    Mat operator()(const Rect &r) const { return (*this)(Range(r.y, r.y + r.height), Range(r.x, r.x + r.width)); }
    void setTo(double v)
    {
        for (int y = 0; y < rows; y++)
            for (int x = 0; x < cols * channels(); x++) set_elem(y, x, v);
    }
    static Mat zeros(int r, int c, int type) { Mat m(r, c, type); m.setTo(0); return m; }
    static Mat ones(int r, int c, int type) { Mat m(r, c, type); m.setTo(1); return m; }
    static Mat eye(int r, int c, int type) { Mat m = zeros(r, c, type); for (int i = 0; i < std::min(r, c); i++) m.set_elem(i, i * m.channels(), 1); return m; }

    void set_elem(int y, int x, double v)
    {
        uchar *p = data + step * (size_t)y;
        switch (depth()) {
            case CV_8U: ((uint8_t *)p)[x] = (uint8_t)v; break;
            case CV_8S: ((int8_t *)p)[x] = (int8_t)v; break;
            case CV_16U: ((uint16_t *)p)[x] = (uint16_t)v; break;
            case CV_16S: ((int16_t *)p)[x] = (int16_t)v; break;
            case CV_32S: ((int32_t *)p)[x] = (int32_t)v; break;
            case CV_32F: ((float *)p)[x] = (float)v; break;
            case CV_64F: ((double *)p)[x] = v; break;
        }
    }

protected:
    int type_ = 0;
    std::shared_ptr<uchar> buf;
    template <typename T> static int depth_of()
    {
        return std::is_same<T, uint8_t>::value ? CV_8U : std::is_same<T, int16_t>::value ? CV_16S : std::is_same<T, int32_t>::value ? CV_32S
             : std::is_same<T, float>::value ? CV_32F : std::is_same<T, double>::value ? CV_64F : CV_8S;
    }
    template <typename T> void from_array(const T *v, int m, int n, bool copy)
    {
        if (copy) { create(m, n, depth_of<T>()); std::memcpy(data, v, sizeof(T) * m * n); }
        else { type_ = depth_of<T>(); rows = m; cols = n; step = sizeof(T) * n; data = (uchar *)v; }
    }
    uchar *at_addr(int i0, size_t esz)
    {
        if (isContinuous() || rows == 1) return data + esz * (size_t)i0;
        if (cols == 1) return data + step * (size_t)i0;
        int i = i0 / cols, j = i0 - i * cols;
        return data + step * (size_t)i + esz * (size_t)j;
    }
    template <typename T> friend class Mat_;
};

template <typename T> class Mat_ : public Mat {
public:
    Mat_() {}
    Mat_(int r, int c) : Mat(r, c, Mat::depth_of<T>()) {}
    Mat_(int r, int c, const T &value) : Mat(r, c, Mat::depth_of<T>()) { setTo((double)value); }
    T &operator()(int y, int x) { return this->template at<T>(y, x); }
};
// MatCommaInitializer_ (mat.hpp): `(Mat_<T>(r, c) << a, b, c ...)` fills row-major; converts to Mat_<T> (hence Mat)
template <typename T> class MatCommaInitializer_ {
public:
    explicit MatCommaInitializer_(const Mat_<T> &m) : mat(m), idx(0) {}
    template <typename T2> MatCommaInitializer_<T> &operator,(T2 v)
    {
        CVSHIM_ASSERT(idx < (int)mat.total());
        mat.template at<T>(idx / mat.cols, idx % mat.cols) = (T)v;
        idx++;
        return *this;
    }
    operator Mat_<T>() const { return mat; }
private:
    Mat_<T> mat;
    int idx;
};
template <typename T, typename T2> static inline MatCommaInitializer_<T> operator<<(const Mat_<T> &m, T2 val)
{
    MatCommaInitializer_<T> ci(m);
    return (ci, val);
}

// ------------------------------------------------------------------------------------------- core functions
static inline void setIdentity(Mat &m, const Scalar &s = Scalar(1))
{
    for (int y = 0; y < m.rows; y++)
        for (int x = 0; x < m.cols; x++) m.set_elem(y, x, y == x ? s.val[0] : 0.0);
}
static inline void absdiff(const Mat &a, const Mat &b, Mat &dst)
{
    CVSHIM_ASSERT(a.rows == b.rows && a.cols == b.cols && a.type() == b.type());
    Mat out(a.rows, a.cols, a.type());
    int n = a.cols * a.channels();
    for (int y = 0; y < a.rows; y++) {
        if (a.depth() == CV_32F) {
            const float *p = a.ptr<float>(y), *q = b.ptr<float>(y);
            float *o = out.ptr<float>(y);
            for (int x = 0; x < n; x++) o[x] = std::fabs(p[x] - q[x]);
        } else if (a.depth() == CV_8U) {
            const uint8_t *p = a.ptr<uint8_t>(y), *q = b.ptr<uint8_t>(y);
            uint8_t *o = out.ptr<uint8_t>(y);
            for (int x = 0; x < n; x++) o[x] = (uint8_t)std::abs((int)p[x] - (int)q[x]);
        } else
            CVSHIM_ASSERT(!"absdiff: depth");
    }
    dst = out;
}
template <typename T, int cn> static inline void absdiff(const Vec<T, cn> &a, const Vec<T, cn> &b, Vec<T, cn> &dst)
{
    for (int i = 0; i < cn; i++) dst.val[i] = std::fabs(a.val[i] - b.val[i]);
}
// cv::solve(A, b, x, DECOMP_SVD) on a float m x n system with one right-hand side (depth_filter.cpp:200)
template <int m, int n> static inline bool solve(const Mat &A, const Vec<float, m> &b, Vec<float, n> &x, int method)
{
    CVSHIM_ASSERT(method == DECOMP_SVD && A.type() == CV_32F && A.rows == m && A.cols == n && A.isContinuous());
    orc::solve_svd_f(A.ptr<float>(), m, n, b.val, 1, x.val);
    return true;
}
static inline bool solve(const Mat &A, const Mat &B, Mat &X, int method)
{
    CVSHIM_ASSERT(method == DECOMP_SVD && A.type() == CV_32F && B.type() == CV_32F && A.rows == B.rows && A.isContinuous() && B.isContinuous());
    Mat out(A.cols, B.cols, CV_32F);
    orc::solve_svd_f(A.ptr<float>(), A.rows, A.cols, B.ptr<float>(), B.cols, out.ptr<float>());
    out.copyTo(X);
    return true;
}

// calib3d: Rodrigues (float vector <-> float matrix; double inside), projectPoints
static inline void Rodrigues(const Vec3f &r, Matx33f &R) { orc::rodrigues_f(r.val, R.val); }
static inline void Rodrigues(const Matx33f &Rf, Vec3f &out)
{
    // matrix -> vector branch of cvRodrigues2 (without its SVD re-orthonormalisation of R; the only caller,
    // PoseManager::get_robot_angles, passes products of exact rotation matrices and is not on the tracking path)
    double R[9];
    for (int i = 0; i < 9; i++) R[i] = Rf.val[i];
    double rx = R[7] - R[5], ry = R[2] - R[6], rz = R[3] - R[1];
    double s = std::sqrt((rx * rx + ry * ry + rz * rz) * 0.25);
    double c = (R[0] + R[4] + R[8] - 1) * 0.5;
    c = c > 1. ? 1. : c < -1. ? -1. : c;
    double theta = std::acos(c);
    if (s < 1e-5) {
        if (c > 0) rx = ry = rz = 0;
        else {
            double t;
            t = (R[0] + 1) * 0.5; rx = std::sqrt(std::max(t, 0.));
            t = (R[4] + 1) * 0.5; ry = std::sqrt(std::max(t, 0.)) * (R[1] < 0 ? -1. : 1.);
            t = (R[8] + 1) * 0.5; rz = std::sqrt(std::max(t, 0.)) * (R[2] < 0 ? -1. : 1.);
            if (std::fabs(rx) < std::fabs(ry) && std::fabs(rx) < std::fabs(rz) && (R[5] > 0) != (ry * rz > 0)) rz = -rz;
            theta /= std::sqrt(rx * rx + ry * ry + rz * rz);
            rx *= theta; ry *= theta; rz *= theta;
        }
    } else {
        double vth = 1 / (2 * s);
        vth *= theta;
        rx *= vth; ry *= vth; rz *= vth;
    }
    out = Vec3f((float)rx, (float)ry, (float)rz);
}
static inline void projectPoints(const std::vector<Point3f> &pts, const Vec3f &rvec, const Vec3f &tvec, const Mat &K, const Mat &dist, Mat &out)
{
    CVSHIM_ASSERT(K.type() == CV_32F && K.rows == 3 && K.cols == 3 && dist.type() == CV_32F && dist.total() == 5);
    CVSHIM_ASSERT(tvec.val[0] == 0 && tvec.val[1] == 0 && tvec.val[2] == 0);
    float d5[5];
    for (int i = 0; i < 5; i++) d5[i] = dist.at<float>(i);
    out.create((int)pts.size(), 1, CV_32FC2);
    static_assert(sizeof(Point3f) == 12, "Point3f layout");
    orc::project_points(pts.empty() ? nullptr : &pts[0].x, (int)pts.size(), rvec.val, K.at<float>(0, 0), K.at<float>(1, 1), K.at<float>(0, 2),
                        K.at<float>(1, 2), d5, out.ptr<float>());
}

// imgproc: Sobel(src, dst, -1, 1, 0) (3x3, BORDER_DEFAULT), matchTemplate(TM_SQDIFF), minMaxLoc
static inline void Sobel(const Mat &src, Mat &dst, int ddepth, int dx, int dy, int ksize = 3, double scale = 1, double delta = 0, int border = BORDER_DEFAULT)
{
    CVSHIM_ASSERT(src.type() == CV_8U && (ddepth == -1 || ddepth == CV_8U) && dx == 1 && dy == 0 && ksize == 3 && scale == 1 && delta == 0 && border == BORDER_DEFAULT);
    Mat out(src.rows, src.cols, CV_8U);
    orc::sobel_x_u8(src.data, src.cols, src.rows, (int)src.step, out.data, (int)out.step);
    dst = out;
}
static inline void matchTemplate(const Mat &image, const Mat &templ, Mat &result, int method)
{
    CVSHIM_ASSERT(method == TM_SQDIFF && image.type() == CV_8U && templ.type() == CV_8U);
    if (!(image.rows >= templ.rows && image.cols >= templ.cols && templ.rows > 0 && templ.cols > 0))
        throw Exception("matchTemplate: the template does not fit into the image");
    std::vector<uint32_t> map;
    int mw, mh;
    orc::ssd_map_u32(image.data, image.cols, image.rows, (int)image.step, templ.data, templ.cols, templ.rows, (int)templ.step, map, mw, mh);
    result.create(mh, mw, CV_32F);
    for (int y = 0; y < mh; y++) {
        float *o = result.ptr<float>(y);
        for (int x = 0; x < mw; x++) o[x] = (float)map[(size_t)y * mw + x];
    }
}
static inline void minMaxLoc(const Mat &m, double *minVal, double *maxVal = nullptr, Point *minLoc = nullptr, Point *maxLoc = nullptr)
{
    CVSHIM_ASSERT(m.type() == CV_32F && !m.empty());
    float mn = m.at<float>(0, 0), mx = mn;
    Point pmn(0, 0), pmx(0, 0);
    for (int y = 0; y < m.rows; y++) {
        const float *p = m.ptr<float>(y);
        for (int x = 0; x < m.cols; x++) {
            if (p[x] < mn) { mn = p[x]; pmn = Point(x, y); }   // first minimum / maximum in raster order
            if (p[x] > mx) { mx = p[x]; pmx = Point(x, y); }
        }
    }
    if (minVal) *minVal = mn;
    if (maxVal) *maxVal = mx;
    if (minLoc) *minLoc = pmn;
    if (maxLoc) *maxLoc = pmx;
}

// features2d: FAST-9/16 with non-maximum suppression
class FastFeatureDetector {
public:
    enum { TYPE_5_8 = 0, TYPE_7_12 = 1, TYPE_9_16 = 2 };
    static Ptr<FastFeatureDetector> create(int threshold = 10, bool nonmaxSuppression = true, int type = TYPE_9_16)
    {
        CVSHIM_ASSERT(nonmaxSuppression && type == TYPE_9_16);
        Ptr<FastFeatureDetector> p(new FastFeatureDetector());
        p->threshold = threshold;
        return p;
    }
    void detect(const Mat &image, std::vector<KeyPoint> &keypoints, const Mat &mask = Mat())
    {
        CVSHIM_ASSERT(image.type() == CV_8U && mask.empty());
        std::vector<orc::FastKp> v;
        orc::fast9_16_nms(image.data, image.cols, image.rows, (int)image.step, threshold, v);
        keypoints.clear();
        for (const orc::FastKp &k : v) keypoints.push_back(KeyPoint((float)k.x, (float)k.y, 7.f, -1, (float)k.score));
    }
private:
    int threshold = 10;
};

// core/optim.hpp: only the interface the reference's callbacks derive from
class MinProblemSolver {
public:
    class Function {
    public:
        virtual ~Function() {}
        virtual int getDims() const = 0;
        virtual double getGradientEps() const { return 1e-3; }
        virtual double calc(const double *x) const = 0;
        virtual void getGradient(const double *, double *) {}
    };
};

// video: KalmanFilter (video/src/kalman.cpp), float.  predict()/correct() write into the existing matrices, like the
// MatExpr assignments of the original do for matrices of matching size — copies of a filter keep sharing their state.
class KalmanFilter {
public:
    Mat statePre, statePost, transitionMatrix, controlMatrix, measurementMatrix, processNoiseCov, measurementNoiseCov, errorCovPre, gain, errorCovPost;
    Mat temp1, temp2, temp3, temp4, temp5;
    KalmanFilter() {}
    KalmanFilter(int dynamParams, int measureParams, int controlParams = 0, int type = CV_32F) { init(dynamParams, measureParams, controlParams, type); }
    void init(int DP, int MP, int CP = 0, int type = CV_32F)
    {
        CVSHIM_ASSERT(DP > 0 && MP > 0 && type == CV_32F);
        CP = std::max(CP, 0);
        assign_zeros(statePre, DP, 1); assign_zeros(statePost, DP, 1);
        assign_zeros(transitionMatrix, DP, DP); setIdentity(transitionMatrix);
        assign_zeros(processNoiseCov, DP, DP); setIdentity(processNoiseCov);
        assign_zeros(measurementMatrix, MP, DP);
        assign_zeros(measurementNoiseCov, MP, MP); setIdentity(measurementNoiseCov);
        assign_zeros(errorCovPre, DP, DP); assign_zeros(errorCovPost, DP, DP);
        assign_zeros(gain, DP, MP);
        if (CP > 0) assign_zeros(controlMatrix, DP, CP); else controlMatrix.release();
        temp1.create(DP, DP, type); temp2.create(MP, DP, type); temp3.create(MP, MP, type); temp4.create(MP, DP, type); temp5.create(MP, 1, type);
    }
    const Mat &predict(const Mat &control = Mat())
    {
        CVSHIM_ASSERT(control.empty());
        const int n = transitionMatrix.rows;
        check(statePost, n, 1); check(errorCovPost, n, n);
        ensure(statePre, n, 1); ensure(errorCovPre, n, n); ensure(temp1, n, n);
        orc::gemm_f(f(transitionMatrix), f(statePost), nullptr, f(statePre), n, n, 1, false);                 // x'(k) = A x(k)
        orc::gemm_f(f(transitionMatrix), f(errorCovPost), nullptr, f(temp1), n, n, n, false);                 // temp1 = A P(k)
        orc::gemm_f(f(temp1), f(transitionMatrix), f(processNoiseCov), f(errorCovPre), n, n, n, true);        // P'(k) = temp1 At + Q
        statePre.copyTo(statePost);
        errorCovPre.copyTo(errorCovPost);
        return statePre;
    }
    const Mat &correct(const Mat &measurement)
    {
        const int n = transitionMatrix.rows, m = measurementMatrix.rows;
        CVSHIM_ASSERT(measurement.type() == CV_32F && (int)measurement.total() == m && measurement.isContinuous());
        ensure(temp2, m, n); ensure(temp3, m, m); ensure(temp4, m, n); ensure(temp5, m, 1); ensure(gain, n, m);
        ensure(statePost, n, 1); ensure(errorCovPost, n, n);
        orc::gemm_f(f(measurementMatrix), f(errorCovPre), nullptr, f(temp2), m, n, n, false);                 // temp2 = H P'(k)
        orc::gemm_f(f(temp2), f(measurementMatrix), f(measurementNoiseCov), f(temp3), m, n, m, true);         // temp3 = temp2 Ht + R
        orc::solve_svd_f(f(temp3), m, m, f(temp2), n, f(temp4));                                             // temp4 = inv(temp3) temp2 = Kt(k)
        for (int i = 0; i < n; i++) for (int j = 0; j < m; j++) f(gain)[i * m + j] = f(temp4)[j * n + i];     // K(k)
        orc::gemm_f(f(measurementMatrix), f(statePre), measurement.ptr<float>(), f(temp5), m, n, 1, false, -1.0);  // temp5 = z(k) - H x'(k)
        orc::gemm_f(f(gain), f(temp5), f(statePre), f(statePost), n, m, 1, false);                            // x(k) = x'(k) + K(k) temp5
        orc::gemm_f(f(gain), f(temp2), f(errorCovPre), f(errorCovPost), n, m, n, false, -1.0);                // P(k) = P'(k) - K(k) temp2
        return statePost;
    }
private:
    static float *f(Mat &m) { return m.ptr<float>(); }
    static void assign_zeros(Mat &m, int r, int c) { m.create(r, c, CV_32F); m.setTo(0); }  // `m = Mat::zeros(r, c, type)` (MatExpr assignment)
    static void ensure(Mat &m, int r, int c) { m.create(r, c, CV_32F); CVSHIM_ASSERT(m.isContinuous()); }
    static void check(const Mat &m, int r, int c) { CVSHIM_ASSERT(m.type() == CV_32F && m.rows == r && m.cols == c && m.isContinuous()); }
};

// video: buildOpticalFlowPyramid / calcOpticalFlowPyrLK.  The returned vector has OpenCV's layout
// [img0, deriv0, img1, deriv1, ...] (level images CV_8U, Scharr derivatives CV_16SC2); the padded level set the
// tracker reads (REFLECT_101 image frame, zero derivative frame — OpenCV keeps these as the border of each Mat's
// parent buffer) hangs on pyramid[0].aux and travels with every copy of the vector's Mats.
static inline int buildOpticalFlowPyramid(const Mat &img, std::vector<Mat> &pyramid, Size winSize, int maxLevel, bool withDerivatives = true,
                                          int pyrBorder = BORDER_REFLECT_101, int derivBorder = BORDER_CONSTANT, bool tryReuseInputImage = true)
{
    (void)tryReuseInputImage;
    CVSHIM_ASSERT(img.type() == CV_8U && withDerivatives && pyrBorder == BORDER_REFLECT_101 && derivBorder == BORDER_CONSTANT);
    CVSHIM_ASSERT(winSize.width > 2 && winSize.height > 2 && winSize.width + 3 <= orc::LKLevel::PAD && winSize.height + 3 <= orc::LKLevel::PAD);
    std::shared_ptr<orc::LKPyramid> p = std::make_shared<orc::LKPyramid>();
    // OpenCV stops early when a level would become smaller than the window; the tracking path never gets there
    orc::build_lk_pyramid(img.data, img.cols, img.rows, (int)img.step, maxLevel, *p);
    pyramid.clear();
    for (int l = 0; l <= maxLevel; l++) {
        orc::LKLevel &lv = p->lv[l];
        Mat im(lv.h, lv.w, CV_8U, lv.img.data());
        Mat de(lv.h, lv.w, CV_16SC2, lv.deriv.data());
        im.aux = p; de.aux = p;   // keeps the level storage alive for as long as any of the headers lives
        pyramid.push_back(im);
        pyramid.push_back(de);
    }
    return maxLevel;
}
static inline void calcOpticalFlowPyrLK(const std::vector<Mat> &prevPyr, const std::vector<Mat> &nextPyr, const std::vector<Point2f> &prevPts,
                                        std::vector<Point2f> &nextPts, std::vector<uchar> &status, std::vector<float> &err,
                                        Size winSize = Size(21, 21), int maxLevel = 3,
                                        TermCriteria criteria = TermCriteria(TermCriteria::COUNT + TermCriteria::EPS, 30, 0.01), int flags = 0,
                                        double minEigThreshold = 1e-4)
{
    CVSHIM_ASSERT(winSize.width == winSize.height && !prevPyr.empty() && !nextPyr.empty());
    std::shared_ptr<orc::LKPyramid> a = std::static_pointer_cast<orc::LKPyramid>(prevPyr[0].aux);
    std::shared_ptr<orc::LKPyramid> b = std::static_pointer_cast<orc::LKPyramid>(nextPyr[0].aux);
    CVSHIM_ASSERT(a && b);  // pyramids must come from buildOpticalFlowPyramid
    maxLevel = std::min(maxLevel, std::min((int)prevPyr.size() / 2 - 1, (int)nextPyr.size() / 2 - 1));
    const size_t n = prevPts.size();
    if (flags & OPTFLOW_USE_INITIAL_FLOW) CVSHIM_ASSERT(nextPts.size() == n);
    else nextPts = prevPts;
    status.assign(n, 1);
    err.assign(n, 0.f);
    int maxCount = (criteria.type & TermCriteria::COUNT) ? std::min(std::max(criteria.maxCount, 0), 100) : 30;
    double eps = (criteria.type & TermCriteria::EPS) ? std::min(std::max(criteria.epsilon, 0.), 10.) : 0.01;
    eps *= eps;
    static_assert(sizeof(Point2f) == 8, "Point2f layout");
    for (size_t i = 0; i < n; i++)
        orc::lk_track_point(*a, *b, winSize.width, maxLevel, maxCount, eps, minEigThreshold, &prevPts[i].x, &nextPts[i].x, &status[i], &err[i]);
}

}  // namespace cv
