// ORACLE — TEST INFRASTRUCTURE ONLY (see ocv_prims.hpp header).  Not part of the product path.
//
// Single-threaded CPU restatement of the reference's tracking hot path
// (eichenberger/stereo-svo-slam, src/lib/*.cpp).  Every function cites the reference file:line it
// follows.  The reference itself cannot be compiled in this environment (it needs the OpenCV C++
// SDK, which is absent — SURVEY.md §8c), so this restatement + cv2-pinned primitives (ocv_prims.hpp)
// is the parity oracle AND the CPU baseline ("port") timed by bench.py.
//
// Pinning status: the reference's own tests hold no golden vectors for this path except the
// exponential_map known-answer (src/test/test_exponential_map.cpp:35-48), checked in
// tests/test_oracle_vs_cv2.py together with the cv2 pins.  Reference-owned loops (alignment,
// refinement, depth filter, grid selection) have no reference-side vectors: "parity unpinned" for
// those, by the reference's own lack of tests; they are restated line by line below.
//
// Build: g++ -O2 -ffp-contract=off -shared -fPIC (oracle/Makefile).
#include "ocv_prims.hpp"
#include <map>
#include <memory>
#include <string>
#include <cstdio>
#include <array>
#include <limits>
#include <chrono>

using namespace orc;

extern "C" {
// Same field order as the reference's CameraSettings (src/include/stereo_slam_types.hpp:16-36)
struct OrcCameraSettings {
    float baseline, fx, fy, cx, cy, k1, k2, k3, p1, p2;
    int grid_height, grid_width, search_x, search_y;
    int window_size_pose_estimator, window_size_opt_flow, window_size_depth_calculator;
    int max_pyramid_levels, min_pyramid_level_pose_estimation;
};
}
typedef OrcCameraSettings Settings;

// ----------------------------------------------------------------------------- small float algebra
// cv::Matx semantics: every product is `s = 0; for k: s += a(i,k)*b(k,j)` in float.
static inline void m33v(const float M[9], const float v[3], float o[3])
{
    float t[3];
    for (int i = 0; i < 3; i++) {
        float s = 0;
        for (int k = 0; k < 3; k++) s += M[i * 3 + k] * v[k];
        t[i] = s;
    }
    o[0] = t[0]; o[1] = t[1]; o[2] = t[2];
}

// pose_manager.cpp:9-17 — PoseManager::set_pose: cached Rodrigues(angles) and Rodrigues(-angles)
struct PoseM {
    float p[6] = {0, 0, 0, 0, 0, 0};  // x y z rx ry rz
    float R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, Ri[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    void set(const float *v)
    {
        for (int i = 0; i < 6; i++) p[i] = v[i];
        float a[3] = {p[3], p[4], p[5]}, na[3] = {-p[3], -p[4], -p[5]};
        rodrigues_f(a, R);
        rodrigues_f(na, Ri);
    }
};

// exponential_map.hpp:12-37 (norm fixed to 1; w unchanged).  `cos(_norm)` on a float: with <opencv2/opencv.hpp> included
// (it pulls in <math.h>, i.e. libstdc++'s global float overloads) this is the float cosine, so both scale factors are float
// and `scalar * Matx33f` multiplies in float — pinned against the reference itself (oracle/_ref, tests/test_ref_pin.py).
static void exponential_map(const float twist[6], float out[6])
{
    const float *v = twist, *w = twist + 3;
    float K[9] = {0, -w[2], w[1], w[2], 0, -w[0], -w[1], w[0], 0};
    float K2[9];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            float s = 0;
            for (int k = 0; k < 3; k++) s += K[i * 3 + k] * K[k * 3 + j];
            K2[i * 3 + j] = s;
        }
    volatile float norm = 1.0f;
    float c1 = 1 - cosf(norm), c2 = norm - sinf(norm);
    float M[9];
    for (int i = 0; i < 9; i++) {
        float e = ((i % 4 == 0) ? 1.f : 0.f) * norm;
        float a = K[i] * c1;
        float b = K2[i] * c2;
        M[i] = (e + a) + b;
    }
    float t[3];
    m33v(M, v, t);
    out[0] = t[0]; out[1] = t[1]; out[2] = t[2];
    out[3] = w[0]; out[4] = w[1]; out[5] = w[2];
}

// transform_keypoints.cpp:11-47
static void project_keypoints(const PoseM &pose, const float *kps3d, int n, const Settings &cs, float *out2)
{
    std::vector<float> in((size_t)n * 3);
    for (int i = 0; i < n; i++)
        for (int k = 0; k < 3; k++) in[i * 3 + k] = kps3d[i * 3 + k] - pose.p[k];
    float rvec[3] = {-pose.p[3], -pose.p[4], -pose.p[5]};
    float dist[5] = {cs.k1, cs.k2, cs.p1, cs.p2, cs.k3};
    project_points(in.data(), n, rvec, cs.fx, cs.fy, cs.cx, cs.cy, dist, out2);
}

// ----------------------------------------------------------------------------- images
struct Img {
    int w = 0, h = 0;
    std::vector<uint8_t> d;
    const uint8_t *row(int y) const { return d.data() + (size_t)y * w; }
};

// stereo_slam.cpp:93-109
static void half_sample(const uint8_t *in, int w, int h, int stride, uint8_t *out)
{
    int ow = w / 2, oh = h / 2;
    for (int j = 0; j < oh; j++) {
        const uint8_t *u = in + (size_t)(2 * j) * stride, *l = in + (size_t)(2 * j + 1) * stride;
        for (int i = 0, x = 0; i < ow; i++, x += 2) out[(size_t)j * ow + i] = (uint8_t)((u[x] + u[x + 1] + l[x] + l[x + 1]) / 4);
    }
}
// stereo_slam.cpp:112-121
static void create_img_pyramid(const uint8_t *img, int w, int h, int stride, int n_levels, std::vector<Img> &pyr)
{
    pyr.resize(n_levels);
    pyr[0].w = w; pyr[0].h = h; pyr[0].d.resize((size_t)w * h);
    for (int y = 0; y < h; y++) memcpy(&pyr[0].d[(size_t)y * w], img + (size_t)y * stride, w);
    for (int i = 1; i < n_levels; i++) {
        pyr[i].w = pyr[i - 1].w / 2; pyr[i].h = pyr[i - 1].h / 2;
        pyr[i].d.resize((size_t)pyr[i].w * pyr[i].h);
        half_sample(pyr[i - 1].d.data(), pyr[i - 1].w, pyr[i - 1].h, pyr[i - 1].w, pyr[i].d.data());
    }
}

struct StereoImage {
    std::vector<Img> left, right;
    LKPyramid opt_flow;
};

// ----------------------------------------------------------------------------- keypoints
enum { KP_FAST = 0, KP_EDGELET = 1 };
struct KF1 { Kalman k; };
struct KpInfo {
    float score = 0;
    int level = 0, type = 0;
    uint64_t keyframe_id = 0;
    size_t keypoint_index = 0;
    bool ignore_during_refinement = false, ignore_completely = false, ignore_temporary = false;
    int outlier_count = 0, inlier_count = 0;
    std::shared_ptr<KF1> kf;  // cv::KalmanFilter copies share their Mats
};
struct KeyPoints {
    std::vector<float> kps2d, kps3d;  // n*2, n*3
    std::vector<KpInfo> info;
    size_t size() const { return info.size(); }
};
struct Frame {
    uint64_t id = 0;
    PoseM pose;
    std::shared_ptr<StereoImage> img;
    KeyPoints kps;
    double time_stamp = 0;
};

// ----------------------------------------------------------------------------- corner detector
// corner_detector.cpp:13-79
static void detect_keypoints(const Img &image, int grid_width, int grid_height, std::vector<float> &kps,
                             std::vector<KpInfo> &infos, int level)
{
    std::vector<FastKp> fast;
    fast9_16_nms(image.d.data(), image.w, image.h, image.w, 6, fast);
    std::vector<uint8_t> edge((size_t)image.w * image.h);
    sobel_x_u8(image.d.data(), image.w, image.h, image.w, edge.data(), image.w);
    int top = 0, bottom = grid_height;
    while (true) {
        int left = 0, right = grid_width;
        while (true) {
            if (right > image.w) break;
            float cx = 0, cy = 0;
            KpInfo info;
            info.score = -1;
            for (const FastKp &k : fast) {
                if (k.x < left || k.x >= right || k.y < top || k.y >= bottom) continue;
                if (info.score < (float)k.score) {
                    info.score = (float)k.score; cx = (float)k.x; cy = (float)k.y; info.type = KP_FAST;
                }
            }
            if (info.score < 0) {
                for (int k = left; k < right; k++)
                    for (int l = top; l < bottom; l++) {
                        uint8_t r = (l < image.h) ? edge[(size_t)l * image.w + k] : 0;
                        if (info.score < r) { info.score = r; cx = (float)k; cy = (float)l; info.type = KP_EDGELET; }
                    }
            }
            info.level = level;
            kps.push_back(cx); kps.push_back(cy);
            infos.push_back(info);
            left += grid_width; right += grid_width;
        }
        bottom += grid_height;
        if (bottom > image.h) break;
        top += grid_height;
    }
}

// ----------------------------------------------------------------------------- image comparison
// image_comparison.cpp:9-91
static float get_intensity_diff(const Img &im1, const Img &im2, const float c1[2], const float c2[2], int patch)
{
    float half = ((float)patch - 1.0f) / 2.0f;
    float s1x = c1[0] - half, s1y = c1[1] - half, s2x = c2[0] - half, s2y = c2[1] - half;
    int ip1x = (int)floor(s1x), ip1y = (int)floor(s1y), ip2x = (int)floor(s2x), ip2y = (int)floor(s2y);
    float x12 = s1x - ip1x, y12 = s1y - ip1y, x22 = s2x - ip2x, y22 = s2y - ip2y;
    float x11 = (float)(1.0 - x12), y11 = (float)(1.0 - y12), x21 = (float)(1.0 - x22), y21 = (float)(1.0 - y22);
    float m1[4] = {x11 * y11, x12 * y11, x11 * y12, x12 * y12};
    float m2[4] = {x21 * y21, x22 * y21, x21 * y22, x22 * y22};
    float intensity = 0;
    if (ip1y >= 0 && ip1y + patch < im1.h && ip2y >= 0 && ip2y + patch < im2.h && ip1x >= 0 && ip1x + patch < im1.w &&
        ip2x >= 0 && ip2x + patch < im2.w) {
        for (int i = 0; i < patch; i++) {
            const uint8_t *s11 = im1.row(i + ip1y) + ip1x, *s12 = im1.row(i + ip1y + 1) + ip1x;
            const uint8_t *s21 = im2.row(i + ip2y) + ip2x, *s22 = im2.row(i + ip2y + 1) + ip2x;
            for (int j = 0; j < patch; j++) {
                float p1[4] = {(float)s11[0], (float)s11[1], (float)s12[0], (float)s12[1]};
                float p2[4] = {(float)s21[0], (float)s21[1], (float)s22[0], (float)s22[1]};
                float i1 = 0, i2 = 0;
                for (int k = 0; k < 4; k++) i1 += m1[k] * p1[k];
                for (int k = 0; k < 4; k++) i2 += m2[k] * p2[k];
                intensity += fabsf(i1 - i2);
                s11++; s12++; s21++; s22++;
            }
        }
    }
    return intensity;
}
// image_comparison.cpp:103-120
static float get_total_intensity_diff(const Img &im1, const Img &im2, const float *k1, const float *k2, int n, int patch)
{
    float diff = 0;
    for (int i = 0; i < n; i++) diff += get_intensity_diff(im1, im2, k1 + 2 * i, k2 + 2 * i, patch);
    return diff;
}

// ----------------------------------------------------------------------------- pose estimator
// pose_estimator.cpp:82-112
static float get_patch_sum(const Img &image, float cx, float cy)
{
    float sx = cx - 0.5f, sy = cy - 0.5f;
    int ipx = (int)floor(sx), ipy = (int)floor(sy);
    float x2 = sx - ipx, y2 = sy - ipy;
    float x1 = (float)(1.0 - x2), y1 = (float)(1.0 - y2);
    const uint8_t *s1 = image.row(ipy) + ipx, *s2 = image.row(ipy + 1) + ipx, *s3 = image.row(ipy + 2) + ipx;
    float intensity = x1 * y1 * s1[0] + y1 * s1[1] + x2 * y1 * s1[2] + x1 * s2[0] + s2[1] + x2 * s2[2] +
                      x1 * y2 * s3[0] + y2 * s3[1] + x2 * y2 * s3[2];
    return intensity;
}

struct AlignStats { int evals[8] = {0}; int grads[8] = {0}; float level_cost[8] = {0}; };

// PoseEstimatorCallback (pose_estimator.cpp:36-68, :226-245, :275-300, :312-539, :541-562)
struct PoseEstimatorCb {
    const std::vector<Img> &cur, &prev;
    const Settings &cs;
    Settings lcs;
    std::vector<float> k2d, k3d, lk2d;  // local (filtered) keypoints, level keypoints
    int level = 0, n = 0;
    std::vector<float> gtj;  // n*16*6
    float inv_hessian[36];
    static const int PATCH = 4;

    PoseEstimatorCb(const std::vector<Img> &cur_, const std::vector<Img> &prev_, const KeyPoints &pk, const Settings &cs_)
        : cur(cur_), prev(prev_), cs(cs_)
    {
        for (size_t i = 0; i < pk.size(); i++) {
            if (pk.info[i].ignore_temporary) continue;
            k2d.push_back(pk.kps2d[2 * i]); k2d.push_back(pk.kps2d[2 * i + 1]);
            for (int k = 0; k < 3; k++) k3d.push_back(pk.kps3d[3 * i + k]);
        }
        n = (int)k2d.size() / 2;
    }
    void set_level(int lv)
    {
        level = lv;
        int divider = 1 << lv;
        lcs = cs;
        lcs.fx /= divider; lcs.fy /= divider; lcs.cx /= divider; lcs.cy /= divider; lcs.baseline /= divider;
        lk2d = k2d;
        if (lv == 0) return;
        for (auto &v : lk2d) v /= divider;
    }
    float do_calc(const PoseM &pm) const
    {
        std::vector<float> kp((size_t)n * 2);
        project_keypoints(pm, k3d.data(), n, lcs, kp.data());
        return get_total_intensity_diff(prev[level], cur[level], lk2d.data(), kp.data(), n, lcs.window_size_pose_estimator);
    }
    void calculate_hessian(const PoseM &pose)
    {
        const Img &pi = prev[level];
        gtj.assign((size_t)n * 16 * 6, 0.f);
        for (int i = 0; i < n; i++) {
            const float fx = lcs.fx, fy = lcs.fy;
            float kx = lk2d[2 * i], ky = lk2d[2 * i + 1];
            kx -= PATCH / 2; ky -= PATCH / 2;
            float kp[3] = {k3d[3 * i] - pose.p[0], k3d[3 * i + 1] - pose.p[1], k3d[3 * i + 2] - pose.p[2]};
            m33v(pose.Ri, kp, kp);
            float x = kp[0], y = kp[1], z = kp[2];
            float J[12] = {-fx / z, 0, fx * x / (z * z), fx * x * y / (z * z), -fx * (1 + (x * x) / (z * z)), fx * y / z,
                           0, -fy / z, fy * y / (z * z), fy * (1 + (y * y) / (z * z)), -fy * x * y / (z * z), -fy * x / z};
            float *it = &gtj[(size_t)i * 16 * 6];
            for (int r = 0; r < PATCH; r++) {
                for (int c = 0; c < PATCH; c++, it += 6) {
                    if ((kx - 2.0) < 0 || (ky - 2.0) < 0 || (kx + 3.0) >= pi.w || (ky + 3.0) >= pi.h) {
                        for (int k = 0; k < 6; k++) it[k] = 0;
                        kx++;
                        continue;
                    }
                    float i1 = get_patch_sum(pi, kx + 1, ky), i2 = get_patch_sum(pi, kx - 1, ky);
                    float i3 = get_patch_sum(pi, kx, ky + 1), i4 = get_patch_sum(pi, kx, ky - 1);
                    float g0 = i1 - i2, g1 = i3 - i4;
                    for (int k = 0; k < 6; k++) {
                        float s = 0;
                        s += g0 * J[k];
                        s += g1 * J[6 + k];
                        it[k] = s;
                    }
                    kx++;
                }
                kx -= PATCH;
                ky++;
            }
        }
        float H[36];
        for (int k = 0; k < 36; k++) H[k] = 0;
        for (size_t e = 0; e < (size_t)n * 16; e++) {
            const float *g = &gtj[e * 6];
            for (int a = 0; a < 6; a++)
                for (int b = 0; b < 6; b++) H[a * 6 + b] = H[a * 6 + b] + g[a] * g[b];
        }
        invert_svd_f(H, 6, inv_hessian);
    }
    void get_gradient(const PoseM &pose, float grad[6])
    {
        const Img &ci = cur[level], &pi = prev[level];
        calculate_hessian(pose);  // `hessian` is never assigned, so it is rebuilt on every call
        std::vector<float> kp((size_t)n * 2);
        project_keypoints(pose, k3d.data(), n, lcs, kp.data());
        std::vector<float> diffs((size_t)n * 16);
        float *diff = diffs.data();
        const int hp = PATCH / 2;
        for (int i = 0; i < n; i++) {
            float kx = kp[2 * i] - hp, ky = kp[2 * i + 1] - hp;
            float rx = lk2d[2 * i] - hp, ry = lk2d[2 * i + 1] - hp;
            for (int r = 0; r < PATCH; r++) {
                for (int c = 0; c < PATCH; c++, kx++, rx++, diff++) {
                    if (!((rx - 1.0) < 0 || (kx - 1.0) < 0 || (ry - 1.0) < 0 || (ky - 1.0) < 0 || (rx + 2.0) > pi.w ||
                          (kx + 2.0) > ci.w || (ry + 2.0) > pi.h || (ky + 2.0) > ci.h)) {
                        float i1 = get_patch_sum(pi, rx, ry);
                        float i2 = get_patch_sum(ci, kx, ky);
                        *diff = i2 - i1;
                    } else
                        *diff = 0;
                }
                ky++; ry++;
                kx -= PATCH; rx -= PATCH;
            }
        }
        float res[6] = {0, 0, 0, 0, 0, 0};
        for (size_t e = 0; e < diffs.size(); e++) {
            const float *g = &gtj[e * 6];
            for (int k = 0; k < 6; k++) res[k] -= g[k] * diffs[e];
        }
        float delta[6];
        for (int i = 0; i < 6; i++) {
            float s = 0;
            for (int k = 0; k < 6; k++) s += inv_hessian[i * 6 + k] * res[k];
            delta[i] = s;
        }
        float pg[6];
        exponential_map(delta, pg);
        float t[3], w[3];
        m33v(pose.R, pg, t);
        m33v(pose.R, pg + 3, w);
        grad[0] = t[0]; grad[1] = t[1]; grad[2] = t[2];
        grad[3] = w[0]; grad[4] = w[1]; grad[5] = w[2];
    }
};

// get_patch_sum reads image rows ipy..ipy+2; the reference guards with the bounds tests above, which admit
// (by construction) reads one row/col past the tested range in rare edge cases — mirror by padding reads:
// Img buffers in this oracle are allocated with slack (see alloc_slack) so such reads stay in-bounds.

// pose_estimator.cpp:166-222 (shared loop counter) ; also pose_refinement.cpp:236-290
template <class CB>
static float gn_driver(CB &cb, const PoseM &guess, PoseM &out, float stop_thr, int *n_evals, int *n_grads)
{
    const int maxIter = 50;
    PoseM x0 = guess, _x;
    float prev_cost = cb.do_calc(x0);
    int evals = 1, grads = 0;
    int i;
    static const bool gn_trace = getenv("SVO_ORACLE_GN_TRACE") != nullptr;   // developer aid: the accept (a) / halve (r) / stop (s) pattern
    for (i = 0; i < maxIter; i++) {
        float gradient[6];
        cb.get_gradient(x0, gradient);
        grads++;
        if (gn_trace) fputc('G', stderr);
        float k = 1.0;
        for (; i < maxIter; i++) {
            float x[6];
            for (int j = 0; j < 6; j++) x[j] = x0.p[j] + (k * gradient[j]);
            _x.set(x);
            float new_cost = cb.do_calc(_x);
            evals++;
            if (new_cost < prev_cost) {
                x0 = _x;
                prev_cost = new_cost;
                if (gn_trace) fputc('a', stderr);
                break;
            } else if (fabs(new_cost - prev_cost) < stop_thr) {
                i = maxIter;
                if (gn_trace) fputc('s', stderr);
                break;
            } else {
                k /= 2;
                if (gn_trace) fputc('r', stderr);
            }
        }
    }
    if (gn_trace) fprintf(stderr, " thr %g\n", (double)stop_thr);
    out = x0;
    if (n_evals) *n_evals = evals;
    if (n_grads) *n_grads = grads;
    return prev_cost;
}

// pose_estimator.cpp:115-130
static float estimate_pose(const std::vector<Img> &cur, const std::vector<Img> &prev, const KeyPoints &pk, const Settings &cs,
                           const PoseM &guess, PoseM &pose, AlignStats *st)
{
    PoseEstimatorCb cb(cur, prev, pk, cs);
    float err = 0;
    PoseM est = guess;
    for (int i = cs.max_pyramid_levels; i > cs.min_pyramid_level_pose_estimation; i--) {
        PoseM ne;
        int lv = i - 1;
        cb.set_level(lv);
        int ev = 0, gr = 0;
        err = gn_driver(cb, est, ne, 1.0f, &ev, &gr);
        if (st && lv < 8) { st->evals[lv] = ev; st->grads[lv] = gr; st->level_cost[lv] = err; }
        est = ne;
    }
    pose = est;
    return err;
}

// ----------------------------------------------------------------------------- pose refiner
// PoseRefinerCallback (pose_refinement.cpp:321-412)
struct PoseRefinerCb {
    const float *k2d, *k3d;
    const std::vector<KpInfo> &info;
    const Settings &cs;
    int n;
    PoseRefinerCb(const float *a, const float *b, const std::vector<KpInfo> &i, const Settings &c) : k2d(a), k3d(b), info(i), cs(c), n((int)i.size()) {}
    float do_calc(const PoseM &pose) const
    {
        std::vector<float> pr((size_t)n * 2);
        project_keypoints(pose, k3d, n, cs, pr.data());
        float tot = 0;
        for (int i = 0; i < n; i++) {
            if (!info[i].ignore_during_refinement && !info[i].ignore_completely && !info[i].ignore_temporary) {
                float d0 = fabsf(pr[2 * i] - k2d[2 * i]), d1 = fabsf(pr[2 * i + 1] - k2d[2 * i + 1]);
                tot += d0 + d1;
            }
        }
        return tot;
    }
    void get_gradient(const PoseM &xp, float grad[6])
    {
        std::vector<float> pr((size_t)n * 2);
        project_keypoints(xp, k3d, n, cs, pr.data());
        float err[6] = {0}, H[36] = {0};
        for (int i = 0; i < n; i++) {
            const float fx = cs.fx, fy = cs.fy;
            float kp[3] = {k3d[3 * i] - xp.p[0], k3d[3 * i + 1] - xp.p[1], k3d[3 * i + 2] - xp.p[2]};
            m33v(xp.Ri, kp, kp);
            float x = kp[0], y = kp[1], z = kp[2];
            if (info[i].ignore_during_refinement || info[i].ignore_completely || info[i].ignore_temporary) continue;
            float J[12] = {-fx / z, 0, fx * x / (z * z), fx * x * y / (z * z), -fx * (1 + (x * x) / (z * z)), fx * y / z,
                           0, -fy / z, fy * y / (z * z), fy * (1 + (y * y) / (z * z)), -fy * x * y / (z * z), -fy * x / z};
            float d[2] = {k2d[2 * i] - pr[2 * i], k2d[2 * i + 1] - pr[2 * i + 1]};
            if ((fabs(d[0]) > 3.0) || (fabs(d[1]) > 3.0)) continue;
            for (int a = 0; a < 6; a++)
                for (int b = 0; b < 6; b++) {
                    float s = 0;
                    s += J[a] * J[b];
                    s += J[6 + a] * J[6 + b];
                    H[a * 6 + b] = H[a * 6 + b] + s;
                }
            for (int a = 0; a < 6; a++) {
                float s = 0;
                s += J[a] * d[0];
                s += J[6 + a] * d[1];
                err[a] = err[a] + s;
            }
        }
        float Hi[36];
        invert_svd_f(H, 6, Hi);
        float tw[6];
        for (int i = 0; i < 6; i++) {
            float s = 0;
            for (int k = 0; k < 6; k++) s += Hi[i * 6 + k] * err[k];
            tw[i] = s;
        }
        exponential_map(tw, grad);
    }
};

// ----------------------------------------------------------------------------- stereo SSD rule
// depth_calculator.cpp:201-239 (mode 0) and depth_filter.cpp:275-326 (mode 1)
static float stereo_disparity(const Img &left, const Img &right, float kx, float ky, const Settings &cs, int mode)
{
    const int window_size = cs.window_size_depth_calculator;
    const int wb = window_size / 2, wa = (window_size + 1) / 2;
    const int sx = cs.search_x, sy = cs.search_y;
    int x = (int)kx, y = (int)ky;
    int x11 = std::max(0, x - wb), x12 = std::min(left.w - 1, x + wa);
    int y11 = std::max(0, y - wb), y12 = std::min(left.h, y + wa);
    if (mode == 1 && (x12 <= 0 || y12 <= 0 || x11 >= (left.w - 1) || y11 >= (left.h - 1))) return -1;
    int x21 = std::max(0, x - wb), x22 = std::min(left.w - 1, x + wa + sx);
    int y21 = std::max(0, y - wb - sy), y22 = std::min(left.h - 1, y + wa + sy);
    if (mode == 1 && (x22 <= 0 || y22 <= 0 || x21 >= (left.w - 1) || y21 >= (left.h - 1))) return -1;
    int tw = x12 - x11, th = y12 - y11, rw = x22 - x21, rh = y22 - y21;
    if (tw <= 0 || th <= 0 || rw < tw || rh < th) return -1;  // cv::matchTemplate would throw here
    std::vector<uint32_t> map;
    int mw, mh;
    ssd_map_u32(right.row(y21) + x21, rw, rh, right.w, left.row(y11) + x11, tw, th, left.w, map, mw, mh);
    uint32_t minv = map[0];
    int mlx = 0, mly = 0;
    for (int k = 0; k < mh; k++)
        for (int j = 0; j < mw; j++)
            if (map[(size_t)k * mw + j] < minv) { minv = map[(size_t)k * mw + j]; mlx = j; mly = k; }
    float minPos = 0;
    int matches = 0;
    for (int j = mlx; j < mw; j++)
        for (int k = mly; k < mh; k++)
            if (map[(size_t)k * mw + j] <= minv) { minPos += j; matches++; }
    minPos = minPos / matches;
    if (mode == 1) return std::max<float>(0.5, minPos);
    return minPos;
}

// ----------------------------------------------------------------------------- the SLAM pipeline
struct KeyFrame : Frame {};

struct OracleSlam {
    Settings cs;
    int W, H;
    std::vector<std::unique_ptr<KeyFrame>> keyframes;  // KeyFrameManager (per-instance counters)
    KeyFrame *keyframe = nullptr;
    std::shared_ptr<Frame> frame;
    std::vector<std::array<float, 6>> trajectory;
    float motion[6] = {0};
    Kalman kf;
    uint64_t keyframe_count = 0;  // depth_calculator.cpp:135 static
    std::map<std::string, std::vector<float>> trace;
    bool tracing = true;
    double stage_ms[8] = {0};  // 0 pyramids 1 alignment 2 klt 3 refine-gn 4 ssd 5 filter 6 keyframe 7 total
    static double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

    OracleSlam(const Settings &s, int w, int h) : cs(s), W(w), H(h)
    {
        // stereo_slam.cpp:29-41
        kf.init(12, 12);
        for (int i = 0; i < 12; i++) { kf.H[i * 12 + i] = 1; kf.Q[i * 12 + i] = 100.0f; kf.Ppost[i * 12 + i] = 1.0f; }
    }
    void tr(const char *name, const float *d, size_t n) { if (tracing) trace[name].assign(d, d + n); }
    void tri(const char *name, const std::vector<int> &v) { if (tracing) trace[name].assign(v.begin(), v.end()); }

    // ---- DepthCalculator (depth_calculator.cpp) ----
    void detect_keypoints_on_each_level(const StereoImage &si, std::vector<std::vector<float>> &kp_pyr,
                                        std::vector<std::vector<KpInfo>> &info_pyr)
    {
        int gw = cs.grid_width, gh = cs.grid_height;
        size_t nl = si.left.size() / 2;
        kp_pyr.resize(nl); info_pyr.resize(nl);
        for (size_t i = 0; i < nl; i++) {
            detect_keypoints(si.left[i], gw, gh, kp_pyr[i], info_pyr[i], (int)i);
            gw /= 2; gh /= 2;
        }
    }
    // depth_calculator.cpp:37-65 (index j reused on every level)
    static void select_best_keypoints(const std::vector<std::vector<float>> &kp_pyr, const std::vector<std::vector<KpInfo>> &info_pyr,
                                      std::vector<float> &kps, std::vector<KpInfo> &info)
    {
        kps = kp_pyr[0]; info = info_pyr[0];
        for (size_t i = 1; i < kp_pyr.size(); i++) {
            const auto &_k = kp_pyr[i];
            const auto &_i = info_pyr[i];
            for (size_t j = 0; j < info.size(); j++) {
                if (j >= _i.size()) continue;  // the reference would read out of bounds here
                if (info[j].type == KP_FAST && _i[j].type == KP_EDGELET) continue;
                if ((info[j].type == _i[j].type) && (info[j].score > _i[j].score)) continue;
                kps[2 * j] = _k[2 * j] * (1 << i);
                kps[2 * j + 1] = _k[2 * j + 1] * (1 << i);
                info[j] = _i[j];
            }
        }
    }
    // depth_calculator.cpp:67-86
    void find_bad_keypoints(Frame &f)
    {
        KeyPoints o;
        for (size_t i = 0; i < f.kps.size(); i++) {
            float x = f.kps.kps2d[2 * i], y = f.kps.kps2d[2 * i + 1];
            if ((x < 0) || (y < 0) || (x > W) || (y > H) || f.kps.info[i].ignore_completely || f.kps.info[i].ignore_during_refinement) continue;
            o.kps2d.push_back(x); o.kps2d.push_back(y);
            for (int k = 0; k < 3; k++) o.kps3d.push_back(f.kps.kps3d[3 * i + k]);
            o.info.push_back(f.kps.info[i]);
        }
        f.kps = o;
    }
    // depth_calculator.cpp:88-130, called with (grid_width, grid_height) into (grid_height, grid_width): SURVEY Q6
    void merge_keypoints(Frame &f, const std::vector<float> &nk, const std::vector<KpInfo> &ni, int grid_height, int grid_width)
    {
        auto &k2 = f.kps.kps2d;
        auto &info = f.kps.info;
        for (int x = 0; x < W; x += grid_width) {
            int left = x, right = left + grid_width;
            for (int y = 0; y < H; y += grid_height) {
                int top = y, bottom = y + grid_height;
                bool match = false;
                for (size_t i = 0; i < info.size(); i++) {
                    float kx = k2[2 * i], ky = k2[2 * i + 1];
                    if (kx > left && kx < right && ky > top && ky < bottom) { match = true; break; }
                }
                if (!match) {
                    for (size_t i = 0; i < ni.size(); i++) {
                        float kx = nk[2 * i], ky = nk[2 * i + 1];
                        if (kx > left && kx < right && ky > top && ky < bottom) {
                            k2.push_back(kx); k2.push_back(ky);
                            info.push_back(ni[i]);
                        }
                    }
                }
            }
        }
    }
    // depth_calculator.cpp:132-392
    void calculate_depth(Frame &f)
    {
        const float fx = cs.fx, fy = cs.fy, cx = cs.cx, cy = cs.cy, baseline = cs.baseline;
        find_bad_keypoints(f);
        std::vector<std::vector<float>> kp_pyr;
        std::vector<std::vector<KpInfo>> info_pyr;
        detect_keypoints_on_each_level(*f.img, kp_pyr, info_pyr);
        std::vector<float> nk;
        std::vector<KpInfo> ni;
        select_best_keypoints(kp_pyr, info_pyr, nk, ni);
        if (tracing) {
            trace["det_kps2d"] = nk;
            std::vector<float> sc, ty, lv;
            for (auto &i : ni) { sc.push_back(i.score); ty.push_back((float)i.type); lv.push_back((float)i.level); }
            trace["det_score"] = sc; trace["det_type"] = ty; trace["det_level"] = lv;
        }
        size_t old_count = f.kps.size();
        merge_keypoints(f, nk, ni, cs.grid_width, cs.grid_height);
        f.kps.kps3d.resize(f.kps.size() * 3);
        const Img &left = f.img->left[0], &right = f.img->right[0];
        std::vector<float> disp_trace;
        for (size_t i = old_count; i < f.kps.size(); i++) {
            float kx = f.kps.kps2d[2 * i], ky = f.kps.kps2d[2 * i + 1];
            float disparity = stereo_disparity(left, right, kx, ky, cs, 0);
            disp_trace.push_back(disparity);
            float _z = baseline / std::max<float>(0.5, disparity);
            float _x = (kx - cx) / fx * _z;
            float _y = (ky - cy) / fy * _z;
            float loc[3] = {_x, _y, _z}, kp3[3];
            m33v(f.pose.R, loc, kp3);
            for (int k = 0; k < 3; k++) f.kps.kps3d[3 * i + k] = kp3[k] + f.pose.p[k];
            KpInfo &in = f.kps.info[i];
            in.keyframe_id = keyframe_count;
            in.keypoint_index = i;
            in.ignore_completely = false; in.ignore_temporary = true; in.ignore_during_refinement = false;
            in.inlier_count = 0; in.outlier_count = 0;
            in.kf = std::make_shared<KF1>();
            Kalman &k = in.kf->k;
            k.init(1, 1);
            k.H[0] = 1; k.Q[0] = 0.0001f; k.Ppost[0] = 1.0f;
            float deviation = (float)(0.5 / (baseline / fx));
            k.Ppost[0] = deviation * deviation;
            k.xpost[0] = 1 / _z;
        }
        if (tracing) { trace["kf_new_disp"] = disp_trace; trace["kf_old_count"] = {(float)old_count}; }
        keyframe_count++;
    }
    // keyframe_manager.cpp:15-32
    KeyFrame *create_keyframe(Frame &f)
    {
        keyframes.emplace_back(new KeyFrame());
        KeyFrame &k = *keyframes.back();
        k.id = keyframes.size() - 1;
        calculate_depth(f);
        k.kps = f.kps; k.pose = f.pose; k.img = f.img;
        k.time_stamp = f.time_stamp;
        return &k;
    }
    // keyframe_manager.cpp:47-74
    bool keyframe_needed(const Frame &f)
    {
        int inside = 0;
        for (size_t i = 0; i < f.kps.size(); i++) {
            float x = f.kps.kps2d[2 * i], y = f.kps.kps2d[2 * i + 1];
            if ((x > 0) && (y > 0) && (x < W) && (y < H) && !f.kps.info[i].ignore_completely) inside++;
        }
        int max_kps = (W / cs.grid_width) * (H / cs.grid_height);
        return inside < 0.66 * max_kps;
    }

    // ---- PoseRefiner::refine_pose (pose_refinement.cpp:62-177) ----
    float refine_pose(Frame &f)
    {
        size_t n = f.kps.size();
        std::vector<float> ref(n * 2), act(n * 2), err(n);
        std::vector<uint8_t> status(n);
        std::vector<int> kfid(n);
        // The reference groups by origin keyframe and calls calcOpticalFlowPyrLK once per keyframe; LK is
        // independent per point, so the grouped calls are equivalent to this per-point loop.
        for (size_t i = 0; i < n; i++) {
            const KpInfo &in = f.kps.info[i];
            KeyFrame *k = keyframes[in.keyframe_id].get();
            ref[2 * i] = k->kps.kps2d[2 * in.keypoint_index]; ref[2 * i + 1] = k->kps.kps2d[2 * in.keypoint_index + 1];
            act[2 * i] = f.kps.kps2d[2 * i]; act[2 * i + 1] = f.kps.kps2d[2 * i + 1];
            kfid[i] = (int)in.keyframe_id;
        }
        tr("klt_prev", ref.data(), ref.size());
        tr("klt_init", act.data(), act.size());
        tri("klt_kf", kfid);
        double t0 = now_ms();
        std::vector<float> lk_max_iters(tracing ? n : 0);
        for (size_t i = 0; i < n; i++) {
            KeyFrame *k = keyframes[f.kps.info[i].keyframe_id].get();
            int it[3];
            lk_track_point(k->img->opt_flow, f.img->opt_flow, cs.window_size_opt_flow, 2, 30, 0.01 * 0.01, 1e-4, &ref[2 * i], &act[2 * i],
                           &status[i], &err[i], tracing ? it : nullptr);
            if (tracing) lk_max_iters[i] = (float)std::max(it[0], std::max(it[1], it[2]));
            if (status[i] == 0) err[i] = std::numeric_limits<float>::infinity();  // optical_flow.cpp:46-50
        }
        stage_ms[2] = now_ms() - t0;
        tr("klt_next", act.data(), act.size());
        tr("klt_max_iters", lk_max_iters.data(), lk_max_iters.size());   // 30 on some level: that level's iteration did not converge
        tr("klt_err", err.data(), err.size());
        if (tracing) { std::vector<float> s(status.begin(), status.end()); trace["klt_status"] = s; }
        for (size_t i = n; i > 0; i--) {
            size_t j = i - 1;
            KpInfo &in = f.kps.info[j];
            float *kp = &f.kps.kps2d[2 * j];
            const float *ne = &act[2 * j];
            float diff = (kp[0] - ne[0]) * (kp[0] - ne[0]) + (kp[1] - ne[1]) * (kp[1] - ne[1]);
            if (err[j] > 20) in.ignore_completely = true;
            else if (diff > 81) in.ignore_during_refinement = true;
            else { in.ignore_during_refinement = false; kp[0] = ne[0]; kp[1] = ne[1]; }
        }
        trace_kps("ref_in", f);
        tr("ref_pose_in", f.pose.p, 6);
        PoseRefinerCb cb(f.kps.kps2d.data(), f.kps.kps3d.data(), f.kps.info, cs);
        PoseM refined;
        int ev = 0, gr = 0;
        t0 = now_ms();
        float ret = gn_driver(cb, f.pose, refined, 0.0001f, &ev, &gr);
        stage_ms[3] = now_ms() - t0;
        f.pose = refined;
        tr("ref_pose_out", f.pose.p, 6);
        if (tracing) { trace["ref_cost"] = {ret}; trace["ref_evals"] = {(float)ev, (float)gr}; }
        return ret;
    }
    void trace_kps(const std::string &pfx, const Frame &f)
    {
        if (!tracing) return;
        trace[pfx + "_kps2d"] = f.kps.kps2d;
        trace[pfx + "_kps3d"] = f.kps.kps3d;
        std::vector<float> fl, io, kfi, kpi, kx, kP;
        for (auto &i : f.kps.info) {
            fl.push_back((float)((i.ignore_during_refinement ? 1 : 0) | (i.ignore_completely ? 2 : 0) | (i.ignore_temporary ? 4 : 0)));
            io.push_back((float)i.inlier_count); io.push_back((float)i.outlier_count);
            kfi.push_back((float)i.keyframe_id); kpi.push_back((float)i.keypoint_index);
            kx.push_back(i.kf ? i.kf->k.xpost[0] : 0.f); kP.push_back(i.kf ? i.kf->k.Ppost[0] : 0.f);
        }
        trace[pfx + "_flags"] = fl; trace[pfx + "_counts"] = io; trace[pfx + "_kfid"] = kfi; trace[pfx + "_kpidx"] = kpi;
        trace[pfx + "_kfx"] = kx; trace[pfx + "_kfP"] = kP;
    }

    // ---- StereoSlam::estimate_pose (stereo_slam.cpp:58-90) ----
    void estimate_pose_stage(Frame *prev)
    {
        trace_kps("align_in", *prev);
        tr("align_pose_in", frame->pose.p, 6);
        PoseM est;
        AlignStats st;
        double t0 = now_ms();
        float cost = estimate_pose(frame->img->left, prev->img->left, prev->kps, cs, frame->pose, est, &st);
        stage_ms[1] = now_ms() - t0;
        tr("align_pose_out", est.p, 6);
        if (tracing) {
            trace["align_cost"] = {cost};
            std::vector<float> e;
            for (int l = 0; l < 8; l++) { e.push_back((float)st.evals[l]); e.push_back((float)st.grads[l]); }
            trace["align_evals"] = e;
        }
        size_t n = prev->kps.size();
        std::vector<float> est_kps(n * 2);
        project_keypoints(est, prev->kps.kps3d.data(), (int)n, cs, est_kps.data());
        frame->pose = est;
        frame->kps.info = prev->kps.info;
        frame->kps.kps3d = prev->kps.kps3d;
        frame->kps.kps2d = est_kps;
        refine_pose(*frame);
    }

    // ---- DepthFilter (depth_filter.cpp) ----
    void update_depth(Frame &f, std::vector<float> &updated)
    {
        const float fx = cs.fx, fy = cs.fy, cx = cs.cx, cy = cs.cy, baseline = cs.baseline;
        size_t n = f.kps.size();
        // calculate_disparities :259-327
        std::vector<float> disp(n);
        double t0 = now_ms();
        for (size_t i = 0; i < n; i++)
            disp[i] = stereo_disparity(f.img->left[0], f.img->right[0], f.kps.kps2d[2 * i], f.kps.kps2d[2 * i + 1], cs, 1);
        stage_ms[4] = now_ms() - t0;
        t0 = now_ms();
        tr("df_disp", disp.data(), n);
        // outlier_check :52-128
        std::vector<float> k3(n * 3);
        for (size_t i = 0; i < n; i++) {
            float d = disp[i];
            float _z = baseline / std::max<float>(d, 0.5);
            float _x = (f.kps.kps2d[2 * i] - cx) / fx * _z;
            float _y = (f.kps.kps2d[2 * i + 1] - cy) / fy * _z;
            float v[3] = {_x, _y, _z}, o[3];
            m33v(f.pose.R, v, o);
            for (int k = 0; k < 3; k++) k3[3 * i + k] = o[k] + f.pose.p[k];
        }
        for (size_t i = 0; i < n; i++) {
            KpInfo &in = f.kps.info[i];
            const KeyFrame *k = keyframes[in.keyframe_id].get();
            const float *ref = &k->kps.kps3d[3 * in.keypoint_index];
            float a[3] = {k3[3 * i] - k->pose.p[0], k3[3 * i + 1] - k->pose.p[1], k3[3 * i + 2] - k->pose.p[2]};
            m33v(k->pose.Ri, a, a);
            float b[3] = {ref[0] - k->pose.p[0], ref[1] - k->pose.p[1], ref[2] - k->pose.p[2]};
            m33v(k->pose.Ri, b, b);
            float disp_ref = baseline / b[2];
            float dsp = baseline / a[2];
            float pixel_distance = dsp - disp_ref;
            float deviation = 0.5;
            if (fabsf(pixel_distance) > 5 * deviation) in.outlier_count++;
            else in.inlier_count++;
        }
        // update_kps3d :130-257
        updated = f.kps.kps3d;
        for (size_t i = 0; i < n; i++) {
            KpInfo &in = f.kps.info[i];
            const KeyFrame *k = keyframes[in.keyframe_id].get();
            const float *c1 = k->pose.p, *c2 = f.pose.p;
            float diff[3] = {fabsf(c1[0] - c2[0]), fabsf(c1[1] - c2[1]), fabsf(c1[2] - c2[2])};
            m33v(k->pose.Ri, diff, diff);
            if (in.ignore_completely || in.ignore_during_refinement) { in.outlier_count++; continue; }
            if (diff[0] < 0.1 && diff[1] < 0.1) continue;
            const float *rf = &k->kps.kps2d[2 * in.keypoint_index];
            float p1[3] = {rf[0] - cx, rf[1] - cy, fx};
            m33v(k->pose.R, p1, p1);
            float p2[3] = {f.kps.kps2d[2 * i] - cx, f.kps.kps2d[2 * i + 1] - cy, fx};
            m33v(f.pose.R, p2, p2);
            float A[6] = {p1[0], -p2[0], p1[1], -p2[1], p1[2], -p2[2]};
            float y[3] = {c2[0] - c1[0], c2[1] - c1[1], c2[2] - c1[2]};
            float l[2];
            solve_svd_f(A, 3, 2, y, 1, l);
            Kalman &kf1 = in.kf->k;
            float deviation = (float)(0.5 / (sqrtf(diff[0] * diff[0] + diff[1] * diff[1])));
            kf1.R[0] = deviation * deviation;
            kf1.predict();
            float Ms[9];
            for (int q = 0; q < 9; q++) Ms[q] = k->pose.Ri[q] * l[0];
            float pc[3] = {p1[0] - c1[0], p1[1] - c1[1], p1[2] - c1[2]}, np[3];
            m33v(Ms, pc, np);
            float _z = np[2];
            float meas = 1 / _z;
            kf1.correct(&meas);
            _z = (float)(1.0 / kf1.xpost[0]);
            float _x = (rf[0] - cx) / fx * _z;
            float _y = (rf[1] - cy) / fy * _z;
            float cp[3] = {_x, _y, _z}, o[3];
            m33v(k->pose.R, cp, o);
            for (int q = 0; q < 3; q++) updated[3 * i + q] = c1[q] + o[q];
        }
        stage_ms[5] = now_ms() - t0;
    }

    // ---- StereoSlam::update_pose (stereo_slam.cpp:296-359) ----
    void update_pose_kf(const float pose[6], const float speed[6], const float pv[6], const float sv[6], double dt, float out[6])
    {
        for (int i = 0; i < 6; i++) kf.A[i * 12 + (i + 6)] = (float)dt;
        kf.predict();
        for (int i = 0; i < 6; i++) { kf.R[i * 12 + i] = pv[i]; kf.R[(i + 6) * 12 + (i + 6)] = sv[i]; }
        float z[12];
        for (int i = 0; i < 6; i++) { z[i] = pose[i]; z[i + 6] = speed[i]; }
        kf.correct(z);
        for (int i = 0; i < 6; i++) out[i] = kf.xpost[i];
    }

    // ---- StereoSlam::new_image (stereo_slam.cpp:123-271) ----
    void new_image(const uint8_t *left, const uint8_t *right, int stride, float time_stamp)
    {
        trace.clear();
        std::shared_ptr<Frame> previous = frame;
        frame = std::make_shared<Frame>();
        frame->time_stamp = time_stamp;
        frame->img = std::make_shared<StereoImage>();
        for (double &m : stage_ms) m = 0;
        double t_start = now_ms(), t0 = t_start;
        create_img_pyramid(left, W, H, stride, cs.max_pyramid_levels, frame->img->left);
        create_img_pyramid(right, W, H, stride, 1, frame->img->right);
        build_lk_pyramid(left, W, H, stride, 2, frame->img->opt_flow);
        stage_ms[0] = now_ms() - t0;
        if (!previous) {
            frame->id = 0;
            float z[6] = {0, 0, 0, 0, 0, 0};
            frame->pose.set(z);
            t0 = now_ms();
            keyframe = create_keyframe(*frame);
            stage_ms[6] = now_ms() - t0;
            for (auto &i : frame->kps.info) i.ignore_temporary = false;
            // the keyframe copy was taken before the flags were cleared (keyframe_manager.cpp:27 precedes
            // stereo_slam.cpp:157-159); later write-backs (stereo_slam.cpp:218-223) refresh it.
        } else {
            frame->id = previous->id + 1;
            float pp[6];
            for (int i = 0; i < 6; i++) pp[i] = kf.xpre[i];  // statePre
            frame->pose.set(pp);
            // remove_outliers :43-56
            {
                KeyPoints u;
                for (size_t i = 0; i < previous->kps.size(); i++) {
                    if (previous->kps.info[i].ignore_completely) continue;
                    u.kps2d.push_back(previous->kps.kps2d[2 * i]); u.kps2d.push_back(previous->kps.kps2d[2 * i + 1]);
                    for (int k = 0; k < 3; k++) u.kps3d.push_back(previous->kps.kps3d[3 * i + k]);
                    u.info.push_back(previous->kps.info[i]);
                }
                previous->kps = u;
            }
            estimate_pose_stage(previous.get());
            trace_kps("df_in", *frame);
            tr("df_pose", frame->pose.p, 6);
            std::vector<float> updated;
            update_depth(*frame, updated);
            for (size_t i = 0; i < frame->kps.size(); i++) {
                KpInfo &info = frame->kps.info[i];
                KeyFrame *k = keyframes[info.keyframe_id].get();
                if (info.outlier_count > info.inlier_count) info.ignore_completely = true;
                if (info.inlier_count > info.outlier_count) info.ignore_temporary = false;
                for (int q = 0; q < 3; q++) {
                    k->kps.kps3d[3 * info.keypoint_index + q] = updated[3 * i + q];
                    frame->kps.kps3d[3 * i + q] = updated[3 * i + q];
                }
                KpInfo &ki = k->kps.info[info.keypoint_index];
                ki.ignore_temporary = info.ignore_temporary;
                ki.ignore_completely = info.ignore_completely;
                ki.inlier_count = info.inlier_count;
                ki.outlier_count = info.outlier_count;
            }
            project_keypoints(frame->pose, frame->kps.kps3d.data(), (int)frame->kps.size(), cs, frame->kps.kps2d.data());
            trace_kps("df_out", *frame);
            bool need = keyframe_needed(*frame);
            if (tracing) trace["kf_needed"] = {need ? 1.f : 0.f};
            if (need) {
                t0 = now_ms();
                keyframe = create_keyframe(*frame);
                stage_ms[6] = now_ms() - t0;
                size_t c = 0;
                for (auto &i : frame->kps.info) if (!i.ignore_temporary) c++;
                if (c < frame->kps.info.size() / 4)
                    for (auto &i : frame->kps.info) i.ignore_temporary = false;
            }
        }
        if (previous) {
            double dt = frame->time_stamp - previous->time_stamp;
            for (int i = 0; i < 6; i++) motion[i] = (float)((double)(frame->pose.p[i] - previous->pose.p[i]) * (1. / dt));
            float pv[6] = {0.1f, 0.1f, 0.1f, 0.1f, 0.1f, 0.1f}, mv[6] = {1, 1, 1, 1, 1, 1}, fp[6];
            tr("kal_pose_in", frame->pose.p, 6);
            update_pose_kf(frame->pose.p, motion, pv, mv, 0.0, fp);
            frame->pose.set(fp);
        }
        trace_kps("out", *frame);
        tr("out_pose", frame->pose.p, 6);
        std::array<float, 6> p;
        for (int i = 0; i < 6; i++) p[i] = frame->pose.p[i];
        trajectory.push_back(p);
        stage_ms[7] = now_ms() - t_start;
    }
};

// ============================================================================= C exports (ctypes)
extern "C" {

void orc_half_sample(const uint8_t *in, int w, int h, int stride, uint8_t *out) { half_sample(in, w, h, stride, out); }
void orc_pyr_down(const uint8_t *in, int w, int h, uint8_t *out) { pyr_down(in, w, h, w, out, (w + 1) / 2, (h + 1) / 2, (w + 1) / 2); }
void orc_scharr(const uint8_t *in, int w, int h, int16_t *out) { scharr_deriv(in, w, h, w, out); }
void orc_sobel_x(const uint8_t *in, int w, int h, uint8_t *out) { sobel_x_u8(in, w, h, w, out, w); }
int orc_fast(const uint8_t *img, int w, int h, int thr, int max, int *out_xys)
{
    std::vector<FastKp> v;
    fast9_16_nms(img, w, h, w, thr, v);
    int n = (int)std::min<size_t>(v.size(), (size_t)max);
    for (int i = 0; i < n; i++) { out_xys[3 * i] = v[i].x; out_xys[3 * i + 1] = v[i].y; out_xys[3 * i + 2] = v[i].score; }
    return (int)v.size();
}
int orc_detect_keypoints(const uint8_t *img, int w, int h, int grid_w, int grid_h, int level, int max, float *xy, float *score, int *type)
{
    Img im; im.w = w; im.h = h; im.d.assign(img, img + (size_t)w * h);
    std::vector<float> k; std::vector<KpInfo> inf;
    detect_keypoints(im, grid_w, grid_h, k, inf, level);
    int n = (int)std::min<size_t>(inf.size(), (size_t)max);
    for (int i = 0; i < n; i++) { xy[2 * i] = k[2 * i]; xy[2 * i + 1] = k[2 * i + 1]; score[i] = inf[i].score; type[i] = inf[i].type; }
    return (int)inf.size();
}
void orc_rodrigues(const float *r, float *R) { rodrigues_f(r, R); }
void orc_project(const OrcCameraSettings *cs, const float *pose6, const float *pts3, int n, float *out2)
{
    PoseM p; p.set(pose6);
    project_keypoints(p, pts3, n, *cs, out2);
}
int orc_invert_svd(const float *A, int n, float *Ainv) { return invert_svd_f(A, n, Ainv) ? 1 : 0; }
void orc_solve_svd(const float *A, int m, int n, const float *B, int nb, float *X) { solve_svd_f(A, m, n, B, nb, X); }
void orc_expmap(const float *tw, float *out) { exponential_map(tw, out); }

void orc_rectify_map(const double *K, const double *D, const double *R, const double *P, int w, int h, float *m1, float *m2)
{
    init_undistort_rectify_map(K, D, R, P, w, h, m1, m2);
}
void orc_remap(const uint8_t *src, int sw, int sh, const float *m1, const float *m2, int w, int h, uint8_t *dst)
{
    remap_linear_u8(src, sw, sh, sw, m1, m2, w, h, dst, w);
}

// generic Kalman for pinning against cv2.KalmanFilter
void *orc_kf_create(int n, int m) { Kalman *k = new Kalman(); k->init(n, m); return k; }
void orc_kf_destroy(void *h) { delete (Kalman *)h; }
float *orc_kf_mat(void *h, int which)
{
    Kalman *k = (Kalman *)h;
    switch (which) {
        case 0: return k->A.data(); case 1: return k->H.data(); case 2: return k->Q.data(); case 3: return k->R.data();
        case 4: return k->Ppre.data(); case 5: return k->Ppost.data(); case 6: return k->xpre.data(); case 7: return k->xpost.data();
        case 8: return k->K.data();
    }
    return nullptr;
}
void orc_kf_predict(void *h) { ((Kalman *)h)->predict(); }
void orc_kf_correct(void *h, const float *z) { ((Kalman *)h)->correct(z); }

void orc_lk(const uint8_t *prev, const uint8_t *next, int w, int h, int win, const float *prev_pts, float *next_pts, int n,
            uint8_t *status, float *err)
{
    LKPyramid a, b;
    build_lk_pyramid(prev, w, h, w, 2, a);
    build_lk_pyramid(next, w, h, w, 2, b);
    for (int i = 0; i < n; i++)
        lk_track_point(a, b, win, 2, 30, 0.01 * 0.01, 1e-4, prev_pts + 2 * i, next_pts + 2 * i, status + i, err + i);
}
// same as orc_lk, plus per point the largest iteration count any level used (30 = that level ran out of iterations)
void orc_lk_iters(const uint8_t *prev, const uint8_t *next, int w, int h, int win, const float *prev_pts, float *next_pts, int n,
                  uint8_t *status, float *err, int *max_iters)
{
    LKPyramid a, b;
    build_lk_pyramid(prev, w, h, w, 2, a);
    build_lk_pyramid(next, w, h, w, 2, b);
    for (int i = 0; i < n; i++) {
        int it[3];
        lk_track_point(a, b, win, 2, 30, 0.01 * 0.01, 1e-4, prev_pts + 2 * i, next_pts + 2 * i, status + i, err + i, it);
        max_iters[i] = std::max(it[0], std::max(it[1], it[2]));
    }
}
void orc_ssd_disparity(const uint8_t *left, const uint8_t *right, int w, int h, const OrcCameraSettings *cs, const float *kps2d, int n,
                       int mode, float *out)
{
    Img l, r; l.w = r.w = w; l.h = r.h = h;
    l.d.assign(left, left + (size_t)w * h); r.d.assign(right, right + (size_t)w * h);
    for (int i = 0; i < n; i++) out[i] = stereo_disparity(l, r, kps2d[2 * i], kps2d[2 * i + 1], *cs, mode);
}
// exact integer SSD map for one keypoint (for comparison with cv2.matchTemplate); returns mw | mh<<16
int orc_ssd_map(const uint8_t *roi, int rw, int rh, int rstride, const uint8_t *tpl, int tw, int th, int tstride, uint32_t *out, int max)
{
    std::vector<uint32_t> m; int mw, mh;
    ssd_map_u32(roi, rw, rh, rstride, tpl, tw, th, tstride, m, mw, mh);
    for (size_t i = 0; i < m.size() && (int)i < max; i++) out[i] = m[i];
    return mw | (mh << 16);
}
// sparse image alignment between two left images (halfSample pyramids built here)
float orc_align(const uint8_t *prev_left, const uint8_t *cur_left, int w, int h, const OrcCameraSettings *cs, const float *kps2d,
                const float *kps3d, int n, const float *pose_in, float *pose_out, int *evals16)
{
    std::vector<Img> pp, cp;
    create_img_pyramid(prev_left, w, h, w, cs->max_pyramid_levels, pp);
    create_img_pyramid(cur_left, w, h, w, cs->max_pyramid_levels, cp);
    KeyPoints k;
    k.kps2d.assign(kps2d, kps2d + 2 * n); k.kps3d.assign(kps3d, kps3d + 3 * n); k.info.resize(n);
    PoseM g, o; g.set(pose_in);
    AlignStats st;
    float c = estimate_pose(cp, pp, k, *cs, g, o, &st);
    for (int i = 0; i < 6; i++) pose_out[i] = o.p[i];
    if (evals16) for (int l = 0; l < 8; l++) { evals16[2 * l] = st.evals[l]; evals16[2 * l + 1] = st.grads[l]; }
    return c;
}
float orc_align_cost(const uint8_t *prev_left, const uint8_t *cur_left, int w, int h, const OrcCameraSettings *cs, const float *kps2d,
                     const float *kps3d, int n, const float *pose, int level, float *grad6)
{
    std::vector<Img> pp, cp;
    create_img_pyramid(prev_left, w, h, w, cs->max_pyramid_levels, pp);
    create_img_pyramid(cur_left, w, h, w, cs->max_pyramid_levels, cp);
    KeyPoints k;
    k.kps2d.assign(kps2d, kps2d + 2 * n); k.kps3d.assign(kps3d, kps3d + 3 * n); k.info.resize(n);
    PoseEstimatorCb cb(cp, pp, k, *cs);
    cb.set_level(level);
    PoseM p; p.set(pose);
    if (grad6) cb.get_gradient(p, grad6);
    return cb.do_calc(p);
}
float orc_refine(const OrcCameraSettings *cs, const float *kps2d, const float *kps3d, const int *flags, int n, const float *pose_in,
                 float *pose_out, int *evals2)
{
    std::vector<KpInfo> info(n);
    for (int i = 0; i < n; i++) {
        info[i].ignore_during_refinement = flags[i] & 1; info[i].ignore_completely = flags[i] & 2; info[i].ignore_temporary = flags[i] & 4;
    }
    PoseRefinerCb cb(kps2d, kps3d, info, *cs);
    PoseM g, o; g.set(pose_in);
    float c = gn_driver(cb, g, o, 0.0001f, evals2, evals2 ? evals2 + 1 : nullptr);
    for (int i = 0; i < 6; i++) pose_out[i] = o.p[i];
    return c;
}

// ---- pipeline ----
void *orc_slam_create(const OrcCameraSettings *cs, int w, int h) { return new OracleSlam(*cs, w, h); }
void orc_slam_destroy(void *h) { delete (OracleSlam *)h; }
void orc_slam_set_tracing(void *h, int on) { ((OracleSlam *)h)->tracing = on != 0; }
void orc_slam_new_image(void *h, const uint8_t *l, const uint8_t *r, int stride, float ts) { ((OracleSlam *)h)->new_image(l, r, stride, ts); }
void orc_slam_update_pose(void *h, const float *pose, const float *speed, const float *pv, const float *sv, double dt, float *out)
{
    ((OracleSlam *)h)->update_pose_kf(pose, speed, pv, sv, dt, out);
}
void orc_slam_get_pose(void *h, float *p6) { OracleSlam *s = (OracleSlam *)h; for (int i = 0; i < 6; i++) p6[i] = s->frame ? s->frame->pose.p[i] : 0; }
void orc_slam_stage_ms(void *h, double *out8) { for (int i = 0; i < 8; i++) out8[i] = ((OracleSlam *)h)->stage_ms[i]; }
int orc_slam_n_kps(void *h) { OracleSlam *s = (OracleSlam *)h; return s->frame ? (int)s->frame->kps.size() : 0; }
int orc_slam_n_keyframes(void *h) { return (int)((OracleSlam *)h)->keyframes.size(); }
int orc_slam_trajectory(void *h, float *out, int max)
{
    OracleSlam *s = (OracleSlam *)h;
    int n = (int)std::min<size_t>(s->trajectory.size(), (size_t)max);
    for (int i = 0; i < n; i++) for (int k = 0; k < 6; k++) out[6 * i + k] = s->trajectory[i][k];
    return (int)s->trajectory.size();
}
// named float trace of the last new_image call; returns element count (or -1 if absent)
int orc_slam_trace(void *h, const char *name, float *out, int max)
{
    OracleSlam *s = (OracleSlam *)h;
    auto it = s->trace.find(name);
    if (it == s->trace.end()) return -1;
    int n = (int)std::min<size_t>(it->second.size(), (size_t)max);
    if (out) memcpy(out, it->second.data(), sizeof(float) * n);
    return (int)it->second.size();
}
// ---- DepthCalculator / KeyFrameManager bookkeeping on its own (the functions OracleSlam uses), for the host-stage tests ----
int orc_select_best(int n_levels, const int *n_per_level, const float *const *xy, const float *const *score, const int *const *type,
                    float *kps2d, float *score_out, int *type_out, int *level_out)
{
    std::vector<std::vector<float>> kp(n_levels);
    std::vector<std::vector<KpInfo>> inf(n_levels);
    for (int l = 0; l < n_levels; l++) {
        kp[l].assign(xy[l], xy[l] + 2 * (size_t)n_per_level[l]);
        inf[l].resize(n_per_level[l]);
        for (int j = 0; j < n_per_level[l]; j++) { inf[l][j].score = score[l][j]; inf[l][j].type = type[l][j]; inf[l][j].level = l; }
    }
    std::vector<float> k;
    std::vector<KpInfo> in;
    OracleSlam::select_best_keypoints(kp, inf, k, in);
    for (size_t j = 0; j < in.size(); j++) {
        kps2d[2 * j] = k[2 * j]; kps2d[2 * j + 1] = k[2 * j + 1];
        score_out[j] = in[j].score; type_out[j] = in[j].type; level_out[j] = in[j].level;
    }
    return (int)in.size();
}
static void flat_frame(Frame &f, int n, const float *kps2d, const uint8_t *flags)
{
    f.kps.kps2d.assign(kps2d, kps2d + 2 * (size_t)n);
    f.kps.kps3d.assign(3 * (size_t)n, 0.f);
    f.kps.info.resize(n);
    for (int i = 0; i < n; i++) {
        f.kps.info[i].keypoint_index = (size_t)i;   // remembers the original position
        if (flags) { f.kps.info[i].ignore_during_refinement = flags[i] & 1; f.kps.info[i].ignore_completely = flags[i] & 2; f.kps.info[i].ignore_temporary = flags[i] & 4; }
    }
}
void orc_find_bad(int w, int h, int n, const float *kps2d, const uint8_t *flags, uint8_t *keep)
{
    Settings cs = {};
    OracleSlam s(cs, w, h);
    Frame f;
    flat_frame(f, n, kps2d, flags);
    s.find_bad_keypoints(f);
    for (int i = 0; i < n; i++) keep[i] = 0;
    for (auto &in : f.kps.info) keep[in.keypoint_index] = 1;
}
// merge as calculate_depth calls it (grid arguments swapped); appended = indices into the new list
int orc_merge(int w, int h, int grid_width, int grid_height, int n_old, const float *old2d, int n_new, const float *new2d, int *appended)
{
    Settings cs = {};
    OracleSlam s(cs, w, h);
    Frame f;
    flat_frame(f, n_old, old2d, nullptr);
    std::vector<float> nk(new2d, new2d + 2 * (size_t)n_new);
    std::vector<KpInfo> ni(n_new);
    for (int i = 0; i < n_new; i++) ni[i].keypoint_index = (size_t)i;
    s.merge_keypoints(f, nk, ni, grid_width, grid_height);
    int c = 0;
    for (size_t i = (size_t)n_old; i < f.kps.size(); i++) appended[c++] = (int)f.kps.info[i].keypoint_index;
    return c;
}
int orc_keyframe_needed(int w, int h, int grid_width, int grid_height, int n, const float *kps2d, const uint8_t *flags)
{
    Settings cs = {};
    cs.grid_width = grid_width; cs.grid_height = grid_height;
    OracleSlam s(cs, w, h);
    Frame f;
    flat_frame(f, n, kps2d, flags);
    return s.keyframe_needed(f) ? 1 : 0;
}

// flat dump of a keypoint set, same layout as oracle/ref_capi.cpp's dump_kps (colour is not modelled: rand())
static int dump_kps(const KeyPoints &k, float *kps2d, float *kps3d, int32_t *info8, float *finfo3, int max)
{
    int n = (int)k.size(), c = std::min(n, max);
    for (int i = 0; i < c; i++) {
        if (kps2d && (size_t)(2 * i + 1) < k.kps2d.size()) { kps2d[2 * i] = k.kps2d[2 * i]; kps2d[2 * i + 1] = k.kps2d[2 * i + 1]; }
        if (kps3d && (size_t)(3 * i + 2) < k.kps3d.size()) for (int q = 0; q < 3; q++) kps3d[3 * i + q] = k.kps3d[3 * i + q];
        const KpInfo &in = k.info[i];
        if (info8) {
            int32_t *o = info8 + 8 * i;
            o[0] = in.level; o[1] = in.type; o[2] = (int32_t)in.keyframe_id; o[3] = (int32_t)in.keypoint_index;
            o[4] = (in.ignore_during_refinement ? 1 : 0) | (in.ignore_completely ? 2 : 0) | (in.ignore_temporary ? 4 : 0);
            o[5] = in.inlier_count; o[6] = in.outlier_count; o[7] = 0;
        }
        if (finfo3) { finfo3[3 * i] = in.score; finfo3[3 * i + 1] = in.kf ? in.kf->k.xpost[0] : 0.f; finfo3[3 * i + 2] = in.kf ? in.kf->k.Ppost[0] : 0.f; }
    }
    return n;
}
int orc_slam_frame(void *h, float *p6, uint64_t *id, double *ts, float *kps2d, float *kps3d, int32_t *info8, float *finfo3, int max)
{
    OracleSlam *s = (OracleSlam *)h;
    if (!s->frame) return -1;
    if (p6) for (int i = 0; i < 6; i++) p6[i] = s->frame->pose.p[i];
    if (id) *id = s->frame->id;
    if (ts) *ts = s->frame->time_stamp;
    return dump_kps(s->frame->kps, kps2d, kps3d, info8, finfo3, max);
}
int orc_slam_keyframe_full(void *h, int k, float *p6, uint64_t *id, float *kps2d, float *kps3d, int32_t *info8, float *finfo3, int max)
{
    OracleSlam *s = (OracleSlam *)h;
    if (k < 0 || k >= (int)s->keyframes.size()) return -1;
    KeyFrame &f = *s->keyframes[k];
    if (p6) for (int i = 0; i < 6; i++) p6[i] = f.pose.p[i];
    if (id) *id = f.id;
    return dump_kps(f.kps, kps2d, kps3d, info8, finfo3, max);
}
// keyframe k: pose (6) ; kps2d/kps3d of the keyframe's own arrays
int orc_slam_keyframe(void *h, int k, float *pose6, float *kps2d, float *kps3d, int max)
{
    OracleSlam *s = (OracleSlam *)h;
    if (k < 0 || k >= (int)s->keyframes.size()) return -1;
    KeyFrame &f = *s->keyframes[k];
    if (pose6) for (int i = 0; i < 6; i++) pose6[i] = f.pose.p[i];
    int n = (int)std::min<size_t>(f.kps.size(), (size_t)max);
    if (kps2d) memcpy(kps2d, f.kps.kps2d.data(), sizeof(float) * 2 * n);
    if (kps3d) memcpy(kps3d, f.kps.kps3d.data(), sizeof(float) * 3 * n);
    return (int)f.kps.size();
}
}  // extern "C"
