// Host facade: the reference's StereoSlam (src/lib/stereo_slam.cpp, src/include/stereo_slam.hpp:27-79)
// re-implemented on top of the svo_* device C-ABI only.  This file contains no CUDA: it owns the
// bookkeeping the reference does on the CPU between device stages — keypoint compaction
// (remove_outliers, stereo_slam.cpp:43-56), keyframe store and keyframe decision (keyframe_manager.cpp),
// cross-level keypoint selection / merge (depth_calculator.cpp:37-130), 3-D initialisation of new keypoints
// (:240-289), write-back into origin keyframes (stereo_slam.cpp:205-226), the 12-state motion Kalman filter
// (:29-41, :296-359) and the trajectory.  Keyframe ids are per instance (the reference's process-global
// statics, depth_calculator.cpp:135 / keyframe_manager.cpp:8, would break multi-sequence use).
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <memory>
#include <vector>
#include <chrono>
#include <cstdlib>
#include <thread>
#include <atomic>

#include "../../include/svo_cuda.h"

// developer trace (SVO_TRACE_KF=1): host time of the keyframe path, stage by stage
static const bool g_trace_kf = getenv("SVO_TRACE_KF") != nullptr;
// developer trace (SVO_HOST_PROFILE=1): host time per frame by phase, printed when a facade instance is destroyed
static const bool g_host_prof = getenv("SVO_HOST_PROFILE") != nullptr;
static inline double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

namespace {

// ---- cv::Rodrigues semantics (double inside, float out) -------------------------------------------------
void rodrigues_f(const float r[3], float R[9])
{
    double rx = r[0], ry = r[1], rz = r[2];
    double theta = std::sqrt(rx * rx + ry * ry + rz * rz);
    double Rd[9];
    if (theta < 2.220446049250313e-16) {
        for (int k = 0; k < 9; k++) Rd[k] = (k % 4 == 0) ? 1.0 : 0.0;
    } else {
        double c = std::cos(theta), s = std::sin(theta), c1 = 1. - c, it = 1. / theta;
        rx *= it; ry *= it; rz *= it;
        double rrt[9] = {rx * rx, rx * ry, rx * rz, rx * ry, ry * ry, ry * rz, rx * rz, ry * rz, rz * rz};
        double r_x[9] = {0, -rz, ry, rz, 0, -rx, -ry, rx, 0};
        for (int k = 0; k < 9; k++) Rd[k] = c * ((k % 4 == 0) ? 1.0 : 0.0) + c1 * rrt[k] + s * r_x[k];
    }
    for (int k = 0; k < 9; k++) R[k] = (float)Rd[k];
}

inline void m33v(const float M[9], const float v[3], float o[3])
{
    float t[3];
    for (int i = 0; i < 3; i++) {
        float s = 0;
        for (int k = 0; k < 3; k++) s += M[i * 3 + k] * v[k];
        t[i] = s;
    }
    o[0] = t[0]; o[1] = t[1]; o[2] = t[2];
}

// ---- cv::KalmanFilter (float matrices, products accumulated in double like cv::gemm) -----------------
struct MotionFilter {
    static const int N = 12;
    float A[N * N], Q[N * N], R[N * N], Ppre[N * N], Ppost[N * N], xpre[N], xpost[N];
    MotionFilter()
    {
        std::memset(this, 0, sizeof(*this));
        for (int i = 0; i < N; i++) {
            A[i * N + i] = 1;            // transitionMatrix = I      (stereo_slam.cpp:34)
            Q[i * N + i] = 100.0f;       // processNoiseCov = 100 I   (:37)
            R[i * N + i] = 1;            // KalmanFilter::init default
            Ppost[i * N + i] = 1.0f;     // errorCovPost = I          (:38)
        }
    }
    static void mul(const float *a, const float *b, const float *c, float *d, int m, int k, int n, bool tb, double alpha)
    {
        std::vector<float> tmp((size_t)m * n);
        for (int i = 0; i < m; i++)
            for (int j = 0; j < n; j++) {
                double s = 0;
                for (int l = 0; l < k; l++) s += (double)a[i * k + l] * (tb ? b[j * k + l] : b[l * n + j]);
                tmp[i * n + j] = (float)(c ? s * alpha + (double)c[i * n + j] : s * alpha);
            }
        std::memcpy(d, tmp.data(), sizeof(float) * m * n);
    }
    void predict()
    {
        mul(A, xpost, nullptr, xpre, N, N, 1, false, 1.0);
        float t1[N * N];
        mul(A, Ppost, nullptr, t1, N, N, N, false, 1.0);
        mul(t1, A, Q, Ppre, N, N, N, true, 1.0);
        std::memcpy(xpost, xpre, sizeof(xpre));
        std::memcpy(Ppost, Ppre, sizeof(Ppre));
    }
    // measurementMatrix = I (stereo_slam.cpp:35): gain = Ppre (Ppre + R)^-1, solved by Gaussian elimination in double
    void correct(const float z[N])
    {
        double S[N][N], B[N][N];  // S X = B with B = Ppre ; gain = X^T
        for (int i = 0; i < N; i++)
            for (int j = 0; j < N; j++) {
                S[i][j] = (double)(float)((double)Ppre[i * N + j] + (double)R[i * N + j]);
                B[i][j] = Ppre[i * N + j];
            }
        for (int c = 0; c < N; c++) {
            int p = c;
            for (int r = c + 1; r < N; r++) if (std::fabs(S[r][c]) > std::fabs(S[p][c])) p = r;
            if (p != c) for (int j = 0; j < N; j++) { std::swap(S[p][j], S[c][j]); std::swap(B[p][j], B[c][j]); }
            double d = S[c][c];
            if (d == 0) continue;
            for (int r = 0; r < N; r++) {
                if (r == c) continue;
                double f = S[r][c] / d;
                if (f == 0) continue;
                for (int j = 0; j < N; j++) { S[r][j] -= f * S[c][j]; B[r][j] -= f * B[c][j]; }
            }
        }
        float K[N * N];
        for (int i = 0; i < N; i++)
            for (int j = 0; j < N; j++) K[j * N + i] = (float)(S[i][i] != 0 ? B[i][j] / S[i][i] : 0.0);  // gain = X^T
        float t5[N];
        for (int i = 0; i < N; i++) t5[i] = (float)(-(double)xpre[i] + (double)z[i]);
        mul(K, t5, xpre, xpost, N, N, 1, false, 1.0);
        mul(K, Ppre, Ppre, Ppost, N, N, N, false, -1.0);
    }
};

struct KeyPoints {
    std::vector<float> kps2d, kps3d;
    std::vector<svo_keypoint_info> info;
    size_t size() const { return info.size(); }
    void clear() { kps2d.clear(); kps3d.clear(); info.clear(); }
};

struct FrameH {
    uint64_t id = 0;
    float pose[6] = {0, 0, 0, 0, 0, 0};
    int slot = -1;
    KeyPoints kps;
    double time_stamp = 0;
};

// ---- DepthCalculator / KeyFrameManager host logic on flat arrays (also exported one by one as svo_host_*) ---------
// depth_calculator.cpp:67-86: a keypoint is dropped when it left the image (note `> width`, `> height`) or carries
// ignore_completely / ignore_during_refinement
inline bool host_is_bad(float x, float y, int W, int H, bool ignore_completely, bool ignore_during_refinement)
{
    return (x < 0) || (y < 0) || (x > W) || (y > H) || ignore_completely || ignore_during_refinement;
}

// depth_calculator.cpp:37-65, one coarser level folded into the running choice: entry j of EVERY level is compared with
// entry j of level 0 (the lists are indexed by position, not by cell — SURVEY Q7); a FAST corner beats an edgelet,
// within a type the coarser level wins unless the finer score is strictly higher
void host_select_level(int lv, int n, const float *xy, const float *sc, const int *ty, std::vector<float> &kps,
                       std::vector<svo_keypoint_info> &info)
{
    if (lv == 0) {
        kps.assign(xy, xy + 2 * (size_t)n);
        info.resize(n);
        for (int j = 0; j < n; j++) {
            std::memset(&info[j], 0, sizeof(svo_keypoint_info));
            info[j].score = sc[j]; info[j].type = ty[j]; info[j].level = 0;
        }
        return;
    }
    for (size_t j = 0; j < info.size(); j++) {
        if ((int)j >= n) continue;  // the reference reads past the coarser level's list here (Q7)
        if (info[j].type == SVO_KP_FAST && ty[j] == SVO_KP_EDGELET) continue;
        if ((info[j].type == ty[j]) && (info[j].score > sc[j])) continue;
        kps[2 * j] = xy[2 * j] * (float)(1 << lv);
        kps[2 * j + 1] = xy[2 * j + 1] * (float)(1 << lv);
        info[j].score = sc[j]; info[j].type = ty[j]; info[j].level = lv;
    }
}

// depth_calculator.cpp:88-130.  The caller passes (grid_width, grid_height) into the parameters named (grid_height,
// grid_width) (:179-180), so the cells walked here are grid_height wide and grid_width tall; cell membership
// is by strict inequalities on all four sides.  Returns, in order, the indices of the new keypoints that are appended.
void host_merge(int W, int H, int cell_w, int cell_h, const float *old2d, size_t n_old, const float *new2d, size_t n_new,
                std::vector<int> &appended)
{
    appended.clear();
    for (int x = 0; x < W; x += cell_w) {
        const int left = x, right = left + cell_w;
        for (int y = 0; y < H; y += cell_h) {
            const int top = y, bottom = y + cell_h;
            bool match = false;
            // kps2d grows while the loops run: keypoints appended for earlier cells are looked at as well
            for (size_t i = 0; i < n_old + appended.size(); i++) {
                const float *p = i < n_old ? old2d + 2 * i : new2d + 2 * (size_t)appended[i - n_old];
                if (p[0] > left && p[0] < right && p[1] > top && p[1] < bottom) { match = true; break; }
            }
            if (match) continue;
            for (size_t i = 0; i < n_new; i++) {
                const float kx = new2d[2 * i], ky = new2d[2 * i + 1];
                if (kx > left && kx < right && ky > top && ky < bottom) appended.push_back((int)i);
            }
        }
    }
}

// keyframe_manager.cpp:47-74
bool host_keyframe_needed(int W, int H, int grid_width, int grid_height, const float *k2, const uint8_t *ignore_completely, size_t stride, size_t n)
{
    int inside = 0;
    for (size_t i = 0; i < n; i++) {
        const float x = k2[2 * i], y = k2[2 * i + 1];
        if ((x > 0) && (y > 0) && (x < W) && (y < H) && !ignore_completely[i * stride]) inside++;
    }
    const int max_kps = (W / grid_width) * (H / grid_height);
    return inside < 0.66 * max_kps;
}

}  // namespace

struct svo_slam {
    svo_camera_settings cs;
    int W = 0, H = 0;
    svo_ctx *ctx = nullptr;
    std::vector<std::unique_ptr<FrameH>> keyframes;
    std::unique_ptr<FrameH> frame, previous;
    std::vector<svo_pose> trajectory;
    float motion[6] = {0, 0, 0, 0, 0, 0};
    MotionFilter kf;
    uint32_t rng = 0x2545F491u;
    char err[256] = "";
    // per-frame scratch (reused)
    svo_track_io io;
    std::vector<float> io_prev2d, io_kps3d, io_ref2d, io_kfstate, io_kps2d;
    std::vector<int> io_kfid, io_kpidx, io_inl, io_outl, io_kltit;
    std::vector<uint8_t> io_kltst;
    long long counters[8] = {0};
    // since creation: frames, tracking frames, keyframes, alignment patches (keypoint x evaluation), LK windows (keypoint x level x
    // iteration), keypoints summed over tracking frames, alignment evaluations, refinement evaluations
    long long totals[8] = {0};
    std::vector<uint8_t> io_flags;
    bool pending = false;
    bool pending_first = false;
    int last_keyframe_created = 0;
    long long dropped_for_capacity = 0;  // new keypoints a keyframe could not take (device keypoint block full)
    double prof_ms[6] = {0, 0, 0, 0, 0, 0};   // SVO_HOST_PROFILE: pack, svo_frame_begin, wait+unpack (svo_track_frame_end), write-back, keyframe, motion filter
    float last_gpu_ms = 0;
    int last_launches = 0;

    int fail(int rc)
    {
        const char *m = svo_last_error(ctx);
        static const char *const generic[] = {"ok", "invalid argument", "CUDA error", "no CUDA device", "capacity exceeded", "call order violated"};
        snprintf(err, sizeof(err), "%s", (m && m[0]) ? m : generic[rc >= 0 && rc <= 5 ? rc : 1]);
        return rc;
    }

    // ---- DepthCalculator host parts -------------------------------------------------------------------------
    // depth_calculator.cpp:67-86
    void find_bad_keypoints(FrameH &f)
    {
        KeyPoints o;
        for (size_t i = 0; i < f.kps.size(); i++) {
            float x = f.kps.kps2d[2 * i], y = f.kps.kps2d[2 * i + 1];
            const svo_keypoint_info &in = f.kps.info[i];
            if (host_is_bad(x, y, W, H, in.ignore_completely, in.ignore_during_refinement)) continue;
            o.kps2d.push_back(x); o.kps2d.push_back(y);
            for (int k = 0; k < 3; k++) o.kps3d.push_back(f.kps.kps3d[3 * i + k]);
            o.info.push_back(in);
        }
        f.kps = std::move(o);
    }

    // depth_calculator.cpp:11-65: per-level detection on the device, cross-level choice by list index on the host
    int detect_and_select(int slot, std::vector<float> &kps, std::vector<svo_keypoint_info> &info)
    {
        int gw = cs.grid_width, gh = cs.grid_height;
        const int nl = cs.max_pyramid_levels / 2;  // `left.size()/2`
        kps.clear(); info.clear();
        for (int lv = 0; lv < nl; lv++) {
            if (gw < 1 || gh < 1) break;
            int lw, lh;
            svo_slot_level_size(ctx, 0, lv, &lw, &lh);
            int cap = (lw / gw + 1) * (lh / gh + 1) + 4;
            std::vector<float> xy((size_t)cap * 2), sc(cap);
            std::vector<int> ty(cap);
            int n = 0;
            int rc = svo_detect_keypoints(ctx, slot, lv, gw, gh, cap, xy.data(), sc.data(), ty.data(), &n);
            if (rc) return rc;
            host_select_level(lv, n, xy.data(), sc.data(), ty.data(), kps, info);
            gw /= 2; gh /= 2;
        }
        return SVO_OK;
    }

    // depth_calculator.cpp:88-130 — called with (grid_width, grid_height) bound to (grid_height, grid_width): Q6
    void merge_keypoints(FrameH &f, const std::vector<float> &nk, const std::vector<svo_keypoint_info> &ni, int grid_height, int grid_width)
    {
        std::vector<int> app;
        host_merge(W, H, grid_width, grid_height, f.kps.kps2d.data(), f.kps.size(), nk.data(), ni.size(), app);
        for (int i : app) {
            f.kps.kps2d.push_back(nk[2 * i]); f.kps.kps2d.push_back(nk[2 * i + 1]);
            f.kps.info.push_back(ni[i]);
        }
    }

    // DepthCalculator::calculate_depth (depth_calculator.cpp:132-392) + KeyFrameManager::create_keyframe (keyframe_manager.cpp:15-32)
    int create_keyframe(FrameH &f)
    {
        const float fx = cs.fx, fy = cs.fy, cx = cs.cx, cy = cs.cy, baseline = cs.baseline;
        const double t0 = g_trace_kf ? now_ms() : 0;
        find_bad_keypoints(f);
        std::vector<float> nk;
        std::vector<svo_keypoint_info> ni;
        int rc = detect_and_select(f.slot, nk, ni);
        if (rc) return rc;
        const double t1 = g_trace_kf ? now_ms() : 0;
        const size_t old_count = f.kps.size();
        merge_keypoints(f, nk, ni, cs.grid_width, cs.grid_height);
        // The reference's lists are unbounded; the device keypoint block is sized at context creation (4 keypoints per grid
        // cell by default).  A keyframe that would exceed it keeps its best-scored NEW keypoints only — tracking goes on
        // instead of every later frame failing with SVO_ERR_CAPACITY (never reached in any tested sequence; counted).
        int cap = 0;
        svo_keypoint_capacity(ctx, &cap);
        if (cap > 0 && f.kps.size() > (size_t)cap) {
            const size_t keep_new = (size_t)cap > old_count ? (size_t)cap - old_count : 0;
            std::vector<size_t> order(f.kps.size() - old_count);
            for (size_t i = 0; i < order.size(); i++) order[i] = old_count + i;
            std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return f.kps.info[a].score > f.kps.info[b].score; });
            order.resize(keep_new);
            std::sort(order.begin(), order.end());
            KeyPoints kept;
            kept.kps2d.assign(f.kps.kps2d.begin(), f.kps.kps2d.begin() + 2 * old_count);
            kept.info.assign(f.kps.info.begin(), f.kps.info.begin() + old_count);
            for (size_t i : order) {
                kept.kps2d.push_back(f.kps.kps2d[2 * i]); kept.kps2d.push_back(f.kps.kps2d[2 * i + 1]);
                kept.info.push_back(f.kps.info[i]);
            }
            kept.kps3d = f.kps.kps3d;
            dropped_for_capacity += (long long)(f.kps.size() - kept.size());
            f.kps = std::move(kept);
        }
        const double t2 = g_trace_kf ? now_ms() : 0;
        const size_t n = f.kps.size();
        f.kps.kps3d.resize(n * 3);
        const size_t n_new = n - old_count;
        std::vector<float> disp(n_new);
        if (n_new > 0) {
            rc = svo_stereo_match(ctx, f.slot, &f.kps.kps2d[2 * old_count], (int)n_new, 0, disp.data());
            if (rc) return rc;
        }
        const double t3 = g_trace_kf ? now_ms() : 0;
        float R[9];
        rodrigues_f(&f.pose[3], R);
        const uint64_t kf_id = keyframes.size();
        for (size_t i = old_count; i < n; i++) {
            const float kx = f.kps.kps2d[2 * i], ky = f.kps.kps2d[2 * i + 1];
            const float disparity = disp[i - old_count];
            float _z = baseline / std::max<float>(0.5, disparity);
            float _x = (kx - cx) / fx * _z;
            float _y = (ky - cy) / fy * _z;
            float loc[3] = {_x, _y, _z}, w3[3];
            m33v(R, loc, w3);
            for (int k = 0; k < 3; k++) f.kps.kps3d[3 * i + k] = w3[k] + f.pose[k];
            svo_keypoint_info &in = f.kps.info[i];
            rng = rng * 1664525u + 1013904223u;
            in.color[0] = (rng >> 8) & 0xFF; in.color[1] = (rng >> 16) & 0xFF; in.color[2] = (rng >> 24) & 0xFF;
            in.keyframe_id = kf_id;
            in.keypoint_index = i;
            in.ignore_completely = 0; in.ignore_temporary = 1; in.ignore_during_refinement = 0;
            in.inlier_count = 0; in.outlier_count = 0;
            float deviation = (float)(0.5 / (baseline / fx));  // depth_calculator.cpp:284-289
            in.kf_variance = deviation * deviation;
            in.kf_inv_depth = 1 / _z;
        }
        std::unique_ptr<FrameH> k(new FrameH());
        k->id = kf_id;
        std::memcpy(k->pose, f.pose, sizeof(f.pose));
        k->slot = f.slot;
        k->kps = f.kps;
        k->time_stamp = f.time_stamp;
        int id = -1;
        rc = svo_keyframe_commit(ctx, f.slot, f.pose, &id);  // the device copies the image set into a slot of the keyframe
        if (rc) return rc;
        if ((uint64_t)id != kf_id) { snprintf(err, sizeof(err), "keyframe id mismatch"); return SVO_ERR_STATE; }
        rc = svo_keyframe_slot(ctx, id, &k->slot);  // the device keeps its own copy of the keyframe's images
        if (rc) return rc;
        // the LK templates of the keypoints this keyframe introduced never change: build them once (optical_flow.cpp:41-44)
        if (n_new > 0 && (rc = svo_keyframe_set_templates(ctx, id, &f.kps.kps2d[2 * old_count], (int)old_count, (int)n_new))) return rc;
        keyframes.push_back(std::move(k));
        last_keyframe_created = 1;
        if (g_trace_kf)
            fprintf(stderr, "[kf] detect %.3f merge %.3f match %.3f init+commit %.3f ms (n_old %zu n_new %zu)\n", t1 - t0, t2 - t1, t3 - t2,
                    now_ms() - t3, old_count, n_new);
        return SVO_OK;
    }

    // keyframe_manager.cpp:47-74
    bool keyframe_needed(const FrameH &f) const
    {
        const uint8_t *ic = f.kps.info.empty() ? nullptr : &f.kps.info[0].ignore_completely;
        return host_keyframe_needed(W, H, cs.grid_width, cs.grid_height, f.kps.kps2d.data(), ic, sizeof(svo_keypoint_info), f.kps.size());
    }

    // StereoSlam::update_pose (stereo_slam.cpp:296-359)
    void update_pose(const float pose[6], const float speed[6], const float pv[6], const float sv[6], double dt, float out[6])
    {
        for (int i = 0; i < 6; i++) kf.A[i * 12 + (i + 6)] = (float)dt;
        kf.predict();
        for (int i = 0; i < 6; i++) { kf.R[i * 12 + i] = pv[i]; kf.R[(i + 6) * 12 + (i + 6)] = sv[i]; }
        float z[12];
        for (int i = 0; i < 6; i++) { z[i] = pose[i]; z[i + 6] = speed[i]; }
        kf.correct(z);
        for (int i = 0; i < 6; i++) out[i] = kf.xpost[i];
    }

    void finish_frame()
    {
        // stereo_slam.cpp:250-270
        if (previous) {
            double dt = frame->time_stamp - previous->time_stamp;
            for (int i = 0; i < 6; i++) motion[i] = (float)((double)(frame->pose[i] - previous->pose[i]) * (1. / dt));
            float pv[6] = {0.1f, 0.1f, 0.1f, 0.1f, 0.1f, 0.1f}, mv[6] = {1, 1, 1, 1, 1, 1}, fp[6];
            update_pose(frame->pose, motion, pv, mv, 0.0, fp);
            std::memcpy(frame->pose, fp, sizeof(fp));
            svo_slot_release(ctx, previous->slot);
            previous.reset();
        }
        svo_pose p = {frame->pose[0], frame->pose[1], frame->pose[2], frame->pose[3], frame->pose[4], frame->pose[5]};
        trajectory.push_back(p);
        totals[0]++;
        totals[2] += last_keyframe_created;
    }

    // Nothing of the instance's state changes unless the frame was enqueued: a failed call (bad stride, capacity, CUDA
    // error) leaves `frame` the current frame and the sequence can go on with the next image.
    int new_image_begin(const uint8_t *left, size_t ls, const uint8_t *right, size_t rs, float ts, bool on_device = false)
    {
        if (pending) { snprintf(err, sizeof(err), "new_image_begin called twice"); return SVO_ERR_STATE; }
        if (ls < (size_t)W || rs < (size_t)W) { snprintf(err, sizeof(err), "row stride smaller than the image width %d", W); return SVO_ERR_INVALID; }
        last_keyframe_created = 0;
        std::unique_ptr<FrameH> nf(new FrameH());
        nf->time_stamp = ts;
        int rc = SVO_OK;
        if (!frame) {  // first frame (stereo_slam.cpp:142-160): pyramids only, keyframe creation in _end
            rc = on_device ? svo_upload_stereo_device(ctx, left, ls, right, rs, &nf->slot)
                           : svo_upload_stereo(ctx, left, ls, right, rs, &nf->slot);  // pyramids (stereo_slam.cpp:135-139)
            if (rc) return fail(rc);
            nf->id = 0;
            frame = std::move(nf);
            pending = true;
            pending_first = true;
            return SVO_OK;
        }
        nf->id = frame->id + 1;
        for (int i = 0; i < 6; i++) nf->pose[i] = kf.xpre[i];  // kf.statePre (Q8)
        // remove_outliers (stereo_slam.cpp:43-56) on what is about to become the previous frame (idempotent)
        {
            KeyPoints u;
            for (size_t i = 0; i < frame->kps.size(); i++) {
                if (frame->kps.info[i].ignore_completely) continue;
                u.kps2d.push_back(frame->kps.kps2d[2 * i]); u.kps2d.push_back(frame->kps.kps2d[2 * i + 1]);
                for (int k = 0; k < 3; k++) u.kps3d.push_back(frame->kps.kps3d[3 * i + k]);
                u.info.push_back(frame->kps.info[i]);
            }
            frame->kps = std::move(u);
        }
        const double tp0 = g_host_prof ? now_ms() : 0;
        const size_t n = frame->kps.size();
        io_prev2d = frame->kps.kps2d;
        io_kps3d = frame->kps.kps3d;
        io_ref2d.resize(n * 2); io_kfid.resize(n); io_kpidx.resize(n); io_flags.resize(n); io_inl.resize(n); io_outl.resize(n); io_kfstate.resize(n * 2);
        io_kps2d.resize(n * 2); io_kltit.assign(n, 0); io_kltst.assign(n, 0);
        for (size_t i = 0; i < n; i++) {
            const svo_keypoint_info &in = frame->kps.info[i];
            const FrameH &k = *keyframes[in.keyframe_id];
            io_ref2d[2 * i] = k.kps.kps2d[2 * in.keypoint_index];
            io_ref2d[2 * i + 1] = k.kps.kps2d[2 * in.keypoint_index + 1];
            io_kfid[i] = (int)in.keyframe_id;
            io_kpidx[i] = (int)in.keypoint_index;
            io_flags[i] = (uint8_t)((in.ignore_during_refinement ? 1 : 0) | (in.ignore_completely ? 2 : 0) | (in.ignore_temporary ? 4 : 0));
            io_inl[i] = in.inlier_count; io_outl[i] = in.outlier_count;
            io_kfstate[2 * i] = in.kf_inv_depth; io_kfstate[2 * i + 1] = in.kf_variance;
        }
        std::memset(&io, 0, sizeof(io));
        io.n = (int)n;
        io.prev_kps2d = io_prev2d.data(); io.kps3d = io_kps3d.data(); io.ref_kps2d = io_ref2d.data(); io.keyframe_id = io_kfid.data();
        io.flags = io_flags.data(); io.inlier_count = io_inl.data(); io.outlier_count = io_outl.data(); io.kf_state = io_kfstate.data();
        io.kps2d = io_kps2d.data();
        io.klt_iters = io_kltit.data(); io.klt_status = io_kltst.data();
        io.keypoint_index = io_kpidx.data();
        std::memcpy(io.pose_prior, nf->pose, sizeof(io.pose_prior));
        // pyramids (stereo_slam.cpp:135-139) + the whole tracking sequence, one CUDA graph launch in steady state
        const double tp1 = g_host_prof ? now_ms() : 0;
        rc = svo_frame_begin(ctx, left, ls, right, rs, on_device ? 1 : 0, frame->slot, &io, &nf->slot);
        if (g_host_prof) { prof_ms[0] += tp1 - tp0; prof_ms[1] += now_ms() - tp1; }
        if (rc) return fail(rc);
        previous = std::move(frame);
        frame = std::move(nf);
        pending = true;
        pending_first = false;
        return SVO_OK;
    }

    int new_image_end()
    {
        if (!pending) { snprintf(err, sizeof(err), "new_image_end without begin"); return SVO_ERR_STATE; }
        pending = false;
        int rc;
        if (pending_first) {
            rc = create_keyframe(*frame);
            if (rc) {  // no first keyframe: the next image starts the sequence again
                if (rc != SVO_ERR_STATE) fail(rc);
                svo_slot_release(ctx, frame->slot);
                frame.reset();
                return rc;
            }
            for (auto &in : frame->kps.info) in.ignore_temporary = 0;  // stereo_slam.cpp:157-159
            last_gpu_ms = 0; last_launches = 0;
            finish_frame();
            return SVO_OK;
        }
        const double te0 = g_host_prof ? now_ms() : 0;
        rc = svo_track_frame_end(ctx, &io);
        if (rc) return fail(rc);
        svo_last_track_timing(ctx, &last_gpu_ms, &last_launches);
        const double te1 = g_host_prof ? now_ms() : 0;
        const size_t n = (size_t)io.n;
        std::memcpy(frame->pose, io.pose_refined, sizeof(frame->pose));
        {
            long long used = 0, ce = 0, ge = 0, it = 0, tr = 0;
            for (size_t i = 0; i < n; i++) {
                if (!previous->kps.info[i].ignore_temporary) used++;
                it += io_kltit[i]; tr += io_kltst[i] ? 1 : 0;
            }
            for (int l = 0; l < 8; l++) { ce += io.align_evals[2 * l]; ge += io.align_evals[2 * l + 1]; }
            counters[0] = (long long)n; counters[1] = used; counters[2] = ce; counters[3] = ge;
            counters[4] = io.refine_evals[0]; counters[5] = io.refine_evals[1]; counters[6] = it; counters[7] = tr;
            totals[1]++; totals[3] += used * (ce + ge); totals[4] += it; totals[5] += (long long)n; totals[6] += ce + ge;
            totals[7] += io.refine_evals[0] + io.refine_evals[1];
        }
        frame->kps.info = previous->kps.info;
        frame->kps.kps3d = io_kps3d;
        frame->kps.kps2d = io_kps2d;
        // flags / counters / filter state come back from the device; write back into the origin keyframes
        // (stereo_slam.cpp:205-226)
        for (size_t i = 0; i < n; i++) {
            svo_keypoint_info &in = frame->kps.info[i];
            in.ignore_during_refinement = (io_flags[i] & 1) ? 1 : 0;
            in.ignore_completely = (io_flags[i] & 2) ? 1 : 0;
            in.ignore_temporary = (io_flags[i] & 4) ? 1 : 0;
            in.inlier_count = io_inl[i]; in.outlier_count = io_outl[i];
            in.kf_inv_depth = io_kfstate[2 * i]; in.kf_variance = io_kfstate[2 * i + 1];
            FrameH &k = *keyframes[in.keyframe_id];
            for (int q = 0; q < 3; q++) k.kps.kps3d[3 * in.keypoint_index + q] = io_kps3d[3 * i + q];
            svo_keypoint_info &ki = k.kps.info[in.keypoint_index];
            ki.ignore_temporary = in.ignore_temporary; ki.ignore_completely = in.ignore_completely;
            ki.inlier_count = in.inlier_count; ki.outlier_count = in.outlier_count;
            ki.kf_inv_depth = in.kf_inv_depth; ki.kf_variance = in.kf_variance;  // cv::KalmanFilter copies share state (Q13)
        }
        const double te2 = g_host_prof ? now_ms() : 0;
        if (g_host_prof) { prof_ms[2] += te1 - te0; prof_ms[3] += te2 - te1; }
        if (keyframe_needed(*frame)) {  // stereo_slam.cpp:231-246
            rc = create_keyframe(*frame);
            if (rc) return rc == SVO_ERR_STATE ? rc : fail(rc);
            size_t c = 0;
            for (auto &in : frame->kps.info) if (!in.ignore_temporary) c++;
            if (c < frame->kps.info.size() / 4)
                for (auto &in : frame->kps.info) in.ignore_temporary = 0;
        }
        const double te3 = g_host_prof ? now_ms() : 0;
        finish_frame();
        if (g_host_prof) { prof_ms[4] += te3 - te2; prof_ms[5] += now_ms() - te3; }
        return SVO_OK;
    }
};

// ================================================================================================ C exports
extern "C" {

static char g_slam_create_err[256] = "";

int svo_slam_create_with_capacity(const svo_camera_settings *settings, int device, int width, int height, int max_keypoints, svo_slam **out)
{
    if (!settings || !out) return SVO_ERR_INVALID;
    svo_slam *s = new svo_slam();
    s->cs = *settings; s->W = width; s->H = height;
    int rc = svo_ctx_create(settings, device, width, height, max_keypoints, &s->ctx);
    if (rc) {
        snprintf(g_slam_create_err, sizeof(g_slam_create_err), "%s", svo_last_error(nullptr));
        delete s;
        return rc;
    }
    *out = s;
    return SVO_OK;
}

int svo_slam_create(const svo_camera_settings *settings, int device, int width, int height, svo_slam **out)
{
    return svo_slam_create_with_capacity(settings, device, width, height, 0, out);
}

int svo_slam_destroy(svo_slam *s)
{
    if (!s) return SVO_ERR_INVALID;
    if (g_host_prof && s->totals[1] > 0) {
        const double f = 1e3 / (double)s->totals[1];
        fprintf(stderr, "[host profile] %lld tracking frames, us per frame: pack %.2f frame_begin %.2f wait+unpack %.2f write-back %.2f keyframe %.2f motion filter %.2f\n",
                s->totals[1], s->prof_ms[0] * f, s->prof_ms[1] * f, s->prof_ms[2] * f, s->prof_ms[3] * f, s->prof_ms[4] * f, s->prof_ms[5] * f);
    }
    if (s->ctx) svo_ctx_destroy(s->ctx);
    delete s;
    return SVO_OK;
}

const char *svo_slam_last_error(svo_slam *s) { return s ? s->err : g_slam_create_err; }
svo_ctx *svo_slam_ctx(svo_slam *s) { return s ? s->ctx : nullptr; }

int svo_slam_new_image_begin(svo_slam *s, const uint8_t *left, size_t ls, const uint8_t *right, size_t rs, float ts)
{
    if (!s || !left || !right) return SVO_ERR_INVALID;
    return s->new_image_begin(left, ls, right, rs, ts);
}
int svo_slam_new_image_device_begin(svo_slam *s, const uint8_t *left, size_t ls, const uint8_t *right, size_t rs, float ts)
{
    if (!s || !left || !right) return SVO_ERR_INVALID;
    return s->new_image_begin(left, ls, right, rs, ts, true);
}
int svo_slam_new_image_end(svo_slam *s)
{
    if (!s) return SVO_ERR_INVALID;
    return s->new_image_end();
}
int svo_slam_new_image(svo_slam *s, const uint8_t *left, size_t ls, const uint8_t *right, size_t rs, float ts)
{
    int rc = svo_slam_new_image_begin(s, left, ls, right, rs, ts);
    if (rc) return rc;
    return svo_slam_new_image_end(s);
}

int svo_slam_set_rectification(svo_slam *s, int which, const double K[9], const double D[5], const double R[9], const double P[9])
{
    if (!s) return SVO_ERR_INVALID;
    if (s->pending) { snprintf(s->err, sizeof(s->err), "set_rectification while a frame is in flight"); return SVO_ERR_STATE; }
    int rc = svo_set_rectification(s->ctx, which, K, D, R, P);
    return rc ? s->fail(rc) : SVO_OK;
}

static int frame_meta(const FrameH *f, uint64_t *id, svo_pose *pose, double *ts, int *n)
{
    if (!f) return SVO_ERR_STATE;
    if (id) *id = f->id;
    if (pose) { pose->x = f->pose[0]; pose->y = f->pose[1]; pose->z = f->pose[2]; pose->rx = f->pose[3]; pose->ry = f->pose[4]; pose->rz = f->pose[5]; }
    if (ts) *ts = f->time_stamp;
    if (n) *n = (int)f->kps.size();
    return SVO_OK;
}
static int frame_kps(const FrameH *f, int max, float *k2, float *k3, svo_keypoint_info *info)
{
    if (!f) return SVO_ERR_STATE;
    size_t n = std::min<size_t>(f->kps.size(), (size_t)std::max(0, max));
    if (k2) std::memcpy(k2, f->kps.kps2d.data(), n * 8);
    if (k3) std::memcpy(k3, f->kps.kps3d.data(), n * 12);
    if (info) std::memcpy(info, f->kps.info.data(), n * sizeof(svo_keypoint_info));
    return SVO_OK;
}

int svo_slam_get_frame(svo_slam *s, uint64_t *id, svo_pose *pose, double *ts, int *n)
{
    if (!s) return SVO_ERR_INVALID;
    return frame_meta(s->frame.get(), id, pose, ts, n);
}
int svo_slam_get_frame_keypoints(svo_slam *s, int max, float *k2, float *k3, svo_keypoint_info *info)
{
    if (!s) return SVO_ERR_INVALID;
    return frame_kps(s->frame.get(), max, k2, k3, info);
}
int svo_slam_get_frame_image(svo_slam *s, int kind, int level, uint8_t *out, size_t stride)
{
    if (!s || !s->frame) return SVO_ERR_STATE;
    return svo_download_level(s->ctx, s->frame->slot, kind, level, out, stride);
}
int svo_slam_keyframe_count(svo_slam *s) { return s ? (int)s->keyframes.size() : 0; }
static const FrameH *kf_at(svo_slam *s, int index)
{
    if (!s || s->keyframes.empty()) return nullptr;
    if (index < 0) index = (int)s->keyframes.size() - 1;
    if (index >= (int)s->keyframes.size()) return nullptr;
    return s->keyframes[index].get();
}
int svo_slam_get_keyframe(svo_slam *s, int index, uint64_t *id, svo_pose *pose, double *ts, int *n)
{
    if (!s) return SVO_ERR_INVALID;
    return frame_meta(kf_at(s, index), id, pose, ts, n);
}
int svo_slam_get_keyframe_keypoints(svo_slam *s, int index, int max, float *k2, float *k3, svo_keypoint_info *info)
{
    if (!s) return SVO_ERR_INVALID;
    return frame_kps(kf_at(s, index), max, k2, k3, info);
}
int svo_slam_get_keyframe_image(svo_slam *s, int index, int kind, int level, uint8_t *out, size_t stride)
{
    const FrameH *k = kf_at(s, index);
    if (!k) return SVO_ERR_STATE;
    return svo_download_level(s->ctx, k->slot, kind, level, out, stride);
}
int svo_slam_get_trajectory(svo_slam *s, int max, svo_pose *out)
{
    if (!s) return 0;
    size_t n = std::min<size_t>(s->trajectory.size(), (size_t)std::max(0, max));
    if (out && n) std::memcpy(out, s->trajectory.data(), n * sizeof(svo_pose));
    return (int)s->trajectory.size();
}
int svo_slam_update_pose(svo_slam *s, const svo_pose *pose, const float speed[6], const float pv[6], const float sv[6], double dt,
                         svo_pose *filtered)
{
    if (!s || !pose || !speed || !pv || !sv) return SVO_ERR_INVALID;
    float p[6] = {pose->x, pose->y, pose->z, pose->rx, pose->ry, pose->rz}, o[6];
    s->update_pose(p, speed, pv, sv, dt, o);
    if (filtered) { filtered->x = o[0]; filtered->y = o[1]; filtered->z = o[2]; filtered->rx = o[3]; filtered->ry = o[4]; filtered->rz = o[5]; }
    return SVO_OK;
}
// ---- host stages on their own (no device, no context): the bookkeeping of DepthCalculator / KeyFrameManager / StereoSlam
// that stays on the CPU, exported so that it can be checked against the reference without a GPU ---------------------------
int svo_host_select_best_keypoints(int n_levels, const int *n_per_level, const float *const *xy, const float *const *score,
                                   const int *const *type, int max_out, float *kps2d, float *score_out, int *type_out, int *level_out,
                                   int *n_out)
{
    if (n_levels < 1 || !n_per_level || !xy || !score || !type || !n_out) return SVO_ERR_INVALID;
    std::vector<float> kps;
    std::vector<svo_keypoint_info> info;
    for (int lv = 0; lv < n_levels; lv++) {
        if (n_per_level[lv] < 0 || (n_per_level[lv] > 0 && (!xy[lv] || !score[lv] || !type[lv]))) return SVO_ERR_INVALID;
        host_select_level(lv, n_per_level[lv], xy[lv], score[lv], type[lv], kps, info);
    }
    *n_out = (int)info.size();
    const size_t n = std::min<size_t>(info.size(), (size_t)std::max(0, max_out));
    for (size_t j = 0; j < n; j++) {
        if (kps2d) { kps2d[2 * j] = kps[2 * j]; kps2d[2 * j + 1] = kps[2 * j + 1]; }
        if (score_out) score_out[j] = info[j].score;
        if (type_out) type_out[j] = info[j].type;
        if (level_out) level_out[j] = info[j].level;
    }
    return SVO_OK;
}
int svo_host_find_bad_keypoints(int width, int height, int n, const float *kps2d, const uint8_t *flags, uint8_t *keep)
{
    if (n < 0 || (n > 0 && (!kps2d || !flags || !keep))) return SVO_ERR_INVALID;
    for (int i = 0; i < n; i++)
        keep[i] = host_is_bad(kps2d[2 * i], kps2d[2 * i + 1], width, height, flags[i] & SVO_FLAG_IGNORE_COMPLETELY, flags[i] & SVO_FLAG_IGNORE_REFINEMENT) ? 0 : 1;
    return SVO_OK;
}
int svo_host_merge_keypoints(int width, int height, int grid_width, int grid_height, int n_old, const float *old_kps2d, int n_new,
                             const float *new_kps2d, int max_out, int *appended, int *n_out)
{
    if (width < 1 || height < 1 || grid_width < 1 || grid_height < 1 || n_old < 0 || n_new < 0 || !n_out || (n_old > 0 && !old_kps2d) ||
        (n_new > 0 && !new_kps2d))
        return SVO_ERR_INVALID;
    std::vector<int> app;
    // calculate_depth hands (grid_width, grid_height) to parameters declared (grid_height, grid_width): cells are
    // grid_height wide and grid_width tall (depth_calculator.cpp:88-90, :179-180)
    host_merge(width, height, grid_height, grid_width, old_kps2d, (size_t)n_old, new_kps2d, (size_t)n_new, app);
    *n_out = (int)app.size();
    for (size_t i = 0; i < app.size() && (int)i < max_out; i++) appended[i] = app[i];
    return SVO_OK;
}
int svo_host_keyframe_needed(int width, int height, int grid_width, int grid_height, int n, const float *kps2d, const uint8_t *flags, int *needed)
{
    if (grid_width < 1 || grid_height < 1 || n < 0 || !needed || (n > 0 && (!kps2d || !flags))) return SVO_ERR_INVALID;
    std::vector<uint8_t> ic((size_t)n);
    for (int i = 0; i < n; i++) ic[i] = (flags[i] & SVO_FLAG_IGNORE_COMPLETELY) ? 1 : 0;
    *needed = host_keyframe_needed(width, height, grid_width, grid_height, kps2d, ic.data(), 1, (size_t)n) ? 1 : 0;
    return SVO_OK;
}
struct svo_motion_filter { MotionFilter f; };
int svo_motion_filter_create(svo_motion_filter **out)
{
    if (!out) return SVO_ERR_INVALID;
    *out = new svo_motion_filter();
    return SVO_OK;
}
int svo_motion_filter_destroy(svo_motion_filter *m) { if (!m) return SVO_ERR_INVALID; delete m; return SVO_OK; }
int svo_motion_filter_update(svo_motion_filter *m, const svo_pose *pose, const float speed[6], const float pv[6], const float sv[6], double dt,
                             svo_pose *filtered, float state_pre12[12])
{
    if (!m || !pose || !speed || !pv || !sv) return SVO_ERR_INVALID;
    MotionFilter &kf = m->f;
    // StereoSlam::update_pose (stereo_slam.cpp:296-359)
    for (int i = 0; i < 6; i++) kf.A[i * 12 + (i + 6)] = (float)dt;
    kf.predict();
    for (int i = 0; i < 6; i++) { kf.R[i * 12 + i] = pv[i]; kf.R[(i + 6) * 12 + (i + 6)] = sv[i]; }
    const float z[12] = {pose->x, pose->y, pose->z, pose->rx, pose->ry, pose->rz, speed[0], speed[1], speed[2], speed[3], speed[4], speed[5]};
    kf.correct(z);
    if (filtered) { filtered->x = kf.xpost[0]; filtered->y = kf.xpost[1]; filtered->z = kf.xpost[2]; filtered->rx = kf.xpost[3]; filtered->ry = kf.xpost[4]; filtered->rz = kf.xpost[5]; }
    if (state_pre12) std::memcpy(state_pre12, kf.xpre, sizeof(kf.xpre));
    return SVO_OK;
}

int svo_slam_last_counters(svo_slam *s, long long *out8)
{
    if (!s || !out8) return SVO_ERR_INVALID;
    for (int k = 0; k < 8; k++) out8[k] = s->counters[k];
    return SVO_OK;
}
int svo_slam_total_counters(svo_slam *s, long long *out8)
{
    if (!s || !out8) return SVO_ERR_INVALID;
    for (int k = 0; k < 8; k++) out8[k] = s->totals[k];
    return SVO_OK;
}

// Many independent sequences, many frames each, one call: worker thread t owns sequences t, t + workers, ... and walks them
// through their frames on its own (begin on all of its sequences, then end on all of them, frame after frame) — there is no
// barrier between frames or between workers, so the GPU always has other sequences' work queued while one finishes.
int svo_slam_reset(svo_slam *s)
{
    if (!s) return SVO_ERR_INVALID;
    if (s->pending) { snprintf(s->err, sizeof(s->err), "svo_slam_reset while a frame is in flight"); return SVO_ERR_STATE; }
    int rc = svo_ctx_reset(s->ctx);
    if (rc) return s->fail(rc);
    s->keyframes.clear();
    s->frame.reset();
    s->previous.reset();
    s->trajectory.clear();
    for (int i = 0; i < 6; i++) s->motion[i] = 0;
    s->kf = MotionFilter();
    s->last_keyframe_created = 0;
    return SVO_OK;   // totals[] keep counting: they describe the work of the handle, not of one sequence
}

int svo_slam_run_many(svo_slam *const *slams, int n, int n_frames, const uint8_t *const *left, const uint8_t *const *right,
                      size_t left_stride, size_t right_stride, const float *time_stamps, int on_device, int workers, int *failed_sequence)
{
    return svo_slam_run_many_restart(slams, n, n_frames, left, right, left_stride, right_stride, time_stamps, nullptr, on_device, workers,
                                     failed_sequence);
}

int svo_slam_run_many_restart(svo_slam *const *slams, int n, int n_frames, const uint8_t *const *left, const uint8_t *const *right,
                              size_t left_stride, size_t right_stride, const float *time_stamps, const uint8_t *restart, int on_device,
                              int workers, int *failed_sequence)
{
    if (!slams || n < 1 || n_frames < 0 || !left || !right || !time_stamps) return SVO_ERR_INVALID;
    for (int i = 0; i < n; i++) if (!slams[i]) return SVO_ERR_INVALID;
    if (workers < 1) workers = 1;
    if (workers > n) workers = n;
    std::atomic<int> rc_first{SVO_OK}, bad{-1};
    // A worker keeps every one of its sequences busy on the device: as soon as frame f of a sequence is finished (results
    // unpacked, bookkeeping done) frame f + 1 of the SAME sequence is enqueued, then the worker turns to its next sequence —
    // at any time all but one of a worker's sequences have a frame queued.
    auto work = [&](int t) {
        int rc = SVO_OK, who = -1;
        auto begin = [&](int i, int f) {
            const size_t k = (size_t)i * n_frames + f;
            if (restart && restart[k] && (rc = svo_slam_reset(slams[i]))) { who = i; return; }
            rc = slams[i]->new_image_begin(left[k], left_stride, right[k], right_stride, time_stamps[k], on_device != 0);
            if (rc) who = i;
        };
        static const bool phased = getenv("SVO_RUN_MANY_PHASED") != nullptr;   // developer switch: begin all, then end all, per frame
        if (phased) {
            for (int f = 0; f < n_frames && rc == SVO_OK; f++) {
                for (int i = t; i < n && rc == SVO_OK; i += workers) begin(i, f);
                for (int i = t; i < n; i += workers) {
                    if (!slams[i]->pending) continue;
                    const int r2 = slams[i]->new_image_end();
                    if (r2 && rc == SVO_OK) { rc = r2; who = i; }
                }
            }
        }
        if (!phased && n_frames > 0)
            for (int i = t; i < n && rc == SVO_OK; i += workers) begin(i, 0);
        for (int f = 0; !phased && f < n_frames && rc == SVO_OK; f++) {
            if (rc_first.load(std::memory_order_relaxed) != SVO_OK) break;
            for (int i = t; i < n && rc == SVO_OK; i += workers) {
                rc = slams[i]->new_image_end();
                if (rc) { who = i; break; }
                if (f + 1 < n_frames) begin(i, f + 1);
            }
        }
        // every sequence whose frame was enqueued is finished, also after a failure (of this worker or of another one)
        for (int i = t; i < n; i += workers) {
            if (!slams[i]->pending) continue;
            const int r2 = slams[i]->new_image_end();
            if (r2 && rc == SVO_OK) { rc = r2; who = i; }
        }
        if (rc) {
            int expect = SVO_OK;
            if (rc_first.compare_exchange_strong(expect, rc)) bad.store(who);
        }
    };
    if (workers == 1) work(0);
    else {
        std::vector<std::thread> th;
        th.reserve(workers);
        for (int t = 0; t < workers; t++) th.emplace_back(work, t);
        for (auto &x : th) x.join();
    }
    if (failed_sequence) *failed_sequence = bad.load();
    return rc_first.load();
}

int svo_slam_dropped_keypoints(svo_slam *s, long long *dropped)
{
    if (!s || !dropped) return SVO_ERR_INVALID;
    *dropped = s->dropped_for_capacity;
    return SVO_OK;
}
int svo_slam_last_stats(svo_slam *s, float *gpu_ms, int *launches, int *keyframe_created)
{
    if (!s) return SVO_ERR_INVALID;
    if (gpu_ms) *gpu_ms = s->last_gpu_ms;
    if (launches) *launches = s->last_launches;
    if (keyframe_created) *keyframe_created = s->last_keyframe_created;
    return SVO_OK;
}
}  // extern "C"
