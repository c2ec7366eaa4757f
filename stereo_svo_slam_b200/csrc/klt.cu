// Pyramidal Lucas-Kanade tracking of keyframe keypoints into the current frame.
//
// Replaces OpticalFlow::calculate_optical_flow -> cv::calcOpticalFlowPyrLK (src/lib/optical_flow.cpp:14-56),
// called once per origin keyframe by PoseRefiner::refine_pose (src/lib/pose_refinement.cpp:102-118), and the
// gating that follows it (pose_refinement.cpp:125-150).  Arithmetic follows OpenCV's LKTrackerInvoker
// (video/src/lkpyramid.cpp, scalar path): Q14 bilinear weights (cvRound), Q5 window intensities, Scharr
// derivatives descaled by 2^14, float 2x2 system, 30 iterations / eps 0.01^2, level-0 error = mean |dI| / 32.
//
// B200 design: one CTA per keypoint, all three levels in one launch.  The previous-image window (plus a
// one-pixel Scharr halo) is staged in shared memory as u8; the Scharr derivative images that OpenCV
// materialises per level (4 B/px of HBM) are never stored: derivatives are computed on the fly from the
// staged tile with OpenCV's border rule (REFLECT_101 inside the level, 0 outside).  The window sums
// (A11, A12, A22, b1, b2, err) are accumulated as exact 64-bit integers — the products are integers in
// OpenCV too — and reduced with warp shuffles + one shared-memory pass, so the only rounding left is the
// final int64 -> float conversion (OpenCV rounds after every add; difference <= 1e-5 px, tolerance 0.01 px).
// Every thread redundantly performs the scalar 2x2 update from the reduced sums, which removes the broadcast
// barrier: one __syncthreads per LK iteration.
#include <cstdlib>
#include "kernels.cuh"

#define KLT_THREADS 256
#define KLT_MAXWIN 31
#define KLT_TILE (KLT_MAXWIN + 3)  // window + bilinear (+1) + Scharr halo (+-1)

struct KltShared {
    uint8_t tile[KLT_TILE * KLT_TILE + 4];
    short gx[(KLT_MAXWIN + 1) * (KLT_MAXWIN + 1)];   // Scharr d/dx at the (win+1)^2 bilinear support positions
    short gy[(KLT_MAXWIN + 1) * (KLT_MAXWIN + 1)];
    short Iw[KLT_MAXWIN * KLT_MAXWIN];
    short Ix[KLT_MAXWIN * KLT_MAXWIN];
    short Iy[KLT_MAXWIN * KLT_MAXWIN];
    long long part[2][3][KLT_THREADS / 32];
    float init[2];
};

__device__ __forceinline__ long long warp_sum_ll(long long v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}

// reduce three int64 per thread over the block; every thread returns the totals. buf alternates per call.
__device__ __forceinline__ void block_sum3(KltShared &sm, int buf, long long &a, long long &b, long long &c)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    a = warp_sum_ll(a); b = warp_sum_ll(b); c = warp_sum_ll(c);
    if (lane == 0) { sm.part[buf][0][warp] = a; sm.part[buf][1][warp] = b; sm.part[buf][2][warp] = c; }
    __syncthreads();
    a = b = c = 0;
#pragma unroll
    for (int w = 0; w < KLT_THREADS / 32; w++) { a += sm.part[buf][0][w]; b += sm.part[buf][1][w]; c += sm.part[buf][2][w]; }
}

__device__ __forceinline__ void lk_weights(float a, float b, int &iw00, int &iw01, int &iw10, int &iw11)
{
    iw00 = __float2int_rn((1.f - a) * (1.f - b) * (float)(1 << 14));
    iw01 = __float2int_rn(a * (1.f - b) * (float)(1 << 14));
    iw10 = __float2int_rn((1.f - a) * b * (float)(1 << 14));
    iw11 = (1 << 14) - iw00 - iw01 - iw10;
}

// calcOpticalFlowPyrLK's two stop tests, evaluated by OpenCV in double on float operands: `dx*dx + dy*dy <= 0.01*0.01` and
// `fabs(dx + prev_dx) < 0.01`.  The double-precision pipe answers in ~25 cycles per dependent operation and these tests close the
// loop-carried chain of every LK iteration, so they are decided in float wherever float can decide them exactly:
//  * |f| < 0.01 (double) for a float f  <=>  |f| <= 0x1.47ae14p-7f, the largest float below the double 0.01 — exact, always;
//  * the squared step is compared in float first; only a value within 1e-4 (relative) of the threshold — float rounding is 1e-7 —
//    goes through the reference's double expression.
__device__ __forceinline__ bool lk_step_converged(float dx, float dy)
{
    const float sf = __fmaf_rn(dx, dx, dy * dy);
    if (sf < 0.9999e-4f) return true;
    if (sf > 1.0001e-4f) return false;
    return (double)dx * (double)dx + (double)dy * (double)dy <= 0.01 * 0.01;
}
__device__ __forceinline__ bool lk_below_eps(float f) { return fabsf(f) <= 0x1.47ae14p-7f; }

template <int WIN>
__global__ void __launch_bounds__(KLT_THREADS) klt_pyr_lk_kernel(KltArgs a)
{
    __shared__ KltShared sm;
    const int i = blockIdx.x;
    const int n = min(*a.n_ptr, a.max_kps);
    if (i >= n) return;
    const int tid = threadIdx.x;
    const int win = WIN > 0 ? WIN : a.cam.win_flow;   // compile-time window => divisions by constants
    const int npx = win * win;
    const float half = (float)(win - 1) * 0.5f;

    // ---- initial guess: given, or project_keypoints(estimated pose) (stereo_slam.cpp:71-75)
    if (tid == 0) {
        float gx, gy;
        if (a.init_pts) { gx = a.init_pts[2 * i]; gy = a.init_pts[2 * i + 1]; }
        else {
            double Rd[9];
            const float *p = a.pose;
            if (a.pose_rd) { for (int k = 0; k < 9; k++) Rd[k] = a.pose_rd[k]; } else dev_rodrigues_d(-p[3], -p[4], -p[5], Rd);
            dev_project(Rd, a.kps3d[3 * i], a.kps3d[3 * i + 1], a.kps3d[3 * i + 2], p[0], p[1], p[2], a.cam.fx, a.cam.fy, a.cam.cx,
                        a.cam.cy, a.cam.k1, a.cam.k2, a.cam.p1, a.cam.p2, a.cam.k3, gx, gy);
        }
        sm.init[0] = gx; sm.init[1] = gy;
    }
    __syncthreads();
    const float init_x = sm.init[0], init_y = sm.init[1];
    const float ppx = a.prev_pts[2 * i], ppy = a.prev_pts[2 * i + 1];
    const LevelDesc *prev_lv = a.keyframe_ids ? (a.kf_lk_table + (size_t)a.keyframe_ids[i] * 2 * SVO_LK_LEVELS) : a.prev_fixed;

    float nx = init_x, ny = init_y;  // nextPts[i]
    int status = 1;
    float err = 0.f;
    int buf = 0;
    int total_iters = 0;

    for (int level = SVO_LK_LEVELS - 1; level >= 0; level--) {
        const LevelDesc I = prev_lv[level];
        const LevelDesc J = a.cur[level];
        const float scale = (float)(1. / (1 << level));
        float px = ppx * scale, py = ppy * scale;
        float qx, qy;
        if (level == SVO_LK_LEVELS - 1) { qx = nx * scale; qy = ny * scale; }
        else { qx = nx * 2.f; qy = ny * 2.f; }
        nx = qx; ny = qy;
        px -= half; py -= half;
        const int ipx = (int)floorf(px), ipy = (int)floorf(py);
        if (ipx < -win || ipx >= I.w || ipy < -win || ipy >= I.h) {
            if (level == 0) { status = 0; err = 0.f; }
            continue;
        }
        int iw00, iw01, iw10, iw11;
        lk_weights(px - (float)ipx, py - (float)ipy, iw00, iw01, iw10, iw11);

        // ---- stage the (win+3)^2 u8 tile: rows ipy-1 .. ipy+win+1, cols ipx-1 .. ipx+win+1 (padded level => no bounds tests)
        const int T = win + 3;
        __syncthreads();  // previous level's readers are done with tile/Iw/Ix/Iy
        for (int k = tid; k < T * T; k += KLT_THREADS) {
            int r = k / T, c = k - r * T;
            sm.tile[k] = I.ptr[(ptrdiff_t)(ipy - 1 + r) * I.pitch + (ipx - 1 + c)];
        }
        __syncthreads();
        // ---- Scharr derivatives at the (win+1)^2 support positions, OpenCV's border rule: REFLECT_101 inside the level
        //      (already in the padded tile), constant 0 outside the level
        const int W1 = win + 1;
        for (int k = tid; k < W1 * W1; k += KLT_THREADS) {
            int y = k / W1, x = k - y * W1;
            const int X = ipx + x, Y = ipy + y;
            int gxv = 0, gyv = 0;
            if (X >= 0 && X < I.w && Y >= 0 && Y < I.h) {
                const uint8_t *t = &sm.tile[(y + 1) * T + (x + 1)];
                int tl = t[-T - 1], tc = t[-T], trr = t[-T + 1], ml = t[-1], mr = t[1], bl = t[T - 1], bc = t[T], br = t[T + 1];
                gxv = 3 * (trr + br) + 10 * mr - (3 * (tl + bl) + 10 * ml);   // [3 10 3]^T x [-1 0 1]
                gyv = 3 * ((bl - tl) + (br - trr)) + 10 * (bc - tc);          // [-1 0 1]^T x [3 10 3]
            }
            sm.gx[k] = (short)gxv; sm.gy[k] = (short)gyv;
        }
        __syncthreads();
        // ---- window template: Iw (Q5), Ix, Iy, and the structure tensor
        long long s11 = 0, s12 = 0, s22 = 0;
        for (int k = tid; k < npx; k += KLT_THREADS) {
            int y = k / win, x = k - y * win;
            const uint8_t *t = &sm.tile[(y + 1) * T + (x + 1)];
            const short *px_ = &sm.gx[y * W1 + x], *py_ = &sm.gy[y * W1 + x];
            int ival = (int)t[0] * iw00 + (int)t[1] * iw01 + (int)t[T] * iw10 + (int)t[T + 1] * iw11;
            int ixv = (int)px_[0] * iw00 + (int)px_[1] * iw01 + (int)px_[W1] * iw10 + (int)px_[W1 + 1] * iw11;
            int iyv = (int)py_[0] * iw00 + (int)py_[1] * iw01 + (int)py_[W1] * iw10 + (int)py_[W1 + 1] * iw11;
            ival = (ival + (1 << 8)) >> 9;
            ixv = (ixv + (1 << 13)) >> 14;
            iyv = (iyv + (1 << 13)) >> 14;
            sm.Iw[k] = (short)ival; sm.Ix[k] = (short)ixv; sm.Iy[k] = (short)iyv;
            s11 += (long long)(ixv * ixv); s12 += (long long)(ixv * iyv); s22 += (long long)(iyv * iyv);
        }
        block_sum3(sm, buf, s11, s12, s22);
        buf ^= 1;
        const float FLT_SCALE = 1.f / (1 << 20);
        float A11 = (float)s11 * FLT_SCALE, A12 = (float)s12 * FLT_SCALE, A22 = (float)s22 * FLT_SCALE;
        float D = A11 * A22 - A12 * A12;
        float minEig = (A22 + A11 - sqrtf((A11 - A22) * (A11 - A22) + 4.f * A12 * A12)) / (float)(2 * win * win);
        if ((double)minEig < 1e-4 || D < 1.1920929e-07f) {
            if (level == 0) status = 0;
            continue;
        }
        D = 1.f / D;
        qx -= half; qy -= half;
        float pdx = 0.f, pdy = 0.f;
        for (int j = 0; j < 30; j++) {
            const int iqx = (int)floorf(qx), iqy = (int)floorf(qy);
            if (iqx < -win || iqx >= J.w || iqy < -win || iqy >= J.h) {
                if (level == 0) status = 0;
                break;
            }
            lk_weights(qx - (float)iqx, qy - (float)iqy, iw00, iw01, iw10, iw11);
            total_iters++;
            long long b1 = 0, b2 = 0, dummy = 0;
            const uint8_t *Jp = J.ptr + (ptrdiff_t)iqy * J.pitch + iqx;
            for (int k = tid; k < npx; k += KLT_THREADS) {
                int y = k / win, x = k - y * win;
                const uint8_t *q = Jp + (ptrdiff_t)y * J.pitch + x;
                int v = (int)q[0] * iw00 + (int)q[1] * iw01 + (int)q[J.pitch] * iw10 + (int)q[J.pitch + 1] * iw11;
                int diff = ((v + (1 << 8)) >> 9) - (int)sm.Iw[k];
                b1 += (long long)(diff * (int)sm.Ix[k]);
                b2 += (long long)(diff * (int)sm.Iy[k]);
            }
            block_sum3(sm, buf, b1, b2, dummy);
            buf ^= 1;
            float fb1 = (float)b1 * FLT_SCALE, fb2 = (float)b2 * FLT_SCALE;
            float dx = (float)((A12 * fb2 - A22 * fb1) * D);
            float dy = (float)((A12 * fb1 - A11 * fb2) * D);
            qx += dx; qy += dy;
            nx = qx + half; ny = qy + half;
            if (lk_step_converged(dx, dy)) break;
            if (j > 0 && lk_below_eps(dx + pdx) && lk_below_eps(dy + pdy)) {
                nx -= dx * 0.5f; ny -= dy * 0.5f;
                break;
            }
            pdx = dx; pdy = dy;
        }
        if (status && level == 0) {
            float ex = nx - half, ey = ny - half;
            const int iex = (int)floorf(ex), iey = (int)floorf(ey);
            if (iex < -win || iex >= J.w || iey < -win || iey >= J.h) { status = 0; continue; }
            lk_weights(ex - (float)iex, ey - (float)iey, iw00, iw01, iw10, iw11);
            long long e = 0, d1 = 0, d2 = 0;
            const uint8_t *Jp = J.ptr + (ptrdiff_t)iey * J.pitch + iex;
            for (int k = tid; k < npx; k += KLT_THREADS) {
                int y = k / win, x = k - y * win;
                const uint8_t *q = Jp + (ptrdiff_t)y * J.pitch + x;
                int v = (int)q[0] * iw00 + (int)q[1] * iw01 + (int)q[J.pitch] * iw10 + (int)q[J.pitch + 1] * iw11;
                int diff = ((v + (1 << 8)) >> 9) - (int)sm.Iw[k];
                e += (long long)abs(diff);
            }
            block_sum3(sm, buf, e, d1, d2);
            buf ^= 1;
            err = (float)e * 1.f / (float)(32 * win * win);
        }
    }

    if (tid == 0) {
        if (status == 0) err = __int_as_float(0x7f800000);  // optical_flow.cpp:46-50
        a.next_pts[2 * i] = nx; a.next_pts[2 * i + 1] = ny;
        a.status[i] = (uint8_t)status;
        a.err[i] = err;
        if (a.iters) a.iters[i] = total_iters;
        if (a.flags) {  // pose_refinement.cpp:125-150
            uint8_t f = a.flags[i];
            float ox = init_x, oy = init_y;
            float d = (init_x - nx) * (init_x - nx) + (init_y - ny) * (init_y - ny);
            if (err > 20) f |= SVO_F_IGN_COMPLETE;
            else if (d > 81) f |= SVO_F_IGN_REFINE;
            else { f &= (uint8_t)~SVO_F_IGN_REFINE; ox = nx; oy = ny; }
            a.flags[i] = f;
            a.kps2d_out[2 * i] = ox; a.kps2d_out[2 * i + 1] = oy;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// Specialisation for the 31x31 window of every shipped configuration (Blender/EuRoC/Econ.yaml): 256 threads,
// thread t owns the 4-pixel run (row t/8, columns 4*(t%8) .. +3) of the window for the whole level, so the template
// (Iw, Ix, Iy of its pixels) lives in registers, horizontally adjacent bilinear taps are shared (10 byte loads for
// 4 pixels instead of 16) and every index is a shift.  Window sums fit 32 bits per thread and are widened to
// 64 bits for the (exact) block reduction.
// ---------------------------------------------------------------------------------------------------------
struct Klt31Shared {
    uint8_t tile[34 * 34 + 4];
    short gx[32 * 32];
    short gy[32 * 32];
    long long part[2][3][KLT_THREADS / 32];
    float init[2];
};

__device__ __forceinline__ void block_sum3_31(Klt31Shared &sm, int buf, long long &a, long long &b, long long &c)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    a = warp_sum_ll(a); b = warp_sum_ll(b); c = warp_sum_ll(c);
    if (lane == 0) { sm.part[buf][0][warp] = a; sm.part[buf][1][warp] = b; sm.part[buf][2][warp] = c; }
    __syncthreads();
    a = b = c = 0;
#pragma unroll
    for (int w = 0; w < KLT_THREADS / 32; w++) { a += sm.part[buf][0][w]; b += sm.part[buf][1][w]; c += sm.part[buf][2][w]; }
}
__device__ __forceinline__ void block_sum2_31(Klt31Shared &sm, int buf, long long &a, long long &b)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    a = warp_sum_ll(a); b = warp_sum_ll(b);
    if (lane == 0) { sm.part[buf][0][warp] = a; sm.part[buf][1][warp] = b; }
    __syncthreads();
    a = b = 0;
#pragma unroll
    for (int w = 0; w < KLT_THREADS / 32; w++) { a += sm.part[buf][0][w]; b += sm.part[buf][1][w]; }
}

__global__ void __launch_bounds__(KLT_THREADS) klt31_kernel(KltArgs a)
{
    __shared__ Klt31Shared sm;
    const int i = blockIdx.x;
    const int n = min(*a.n_ptr, a.max_kps);
    if (i >= n) return;
    const int tid = threadIdx.x;
    constexpr int win = 31, T = 34, W1 = 32;
    const float half = 15.0f;
    const int ry = tid >> 3, rx = (tid & 7) << 2;      // this thread's run: row ry, columns rx .. rx+3
    const bool have_row = ry < win;                     // rows 31 (threads 248..255) are idle
    const int npx_run = (rx + 4 <= win) ? 4 : 3;        // the last run of a row has 3 pixels

    if (tid == 0) {
        float gx, gy;
        if (a.init_pts) { gx = a.init_pts[2 * i]; gy = a.init_pts[2 * i + 1]; }
        else {
            double Rd[9];
            const float *p = a.pose;
            if (a.pose_rd) { for (int k = 0; k < 9; k++) Rd[k] = a.pose_rd[k]; } else dev_rodrigues_d(-p[3], -p[4], -p[5], Rd);
            dev_project(Rd, a.kps3d[3 * i], a.kps3d[3 * i + 1], a.kps3d[3 * i + 2], p[0], p[1], p[2], a.cam.fx, a.cam.fy, a.cam.cx,
                        a.cam.cy, a.cam.k1, a.cam.k2, a.cam.p1, a.cam.p2, a.cam.k3, gx, gy);
        }
        sm.init[0] = gx; sm.init[1] = gy;
    }
    __syncthreads();
    const float init_x = sm.init[0], init_y = sm.init[1];
    const float ppx = a.prev_pts[2 * i], ppy = a.prev_pts[2 * i + 1];
    const LevelDesc *prev_lv = a.keyframe_ids ? (a.kf_lk_table + (size_t)a.keyframe_ids[i] * 2 * SVO_LK_LEVELS) : a.prev_fixed;

    float nx = init_x, ny = init_y;
    int status = 1;
    float err = 0.f;
    int buf = 0, total_iters = 0;

    for (int level = SVO_LK_LEVELS - 1; level >= 0; level--) {
        const LevelDesc I = prev_lv[level];
        const LevelDesc J = a.cur[level];
        const float scale = (float)(1. / (1 << level));
        float px = ppx * scale, py = ppy * scale;
        float qx, qy;
        if (level == SVO_LK_LEVELS - 1) { qx = nx * scale; qy = ny * scale; }
        else { qx = nx * 2.f; qy = ny * 2.f; }
        nx = qx; ny = qy;
        px -= half; py -= half;
        const int ipx = (int)floorf(px), ipy = (int)floorf(py);
        if (ipx < -win || ipx >= I.w || ipy < -win || ipy >= I.h) {
            if (level == 0) { status = 0; err = 0.f; }
            continue;
        }
        int iw00, iw01, iw10, iw11;
        lk_weights(px - (float)ipx, py - (float)ipy, iw00, iw01, iw10, iw11);

        __syncthreads();  // previous level's readers are done with tile/gx/gy
        for (int k = tid; k < T * T; k += KLT_THREADS) {
            int r = k / T, c = k - r * T;
            sm.tile[k] = I.ptr[(ptrdiff_t)(ipy - 1 + r) * I.pitch + (ipx - 1 + c)];
        }
        __syncthreads();
        for (int k = tid; k < W1 * W1; k += KLT_THREADS) {
            const int y = k >> 5, x = k & 31;
            const int X = ipx + x, Y = ipy + y;
            int gxv = 0, gyv = 0;
            if (X >= 0 && X < I.w && Y >= 0 && Y < I.h) {
                const uint8_t *t = &sm.tile[(y + 1) * T + (x + 1)];
                int tl = t[-T - 1], tc = t[-T], trr = t[-T + 1], ml = t[-1], mr = t[1], bl = t[T - 1], bc = t[T], br = t[T + 1];
                gxv = 3 * (trr + br) + 10 * mr - (3 * (tl + bl) + 10 * ml);
                gyv = 3 * ((bl - tl) + (br - trr)) + 10 * (bc - tc);
            }
            sm.gx[k] = (short)gxv; sm.gy[k] = (short)gyv;
        }
        __syncthreads();
        // ---- template of this thread's run, kept in registers for the whole level
        int Iw[4] = {0, 0, 0, 0}, Ix[4] = {0, 0, 0, 0}, Iy[4] = {0, 0, 0, 0};
        long long s11 = 0, s12 = 0, s22 = 0;
        if (have_row) {
            const uint8_t *t0 = &sm.tile[(ry + 1) * T + (rx + 1)];
            const short *g0 = &sm.gx[ry * W1 + rx], *h0 = &sm.gy[ry * W1 + rx];
            int a0[5], a1[5], gx0[5], gx1[5], gy0[5], gy1[5];
#pragma unroll
            for (int j = 0; j < 5; j++) {
                a0[j] = t0[j]; a1[j] = t0[T + j];
                // columns up to rx+4 <= 32: gx/gy have 32 columns (index 31 max) — the 5th tap of the last run is unused
                const int jj = (rx + j < W1) ? j : 0;
                gx0[j] = g0[jj]; gx1[j] = g0[W1 + jj]; gy0[j] = h0[jj]; gy1[j] = h0[W1 + jj];
            }
            int acc11 = 0, acc12 = 0, acc22 = 0;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                if (j < npx_run) {
                    int ival = a0[j] * iw00 + a0[j + 1] * iw01 + a1[j] * iw10 + a1[j + 1] * iw11;
                    int ixv = gx0[j] * iw00 + gx0[j + 1] * iw01 + gx1[j] * iw10 + gx1[j + 1] * iw11;
                    int iyv = gy0[j] * iw00 + gy0[j + 1] * iw01 + gy1[j] * iw10 + gy1[j + 1] * iw11;
                    ival = (ival + (1 << 8)) >> 9;
                    ixv = (ixv + (1 << 13)) >> 14;
                    iyv = (iyv + (1 << 13)) >> 14;
                    Iw[j] = ival; Ix[j] = ixv; Iy[j] = iyv;
                    acc11 += ixv * ixv; acc12 += ixv * iyv; acc22 += iyv * iyv;   // <= 4 * 4080^2 < 2^31
                }
            }
            s11 = acc11; s12 = acc12; s22 = acc22;
        }
        block_sum3_31(sm, buf, s11, s12, s22);
        buf ^= 1;
        const float FLT_SCALE = 1.f / (1 << 20);
        float A11 = (float)s11 * FLT_SCALE, A12 = (float)s12 * FLT_SCALE, A22 = (float)s22 * FLT_SCALE;
        float D = A11 * A22 - A12 * A12;
        float minEig = (A22 + A11 - sqrtf((A11 - A22) * (A11 - A22) + 4.f * A12 * A12)) / (float)(2 * win * win);
        if ((double)minEig < 1e-4 || D < 1.1920929e-07f) {
            if (level == 0) status = 0;
            continue;
        }
        D = 1.f / D;
        qx -= half; qy -= half;
        float pdx = 0.f, pdy = 0.f;
        const ptrdiff_t run_off = (ptrdiff_t)ry * J.pitch + rx;
        for (int j = 0; j < 30; j++) {
            const int iqx = (int)floorf(qx), iqy = (int)floorf(qy);
            if (iqx < -win || iqx >= J.w || iqy < -win || iqy >= J.h) {
                if (level == 0) status = 0;
                break;
            }
            lk_weights(qx - (float)iqx, qy - (float)iqy, iw00, iw01, iw10, iw11);
            total_iters++;
            long long b1 = 0, b2 = 0;
            if (have_row) {
                const uint8_t *q = J.ptr + (ptrdiff_t)iqy * J.pitch + iqx + run_off;
                int c0[5], c1[5];
#pragma unroll
                for (int t = 0; t < 5; t++) { c0[t] = q[t]; c1[t] = q[J.pitch + t]; }
                int acc1 = 0, acc2 = 0;
#pragma unroll
                for (int t = 0; t < 4; t++) {
                    int v = c0[t] * iw00 + c0[t + 1] * iw01 + c1[t] * iw10 + c1[t + 1] * iw11;
                    int diff = ((v + (1 << 8)) >> 9) - Iw[t];
                    if (t >= npx_run) diff = 0;
                    acc1 += diff * Ix[t]; acc2 += diff * Iy[t];   // <= 4 * 8160 * 4080 < 2^31
                }
                b1 = acc1; b2 = acc2;
            }
            block_sum2_31(sm, buf, b1, b2);
            buf ^= 1;
            float fb1 = (float)b1 * FLT_SCALE, fb2 = (float)b2 * FLT_SCALE;
            float dx = (float)((A12 * fb2 - A22 * fb1) * D);
            float dy = (float)((A12 * fb1 - A11 * fb2) * D);
            qx += dx; qy += dy;
            nx = qx + half; ny = qy + half;
            if (lk_step_converged(dx, dy)) break;
            if (j > 0 && lk_below_eps(dx + pdx) && lk_below_eps(dy + pdy)) {
                nx -= dx * 0.5f; ny -= dy * 0.5f;
                break;
            }
            pdx = dx; pdy = dy;
        }
        if (status && level == 0) {
            float ex = nx - half, ey = ny - half;
            const int iex = (int)floorf(ex), iey = (int)floorf(ey);
            if (iex < -win || iex >= J.w || iey < -win || iey >= J.h) { status = 0; continue; }
            lk_weights(ex - (float)iex, ey - (float)iey, iw00, iw01, iw10, iw11);
            long long e = 0, d1 = 0;
            if (have_row) {
                const uint8_t *q = J.ptr + (ptrdiff_t)iey * J.pitch + iex + run_off;
                int c0[5], c1[5];
#pragma unroll
                for (int t = 0; t < 5; t++) { c0[t] = q[t]; c1[t] = q[J.pitch + t]; }
                int acc = 0;
#pragma unroll
                for (int t = 0; t < 4; t++) {
                    int v = c0[t] * iw00 + c0[t + 1] * iw01 + c1[t] * iw10 + c1[t + 1] * iw11;
                    int diff = ((v + (1 << 8)) >> 9) - Iw[t];
                    if (t < npx_run) acc += abs(diff);
                }
                e = acc;
            }
            block_sum2_31(sm, buf, e, d1);
            buf ^= 1;
            err = (float)e * 1.f / (float)(32 * win * win);
        }
    }

    if (tid == 0) {
        if (status == 0) err = __int_as_float(0x7f800000);  // optical_flow.cpp:46-50
        a.next_pts[2 * i] = nx; a.next_pts[2 * i + 1] = ny;
        a.status[i] = (uint8_t)status;
        a.err[i] = err;
        if (a.iters) a.iters[i] = total_iters;
        if (a.flags) {  // pose_refinement.cpp:125-150
            uint8_t f = a.flags[i];
            float ox = init_x, oy = init_y;
            float d = (init_x - nx) * (init_x - nx) + (init_y - ny) * (init_y - ny);
            if (err > 20) f |= SVO_F_IGN_COMPLETE;
            else if (d > 81) f |= SVO_F_IGN_REFINE;
            else { f &= (uint8_t)~SVO_F_IGN_REFINE; ox = nx; oy = ny; }
            a.flags[i] = f;
            a.kps2d_out[2 * i] = ox; a.kps2d_out[2 * i + 1] = oy;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// Warp-per-keypoint kernel for the 31x31 window: lane l owns window row l for the whole level (31 template pixels
// = Iw/Ix/Iy in registers), no block barrier and no shared memory anywhere.  The template is built from the keyframe's
// image level and its Scharr derivative level (materialised once per keyframe, lk_scharr_kernel): a lane fetches its own
// row and takes the row below from lane l+1 with shuffles.  Per LK iteration a lane fetches its 32-byte row of the next
// image as aligned words (+ funnel shift), gets the row below the same way, evaluates the Q14 bilinear taps two at a time
// on the 2-way 16x8-bit dot-product unit (IDP.2A: weights are <= 2^14, pixels 8 bit), and the two window sums are
// reduced exactly with integer warp reductions (REDUX on 16-bit halves).  Every lane then applies the 2x2 update
// redundantly.  4 keypoints per 128-thread CTA; the footprint per keypoint is one warp.
// ---------------------------------------------------------------------------------------------------------
#define KLTW_WARPS 4
#define KLTW_REG_ROWS 48     // staged search region per warp: 32 window rows + 8 above + 8 below
#define KLTW_REG_PITCH 17    // words per staged row: 64 bytes + 1 word (odd pitch: the 32 rows of a warp-wide load hit 32 banks)

// two-way dot product of SIGNED 16-bit halves of a with UNSIGNED bytes of b (lower / upper pair), plus c — the CUDA intrinsics
// __dp2a_lo/hi only come as all-signed or all-unsigned
__device__ __forceinline__ int dp2a_lo_s16u8(uint32_t a, uint32_t b, int c)
{
    int d;
    asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ int dp2a_hi_s16u8(uint32_t a, uint32_t b, int c)
{
    int d;
    asm("dp2a.hi.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

// exact warp-wide sum of a 32-bit signed value per lane (|v| < 2^31), returned as 64 bit to every lane
__device__ __forceinline__ long long warp_sum_i32_exact(int v)
{
    const int lo = v & 0xFFFF, hi = v >> 16;   // v == (hi << 16) + lo
    const int slo = __reduce_add_sync(0xffffffffu, lo);
    const int shi = __reduce_add_sync(0xffffffffu, hi);
    return ((long long)shi << 16) + (long long)slo;
}

// Template of one window row (lane = row; lane 31 only feeds lane 30 with the row below): the 31 bilinear samples of the
// keyframe level (Iw, Q5) and of its Scharr derivative level (Ix, Iy) at the sub-pixel position given by the weights, plus this
// lane's share of the structure-tensor sums.  Used per frame (keyframes without cached templates) and once per keyframe by
// klt_template_kernel — the same code, so both ways give the same bits.
__device__ __forceinline__ void klt31_build_row(const LevelDesc &I, const LevelDesc &Dv, int ipx, int ipy, int iw00, int iw01, int iw10, int iw11,
                                                int lane, bool have_row, int (&Iw)[31], int (&Ix)[31], int (&Iy)[31], int &acc11, int &acc12,
                                                int &acc22)
{
    // image row ipy + lane, pixels ipx .. ipx + 31 (inside the level's REFLECT_101 frame), as 8 realigned words
    const uint8_t *rowp = I.ptr + (ptrdiff_t)(ipy + lane) * I.pitch + ipx;
    const uintptr_t addr = reinterpret_cast<uintptr_t>(rowp);
    const uint32_t *base = reinterpret_cast<const uint32_t *>(addr & ~(uintptr_t)3);
    const int sh = (int)(addr & 3) * 8;
    uint32_t wv[9], ta[8], tb[8];
#pragma unroll
    for (int k = 0; k < 9; k++) wv[k] = base[k];
#pragma unroll
    for (int k = 0; k < 8; k++) ta[k] = __funnelshift_r(wv[k], wv[k + 1], sh);
#pragma unroll
    for (int k = 0; k < 8; k++) tb[k] = __shfl_down_sync(0xffffffffu, ta[k], 1);
    // derivative row (Ix | Iy << 16 per pixel; zero outside the image), same pixels
    const uint32_t *drow = reinterpret_cast<const uint32_t *>(Dv.ptr + (ptrdiff_t)(ipy + lane) * Dv.pitch) + ipx;
    uint32_t da = drow[0], db = __shfl_down_sync(0xffffffffu, da, 1);
    int pa = (int)(ta[0] & 255u), pb = (int)(tb[0] & 255u);
    acc11 = 0; acc12 = 0; acc22 = 0;
#pragma unroll
    for (int x = 0; x < 31; x++) {
        const uint32_t na_w = drow[x + 1];
        const uint32_t nb_w = __shfl_down_sync(0xffffffffu, na_w, 1);
        const int na = (int)((ta[(x + 1) >> 2] >> (8 * ((x + 1) & 3))) & 255u), nb = (int)((tb[(x + 1) >> 2] >> (8 * ((x + 1) & 3))) & 255u);
        const int pgx0 = (int)(short)(da & 0xffffu), pgy0 = (int)da >> 16, pgx1 = (int)(short)(db & 0xffffu), pgy1 = (int)db >> 16;
        const int ngx0 = (int)(short)(na_w & 0xffffu), ngy0 = (int)na_w >> 16, ngx1 = (int)(short)(nb_w & 0xffffu), ngy1 = (int)nb_w >> 16;
        int ival = pa * iw00 + na * iw01 + pb * iw10 + nb * iw11;
        int ixv = pgx0 * iw00 + ngx0 * iw01 + pgx1 * iw10 + ngx1 * iw11;
        int iyv = pgy0 * iw00 + ngy0 * iw01 + pgy1 * iw10 + ngy1 * iw11;
        ival = (ival + (1 << 8)) >> 9;
        ixv = (ixv + (1 << 13)) >> 14;
        iyv = (iyv + (1 << 13)) >> 14;
        if (!have_row) { ival = 0; ixv = 0; iyv = 0; }
        Iw[x] = ival; Ix[x] = ixv; Iy[x] = iyv;
        acc11 += ixv * ixv; acc12 += ixv * iyv; acc22 += iyv * iyv;   // <= 31 * 4080^2 < 2^31
        pa = na; pb = nb; da = na_w; db = nb_w;
    }
}

// Once per KEYFRAME: the templates of its new keypoints on all three levels, stored so that a lane fetches its row with
// twelve coalesced 16-byte loads: data[((keypoint * 3 + level) * 12 + chunk) * 32 + lane], chunks 0-3 = Iw[0..31) as int16 (one
// pad), 4-7 = Ix, 8-11 = Iy; hdr[keypoint * 3 + level] = (A11, A12, A22, flag) with flag 1 = window outside the level.
// The keyframe position of a keypoint never changes, so what calcOpticalFlowPyrLK rebuilds for every frame (and the first
// versions of klt31w_kernel did too: 46 % of the kernel's time) is a per-keyframe constant.
__global__ void __launch_bounds__(32 * KLTW_WARPS) klt_template_kernel(KltTemplateArgs a)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int i = blockIdx.x * KLTW_WARPS + wid;
    if (i >= a.n) return;
    constexpr int win = 31;
    const float half = 15.0f;
    const bool have_row = lane < win;
    const float ppx = a.kps2d[2 * i], ppy = a.kps2d[2 * i + 1];
    for (int level = 0; level < SVO_LK_LEVELS; level++) {
        const LevelDesc I = a.lk[level], Dv = a.lkd[level];
        const float scale = (float)(1. / (1 << level));
        float px = ppx * scale, py = ppy * scale;
        px -= half; py -= half;
        const int ipx = (int)floorf(px), ipy = (int)floorf(py);
        float4 h = make_float4(0.f, 0.f, 0.f, __int_as_float(1));
        if (!(ipx < -win || ipx >= I.w || ipy < -win || ipy >= I.h)) {
            int iw00, iw01, iw10, iw11;
            lk_weights(px - (float)ipx, py - (float)ipy, iw00, iw01, iw10, iw11);
            int Iw[31], Ix[31], Iy[31], acc11, acc12, acc22;
            klt31_build_row(I, Dv, ipx, ipy, iw00, iw01, iw10, iw11, lane, have_row, Iw, Ix, Iy, acc11, acc12, acc22);
            const long long s11 = warp_sum_i32_exact(acc11), s12 = warp_sum_i32_exact(acc12), s22 = warp_sum_i32_exact(acc22);
            const float FLT_SCALE = 1.f / (1 << 20);
            h = make_float4((float)s11 * FLT_SCALE, (float)s12 * FLT_SCALE, (float)s22 * FLT_SCALE, __int_as_float(0));
            uint4 *dst = a.data + ((size_t)(i * SVO_LK_LEVELS + level) * KLT_TPL_CHUNKS) * 32 + lane;
#pragma unroll
            for (int c = 0; c < 4; c++) {
                uint32_t w[3][4];
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const int x0 = 8 * c + 2 * q, x1 = x0 + 1;
                    const int a0 = Iw[x0], a1 = x1 < 31 ? Iw[x1] : 0, b0 = Ix[x0], b1 = x1 < 31 ? Ix[x1] : 0, c0 = Iy[x0], c1 = x1 < 31 ? Iy[x1] : 0;
                    w[0][q] = ((uint32_t)a0 & 0xffffu) | ((uint32_t)a1 << 16);
                    w[1][q] = ((uint32_t)b0 & 0xffffu) | ((uint32_t)b1 << 16);
                    w[2][q] = ((uint32_t)c0 & 0xffffu) | ((uint32_t)c1 << 16);
                }
                dst[(size_t)c * 32] = make_uint4(w[0][0], w[0][1], w[0][2], w[0][3]);
                dst[(size_t)(4 + c) * 32] = make_uint4(w[1][0], w[1][1], w[1][2], w[1][3]);
                dst[(size_t)(8 + c) * 32] = make_uint4(w[2][0], w[2][1], w[2][2], w[2][3]);
            }
        }
        if (lane == 0) a.hdr[i * SVO_LK_LEVELS + level] = h;
    }
}

void launch_klt_templates(const KltTemplateArgs &a, cudaStream_t st)
{
    if (a.n <= 0) return;
    klt_template_kernel<<<(a.n + KLTW_WARPS - 1) / KLTW_WARPS, 32 * KLTW_WARPS, 0, st>>>(a);
}

// WARPS keypoints per CTA.  The warps of a CTA never talk to each other; what the CTA size decides is when a finished warp's registers
// and shared memory go back to the SM: a CTA lives as long as its slowest warp (27 LK iterations per keypoint on average, up to 90),
// so with many frames in flight one-warp CTAs return their slot as soon as their keypoint is done (WARPS = 1, 16 CTAs per SM).
template <int WARPS>
__global__ void __launch_bounds__(32 * WARPS, 16 / WARPS) klt31w_kernel(KltArgs a)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int i = blockIdx.x * WARPS + wid;
    const int n = min(*a.n_ptr, a.max_kps);
    if (i >= n) return;   // whole warp exits together
    constexpr int win = 31;
    const float half = 15.0f;
    const bool have_row = lane < win;

    float init_x, init_y;
    if (a.init_pts) { init_x = a.init_pts[2 * i]; init_y = a.init_pts[2 * i + 1]; }
    else {
        double Rd[9];
        const float *p = a.pose;
        if (a.pose_rd) { for (int k = 0; k < 9; k++) Rd[k] = a.pose_rd[k]; } else dev_rodrigues_d(-p[3], -p[4], -p[5], Rd);
        dev_project(Rd, a.kps3d[3 * i], a.kps3d[3 * i + 1], a.kps3d[3 * i + 2], p[0], p[1], p[2], a.cam.fx, a.cam.fy, a.cam.cx, a.cam.cy,
                    a.cam.k1, a.cam.k2, a.cam.p1, a.cam.p2, a.cam.k3, init_x, init_y);
    }
    const float ppx = a.prev_pts[2 * i], ppy = a.prev_pts[2 * i + 1];
    const LevelDesc *prev_lv = a.keyframe_ids ? (a.kf_lk_table + (size_t)a.keyframe_ids[i] * 2 * SVO_LK_LEVELS) : a.prev_fixed;
    const LevelDesc *prev_dv = a.keyframe_ids ? (prev_lv + SVO_LK_LEVELS) : a.prev_fixed_deriv;
    // template cache of the origin keyframe (svo_keyframe_set_templates): slot of this keypoint, if it has one
    const uint4 *tpl_data = nullptr;
    const float4 *tpl_hdr = nullptr;
    int tpl_slot = 0;
    if (a.kf_tpl_table && a.kp_index && a.keyframe_ids) {
        const KfTemplates t = a.kf_tpl_table[a.keyframe_ids[i]];
        const int idx = a.kp_index[i] - t.first;
        if (t.data && idx >= 0 && idx < t.count) { tpl_data = t.data; tpl_hdr = t.hdr; tpl_slot = idx; }
    }

    float nx = init_x, ny = init_y;
    int status = 1;
    float err = 0.f;
    int total_iters = 0;

    // Search region of the current level in shared memory: KLTW_REG_ROWS rows x 64 bytes around the start position (the
    // window moves by a fraction of a pixel per iteration), re-staged if the window ever leaves it.  A lane reads ITS row
    // of the window, so a 4-byte global load of the warp touches 32 cache lines (32 L1 wavefronts); the same load from
    // shared memory with an odd word pitch is one conflict-free wavefront — 9 loads per iteration, ~27 iterations per keypoint.
    __shared__ uint32_t s_win[WARPS][KLTW_REG_ROWS * KLTW_REG_PITCH];
    uint32_t *swin = s_win[wid];
    int reg_x0 = 0, reg_y0 = 0;
    bool staged = false;
    auto stage = [&](const LevelDesc &J, int ox, int oy) {
        // origin: 8 px left of / above the window, 16-byte aligned in x, clamped to the level's padded buffer
        // (rows -32 .. h + 31, bytes -32 .. pitch - 33 of every row)
        const int x0 = min(max(((ox - 8) >> 4) << 4, -SVO_LK_PAD), J.pitch - SVO_LK_PAD - 64);
        const int y0 = min(max(oy - 8, -SVO_LK_PAD), J.h + SVO_LK_PAD - KLTW_REG_ROWS);
        __syncwarp();   // every lane is done with the previous region
        uint4 v[KLTW_REG_ROWS * 4 / 32];
#pragma unroll
        for (int t = 0; t < KLTW_REG_ROWS * 4 / 32; t++) {
            const int idx = lane + 32 * t, r = idx >> 2, c = idx & 3;
            v[t] = *reinterpret_cast<const uint4 *>(J.ptr + (ptrdiff_t)(y0 + r) * J.pitch + x0 + 16 * c);
        }
#pragma unroll
        for (int t = 0; t < KLTW_REG_ROWS * 4 / 32; t++) {
            const int idx = lane + 32 * t, r = idx >> 2, c = idx & 3;
            uint32_t *d = swin + r * KLTW_REG_PITCH + 4 * c;
            d[0] = v[t].x; d[1] = v[t].y; d[2] = v[t].z; d[3] = v[t].w;
        }
        __syncwarp();
        reg_x0 = x0; reg_y0 = y0; staged = true;
    };
    // window origin (ox, oy) inside the staged region?  9 words from the aligned-down byte must fit the 64-byte row,
    // rows oy .. oy + 31 the KLTW_REG_ROWS rows
    auto ensure_staged = [&](const LevelDesc &J, int ox, int oy) {
        const int rx = ox - reg_x0, ry = oy - reg_y0;
        if (!staged || rx < 0 || rx > 31 || ry < 0 || ry > KLTW_REG_ROWS - 32) stage(J, ox, oy);
    };

    for (int level = SVO_LK_LEVELS - 1; level >= 0; level--) {
        const LevelDesc I = prev_lv[level];
        const LevelDesc Dv = prev_dv[level];
        const LevelDesc J = a.cur[level];
        staged = false;
        const float scale = (float)(1. / (1 << level));
        float px = ppx * scale, py = ppy * scale;
        float qx, qy;
        if (level == SVO_LK_LEVELS - 1) { qx = nx * scale; qy = ny * scale; }
        else { qx = nx * 2.f; qy = ny * 2.f; }
        nx = qx; ny = qy;
        px -= half; py -= half;
        const int ipx = (int)floorf(px), ipy = (int)floorf(py);
        if (ipx < -win || ipx >= I.w || ipy < -win || ipy >= I.h) {
            if (level == 0) { status = 0; err = 0.f; }
            continue;
        }
        int iw00, iw01, iw10, iw11;
        lk_weights(px - (float)ipx, py - (float)ipy, iw00, iw01, iw10, iw11);

        // ---- template of this lane's row in registers: fetched from the keyframe's template cache, else built here
        int Iw[31], Ix[31], Iy[31];
        float A11, A12, A22;
        if (tpl_data) {
            const uint4 *src = tpl_data + ((size_t)(tpl_slot * SVO_LK_LEVELS + level) * KLT_TPL_CHUNKS) * 32 + lane;
            const float4 h = tpl_hdr[tpl_slot * SVO_LK_LEVELS + level];
            A11 = h.x; A12 = h.y; A22 = h.z;
            if (level > 0) {   // next level's template (48 lines of 128 bytes) on its way from DRAM while this level iterates
                const char *nxt = reinterpret_cast<const char *>(tpl_data + ((size_t)(tpl_slot * SVO_LK_LEVELS + level - 1) * KLT_TPL_CHUNKS) * 32);
                asm volatile("prefetch.global.L2 [%0];" ::"l"(nxt + 128 * lane));
                if (lane < 16) asm volatile("prefetch.global.L2 [%0];" ::"l"(nxt + 128 * (32 + lane)));
            }
#pragma unroll
            for (int c = 0; c < 4; c++) {
                const uint4 wa = src[(size_t)c * 32], wb = src[(size_t)(4 + c) * 32], wc = src[(size_t)(8 + c) * 32];
                const uint32_t ua[4] = {wa.x, wa.y, wa.z, wa.w}, ub[4] = {wb.x, wb.y, wb.z, wb.w}, uc[4] = {wc.x, wc.y, wc.z, wc.w};
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const int x0 = 8 * c + 2 * q, x1 = x0 + 1;
                    Iw[x0] = (int)(short)(ua[q] & 0xffffu); Ix[x0] = (int)(short)(ub[q] & 0xffffu); Iy[x0] = (int)(short)(uc[q] & 0xffffu);
                    if (x1 < 31) { Iw[x1] = (int)ua[q] >> 16; Ix[x1] = (int)ub[q] >> 16; Iy[x1] = (int)uc[q] >> 16; }
                }
            }
        } else {
            int acc11, acc12, acc22;
            klt31_build_row(I, Dv, ipx, ipy, iw00, iw01, iw10, iw11, lane, have_row, Iw, Ix, Iy, acc11, acc12, acc22);
            const long long s11 = warp_sum_i32_exact(acc11), s12 = warp_sum_i32_exact(acc12), s22 = warp_sum_i32_exact(acc22);
            const float FLT_SCALE0 = 1.f / (1 << 20);
            A11 = (float)s11 * FLT_SCALE0; A12 = (float)s12 * FLT_SCALE0; A22 = (float)s22 * FLT_SCALE0;
        }
        const float FLT_SCALE = 1.f / (1 << 20);
        float D = A11 * A22 - A12 * A12;
        float minEig = (A22 + A11 - sqrtf((A11 - A22) * (A11 - A22) + 4.f * A12 * A12)) / (float)(2 * win * win);
        if ((double)minEig < 1e-4 || D < 1.1920929e-07f) {
            if (level == 0) status = 0;
            continue;
        }
        D = 1.f / D;
        qx -= half; qy -= half;
        float pdx = 0.f, pdy = 0.f;
        // the template value enters the window passes as the accumulator of the bilinear dot product: with
        // c = 2^8 - 2^9 Iw,  (taps + c) >> 9  ==  ((taps + 2^8) >> 9) - Iw  exactly (2^9 Iw is a multiple of the divisor)
#pragma unroll
        for (int x = 0; x < 31; x++) Iw[x] = (1 << 8) - (Iw[x] << 9);
        // one evaluation pass over the window at integer origin (ox, oy) with weights w*: MODE 0 -> b1/b2, MODE 1 -> sum |diff|
        auto window_pass = [&](int ox, int oy, int w00, int w01, int w10, int w11, int &o1, int &o2, bool want_err) {
            ensure_staged(J, ox, oy);
            // row oy + lane, bytes ox .. ox + 32 (region origin and level rows are 16-byte aligned, so ox and rx share their low bits)
            const int rx = ox - reg_x0, ry = oy - reg_y0;
            const uint32_t *base = swin + (ry + lane) * KLTW_REG_PITCH + (rx >> 2);
            const int sh = (rx & 3) * 8;
            uint32_t wv[9];
#pragma unroll
            for (int k = 0; k < 9; k++) wv[k] = base[k];
            uint32_t top[9], bot[9];
#pragma unroll
            for (int k = 0; k < 8; k++) top[k] = __funnelshift_r(wv[k], wv[k + 1], sh);
            top[8] = 0;
#pragma unroll
            for (int k = 0; k < 8; k++) bot[k] = __shfl_down_sync(0xffffffffu, top[k], 1);
            bot[8] = 0;
            // Q14 weights as SIGNED 16-bit pairs: OpenCV takes iw11 as the remainder 2^14 - iw00 - iw01 - iw10, which is -1
            // when the three rounded weights add up to 2^14 + 1 (fractions of a few 1e-5 of a pixel)
            const uint32_t Wt = ((uint32_t)w00 & 0xffffu) | ((uint32_t)w01 << 16), Wb = ((uint32_t)w10 & 0xffffu) | ((uint32_t)w11 << 16);
            int acc1 = 0, acc2 = 0;
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const uint32_t ts = __funnelshift_r(top[k], top[k + 1], 8), bs = __funnelshift_r(bot[k], bot[k + 1], 8);
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const int x = 4 * k + q;
                    if (x < 31) {
                        int v;
                        // rounding constant and template value ride in as the accumulator of the first dot product
                        const int c0 = Iw[x];
                        if (q == 0) v = dp2a_lo_s16u8(Wt, top[k], dp2a_lo_s16u8(Wb, bot[k], c0));
                        else if (q == 1) v = dp2a_lo_s16u8(Wt, ts, dp2a_lo_s16u8(Wb, bs, c0));
                        else if (q == 2) v = dp2a_hi_s16u8(Wt, top[k], dp2a_hi_s16u8(Wb, bot[k], c0));
                        else v = dp2a_hi_s16u8(Wt, ts, dp2a_hi_s16u8(Wb, bs, c0));
                        const int diff = v >> 9;
                        if (want_err) acc1 += abs(diff);
                        else { acc1 += diff * Ix[x]; acc2 += diff * Iy[x]; }   // <= 31 * 8160 * 4080 < 2^31
                    }
                }
            }
            if (!have_row) { acc1 = 0; acc2 = 0; }
            o1 = acc1; o2 = acc2;
        };
        for (int j = 0; j < 30; j++) {
            const int iqx = (int)floorf(qx), iqy = (int)floorf(qy);
            if (iqx < -win || iqx >= J.w || iqy < -win || iqy >= J.h) {
                if (level == 0) status = 0;
                break;
            }
            lk_weights(qx - (float)iqx, qy - (float)iqy, iw00, iw01, iw10, iw11);
            total_iters++;
            int p1, p2;
            window_pass(iqx, iqy, iw00, iw01, iw10, iw11, p1, p2, false);
            const long long b1 = warp_sum_i32_exact(p1), b2 = warp_sum_i32_exact(p2);
            float fb1 = (float)b1 * FLT_SCALE, fb2 = (float)b2 * FLT_SCALE;
            float dx = (float)((A12 * fb2 - A22 * fb1) * D);
            float dy = (float)((A12 * fb1 - A11 * fb2) * D);
            qx += dx; qy += dy;
            nx = qx + half; ny = qy + half;
            if (lk_step_converged(dx, dy)) break;
            if (j > 0 && lk_below_eps(dx + pdx) && lk_below_eps(dy + pdy)) {
                nx -= dx * 0.5f; ny -= dy * 0.5f;
                break;
            }
            pdx = dx; pdy = dy;
        }
        if (status && level == 0) {
            float ex = nx - half, ey = ny - half;
            const int iex = (int)floorf(ex), iey = (int)floorf(ey);
            if (iex < -win || iex >= J.w || iey < -win || iey >= J.h) { status = 0; continue; }
            lk_weights(ex - (float)iex, ey - (float)iey, iw00, iw01, iw10, iw11);
            int e1, e2;
            window_pass(iex, iey, iw00, iw01, iw10, iw11, e1, e2, true);
            const long long e = warp_sum_i32_exact(e1);
            err = (float)e * 1.f / (float)(32 * win * win);
        }
    }

    if (lane == 0) {
        if (status == 0) err = __int_as_float(0x7f800000);  // optical_flow.cpp:46-50
        a.next_pts[2 * i] = nx; a.next_pts[2 * i + 1] = ny;
        a.status[i] = (uint8_t)status;
        a.err[i] = err;
        if (a.iters) a.iters[i] = total_iters;
        if (a.flags) {  // pose_refinement.cpp:125-150
            uint8_t f = a.flags[i];
            float ox = init_x, oy = init_y;
            float d = (init_x - nx) * (init_x - nx) + (init_y - ny) * (init_y - ny);
            if (err > 20) f |= SVO_F_IGN_COMPLETE;
            else if (d > 81) f |= SVO_F_IGN_REFINE;
            else { f &= (uint8_t)~SVO_F_IGN_REFINE; ox = nx; oy = ny; }
            a.flags[i] = f;
            a.kps2d_out[2 * i] = ox; a.kps2d_out[2 * i + 1] = oy;
        }
    }
}

// Two warps per keypoint (a sequence that has the device to itself): what a lone frame waits for in the KLT stage is its SLOWEST
// keypoint — up to 90 LK iterations against 10 on average — and an iteration of klt31w_kernel is ~350 instructions of ONE warp
// with nothing else on its scheduler to hide their latency.  Here lane l of warp h owns columns 16h .. 16h + 15 of window row l:
// half the template in registers, half the window pass per iteration; the two partial sums meet through shared memory (exact
// integers, so the order of the additions is immaterial) behind one CTA barrier per iteration, and both warps then take the same
// 2x2 update redundantly.  Same bits as klt31w_kernel; twice the warps, so it is not the shape for many sequences per device.
__global__ void __launch_bounds__(64, 8) klt31h_kernel(KltArgs a)
{
    const int lane = threadIdx.x & 31, hw = threadIdx.x >> 5;
    const int i = blockIdx.x;
    const int n = min(*a.n_ptr, a.max_kps);
    if (i >= n) return;   // whole CTA exits together
    constexpr int win = 31;
    const float half = 15.0f;
    const bool have_row = lane < win;
    const int x0h = 16 * hw;                // first window column of this warp

    float init_x, init_y;
    if (a.init_pts) { init_x = a.init_pts[2 * i]; init_y = a.init_pts[2 * i + 1]; }
    else {
        double Rd[9];
        const float *p = a.pose;
        if (a.pose_rd) { for (int k = 0; k < 9; k++) Rd[k] = a.pose_rd[k]; } else dev_rodrigues_d(-p[3], -p[4], -p[5], Rd);
        dev_project(Rd, a.kps3d[3 * i], a.kps3d[3 * i + 1], a.kps3d[3 * i + 2], p[0], p[1], p[2], a.cam.fx, a.cam.fy, a.cam.cx, a.cam.cy,
                    a.cam.k1, a.cam.k2, a.cam.p1, a.cam.p2, a.cam.k3, init_x, init_y);
    }
    const float ppx = a.prev_pts[2 * i], ppy = a.prev_pts[2 * i + 1];
    const LevelDesc *prev_lv = a.keyframe_ids ? (a.kf_lk_table + (size_t)a.keyframe_ids[i] * 2 * SVO_LK_LEVELS) : a.prev_fixed;
    const LevelDesc *prev_dv = a.keyframe_ids ? (prev_lv + SVO_LK_LEVELS) : a.prev_fixed_deriv;
    const uint4 *tpl_data = nullptr;
    const float4 *tpl_hdr = nullptr;
    int tpl_slot = 0;
    if (a.kf_tpl_table && a.kp_index && a.keyframe_ids) {
        const KfTemplates t = a.kf_tpl_table[a.keyframe_ids[i]];
        const int idx = a.kp_index[i] - t.first;
        if (t.data && idx >= 0 && idx < t.count) { tpl_data = t.data; tpl_hdr = t.hdr; tpl_slot = idx; }
    }

    float nx = init_x, ny = init_y;
    int status = 1;
    float err = 0.f;
    int total_iters = 0;

    __shared__ uint32_t swin[KLTW_REG_ROWS * KLTW_REG_PITCH + KLTW_REG_PITCH];
    __shared__ long long s_part[2][2][2];   // [buffer][warp][b1, b2]
    int pbuf = 0;
    int reg_x0 = 0, reg_y0 = 0;
    bool staged = false;
    auto stage = [&](const LevelDesc &J, int ox, int oy) {
        const int x0 = min(max(((ox - 8) >> 4) << 4, -SVO_LK_PAD), J.pitch - SVO_LK_PAD - 64);
        const int y0 = min(max(oy - 8, -SVO_LK_PAD), J.h + SVO_LK_PAD - KLTW_REG_ROWS);
        __syncthreads();   // both warps are done with the previous region
        uint4 v[KLTW_REG_ROWS * 4 / 64];
#pragma unroll
        for (int t = 0; t < KLTW_REG_ROWS * 4 / 64; t++) {
            const int idx = (int)threadIdx.x + 64 * t, r = idx >> 2, c = idx & 3;
            v[t] = *reinterpret_cast<const uint4 *>(J.ptr + (ptrdiff_t)(y0 + r) * J.pitch + x0 + 16 * c);
        }
#pragma unroll
        for (int t = 0; t < KLTW_REG_ROWS * 4 / 64; t++) {
            const int idx = (int)threadIdx.x + 64 * t, r = idx >> 2, c = idx & 3;
            uint32_t *d = swin + r * KLTW_REG_PITCH + 4 * c;
            d[0] = v[t].x; d[1] = v[t].y; d[2] = v[t].z; d[3] = v[t].w;
        }
        __syncthreads();
        reg_x0 = x0; reg_y0 = y0; staged = true;
    };
    auto ensure_staged = [&](const LevelDesc &J, int ox, int oy) {   // CTA uniform: both warps carry the same positions
        const int rx = ox - reg_x0, ry = oy - reg_y0;
        if (!staged || rx < 0 || rx > 31 || ry < 0 || ry > KLTW_REG_ROWS - 32) stage(J, ox, oy);
    };
    // CTA-wide exact sum of one value pair per lane (both warps get both totals)
    auto cta_sum2 = [&](int p1, int p2, long long &o1, long long &o2) {
        const long long w1 = warp_sum_i32_exact(p1), w2 = warp_sum_i32_exact(p2);
        if (lane == 0) { s_part[pbuf][hw][0] = w1; s_part[pbuf][hw][1] = w2; }
        __syncthreads();
        o1 = s_part[pbuf][0][0] + s_part[pbuf][1][0];
        o2 = s_part[pbuf][0][1] + s_part[pbuf][1][1];
        pbuf ^= 1;   // a buffer is written again two sums later, i.e. after the barrier that follows every reader of this one
    };

    for (int level = SVO_LK_LEVELS - 1; level >= 0; level--) {
        const LevelDesc I = prev_lv[level];
        const LevelDesc Dv = prev_dv[level];
        const LevelDesc J = a.cur[level];
        staged = false;
        const float scale = (float)(1. / (1 << level));
        float px = ppx * scale, py = ppy * scale;
        float qx, qy;
        if (level == SVO_LK_LEVELS - 1) { qx = nx * scale; qy = ny * scale; }
        else { qx = nx * 2.f; qy = ny * 2.f; }
        nx = qx; ny = qy;
        px -= half; py -= half;
        const int ipx = (int)floorf(px), ipy = (int)floorf(py);
        if (ipx < -win || ipx >= I.w || ipy < -win || ipy >= I.h) {
            if (level == 0) { status = 0; err = 0.f; }
            continue;
        }
        int iw00, iw01, iw10, iw11;
        lk_weights(px - (float)ipx, py - (float)ipy, iw00, iw01, iw10, iw11);

        // ---- this lane's HALF row of the template: columns x0h .. x0h + 15 (column 31 does not exist: zero)
        int Iw[16], Ix[16], Iy[16];
        float A11, A12, A22;
        if (tpl_data) {
            const uint4 *src = tpl_data + ((size_t)(tpl_slot * SVO_LK_LEVELS + level) * KLT_TPL_CHUNKS) * 32 + lane;
            const float4 h = tpl_hdr[tpl_slot * SVO_LK_LEVELS + level];
            A11 = h.x; A12 = h.y; A22 = h.z;
            if (level > 0) {
                const char *nxt = reinterpret_cast<const char *>(tpl_data + ((size_t)(tpl_slot * SVO_LK_LEVELS + level - 1) * KLT_TPL_CHUNKS) * 32);
                if (threadIdx.x < 48) asm volatile("prefetch.global.L2 [%0];" ::"l"(nxt + 128 * threadIdx.x));
            }
#pragma unroll
            for (int c = 0; c < 2; c++) {
                const uint4 wa = src[(size_t)(2 * hw + c) * 32], wb = src[(size_t)(4 + 2 * hw + c) * 32], wc = src[(size_t)(8 + 2 * hw + c) * 32];
                const uint32_t ua[4] = {wa.x, wa.y, wa.z, wa.w}, ub[4] = {wb.x, wb.y, wb.z, wb.w}, uc[4] = {wc.x, wc.y, wc.z, wc.w};
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const int xl = 8 * c + 2 * q;
                    Iw[xl] = (int)(short)(ua[q] & 0xffffu); Ix[xl] = (int)(short)(ub[q] & 0xffffu); Iy[xl] = (int)(short)(uc[q] & 0xffffu);
                    Iw[xl + 1] = (int)ua[q] >> 16; Ix[xl + 1] = (int)ub[q] >> 16; Iy[xl + 1] = (int)uc[q] >> 16;   // the pad of column 31 is stored as 0
                }
            }
        } else {
            // keyframe without cached templates (stage entry points): every warp builds whole rows and keeps its half
            int fIw[31], fIx[31], fIy[31], acc11, acc12, acc22;
            klt31_build_row(I, Dv, ipx, ipy, iw00, iw01, iw10, iw11, lane, have_row, fIw, fIx, fIy, acc11, acc12, acc22);
            const long long s11 = warp_sum_i32_exact(acc11), s12 = warp_sum_i32_exact(acc12), s22 = warp_sum_i32_exact(acc22);
            const float FLT_SCALE0 = 1.f / (1 << 20);
            A11 = (float)s11 * FLT_SCALE0; A12 = (float)s12 * FLT_SCALE0; A22 = (float)s22 * FLT_SCALE0;
#pragma unroll
            for (int xl = 0; xl < 16; xl++) {
                Iw[xl] = hw ? (xl < 15 ? fIw[16 + (xl < 15 ? xl : 0)] : 0) : fIw[xl];
                Ix[xl] = hw ? (xl < 15 ? fIx[16 + (xl < 15 ? xl : 0)] : 0) : fIx[xl];
                Iy[xl] = hw ? (xl < 15 ? fIy[16 + (xl < 15 ? xl : 0)] : 0) : fIy[xl];
            }
        }
        const float FLT_SCALE = 1.f / (1 << 20);
        float D = A11 * A22 - A12 * A12;
        float minEig = (A22 + A11 - sqrtf((A11 - A22) * (A11 - A22) + 4.f * A12 * A12)) / (float)(2 * win * win);
        if ((double)minEig < 1e-4 || D < 1.1920929e-07f) {
            if (level == 0) status = 0;
            continue;
        }
        D = 1.f / D;
        qx -= half; qy -= half;
        float pdx = 0.f, pdy = 0.f;
#pragma unroll
        for (int x = 0; x < 16; x++) Iw[x] = (1 << 8) - (Iw[x] << 9);
        // this lane's half of one evaluation pass over the window at integer origin (ox, oy)
        auto window_pass = [&](int ox, int oy, int w00, int w01, int w10, int w11, int &o1, int &o2, bool want_err) {
            ensure_staged(J, ox, oy);
            const int rx = ox - reg_x0 + x0h, ry = oy - reg_y0;
            const uint32_t *base = swin + (ry + lane) * KLTW_REG_PITCH + (rx >> 2);
            const int sh = (rx & 3) * 8;
            uint32_t wv[6];
#pragma unroll
            for (int k = 0; k < 6; k++) wv[k] = base[k];
            uint32_t top[5], bot[5];
#pragma unroll
            for (int k = 0; k < 5; k++) top[k] = __funnelshift_r(wv[k], wv[k + 1], sh);
#pragma unroll
            for (int k = 0; k < 5; k++) bot[k] = __shfl_down_sync(0xffffffffu, top[k], 1);
            const uint32_t Wt = ((uint32_t)w00 & 0xffffu) | ((uint32_t)w01 << 16), Wb = ((uint32_t)w10 & 0xffffu) | ((uint32_t)w11 << 16);
            int acc1 = 0, acc2 = 0;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const uint32_t ts = __funnelshift_r(top[k], top[k + 1], 8), bs = __funnelshift_r(bot[k], bot[k + 1], 8);
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const int xl = 4 * k + q;
                    int v;
                    const int c0 = Iw[xl];
                    if (q == 0) v = dp2a_lo_s16u8(Wt, top[k], dp2a_lo_s16u8(Wb, bot[k], c0));
                    else if (q == 1) v = dp2a_lo_s16u8(Wt, ts, dp2a_lo_s16u8(Wb, bs, c0));
                    else if (q == 2) v = dp2a_hi_s16u8(Wt, top[k], dp2a_hi_s16u8(Wb, bot[k], c0));
                    else v = dp2a_hi_s16u8(Wt, ts, dp2a_hi_s16u8(Wb, bs, c0));
                    int diff = v >> 9;
                    if (xl == 15 && hw) diff = 0;          // column 31 is not part of the window
                    if (want_err) acc1 += abs(diff);
                    else { acc1 += diff * Ix[xl]; acc2 += diff * Iy[xl]; }
                }
            }
            if (!have_row) { acc1 = 0; acc2 = 0; }
            o1 = acc1; o2 = acc2;
        };
        for (int j = 0; j < 30; j++) {
            const int iqx = (int)floorf(qx), iqy = (int)floorf(qy);
            if (iqx < -win || iqx >= J.w || iqy < -win || iqy >= J.h) {
                if (level == 0) status = 0;
                break;
            }
            lk_weights(qx - (float)iqx, qy - (float)iqy, iw00, iw01, iw10, iw11);
            total_iters++;
            int p1, p2;
            window_pass(iqx, iqy, iw00, iw01, iw10, iw11, p1, p2, false);
            long long b1, b2;
            cta_sum2(p1, p2, b1, b2);
            float fb1 = (float)b1 * FLT_SCALE, fb2 = (float)b2 * FLT_SCALE;
            float dx = (float)((A12 * fb2 - A22 * fb1) * D);
            float dy = (float)((A12 * fb1 - A11 * fb2) * D);
            qx += dx; qy += dy;
            nx = qx + half; ny = qy + half;
            if (lk_step_converged(dx, dy)) break;
            if (j > 0 && lk_below_eps(dx + pdx) && lk_below_eps(dy + pdy)) {
                nx -= dx * 0.5f; ny -= dy * 0.5f;
                break;
            }
            pdx = dx; pdy = dy;
        }
        if (status && level == 0) {
            float ex = nx - half, ey = ny - half;
            const int iex = (int)floorf(ex), iey = (int)floorf(ey);
            if (iex < -win || iex >= J.w || iey < -win || iey >= J.h) { status = 0; continue; }
            lk_weights(ex - (float)iex, ey - (float)iey, iw00, iw01, iw10, iw11);
            int e1, e2;
            window_pass(iex, iey, iw00, iw01, iw10, iw11, e1, e2, true);
            long long e, dummy;
            cta_sum2(e1, 0, e, dummy);
            err = (float)e * 1.f / (float)(32 * win * win);
        }
    }

    if (threadIdx.x == 0) {
        if (status == 0) err = __int_as_float(0x7f800000);  // optical_flow.cpp:46-50
        a.next_pts[2 * i] = nx; a.next_pts[2 * i + 1] = ny;
        a.status[i] = (uint8_t)status;
        a.err[i] = err;
        if (a.iters) a.iters[i] = total_iters;
        if (a.flags) {  // pose_refinement.cpp:125-150
            uint8_t f = a.flags[i];
            float ox = init_x, oy = init_y;
            float d = (init_x - nx) * (init_x - nx) + (init_y - ny) * (init_y - ny);
            if (err > 20) f |= SVO_F_IGN_COMPLETE;
            else if (d > 81) f |= SVO_F_IGN_REFINE;
            else { f &= (uint8_t)~SVO_F_IGN_REFINE; ox = nx; oy = ny; }
            a.flags[i] = f;
            a.kps2d_out[2 * i] = ox; a.kps2d_out[2 * i + 1] = oy;
        }
    }
}

static int g_klt_variant = -1;  // SVO_KLT_VARIANT=block selects the CTA-per-keypoint kernel (A/B measurements)
static int g_klt_warps = -1;    // SVO_KLT_WARPS=1|4: keypoints per CTA of the warp-per-keypoint kernel

void launch_klt(const KltArgs &a, bool wide, cudaStream_t st)
{
    if (a.max_kps <= 0) return;
    if (g_klt_variant < 0) {
        const char *e = getenv("SVO_KLT_VARIANT");
        g_klt_variant = (e && e[0] == 'b') ? 1 : 0;
    }
    if (g_klt_warps < 0) {
        const char *e = getenv("SVO_KLT_WARPS");
        g_klt_warps = (e && atoi(e) == 4) ? 4 : 1;
    }
    static const bool no_split = getenv("SVO_KLT_NO_SPLIT") != nullptr;   // developer A/B switch
    // two warps per keypoint: -8 % for a lone frame of a few hundred keypoints (the stage lasts as long as its slowest keypoint), a
    // loss once the keypoints outnumber the warp slots (C4, 3 000 keypoints: 59 -> 65 us)
    if (a.cam.win_flow == 31 && g_klt_variant == 0 && wide && !no_split && a.max_kps <= 1024) klt31h_kernel<<<a.max_kps, 64, 0, st>>>(a);
    else if (a.cam.win_flow == 31 && g_klt_variant == 0 && g_klt_warps == 4) klt31w_kernel<4><<<(a.max_kps + 3) / 4, 128, 0, st>>>(a);
    else if (a.cam.win_flow == 31 && g_klt_variant == 0) klt31w_kernel<1><<<a.max_kps, 32, 0, st>>>(a);
    else if (a.cam.win_flow == 31) klt31_kernel<<<a.max_kps, KLT_THREADS, 0, st>>>(a);
    else klt_pyr_lk_kernel<0><<<a.max_kps, KLT_THREADS, 0, st>>>(a);
}
