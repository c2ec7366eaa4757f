// Shared device/host helpers for the sm_100a tracking kernels.
// All translation units are compiled with -fmad=false: the reference's float arithmetic is plain scalar
// C++ without FMA contraction (src/Makefile:5-6), and per-keypoint values are reproduced operation by
// operation so that only cross-keypoint reductions differ in rounding.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
#include "../../include/svo_cuda.h"

#define SVO_MAX_LEVELS 8
#define SVO_LK_LEVELS 3
#define SVO_LK_PAD 32  // border (px) materialised around every LK pyramid level (>= winSize + 1)

// ---- flags (bit layout shared by host and device; mirrors KeyPointInformation's three booleans,
//      src/include/stereo_slam_types.hpp:86-100)
#define SVO_F_IGN_REFINE 1u
#define SVO_F_IGN_COMPLETE 2u
#define SVO_F_IGN_TEMP 4u

struct DevCam {  // camera + algorithm settings, passed by value to kernels
    float baseline, fx, fy, cx, cy, k1, k2, k3, p1, p2;
    int grid_height, grid_width, search_x, search_y;
    int win_pose, win_flow, win_depth, max_levels, min_level;
};

static inline DevCam make_devcam(const svo_camera_settings &s)
{
    DevCam c;
    c.baseline = s.baseline; c.fx = s.fx; c.fy = s.fy; c.cx = s.cx; c.cy = s.cy;
    c.k1 = s.k1; c.k2 = s.k2; c.k3 = s.k3; c.p1 = s.p1; c.p2 = s.p2;
    c.grid_height = s.grid_height; c.grid_width = s.grid_width; c.search_x = s.search_x; c.search_y = s.search_y;
    c.win_pose = s.window_size_pose_estimator; c.win_flow = s.window_size_opt_flow;
    c.win_depth = s.window_size_depth_calculator; c.max_levels = s.max_pyramid_levels;
    c.min_level = s.min_pyramid_level_pose_estimation;
    return c;
}

#ifdef __CUDACC__
// cv::Rodrigues (vector -> matrix) in double — same expression order as OpenCV (pose_manager.cpp:15-16).
// (not inlined: it is used at three places of each solver kernel; measured neutral in time, 4 KB less code per kernel)
static __device__ __noinline__ void dev_rodrigues_d(float r0, float r1, float r2, double R[9])
{
    double rx = r0, ry = r1, rz = r2;
    double theta = sqrt(rx * rx + ry * ry + rz * rz);
    if (theta < 2.220446049250313e-16) {
        R[0] = 1; R[1] = 0; R[2] = 0; R[3] = 0; R[4] = 1; R[5] = 0; R[6] = 0; R[7] = 0; R[8] = 1;
        return;
    }
    double s, c;
    sincos(theta, &s, &c);
    double c1 = 1. - c;
    double it = 1. / theta;
    rx *= it; ry *= it; rz *= it;
    // OpenCV writes R = c I + c1 r r^T + s [r]x out in full — R[1] = c * 0 + c1 * (rx * ry) + s * (-rz) and so on.  c * 1 is c, adding a
    // product with 0 changes nothing and a + s * (-b) is a - s * b, all exactly, so the 24 operations below give the same bits as
    // the 63 of the full expressions (up to the sign of a zero entry) — this runs on single lanes inside every gradient round
    const double xx = c1 * (rx * rx), xy = c1 * (rx * ry), xz = c1 * (rx * rz), yy = c1 * (ry * ry), yz = c1 * (ry * rz), zz = c1 * (rz * rz);
    const double sx = s * rx, sy = s * ry, sz = s * rz;
    R[0] = c + xx;  R[1] = xy - sz; R[2] = xz + sy;
    R[3] = xy + sz; R[4] = c + yy;  R[5] = yz - sx;
    R[6] = xz - sy; R[7] = yz + sx; R[8] = c + zz;
}
__device__ __forceinline__ void dev_rodrigues_f(float r0, float r1, float r2, float R[9])
{
    double Rd[9];
    dev_rodrigues_d(r0, r1, r2, Rd);
#pragma unroll
    for (int k = 0; k < 9; k++) R[k] = (float)Rd[k];
}

// cv::Matx33f * cv::Vec3f : s = 0; s += a(i,k)*b(k)  (float, sequential)
__device__ __forceinline__ void dev_m33v(const float *M, float v0, float v1, float v2, float &o0, float &o1, float &o2)
{
    float s;
    s = 0.f; s += M[0] * v0; s += M[1] * v1; s += M[2] * v2; o0 = s;
    s = 0.f; s += M[3] * v0; s += M[4] * v1; s += M[5] * v2; o1 = s;
    s = 0.f; s += M[6] * v0; s += M[7] * v1; s += M[8] * v2; o2 = s;
}

// project_keypoints (transform_keypoints.cpp:11-47): float (P - t), then cv::projectPoints in double with
// rotation Rd = Rodrigues(-r) (double, NOT rounded to float) and tvec = 0.
__device__ __forceinline__ void dev_project(const double *Rd, float px, float py, float pz, float tx, float ty, float tz,
                                            float fx, float fy, float cx, float cy, float k1, float k2, float p1, float p2,
                                            float k3, float &u, float &v)
{
    double X = (double)(px - tx), Y = (double)(py - ty), Z = (double)(pz - tz);
    double x = Rd[0] * X + Rd[1] * Y + Rd[2] * Z + 0.0;
    double y = Rd[3] * X + Rd[4] * Y + Rd[5] * Z + 0.0;
    double z = Rd[6] * X + Rd[7] * Y + Rd[8] * Z + 0.0;
    z = z ? 1. / z : 1;
    x *= z; y *= z;
    double r2 = x * x + y * y, r4 = r2 * r2, r6 = r4 * r2;
    double a1 = 2 * x * y, a2 = r2 + 2 * x * x, a3 = r2 + 2 * y * y;
    double cdist = 1 + (double)k1 * r2 + (double)k2 * r4 + (double)k3 * r6;
    double xd0 = x * cdist * 1.0 + (double)p1 * a1 + (double)p2 * a2;
    double yd0 = y * cdist * 1.0 + (double)p1 * a3 + (double)p2 * a1;
    u = (float)(xd0 * (double)fx + (double)cx);
    v = (float)(yd0 * (double)fy + (double)cy);
}

// dev_project with a fast path for a camera without lens distortion (k1 = k2 = p1 = p2 = k3 = 0, all shipped settings): with
// zero coefficients cv::projectPoints' distortion polynomial multiplies by exactly 1 and adds exactly 0, so dropping it leaves
// every finite result bit-identical — and takes a dozen dependent double-precision operations out of every cost evaluation.
__device__ __forceinline__ bool dev_cam_nodist(const DevCam &cam)
{
    return cam.k1 == 0.f && cam.k2 == 0.f && cam.p1 == 0.f && cam.p2 == 0.f && cam.k3 == 0.f;
}
// (the full model out of line: no shipped camera takes it)
static __device__ __noinline__ void dev_project_dist(const double *Rd, float px, float py, float pz, float tx, float ty, float tz,
                                                     float fx, float fy, float cx, float cy, float k1, float k2, float p1, float p2,
                                                     float k3, float *uv)
{
    dev_project(Rd, px, py, pz, tx, ty, tz, fx, fy, cx, cy, k1, k2, p1, p2, k3, uv[0], uv[1]);
}
__device__ __forceinline__ void dev_project_nd(bool nodist, const double *Rd, float px, float py, float pz, float tx, float ty, float tz,
                                               float fx, float fy, float cx, float cy, const DevCam &cam, float &u, float &v)
{
    if (!nodist) {
        float uv[2];
        dev_project_dist(Rd, px, py, pz, tx, ty, tz, fx, fy, cx, cy, cam.k1, cam.k2, cam.p1, cam.p2, cam.k3, uv);
        u = uv[0]; v = uv[1];
        return;
    }
    double X = (double)(px - tx), Y = (double)(py - ty), Z = (double)(pz - tz);
    double x = Rd[0] * X + Rd[1] * Y + Rd[2] * Z + 0.0;
    double y = Rd[3] * X + Rd[4] * Y + Rd[5] * Z + 0.0;
    double z = Rd[6] * X + Rd[7] * Y + Rd[8] * Z + 0.0;
    z = z ? 1. / z : 1;
    x *= z; y *= z;
    u = (float)(x * (double)fx + (double)cx);
    v = (float)(y * (double)fy + (double)cy);
}

// exponential_map.hpp:12-37 — norm fixed to 1, w unchanged.  `cos(_norm)` / `sin(_norm)` take a float there, and under
// <opencv2/opencv.hpp> (which pulls in <math.h>, hence libstdc++'s global float overloads) they are float calls, so the two
// scale factors are the FLOAT values 1 - cosf(1) and 1 - sinf(1) and `scalar * Matx33f` rounds float products — pinned against
// the reference itself (oracle/_ref, tests/test_ref_pin.py).
__device__ __forceinline__ void dev_expmap(const float tw[6], float out[6])
{
    const float C1 = 0x1.d6bbp-2f;   // 1.f - cosf(1.f)
    const float C2 = 0x1.44aaep-3f;  // 1.f - sinf(1.f)
    float w0 = tw[3], w1 = tw[4], w2 = tw[5];
    // M = I + C1 K + C2 K K with K = [w]x, formed by the reference as generic 3x3 float products (s = 0; s += K(i,k) * K(k,j)).  The
    // zero entries of K contribute exact zeros, so each entry of K K is one product or a sum of two, written out here in the
    // reference's order of additions: the same bits with a third of the operations.
    const float k00 = (-(w2 * w2)) - (w1 * w1), k11 = (-(w2 * w2)) - (w0 * w0), k22 = (-(w1 * w1)) - (w0 * w0);
    const float k01 = w1 * w0, k02 = w2 * w0, k12 = w2 * w1;
    float M[9];
    M[0] = 1.f + k00 * C2;            M[1] = (-w2) * C1 + k01 * C2;      M[2] = w1 * C1 + k02 * C2;
    M[3] = w2 * C1 + k01 * C2;        M[4] = 1.f + k11 * C2;             M[5] = (-w0) * C1 + k12 * C2;
    M[6] = (-w1) * C1 + k02 * C2;     M[7] = w0 * C1 + k12 * C2;         M[8] = 1.f + k22 * C2;
    dev_m33v(M, tw[0], tw[1], tw[2], out[0], out[1], out[2]);
    out[3] = w0; out[4] = w1; out[5] = w2;
}

// cv::Matx66f::inv(DECOMP_SVD) (pose_estimator.cpp:405, pose_refinement.cpp:398) for a Hessian the double LDL^T of the solvers
// rejects (not numerically positive definite: fewer than three usable keypoints, collinear points, ...).  The reference
// then still takes the step its FLOAT pseudo-inverse gives: one-sided Jacobi SVD (OpenCV JacobiSVDImpl_<float>: rotations in
// float, norms and dot products in double), 1/w for every singular value above 2*DBL_EPSILON*sum(w) (SVBkSbImpl_), and an
// all-zero inverse when cv::invert reports singularity (smallest singular value exactly 0, or largest < FLT_EPSILON).
// Same operation order as OpenCV so that an identical float H gives the identical inverse; one thread, rarely executed.
static __device__ __noinline__ bool dev_invert_svd6(const float *A /*6x6 row-major*/, float *Ainv)
{
    const int n = 6;
    float At[36], Vt[36];
    double W[6];
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) At[i * n + j] = A[j * n + i];
    const float eps = 1.1920929e-07f * 2;
    for (int i = 0; i < n; i++) {
        double sd = 0;
        for (int k = 0; k < n; k++) { const float t = At[i * n + k]; sd += (double)t * t; }
        W[i] = sd;
        for (int k = 0; k < n; k++) Vt[i * n + k] = 0.f;
        Vt[i * n + i] = 1.f;
    }
    for (int iter = 0; iter < 30; iter++) {
        bool changed = false;
        for (int i = 0; i < n - 1; i++)
            for (int j = i + 1; j < n; j++) {
                float *Ai = At + i * n, *Aj = At + j * n;
                double a = W[i], p = 0, b = W[j];
                for (int k = 0; k < n; k++) p += (double)Ai[k] * Aj[k];
                if (fabs(p) <= eps * sqrt(a * b)) continue;
                p *= 2;
                const double beta = a - b, gamma = hypot(p, beta);
                float c, s;
                if (beta < 0) {
                    const double delta = (gamma - beta) * 0.5;
                    s = (float)sqrt(delta / gamma);
                    c = (float)(p / (gamma * s * 2));
                } else {
                    c = (float)sqrt((gamma + beta) / (gamma * 2));
                    s = (float)(p / (gamma * c * 2));
                }
                a = b = 0;
                for (int k = 0; k < n; k++) {
                    const float t0 = c * Ai[k] + s * Aj[k];
                    const float t1 = -s * Ai[k] + c * Aj[k];
                    Ai[k] = t0; Aj[k] = t1;
                    a += (double)t0 * t0; b += (double)t1 * t1;
                }
                W[i] = a; W[j] = b;
                changed = true;
                float *Vi = Vt + i * n, *Vj = Vt + j * n;
                for (int k = 0; k < n; k++) {
                    const float t0 = c * Vi[k] + s * Vj[k];
                    const float t1 = -s * Vi[k] + c * Vj[k];
                    Vi[k] = t0; Vj[k] = t1;
                }
            }
        if (!changed) break;
    }
    for (int i = 0; i < n; i++) {
        double sd = 0;
        for (int k = 0; k < n; k++) { const float t = At[i * n + k]; sd += (double)t * t; }
        W[i] = sqrt(sd);
    }
    for (int i = 0; i < n - 1; i++) {   // descending singular values
        int j = i;
        for (int k = i + 1; k < n; k++) if (W[j] < W[k]) j = k;
        if (i != j) {
            const double t = W[i]; W[i] = W[j]; W[j] = t;
            for (int k = 0; k < n; k++) {
                float u = At[i * n + k]; At[i * n + k] = At[j * n + k]; At[j * n + k] = u;
                u = Vt[i * n + k]; Vt[i * n + k] = Vt[j * n + k]; Vt[j * n + k] = u;
            }
        }
    }
    float w[6];
    for (int i = 0; i < n; i++) {
        w[i] = (float)W[i];
        const float sc = (float)(W[i] > 1.17549435e-38 ? 1 / W[i] : 0.);   // unit left vectors (zero singular value: left as zero)
        for (int k = 0; k < n; k++) At[i * n + k] *= sc;
    }
    // back substitution with the identity as right-hand side: X = sum_i v_i (u_i^T / w_i)
    double threshold = 0;
    for (int i = 0; i < n; i++) threshold += w[i];
    threshold *= (double)(float)(2.220446049250313e-16 * 2);
    for (int i = 0; i < n * n; i++) Ainv[i] = 0.f;
    for (int i = 0; i < n; i++) {
        double wi = w[i];
        if (fabs(wi) <= threshold) continue;
        wi = 1 / wi;
        double buf[6];
        for (int j = 0; j < n; j++) buf[j] = At[i * n + j] * wi;
        for (int r = 0; r < n; r++) {
            const float sv = Vt[i * n + r];
            for (int j = 0; j < n; j++) Ainv[r * n + j] = (float)(Ainv[r * n + j] + (double)sv * buf[j]);
        }
    }
    const double ratio = w[0] >= 1.1920929e-07f ? (double)(w[n - 1] / w[0]) : 0;
    if (ratio == 0) {
        for (int i = 0; i < n * n; i++) Ainv[i] = 0.f;
        return false;
    }
    return true;
}

// The Gauss-Newton step of a solver whose LDL^T failed: delta = inv(H) * b as the reference computes it (float pseudo-inverse,
// Matx66f * Matx61f in float).  Hu: 21 upper-triangular sums, b: 6 sums (both double, rounded to float here).
static __device__ __noinline__ void dev_pinv_step(const double *Hu, const double *b, float delta[6])
{
    float H[36], Hi[36];
    int k = 0;
    for (int i = 0; i < 6; i++)
        for (int j = i; j < 6; j++) { H[i * 6 + j] = (float)Hu[k]; H[j * 6 + i] = (float)Hu[k]; k++; }
    dev_invert_svd6(H, Hi);   // all zeros when singular
    for (int i = 0; i < 6; i++) {
        float s = 0.f;
        for (int q = 0; q < 6; q++) s += Hi[i * 6 + q] * (float)b[q];
        delta[i] = s;
    }
}

__device__ __forceinline__ int dev_reflect101(int p, int len)
{
    if (len == 1) return 0;
    while (p < 0 || p >= len) p = (p < 0) ? -p : 2 * (len - 1) - p;
    return p;
}

// Warp-wide sums of N <= 32 per-lane values at once: lane l ends with the sum over all lanes of value l.  Every step halves the
// number of values a lane still carries (the half whose index bit matches the lane's), so 31 shuffles do what N x 5 do when each
// value is reduced on its own — and the additions pair up exactly as in the shfl_down tree (l with l^16, then ^8, ^4, ^2, ^1),
// so the result has the same bits.
template <int N>
__device__ __forceinline__ float warp_sum_scatter(const float (&v)[N], int lane)
{
    float a16[16], a8[8], a4[4], a2[2];
    const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4, b1 = lane & 2, b0 = lane & 1;
#pragma unroll
    for (int k = 0; k < 16; k++) {
        const float lo = k < N ? v[k] : 0.f, hi = k + 16 < N ? v[k + 16] : 0.f;
        const float mine = b4 ? hi : lo, send = b4 ? lo : hi;
        a16[k] = mine + __shfl_xor_sync(0xffffffffu, send, 16);
    }
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const float mine = b3 ? a16[k + 8] : a16[k], send = b3 ? a16[k] : a16[k + 8];
        a8[k] = mine + __shfl_xor_sync(0xffffffffu, send, 8);
    }
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const float mine = b2 ? a8[k + 4] : a8[k], send = b2 ? a8[k] : a8[k + 4];
        a4[k] = mine + __shfl_xor_sync(0xffffffffu, send, 4);
    }
#pragma unroll
    for (int k = 0; k < 2; k++) {
        const float mine = b1 ? a4[k + 2] : a4[k], send = b1 ? a4[k] : a4[k + 2];
        a2[k] = mine + __shfl_xor_sync(0xffffffffu, send, 2);
    }
    const float mine = b0 ? a2[1] : a2[0], send = b0 ? a2[0] : a2[1];
    return mine + __shfl_xor_sync(0xffffffffu, send, 1);
}

// block-wide sum of `NV` doubles per thread; result valid in ALL threads of warp 0 lane 0 -> written to out[]
// (fixed reduction tree => deterministic). scratch must hold NV * 32 doubles.
template <int NV>
__device__ __forceinline__ void block_reduce_sum(double (&v)[NV], double *scratch, double *out)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int k = 0; k < NV; k++) {
        double x = v[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
        if (lane == 0) scratch[k * 32 + warp] = x;
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int k = 0; k < NV; k++) {
            double x = (lane < nwarps) ? scratch[k * 32 + lane] : 0.0;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
            if (lane == 0) out[k] = x;
        }
    }
    __syncthreads();
}
// ---- thread-block clusters: mbarrier, 1-D TMA bulk copies and distributed-shared-memory pushes (alignment and refinement solvers)
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, uint32_t parity)
{
    uint32_t ok = 0;
    const uint32_t addr = smem_u32(bar);
    while (!ok) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(ok)
            : "r"(addr), "r"(parity)
            : "memory");
    }
}
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, uint32_t bytes, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ unsigned cluster_ctarank()
{
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// store a double into the shared memory of CTA `rank` of this cluster (distributed shared memory)
__device__ __forceinline__ void dsmem_store_f64(void *local_smem_ptr, unsigned rank, double v)
{
    uint32_t local = smem_u32(local_smem_ptr), remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(rank));
    asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(remote), "d"(v) : "memory");
}

// Push a double into CTA `rank`'s shared memory and signal that CTA's mbarrier with the 8 transferred bytes
// (st.async + complete_tx): data and notification travel together, no cluster-scope fence or barrier is needed.
__device__ __forceinline__ void dsmem_push_f64(void *local_slot, unsigned long long *local_bar, unsigned rank, double v)
{
    uint32_t slot = smem_u32(local_slot), bar = smem_u32(local_bar), rslot, rbar;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rslot) : "r"(slot), "r"(rank));
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rbar) : "r"(bar), "r"(rank));
    asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];" ::"r"(rslot), "l"(__double_as_longlong(v)), "r"(rbar)
                 : "memory");
}

// 2x6 Jacobian of the projection wrt the twist (pose_estimator.cpp:366-380, pose_refinement.cpp:354-366), the reference's own
// expressions with their fourteen divisions.  Taking 1/Z once and multiplying through saves a third of the instructions of a
// gradient round (measured: -0.25 us per round), but the entries then differ by an ulp, and with ONE usable keypoint the reference's
// float SVD finds H = J^T J exactly singular and takes no step, which only the reference's own bits reproduce
// (tests/test_gpu_sequences.py::test_rank_deficient_hessian[one]) — so the divisions stay.
__device__ __forceinline__ void dev_jacobian(float fx, float fy, float X, float Y, float Z, float (&J)[12])
{
    J[0] = -fx / Z; J[1] = 0.f; J[2] = fx * X / (Z * Z); J[3] = fx * X * Y / (Z * Z);
    J[4] = -fx * (1 + (X * X) / (Z * Z)); J[5] = fx * Y / Z;
    J[6] = 0.f; J[7] = -fy / Z; J[8] = fy * Y / (Z * Z); J[9] = fy * (1 + (Y * Y) / (Z * Z));
    J[10] = -fy * X * Y / (Z * Z); J[11] = -fy * X / Z;
}

// 1 / d for the pivots of the 6x6 solve: hardware reciprocal approximation (MUFU.RCP64H) refined by two Newton steps in fused
// arithmetic — a third of the latency of the IEEE division, which sits six times on the critical path of every gradient round
// of both solvers
__device__ __forceinline__ double dev_rcp_fast(double d)
{
    double x;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(d));
    x = fma(x, fma(-d, x, 1.0), x);
    x = fma(x, fma(-d, x, 1.0), x);   // >= 40 good bits even from a 10-bit seed: far below what the float step delta resolves
    return x;
}

// 6x6 SPD solve in double (LDL^T), the solvers' own arithmetic (the reference inverts in float: any accurate solve is within the
// pose tolerance), so it is written for latency: explicitly fused multiply-adds (this file is compiled with -fmad=false for the
// reference-exact float code) and one fast reciprocal per pivot.  Returns false when H is not numerically positive definite.
__device__ __forceinline__ bool dev_solve6(const double *Hu /*21 upper-tri row-major*/, const double *b, double *x)
{
    double A[6][6];
    {
        int k = 0;
#pragma unroll
        for (int i = 0; i < 6; i++)
#pragma unroll
            for (int j = i; j < 6; j++) { A[i][j] = Hu[k]; A[j][i] = Hu[k]; k++; }
    }
    double maxd = 0;
#pragma unroll
    for (int i = 0; i < 6; i++) maxd = fmax(maxd, A[i][i]);
    if (!(maxd > 0)) return false;
    // W[i][j] = L[i][j] * D[j] (the un-normalised column), L[i][j] = W[i][j] / D[j]
    double L[6][6], W[6][6], Dinv[6];
    bool ok = true;
#pragma unroll
    for (int j = 0; j < 6; j++) {
        double d = A[j][j];
#pragma unroll
        for (int q = 0; q < j; q++) d = fma(-W[j][q], L[j][q], d);
        if (!(d > 1e-13 * maxd)) { ok = false; d = 1.0; }
        Dinv[j] = dev_rcp_fast(d);
#pragma unroll
        for (int i = j + 1; i < 6; i++) {
            double t = A[i][j];
#pragma unroll
            for (int q = 0; q < j; q++) t = fma(-W[i][q], L[j][q], t);
            W[i][j] = t;
            L[i][j] = t * Dinv[j];
        }
    }
    if (!ok) return false;
    double y[6];
#pragma unroll
    for (int i = 0; i < 6; i++) {
        double t = b[i];
#pragma unroll
        for (int q = 0; q < i; q++) t = fma(-L[i][q], y[q], t);
        y[i] = t;
    }
#pragma unroll
    for (int i = 5; i >= 0; i--) {
        double t = y[i] * Dinv[i];
#pragma unroll
        for (int q = i + 1; q < 6; q++) t = fma(-L[q][i], x[q], t);
        x[i] = t;
    }
    return true;
}

// Sum of N values spaced `stride` apart as a balanced binary tree, in double.  The solvers' cross-warp and cross-CTA sums are
// chains of dependent double-precision additions on the critical path of every evaluation round (the FP64 pipe answers in ~25
// cycles): a tree of depth log2(N) instead of a chain of length N.  The grouping is fixed (deterministic) and is the solvers' own —
// the reference adds floats sequentially, and a double holds these sums of a few thousand floats exactly in all but pathological
// cases, so the grouping does not even show in the bits.
template <int N, typename T>
__device__ __forceinline__ double tree_sum(const T *v, int stride = 1)
{
    if constexpr (N == 1) return (double)v[0];
    else return tree_sum<N / 2, T>(v, stride) + tree_sum<N - N / 2, T>(v + (size_t)(N / 2) * stride, stride);
}

// developer aid (SVO_SOLVER_TRACE): one thread stamps the phases of a solver kernel with the SM clock
struct SolverTrace {
    unsigned long long *buf; int n; bool on;
    __device__ __forceinline__ SolverTrace(unsigned long long *b, bool writer) : buf(b), n(0), on(b != nullptr && writer) {}
    __device__ __forceinline__ void stamp(int tag, int extra = 0)
    {
        if (on && n < 1000) { buf[1 + n] = ((unsigned long long)clock64() << 16) | ((unsigned long long)(extra & 255) << 8) | (unsigned)(tag & 255); n++; }
    }
    __device__ __forceinline__ void finish() { if (on) buf[0] = (unsigned long long)n; }
};
#endif  // __CUDACC__

#define SVO_CUDA_CHECK(ctx_err, expr)                                                                   \
    do {                                                                                                \
        cudaError_t _e = (expr);                                                                        \
        if (_e != cudaSuccess) {                                                                        \
            snprintf((ctx_err), 256, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return SVO_ERR_CUDA;                                                                        \
        }                                                                                               \
    } while (0)
