// Sparse direct image alignment — PoseEstimator (src/lib/pose_estimator.cpp) as ONE persistent kernel.
//
// Reference structure replaced here (all per frame, on the CPU, single-threaded):
//   estimate_pose            :115-130   levels max-1 .. min
//   estimate_pose_at_level   :166-222   <= 50 cost evaluations per level, accept / halve / stop
//   do_calc                  :275-300   project_keypoints + get_total_intensity_diff (image_comparison.cpp:9-120)
//   calculate_hessian        :312-416   box-sum gradients x 2x6 Jacobian, H = sum (gJ)^T(gJ), rebuilt every call (Q1)
//   get_gradient             :418-539   residuals, b = -sum (gJ) r, delta = H^-1 b, exponential_map, rotate to world
//
// B200 design: one thread-block CLUSTER (1, 2, 4, 8 or 16 CTAs = SMs) owns the whole solve of one frame.  Per level every
// CTA stages the two level images in its shared memory with TMA bulk copies (cp.async.bulk + mbarrier); four lanes share a
// keypoint (lane r owns row r of its 4x4 patch) and do the photometric residuals and 1x6 Jacobian rows in registers with the
// reference's exact float operation order; the 21+6+1 normal-equation terms are reduced with warp shuffles, one pass per
// CTA and one DSMEM exchange per evaluation (st.async + mbarrier complete_tx) in a fixed tree (deterministic).  The 6x6
// solve (double LDL^T; the reference's float pseudo-inverse when H is not positive definite), the exponential map and the
// accept/halve/stop decisions run in the kernel, redundantly in every CTA: there is no host round trip anywhere inside a
// frame.  Pose-independent reference-patch terms (box-sum gradients and reference box sums, 48 floats + validity mask per
// keypoint) are computed once per level and cached in an L2-resident scratch array laid out [term][keypoint].
//
// Bound: dependency latency (serial evaluations) and shared-memory/ALU throughput, not HBM.
#include "kernels.cuh"

#define CLMAX 16             // largest cluster: CTAs (= SMs) cooperating on one frame; the kernel is instantiated for 1, 2, 4, 8, 16
                             // (16 is a non-portable cluster size: opt-in attribute, for frames with thousands of keypoints)
#define ALIGN_THREADS 256    // per CTA: 8 warps x 8 keypoints x 4 patch rows
#define ALIGN_WARPS (ALIGN_THREADS / 32)
#define NGRAD 27             // 21 (H upper triangle) + 6 (b)

#define ALIGN_DEPTH 16       // halving steps whose rotation matrices are prepared right after a gradient (x0 + 2^-j grad, j < 16)
struct AlignHdr {
    unsigned long long bar;                // TMA completion barrier (level images)
    unsigned long long xbar[3];            // DSMEM exchange barriers: cost (two, alternating) and gradient
    double Rd[2 * ALIGN_DEPTH + 2][9];     // Rodrigues(-r): two tables of ALIGN_DEPTH step sizes (k = 2^-j) + 2 spare slots
    double warp_cost[ALIGN_WARPS];         // this CTA's per-warp cost partials
    double cl_cost[2][CLMAX];              // per-CTA cost partials of the whole cluster ([group][CTA of the group], written through DSMEM), double buffered
    float warp_grad[ALIGN_WARPS][NGRAD];   // this CTA's per-warp partials of H (21) and b (6)
    double cl_grad[CLMAX][NGRAD];             // per-CTA partials of the whole cluster (written through DSMEM)
    double red_out[NGRAD];
    float grad[6];
    int n;
};
#define HDR_BYTES ((sizeof(AlignHdr) + 127) / 128 * 128)

// developer aid: same wait, but after ~2 s of polling it records where it was stuck in dbg[] (mapped host memory) and traps
__device__ __noinline__ void mbar_wait_dbg(unsigned long long *bar, uint32_t parity, int *dbg, int tag, int crank, int level, int mode, int evals, int grads)
{
    uint32_t ok = 0;
    const uint32_t addr = smem_u32(bar);
    long long t0 = clock64();
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
        if (!ok && clock64() - t0 > 4000000000ll) {
            if (atomicAdd(dbg, 1) < 4) {
                int *q = dbg + 1 + 10 * (crank & 3);
                q[0] = tag; q[1] = crank; q[2] = threadIdx.x; q[3] = (int)parity; q[4] = level; q[5] = mode; q[6] = evals; q[7] = grads;
                unsigned long long st;
                asm volatile("ld.shared.b64 %0, [%1];" : "=l"(st) : "r"(addr));
                q[8] = (int)(st & 0xffffffffu); q[9] = (int)(st >> 32);
                __threadfence_system();
            }
            __nanosleep(1000000);
            __trap();
        }
    }
}

// exact u8 -> float without the slow I2F path: 0x4B000000 | v is the float 8388608 + v
__device__ __forceinline__ float u8f(unsigned v) { return __uint_as_float(0x4B000000u | v) - 8388608.0f; }

// five consecutive pixels at any byte address as exact floats: two aligned word loads instead of five byte loads (the lanes of a
// warp read scattered patches, so every load instruction costs its wavefronts on the one shared-memory pipe of the SM), the bytes
// picked out with PRMT straight into the 0x4B000000 | v form of u8f.  Reads up to 3 bytes past the fifth pixel: inside the
// 16 bytes of slack every staged / stored level has.
__device__ __forceinline__ void load5f(const uint8_t *p, float (&o)[5])
{
    const uintptr_t addr = reinterpret_cast<uintptr_t>(p);
    const uint32_t *w = reinterpret_cast<const uint32_t *>(addr & ~(uintptr_t)3);
    const uint32_t w0 = w[0], w1 = w[1];
    const int sh = (int)(addr & 3) * 8;
    const uint32_t t0 = __funnelshift_r(w0, w1, sh);        // pixels 0..3
    const uint32_t t1 = sh ? (w1 >> sh) : w1;                // pixel 4 in its low byte (sh == 0: first byte of the second word)
    o[0] = __uint_as_float(__byte_perm(t0, 0x4B000000u, 0x7440)) - 8388608.0f;
    o[1] = __uint_as_float(__byte_perm(t0, 0x4B000000u, 0x7441)) - 8388608.0f;
    o[2] = __uint_as_float(__byte_perm(t0, 0x4B000000u, 0x7442)) - 8388608.0f;
    o[3] = __uint_as_float(__byte_perm(t0, 0x4B000000u, 0x7443)) - 8388608.0f;
    o[4] = __uint_as_float(__byte_perm(t1, 0x4B000000u, 0x7440)) - 8388608.0f;
}

// get_patch_sum (pose_estimator.cpp:82-112), operation for operation
__device__ __forceinline__ float patch_sum(const uint8_t *img, int pitch, float cx, float cy)
{
    float sx = cx - 0.5f, sy = cy - 0.5f;
    float fxf = floorf(sx), fyf = floorf(sy);
    int ipx = (int)fxf, ipy = (int)fyf;
    float x2 = sx - fxf, y2 = sy - fyf;   // (float)ipx == fxf exactly
    float x1 = 1.0f - x2, y1 = 1.0f - y2; // == (float)(1.0 - (double)x2): x2 in [0,1) makes the float subtraction exact or equally rounded
    const uint8_t *s1 = img + ipy * pitch + ipx;
    const uint8_t *s2 = s1 + pitch;
    const uint8_t *s3 = s2 + pitch;
    float v = x1 * y1 * u8f(s1[0]);
    v = v + y1 * u8f(s1[1]);
    v = v + x2 * y1 * u8f(s1[2]);
    v = v + x1 * u8f(s2[0]);
    v = v + u8f(s2[1]);
    v = v + x2 * u8f(s2[2]);
    v = v + x1 * y2 * u8f(s3[0]);
    v = v + y2 * u8f(s3[1]);
    v = v + x2 * y2 * u8f(s3[2]);
    return v;
}

// six consecutive pixels at any byte address as exact floats (three aligned words; reads up to 11 bytes past p: inside the 16
// bytes of slack every staged / stored level has)
__device__ __forceinline__ void load6f(const uint8_t *p, float (&o)[6])
{
    const uintptr_t addr = reinterpret_cast<uintptr_t>(p);
    const uint32_t *w = reinterpret_cast<const uint32_t *>(addr & ~(uintptr_t)3);
    const uint32_t w0 = w[0], w1 = w[1], w2 = w[2];
    const int sh = (int)(addr & 3) * 8;
    const uint32_t t0 = __funnelshift_r(w0, w1, sh), t1 = __funnelshift_r(w1, w2, sh);
    o[0] = __uint_as_float(__byte_perm(t0, 0x4B000000u, 0x7440)) - 8388608.0f;
    o[1] = __uint_as_float(__byte_perm(t0, 0x4B000000u, 0x7441)) - 8388608.0f;
    o[2] = __uint_as_float(__byte_perm(t0, 0x4B000000u, 0x7442)) - 8388608.0f;
    o[3] = __uint_as_float(__byte_perm(t0, 0x4B000000u, 0x7443)) - 8388608.0f;
    o[4] = __uint_as_float(__byte_perm(t1, 0x4B000000u, 0x7440)) - 8388608.0f;
    o[5] = __uint_as_float(__byte_perm(t1, 0x4B000000u, 0x7441)) - 8388608.0f;
}

// get_patch_sum at the four pixels (cx, cy), (cx + 1, cy), ... of one patch row (the positions the residual loop of get_gradient
// visits by repeated `+= 1.f`): their 3x3 footprints overlap, so the 3 x 6 pixels are fetched once (nine word loads instead of
// thirty-six byte loads); every sum is then formed from its own weights in patch_sum's operation order — the same bits.  Returns
// false (nothing computed) when the four footprints are not simply one pixel apart (a float increment that crossed a binade and
// changed the integer part differently): the caller then takes the one-by-one route.
__device__ __forceinline__ bool patch_sum_row4(const uint8_t *img, int pitch, float cx, float cy, float (&out)[4])
{
    float px[4] = {cx, 0.f, 0.f, 0.f};
#pragma unroll
    for (int c = 1; c < 4; c++) px[c] = px[c - 1] + 1.f;
    const float sy = cy - 0.5f, fyf = floorf(sy);
    const int ipy = (int)fyf;
    const float y2 = sy - fyf, y1 = 1.0f - y2;
    int ipx[4];
    float x2[4];
#pragma unroll
    for (int c = 0; c < 4; c++) {
        const float sx = px[c] - 0.5f, fxf = floorf(sx);
        ipx[c] = (int)fxf;
        x2[c] = sx - fxf;
    }
    if (ipx[1] != ipx[0] + 1 || ipx[2] != ipx[0] + 2 || ipx[3] != ipx[0] + 3) return false;
    const uint8_t *s1 = img + ipy * pitch + ipx[0];
    float r0[6], r1[6], r2[6];
    load6f(s1, r0); load6f(s1 + pitch, r1); load6f(s1 + 2 * pitch, r2);
#pragma unroll
    for (int c = 0; c < 4; c++) {
        const float x1 = 1.0f - x2[c];
        float v = x1 * y1 * r0[c];
        v = v + y1 * r0[c + 1];
        v = v + x2[c] * y1 * r0[c + 2];
        v = v + x1 * r1[c];
        v = v + r1[c + 1];
        v = v + x2[c] * r1[c + 2];
        v = v + x1 * y2 * r2[c];
        v = v + y2 * r2[c + 1];
        v = v + x2[c] * y2 * r2[c + 2];
        out[c] = v;
    }
    return true;
}

// _get_intensity_diff (image_comparison.cpp:9-91); PATCH > 0: compile-time window (footprints cached in registers)
template <int PATCH>
__device__ __forceinline__ float intensity_diff(const uint8_t *im1, const uint8_t *im2, int w, int h, int pitch, float c1x, float c1y,
                                                float c2x, float c2y, int patch_rt)
{
    const int patch = PATCH > 0 ? PATCH : patch_rt;
    float half = ((float)patch - 1.0f) / 2.0f;
    float s1x = c1x - half, s1y = c1y - half, s2x = c2x - half, s2y = c2y - half;
    float f1x = floorf(s1x), f1y = floorf(s1y), f2x = floorf(s2x), f2y = floorf(s2y);
    int ip1x = (int)f1x, ip1y = (int)f1y, ip2x = (int)f2x, ip2y = (int)f2y;
    float x12 = s1x - f1x, y12 = s1y - f1y, x22 = s2x - f2x, y22 = s2y - f2y;
    float x11 = 1.0f - x12, y11 = 1.0f - y12;
    float x21 = 1.0f - x22, y21 = 1.0f - y22;
    float m11 = x11 * y11, m12 = x12 * y11, m13 = x11 * y12, m14 = x12 * y12;
    float m21 = x21 * y21, m22 = x22 * y21, m23 = x21 * y22, m24 = x22 * y22;
    float intensity = 0.f;
    if (ip1y >= 0 && ip1y + patch < h && ip2y >= 0 && ip2y + patch < h && ip1x >= 0 && ip1x + patch < w && ip2x >= 0 &&
        ip2x + patch < w) {
        if (PATCH > 0) {
            float a[PATCH + 1][PATCH + 1], b[PATCH + 1][PATCH + 1];
            const uint8_t *p1 = im1 + ip1y * pitch + ip1x, *p2 = im2 + ip2y * pitch + ip2x;
#pragma unroll
            for (int i = 0; i <= PATCH; i++)
#pragma unroll
                for (int j = 0; j <= PATCH; j++) { a[i][j] = u8f(p1[i * pitch + j]); b[i][j] = u8f(p2[i * pitch + j]); }
#pragma unroll
            for (int i = 0; i < PATCH; i++)
#pragma unroll
                for (int j = 0; j < PATCH; j++) {
                    float i1 = 0.f, i2 = 0.f;
                    i1 += m11 * a[i][j]; i1 += m12 * a[i][j + 1]; i1 += m13 * a[i + 1][j]; i1 += m14 * a[i + 1][j + 1];
                    i2 += m21 * b[i][j]; i2 += m22 * b[i][j + 1]; i2 += m23 * b[i + 1][j]; i2 += m24 * b[i + 1][j + 1];
                    intensity += fabsf(i1 - i2);
                }
        } else {
            for (int i = 0; i < patch; i++) {
                const uint8_t *s11 = im1 + (i + ip1y) * pitch + ip1x, *s12 = s11 + pitch;
                const uint8_t *s21 = im2 + (i + ip2y) * pitch + ip2x, *s22 = s21 + pitch;
                for (int j = 0; j < patch; j++) {
                    float i1 = 0.f, i2 = 0.f;
                    i1 += m11 * u8f(s11[j]); i1 += m12 * u8f(s11[j + 1]); i1 += m13 * u8f(s12[j]); i1 += m14 * u8f(s12[j + 1]);
                    i2 += m21 * u8f(s21[j]); i2 += m22 * u8f(s21[j + 1]); i2 += m23 * u8f(s22[j]); i2 += m24 * u8f(s22[j + 1]);
                    intensity += fabsf(i1 - i2);
                }
            }
        }
    }
    return intensity;
}

// One row (4 pixels) of _get_intensity_diff (image_comparison.cpp:9-91) for the 4x4 window of the shipped configurations, in two
// halves: the reference-image half (bilinear samples i1 of the previous image at the keypoint: pose-independent, so a lane keeps
// them in registers for a whole level) and the current-image half (samples at the projection, |i1 - i2|).  Same operations on the
// same values as the one-piece routine: the same bits.  Other window sizes use the generic whole-patch routine on lane 0 of the
// quad (d0 carries the whole patch sum, in the reference's order).
__device__ __forceinline__ void ref_row4(const uint8_t *im1, int w, int h, int pitch, float c1x, float c1y, int row, float (&i1)[4], bool &ok)
{
    float half = ((float)4 - 1.0f) / 2.0f;
    float s1x = c1x - half, s1y = c1y - half;
    float f1x = floorf(s1x), f1y = floorf(s1y);
    int ip1x = (int)f1x, ip1y = (int)f1y;
    float x12 = s1x - f1x, y12 = s1y - f1y;
    float x11 = 1.0f - x12, y11 = 1.0f - y12;
    float m11 = x11 * y11, m12 = x12 * y11, m13 = x11 * y12, m14 = x12 * y12;
    ok = ip1y >= 0 && ip1y + 4 < h && ip1x >= 0 && ip1x + 4 < w;
#pragma unroll
    for (int j = 0; j < 4; j++) i1[j] = 0.f;
    if (ok) {
        const uint8_t *p1 = im1 + (ip1y + row) * pitch + ip1x;
        float a0[5], a1[5];
        load5f(p1, a0); load5f(p1 + pitch, a1);
#pragma unroll
        for (int j = 0; j < 4; j++) {
            float t = 0.f;
            t += m11 * a0[j]; t += m12 * a0[j + 1]; t += m13 * a1[j]; t += m14 * a1[j + 1];
            i1[j] = t;
        }
    }
}
__device__ __forceinline__ void cur_row4(const uint8_t *im2, int w, int h, int pitch, float c2x, float c2y, int row, const float (&i1)[4], bool ok1,
                                         float &d0, float &d1, float &d2, float &d3)
{
    d0 = d1 = d2 = d3 = 0.f;
    float half = ((float)4 - 1.0f) / 2.0f;
    float s2x = c2x - half, s2y = c2y - half;
    float f2x = floorf(s2x), f2y = floorf(s2y);
    int ip2x = (int)f2x, ip2y = (int)f2y;
    float x22 = s2x - f2x, y22 = s2y - f2y;
    float x21 = 1.0f - x22, y21 = 1.0f - y22;
    float m21 = x21 * y21, m22 = x22 * y21, m23 = x21 * y22, m24 = x22 * y22;
    if (ok1 && ip2y >= 0 && ip2y + 4 < h && ip2x >= 0 && ip2x + 4 < w) {
        const uint8_t *p2 = im2 + (ip2y + row) * pitch + ip2x;
        float b0[5], b1[5];
        load5f(p2, b0); load5f(p2 + pitch, b1);
        float d[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            float i2 = 0.f;
            i2 += m21 * b0[j]; i2 += m22 * b0[j + 1]; i2 += m23 * b1[j]; i2 += m24 * b1[j + 1];
            d[j] = fabsf(i1[j] - i2);
        }
        d0 = d[0]; d1 = d[1]; d2 = d[2]; d3 = d[3];
    }
}
// window sizes other than 4: the generic whole-patch routine on lane 0 of the quad (the whole patch sum in the reference's order);
// kept out of line — no shipped configuration takes it, and the solver loop has to stay small
static __device__ __noinline__ float intensity_diff_generic(const uint8_t *im1, const uint8_t *im2, int w, int h, int pitch, float c1x, float c1y,
                                                            float c2x, float c2y, int patch)
{
    return intensity_diff<0>(im1, im2, w, h, pitch, c1x, c1y, c2x, c2y, patch);
}

// One thread-block CLUSTER (8 CTAs x 256 threads) = one frame.  Four lanes share a keypoint: lane r of the quad owns
// row r of its 4x4 patch.  kSmem: level images staged in every CTA's shared memory (TMA), else read via L1/L2.
//
// G > 1 (svo_set_solver_width): the cluster has G groups of CL CTAs.  Every group covers ALL keypoints exactly like a cluster of CL
// CTAs would (same keypoint -> thread mapping, same reduction trees: the same bits); gradients are evaluated by every group for
// itself, and in a cost round group g evaluates trial pose x0 + 2^-(j+g) grad.  A line search of a running sequence is a run of
// rejected-and-halved steps ("G a G r r a G r r r r r r s": 22 cost evaluations for 7 gradients per C3 frame), every trial pose and
// its matrix is known once the gradient is, and a rejected step decides nothing but "try the next one": the G costs of a round are
// exchanged across the whole cluster and every thread replays the accept / halve / stop rule of estimate_pose_at_level over them
// in order.  Same evaluations, decisions and counters as the sequential search; a run of rejections costs 1/G of the rounds.
template <bool kSmem, int CL, int G>
__global__ void __launch_bounds__(ALIGN_THREADS, 1) sparse_align_kernel(AlignArgs a)
{
    static_assert(CL * G <= CLMAX, "cluster too large");
    constexpr int KPS_PER_PASS = CL * ALIGN_WARPS * 8;
    extern __shared__ __align__(128) uint8_t smem_raw[];
    AlignHdr *hdr = reinterpret_cast<AlignHdr *>(smem_raw);
    uint8_t *img_area = smem_raw + HDR_BYTES;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int quad = lane >> 2, row = lane & 3;
    const unsigned prank = cluster_ctarank();                       // rank in the physical cluster of CL * G CTAs
    const unsigned grp = G > 1 ? prank / CL : 0, crank = G > 1 ? prank % CL : prank;   // group (trial pose of a cost round), rank in the group
    const DevCam cam = a.cam;
    if (tid == 0) {
        if (kSmem) mbar_init(&hdr->bar, 1);
        for (int k = 0; k < 3; k++) mbar_init(&hdr->xbar[k], 1);
        hdr->n = min(*a.n_ptr, a.max_kps);
        if (prank == 0) for (int k = 0; k < 16; k++) a.evals_out[k] = 0;
    }
    __syncthreads();
    cluster_sync_all();   // every CTA's shared memory is live before anyone writes into it remotely
    const int n = hdr->n;
    SolverTrace trc(a.trace, tid == 0 && prank == 0);
    trc.stamp(0);
    const int npass = (n + KPS_PER_PASS - 1) / KPS_PER_PASS;
    const bool nodist = dev_cam_nodist(cam);
    uint32_t phase = 0, xphase[3] = {0, 0, 0};
    // cost-exchange buffer / barrier of the NEXT evaluation.  It must keep alternating across level boundaries: a CTA may
    // push evaluation e+2 into a peer only after passing the wait of e+1, and every peer pushes e+1 only after all its
    // threads have read the table of e — so buffer e % 2 is free again exactly two evaluations later, never one.
    int cbuf = 0;
    float *scr = a.scratch;                 // [13][4 * max_kps]: per (keypoint,row): 4 x (g0, g1, Sprev) + mask
    const size_t SN = (size_t)4 * a.max_kps;

    // solver state: identical in every thread of every CTA (recomputed redundantly from the broadcast partial sums)
    float x0[6], xt[6], grad[6];
#pragma unroll
    for (int k = 0; k < 6; k++) { x0[k] = a.pose_in[k]; xt[k] = x0[k]; grad[k] = 0.f; }
    float prev_cost = 0.f;
    int x0slot = 2 * ALIGN_DEPTH, sp = 1, tb = 0, xtslot = 0;

    const int lv_hi = (a.probe_level >= 0) ? a.probe_level : cam.max_levels - 1;
    const int lv_lo = (a.probe_level >= 0) ? a.probe_level : cam.min_level;

    for (int level = lv_hi; level >= lv_lo; level--) {
        const LevelDesc P = a.prev[level], C = a.cur[level];
        const int w = P.w, h = P.h;
        const uint8_t *pimg, *cimg;
        if (kSmem) {
            const uint32_t bytes16 = ((uint32_t)(w * h) + 15u) & ~15u;
            __syncthreads();  // everyone finished reading the previous level's tiles
            if (tid == 0) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_expect_tx(&hdr->bar, 2 * bytes16);
                tma_load_1d(img_area, P.ptr, bytes16, &hdr->bar);
                tma_load_1d(img_area + bytes16, C.ptr, bytes16, &hdr->bar);
            }
            pimg = img_area; cimg = img_area + bytes16;
        } else {
            pimg = P.ptr; cimg = C.ptr;
        }
        const int pitch = w;  // halfSample levels are stored with pitch == width
        // the reference's footprint tests promote to double — ((double)x - 2.0) < 0, ((double)x + 3.0) >= w, ... — and a float plus
        // or minus a small integer is exact in double, so they are exactly x < 2.f, x >= (float)(w - 3), ...: float compares, no
        // conversions and no trip through the double-precision pipe (12 per keypoint row and gradient evaluation)
        const float wm3 = (float)(w - 3), hm3 = (float)(h - 3), wm2 = (float)(w - 2), hm2 = (float)(h - 2);
        trc.stamp(1, level);

        // setLevel (pose_estimator.cpp:541-562)
        const int divider = 1 << level;
        const float fdiv = (float)divider;
        const float lfx = cam.fx / fdiv, lfy = cam.fy / fdiv, lcx = cam.cx / fdiv, lcy = cam.cy / fdiv;

        // Rodrigues matrices: after every gradient, 8 warps compute in parallel the matrices of the 8 trial poses
        // x0 + 2^-j grad (table tb); cost evaluations then find theirs ready, and an accepted pose keeps its slot as
        // "the matrix of x0" while the next gradient fills the other table.  j >= 8 falls back to a spare slot.
        if (level == lv_hi) {
            if (tid == 0) dev_rodrigues_d(-x0[3], -x0[4], -x0[5], hdr->Rd[2 * ALIGN_DEPTH]);   // overlaps the TMA copy
            x0slot = 2 * ALIGN_DEPTH; sp = 1; tb = 0;
        }
        if (kSmem) {
            if (a.dbg) mbar_wait_dbg(&hdr->bar, phase, a.dbg, 1, (int)prank, level, -1, 0, 0);
            else mbar_wait(&hdr->bar, phase);
            phase ^= 1;
        }
        trc.stamp(2, level);

        // ---- per-level cache of the pose-independent reference terms of this lane's patch row
        //      (calculate_hessian :346-395 image part, get_gradient :449-460 reference part)
        for (int pass = 0; pass < npass; pass++) {
            const int i = ((pass * ALIGN_WARPS + warp) * CL + (int)crank) * 8 + quad;
            if (i >= n) continue;
            if (a.flags && (a.flags[i] & SVO_F_IGN_TEMP)) continue;
            float bx = a.kps2d[2 * i], by = a.kps2d[2 * i + 1];
            if (level != 0) { bx = bx / fdiv; by = by / fdiv; }
            float kx = bx - 2.f, ky = by - 2.f;   // Hessian loop coordinates
            float rx = bx - 2.f, ry = by - 2.f;   // residual loop reference coordinates
            // replay the reference's float increments of the rows above this lane's row
            for (int r = 0; r < row; r++) {
                kx += 1.f; kx += 1.f; kx += 1.f; kx += 1.f; kx -= 4.f; ky += 1.f;
                rx += 1.f; rx += 1.f; rx += 1.f; rx += 1.f; ry += 1.f; rx -= 4.f;
            }
            unsigned mask = 0;
            const size_t slot = (size_t)4 * i + row;
            // (a loop, not four unrolled copies of five inlined patch sums: this code runs once per level, i.e. always from a cold
            // instruction cache, and what it costs is the fetch of its own instructions)
#pragma unroll 1
            for (int c = 0; c < 4; c++) {
                float g0 = 0.f, g1 = 0.f;
                if (!(kx < 2.f || ky < 2.f || kx >= wm3 || ky >= hm3)) {   // ((double)kx - 2.0) < 0 || ... || ((double)kx + 3.0) >= w: see wm3
                    float i1 = patch_sum(pimg, pitch, kx + 1.f, ky), i2 = patch_sum(pimg, pitch, kx - 1.f, ky);
                    float i3 = patch_sum(pimg, pitch, kx, ky + 1.f), i4 = patch_sum(pimg, pitch, kx, ky - 1.f);
                    g0 = i1 - i2; g1 = i3 - i4;
                    mask |= 1u << c;
                }
                scr[(size_t)c * SN + slot] = g0;
                scr[(size_t)(4 + c) * SN + slot] = g1;
                kx += 1.f;
                float sp = 0.f;
                if (!(rx < 1.f || ry < 1.f || rx > wm2 || ry > hm2)) {
                    sp = patch_sum(pimg, pitch, rx, ry);
                    mask |= 1u << (4 + c);
                }
                scr[(size_t)(8 + c) * SN + slot] = sp;
                rx += 1.f;
            }
            scr[(size_t)12 * SN + slot] = __uint_as_float(mask);
        }
        // (each thread only ever reads back the scratch entries it wrote itself)
        __syncthreads();  // Rd of x0 visible
        trc.stamp(3, level);

        // pose-independent inputs of this lane's keypoint of the first pass (all of them for up to 64 x CL keypoints), kept in
        // registers for the whole level: the ~30 evaluations of a level re-read and re-divide nothing
        const int i_first = (warp * CL + (int)crank) * 8 + quad;
        const bool act_first = (i_first < n) && !(a.flags && (a.flags[i_first] & SVO_F_IGN_TEMP));
        float bx_first = 0.f, by_first = 0.f, P_first[3] = {0.f, 0.f, 0.f};
        if (act_first) {
            bx_first = a.kps2d[2 * i_first]; by_first = a.kps2d[2 * i_first + 1];
            if (level != 0) { bx_first = bx_first / fdiv; by_first = by_first / fdiv; }   // setLevel :558-561
            P_first[0] = a.kps3d[3 * i_first]; P_first[1] = a.kps3d[3 * i_first + 1]; P_first[2] = a.kps3d[3 * i_first + 2];
        }
        // ... and so do the bilinear samples of this lane's row of the reference patch (4x4 window)
        const bool win4 = cam.win_pose == 4;
        float i1_first[4] = {0.f, 0.f, 0.f, 0.f};
        bool ok1_first = false;
        if (act_first && win4) ref_row4(pimg, w, h, pitch, bx_first, by_first, row, i1_first, ok1_first);

        // ---- Gauss-Newton driver (estimate_pose_at_level :166-222).
        //  mode 0: cost at x0 (initial)   mode 1: gradient at x0   mode 2: cost at xt   mode 3: level done
        int mode = 0, it = 0, n_evals = 0, n_grads = 0, jstep = 0;
        float kstep = 1.f;
        while (mode != 3) {
            if (mode != 1) {
                // ---------------- do_calc: bilinear SAD (prev @ reference position, cur @ projection)
                // trials of this round: group g takes halving step jstep + g as long as its matrix is in the table, the iteration
                // budget allows it (it + g < 50) and the pose is a trial at all (mode 2, table slot); otherwise every group
                // evaluates the same pose (the bits are the same) and group 0's result is used
                int nvalid = 1;
                if (G > 1 && mode == 2 && xtslot < 2 * ALIGN_DEPTH) { while (nvalid < G && jstep + nvalid < ALIGN_DEPTH && it + nvalid < 50) nvalid++; }
                const int g_mine = (G > 1 && nvalid > 1) ? (int)grp : 0;
                const bool mine = g_mine < nvalid;
                float xg[6], kg = kstep;
                for (int q = 0; q < g_mine; q++) kg = kg / 2;
#pragma unroll
                for (int k = 0; k < 6; k++) xg[k] = (mode == 2) ? (g_mine == 0 ? xt[k] : x0[k] + (kg * grad[k])) : x0[k];
                const double *Rd = hdr->Rd[(mode == 2) ? (mine ? xtslot + g_mine : xtslot) : x0slot];
                const float tx = xg[0], ty = xg[1], tz = xg[2];
                double part = 0.0;
                // (one loop body for the register-held first pass and the later ones)
                for (int pass = 0; mine && pass < npass; pass++) {
                    const int i = ((pass * ALIGN_WARPS + warp) * CL + (int)crank) * 8 + quad;
                    const bool active = pass == 0 ? act_first : ((i < n) && !(a.flags && (a.flags[i] & SVO_F_IGN_TEMP)));
                    float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
                    if (active) {
                        float bx = bx_first, by = by_first, Px = P_first[0], Py = P_first[1], Pz = P_first[2];
                        float i1[4] = {i1_first[0], i1_first[1], i1_first[2], i1_first[3]};
                        bool ok1 = ok1_first;
                        if (pass != 0) {
                            bx = a.kps2d[2 * i]; by = a.kps2d[2 * i + 1];
                            if (level != 0) { bx = bx / fdiv; by = by / fdiv; }
                            Px = a.kps3d[3 * i]; Py = a.kps3d[3 * i + 1]; Pz = a.kps3d[3 * i + 2];
                            if (win4) ref_row4(pimg, w, h, pitch, bx, by, row, i1, ok1);
                        }
                        float u, v;
                        dev_project_nd(nodist, Rd, Px, Py, Pz, tx, ty, tz, lfx, lfy, lcx, lcy, cam, u, v);
                        if (win4) cur_row4(cimg, w, h, pitch, u, v, row, i1, ok1, d0, d1, d2, d3);
                        else if (row == 0) d0 = intensity_diff_generic(pimg, cimg, w, h, pitch, bx, by, u, v, cam.win_pose);
                    }
                    // the reference adds the 16 |dI| terms of a patch sequentially in raster order: chain the four rows
                    float run = 0.f;
#pragma unroll
                    for (int rr = 0; rr < 4; rr++) {
                        float mine_sum = run;
                        mine_sum += d0; mine_sum += d1; mine_sum += d2; mine_sum += d3;
                        run = __shfl_sync(0xffffffffu, mine_sum, (lane & ~3) | rr);
                    }
                    if (row == 0) part += (double)run;   // one lane per keypoint carries the patch cost
                }
                // only the eight quad leaders of a warp (row == 0) carry a value: three butterfly steps instead of five
                part += __shfl_xor_sync(0xffffffffu, part, 4);
                part += __shfl_xor_sync(0xffffffffu, part, 8);
                part += __shfl_xor_sync(0xffffffffu, part, 16);
                if (lane == 0) hdr->warp_cost[warp] = part;
                __syncthreads();
                if (tid < CL * G) {   // lane t pushes this CTA's partial into CTA t's table and signals CTA t's barrier (every CTA pushes in
                                      // every round, with or without a trial of its own: see the note on cbuf)
                    const double cta = tree_sum<ALIGN_WARPS>(hdr->warp_cost);
                    dsmem_push_f64(&hdr->cl_cost[cbuf][grp * CL + crank], &hdr->xbar[cbuf], (unsigned)tid, cta);
                    if (tid == 0) mbar_expect_tx(&hdr->xbar[cbuf], CL * G * 8);
                }
                if (a.dbg) mbar_wait_dbg(&hdr->xbar[cbuf], xphase[cbuf], a.dbg, 2 + cbuf, (int)prank, level, mode, n_evals, n_grads);
                else mbar_wait(&hdr->xbar[cbuf], xphase[cbuf]);
                xphase[cbuf] ^= 1;
                const double *table = hdr->cl_cost[cbuf];
                cbuf ^= 1;
                trc.stamp(4, nvalid);
                // replay of the sequential driver over the costs of this round, in order
                for (int g = 0; g < nvalid && mode != 3 && mode != 1; g++) {
                    const float cost = (float)tree_sum<CL>(table + g * CL);
                    n_evals++;
                    if (mode == 0) {
                        prev_cost = cost;
                        mode = 1;   // it == 0 < 50
                    } else if (cost < prev_cost) {
#pragma unroll
                        for (int k = 0; k < 6; k++) x0[k] = xt[k];
                        prev_cost = cost;
                        it++;       // outer loop increment after `break`
                        mode = (it < 50) ? 1 : 3;
                        x0slot = xtslot;                 // the accepted pose keeps its matrix
                        if (xtslot >= 2 * ALIGN_DEPTH) sp ^= 1; else tb ^= 1;   // ... and the next gradient / fallback writes elsewhere
                    } else if (fabsf(cost - prev_cost) < 1.0f) {
                        mode = 3;
                    } else {
                        kstep = kstep / 2;
                        jstep++;
                        it++;       // inner loop increment
                        if (it < 50) {
#pragma unroll
                            for (int k = 0; k < 6; k++) xt[k] = x0[k] + (kstep * grad[k]);
                            if (jstep < ALIGN_DEPTH) {
                                xtslot = tb * ALIGN_DEPTH + jstep;
                            } else {   // rare: ALIGN_DEPTH halvings and more — compute on demand (g is the last trial of its round here)
                                xtslot = 2 * ALIGN_DEPTH + sp;
                                if (tid == 0) dev_rodrigues_d(-xt[3], -xt[4], -xt[5], hdr->Rd[xtslot]);
                                __syncthreads();
                            }
                            mode = 2;
                        } else {
                            mode = 3;
                        }
                    }
                }
            } else {
                // ---------------- get_gradient at x0 (Hessian rebuilt on every call); every group for itself
                const double *Rd = hdr->Rd[x0slot];
                const float tx = x0[0], ty = x0[1], tz = x0[2];
                float Rif[9];
#pragma unroll
                for (int k = 0; k < 9; k++) Rif[k] = (float)Rd[k];   // inv_rot_mat = float(Rodrigues(-r))
                float acc[NGRAD];
#pragma unroll
                for (int k = 0; k < NGRAD; k++) acc[k] = 0.f;
                for (int pass = 0; pass < npass; pass++) {
                    const int i = ((pass * ALIGN_WARPS + warp) * CL + (int)crank) * 8 + quad;
                    if (pass == 0 ? !act_first : ((i >= n) || (a.flags && (a.flags[i] & SVO_F_IGN_TEMP)))) continue;
                    const size_t slot = (size_t)4 * i + row;
                    const unsigned mask = __float_as_uint(scr[(size_t)12 * SN + slot]);
                    float g0v[4], g1v[4], spv[4];
#pragma unroll
                    for (int c = 0; c < 4; c++) { g0v[c] = scr[(size_t)c * SN + slot]; g1v[c] = scr[(size_t)(4 + c) * SN + slot]; spv[c] = scr[(size_t)(8 + c) * SN + slot]; }
                    const float Px = pass == 0 ? P_first[0] : a.kps3d[3 * i], Py = pass == 0 ? P_first[1] : a.kps3d[3 * i + 1],
                                Pz = pass == 0 ? P_first[2] : a.kps3d[3 * i + 2];
                    float u, v;
                    dev_project_nd(nodist, Rd, Px, Py, Pz, tx, ty, tz, lfx, lfy, lcx, lcy, cam, u, v);
                    float X, Y, Z;
                    dev_m33v(Rif, Px - tx, Py - ty, Pz - tz, X, Y, Z);
                    float J[12];
                    dev_jacobian(lfx, lfy, X, Y, Z, J);
                    float qx = u - 2.f, qy = v - 2.f;       // residual loop, current-image coordinates
                    for (int r = 0; r < row; r++) { qx += 1.f; qx += 1.f; qx += 1.f; qx += 1.f; qy += 1.f; qx -= 4.f; }
                    // residuals of the four pixels of this row: all four inside the image (the usual case) -> one shared fetch
                    float diffv[4] = {0.f, 0.f, 0.f, 0.f};
                    {
                        float qxc[4] = {qx, 0.f, 0.f, 0.f};
#pragma unroll
                        for (int c = 1; c < 4; c++) qxc[c] = qxc[c - 1] + 1.f;
                        bool all_in = ((mask >> 4) & 15u) == 15u && !(qy < 1.f || qy > hm2);
#pragma unroll
                        for (int c = 0; c < 4; c++) all_in = all_in && !(qxc[c] < 1.f || qxc[c] > wm2);
                        float ps[4];
                        if (all_in && patch_sum_row4(cimg, pitch, qx, qy, ps)) {
#pragma unroll
                            for (int c = 0; c < 4; c++) diffv[c] = ps[c] - spv[c];
                        } else {   // a row that leaves the image: pixel by pixel (rare: a loop, not four more inlined patch sums)
#pragma unroll 1
                            for (int c = 0; c < 4; c++) {
                                const float qc = c == 0 ? qxc[0] : c == 1 ? qxc[1] : c == 2 ? qxc[2] : qxc[3];
                                const float sv = c == 0 ? spv[0] : c == 1 ? spv[1] : c == 2 ? spv[2] : spv[3];
                                float dv = 0.f;
                                if ((mask & (1u << (4 + c))) && !(qc < 1.f || qy < 1.f || qc > wm2 || qy > hm2)) dv = patch_sum(cimg, pitch, qc, qy) - sv;
                                if (c == 0) diffv[0] = dv; else if (c == 1) diffv[1] = dv; else if (c == 2) diffv[2] = dv; else diffv[3] = dv;
                            }
                        }
                    }
#pragma unroll
                    for (int c = 0; c < 4; c++) {
                        float gj[6];
#pragma unroll
                        for (int k = 0; k < 6; k++) {
                            const float t = __fmaf_rn(g1v[c], J[6 + k], g0v[c] * J[k]);   // enters H and b only (regrouped sums)
                            gj[k] = (mask & (1u << c)) ? t : 0.f;
                        }
                        const float diff = diffv[c];
                        int hk = 0;
#pragma unroll
                        for (int p = 0; p < 6; p++)
#pragma unroll
                            for (int q = p; q < 6; q++) { acc[hk] = __fmaf_rn(gj[p], gj[q], acc[hk]); hk++; }   // sum order differs from the reference anyway
#pragma unroll
                        for (int p = 0; p < 6; p++) acc[21 + p] = __fmaf_rn(-gj[p], diff, acc[21 + p]);
                    }
                }
                // warp reduce (float) -> CTA partial (double, 27 threads) -> every CTA's table through DSMEM
                trc.stamp(7);
                {
                    const float t = warp_sum_scatter<NGRAD>(acc, lane);   // lane k: warp sum of term k
                    if (lane < NGRAD) hdr->warp_grad[warp][lane] = t;
                }
                __syncthreads();
                if (tid < NGRAD) {
                    const double t = tree_sum<ALIGN_WARPS>(&hdr->warp_grad[0][tid], NGRAD);
#pragma unroll
                    for (unsigned dst = 0; dst < CL; dst++) dsmem_push_f64(&hdr->cl_grad[crank][tid], &hdr->xbar[2], grp * CL + dst, t);
                    if (tid == 0) mbar_expect_tx(&hdr->xbar[2], CL * NGRAD * 8);
                }
                if (a.dbg) mbar_wait_dbg(&hdr->xbar[2], xphase[2], a.dbg, 4, (int)prank, level, mode, n_evals, n_grads);
                else mbar_wait(&hdr->xbar[2], xphase[2]);
                xphase[2] ^= 1;
                if (tid < NGRAD) {
                    hdr->red_out[tid] = tree_sum<CL>(&hdr->cl_grad[0][tid], NGRAD);
                }
                __syncthreads();
                trc.stamp(8);
                n_grads++;
                // every CTA solves the same 6x6 system redundantly (no further cluster traffic), and so do 16 lanes of it (two in each
                // warp) side by side, each going straight on to the matrix of "its" trial pose x0 + 2^-j grad, j = warp + 8 * lane:
                // no block barrier between the solve and the Rodrigues formulas
                if (lane < ALIGN_DEPTH / ALIGN_WARPS) {
                    double dx[6];
                    float delta[6], pg[6], Rf[9];
                    if (dev_solve6(hdr->red_out, hdr->red_out + 21, dx)) {
#pragma unroll
                        for (int k = 0; k < 6; k++) delta[k] = (float)dx[k];
                    } else {
                        dev_pinv_step(hdr->red_out, hdr->red_out + 21, delta);   // rank-deficient H: the reference's pseudo-inverse step
                    }
                    trc.stamp(9);
                    dev_expmap(delta, pg);
                    // rot_mat = float(Rodrigues(r)) == transpose of float(Rodrigues(-r)) bit for bit
#pragma unroll
                    for (int p = 0; p < 3; p++)
#pragma unroll
                        for (int q = 0; q < 3; q++) Rf[p * 3 + q] = Rif[q * 3 + p];
                    float g[6];
                    dev_m33v(Rf, pg[0], pg[1], pg[2], g[0], g[1], g[2]);
                    dev_m33v(Rf, pg[3], pg[4], pg[5], g[3], g[4], g[5]);
                    if (tid == 0) {
#pragma unroll
                        for (int k = 0; k < 6; k++) hdr->grad[k] = g[k];
                    }
                    trc.stamp(10);
                    const int j = warp + ALIGN_WARPS * lane;
                    float kj = 1.f;
                    for (int q = 0; q < j; q++) kj = kj / 2;   // the driver's own sequence of halvings
                    dev_rodrigues_d(-(x0[3] + (kj * g[3])), -(x0[4] + (kj * g[4])), -(x0[5] + (kj * g[5])), hdr->Rd[tb * ALIGN_DEPTH + j]);
                    trc.stamp(11);
                }
                __syncthreads();
#pragma unroll
                for (int k = 0; k < 6; k++) grad[k] = hdr->grad[k];
                kstep = 1.f;
                jstep = 0;
#pragma unroll
                for (int k = 0; k < 6; k++) xt[k] = x0[k] + (kstep * grad[k]);
                xtslot = tb * ALIGN_DEPTH;
                mode = 2;
                trc.stamp(5);
                if (a.probe_level >= 0) {
                    if (tid == 0 && prank == 0)
                        for (int k = 0; k < 6; k++) a.probe_grad[k] = grad[k];
                    mode = 3;
                }
            }
        }
        if (tid == 0 && prank == 0 && level < 8) { a.evals_out[2 * level] = n_evals; a.evals_out[2 * level + 1] = n_grads; }
    }
    if (tid == 0 && prank == 0) {
        for (int k = 0; k < 6; k++) a.pose_out[k] = x0[k];
        *a.cost_out = prev_cost;
    }
    if (a.rd_out && prank == 0 && tid < 9) a.rd_out[tid] = hdr->Rd[x0slot][tid];   // the matrix of the final pose, for KLT / refinement
    trc.stamp(6);
    trc.finish();
    cluster_sync_all();   // no CTA exits while others may still write into its shared memory
}

static bool align_levels_fit(const AlignArgs &a, size_t &need)
{
    need = 0;
    int hi = a.probe_level >= 0 ? a.probe_level : a.cam.max_levels - 1;
    int lo = a.probe_level >= 0 ? a.probe_level : a.cam.min_level;
    bool fit = true;
    for (int l = lo; l <= hi; l++) {
        size_t b = (((size_t)a.prev[l].w * a.prev[l].h) + 15) & ~(size_t)15;
        if (2 * b > 200 * 1024) fit = false;
        need = need > 2 * b ? need : 2 * b;
    }
    need += 16;   // load5f reads whole words: up to 3 bytes past the last pixel of the second image
    return fit;
}

size_t align_smem_bytes(const AlignArgs &a)
{
    size_t need;
    bool fit = align_levels_fit(a, need);
    return HDR_BYTES + (fit ? need : 0);
}

size_t align_scratch_floats(int max_kps) { return (size_t)13 * 4 * max_kps; }

cudaError_t align_init_device()
{
    cudaError_t e = cudaSuccess;
#define ALIGN_OPT_IN(c, g) \
    if (e == cudaSuccess) e = cudaFuncSetAttribute(sparse_align_kernel<true, c, g>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024)); \
    if (e == cudaSuccess && (c) * (g) > 8) e = cudaFuncSetAttribute(sparse_align_kernel<true, c, g>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1); \
    if (e == cudaSuccess && (c) * (g) > 8) e = cudaFuncSetAttribute(sparse_align_kernel<false, c, g>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1)
    ALIGN_OPT_IN(1, 1); ALIGN_OPT_IN(2, 1); ALIGN_OPT_IN(4, 1); ALIGN_OPT_IN(8, 1); ALIGN_OPT_IN(16, 1);
    ALIGN_OPT_IN(8, 2); ALIGN_OPT_IN(4, 2); ALIGN_OPT_IN(4, 4);
#undef ALIGN_OPT_IN
    return e;
}

template <int CL, int G>
static cudaError_t launch_align_cl(const AlignArgs &a, bool fit, size_t need, cudaStream_t st)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(CL * G, 1, 1);
    cfg.blockDim = dim3(ALIGN_THREADS, 1, 1);
    cfg.dynamicSmemBytes = HDR_BYTES + (fit ? need : 0);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL * G; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (fit) return cudaLaunchKernelEx(&cfg, sparse_align_kernel<true, CL, G>, a);
    return cudaLaunchKernelEx(&cfg, sparse_align_kernel<false, CL, G>, a);
}

// cluster = number of SMs that share one evaluation of a frame's solve: 8 minimises the latency of a single sequence, 1-2
// minimise the SM time per frame when many sequences are in flight (the solve is a chain of ~28 dependent evaluations,
// so SMs x duration is what a frame costs the GPU).  groups > 1: that many such clusters side by side, one trial pose of the
// line search each (svo_set_solver_width); the results do not depend on it.
cudaError_t launch_align(const AlignArgs &a, cudaStream_t st)
{
    size_t need;
    bool fit = align_levels_fit(a, need);
    if (a.groups == 2 && a.cluster == 8) return launch_align_cl<8, 2>(a, fit, need, st);
    if (a.groups == 2 && a.cluster == 4) return launch_align_cl<4, 2>(a, fit, need, st);
    if (a.groups == 4 && a.cluster == 4) return launch_align_cl<4, 4>(a, fit, need, st);
    switch (a.cluster) {
    case 1: return launch_align_cl<1, 1>(a, fit, need, st);
    case 2: return launch_align_cl<2, 1>(a, fit, need, st);
    case 4: return launch_align_cl<4, 1>(a, fit, need, st);
    case 16: return launch_align_cl<16, 1>(a, fit, need, st);
    default: return launch_align_cl<8, 1>(a, fit, need, st);
    }
}
