// Sparse direct image alignment — PoseEstimator (src/lib/pose_estimator.cpp) as ONE persistent kernel.
//
// Reference structure replaced here (all per frame, on the CPU, single-threaded):
//   estimate_pose            :115-130   levels max-1 .. min
//   estimate_pose_at_level   :166-222   <= 50 cost evaluations per level, accept / halve / stop
//   do_calc                  :275-300   project_keypoints + get_total_intensity_diff (image_comparison.cpp:9-120)
//   calculate_hessian        :312-416   box-sum gradients x 2x6 Jacobian, H = sum (gJ)^T(gJ), rebuilt every call (Q1)
//   get_gradient             :418-539   residuals, b = -sum (gJ) r, delta = H^-1 b, exponential_map, rotate to world
//
// B200 design: one CTA owns the whole solve of one frame.  Per level the two level images are staged in
// shared memory with TMA bulk copies (cp.async.bulk + mbarrier), every thread owns keypoints (strided), does
// the photometric residuals and 1x6 Jacobian rows in registers with the reference's exact float operation
// order, and the 21+6+1 normal-equation terms are reduced with warp shuffles and one block-level pass in a
// fixed tree (deterministic).  The 6x6 solve, the exponential map and the accept/halve/stop decisions run in
// the kernel: there is no host round trip anywhere inside a frame.  Pose-independent reference-patch terms
// (box-sum gradients and reference box sums, 48 floats + validity mask per keypoint) are computed once per
// level and cached in an L2-resident scratch array laid out [term][keypoint] for coalesced access.
//
// Bound: dependency latency (serial evaluations) and shared-memory/ALU throughput, not HBM.
#include "kernels.cuh"

#define ALIGN_THREADS 512
#define NRED 28  // 21 (H upper triangle) + 6 (b) + 1 (cost)
#define SCR_TERMS 49

struct AlignHdr {
    unsigned long long bar;
    double red_out[NRED];
    double Rd[9];    // Rodrigues(-r) in double: projection rotation
    float Rf[9];     // rot_mat      = float(Rodrigues(r))
    float Rif[9];    // inv_rot_mat  = float(Rodrigues(-r))
    float x0[6], xt[6], grad[6];
    float prev_cost, new_cost;
    int n, ctrl;
};
#define HDR_BYTES ((sizeof(AlignHdr) + 127) / 128 * 128)
#define RED_BYTES (NRED * 32 * sizeof(double))

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

// get_patch_sum (pose_estimator.cpp:82-112), operation for operation
__device__ __forceinline__ float patch_sum(const uint8_t *img, int pitch, float cx, float cy)
{
    float sx = cx - 0.5f, sy = cy - 0.5f;
    float fxf = floorf(sx), fyf = floorf(sy);
    int ipx = (int)fxf, ipy = (int)fyf;
    float x2 = sx - (float)ipx, y2 = sy - (float)ipy;
    float x1 = (float)(1.0 - (double)x2), y1 = (float)(1.0 - (double)y2);
    const uint8_t *s1 = img + ipy * pitch + ipx;
    const uint8_t *s2 = s1 + pitch;
    const uint8_t *s3 = s2 + pitch;
    float v = x1 * y1 * (float)s1[0];
    v = v + y1 * (float)s1[1];
    v = v + x2 * y1 * (float)s1[2];
    v = v + x1 * (float)s2[0];
    v = v + (float)s2[1];
    v = v + x2 * (float)s2[2];
    v = v + x1 * y2 * (float)s3[0];
    v = v + y2 * (float)s3[1];
    v = v + x2 * y2 * (float)s3[2];
    return v;
}

// _get_intensity_diff (image_comparison.cpp:9-91)
__device__ __forceinline__ float intensity_diff(const uint8_t *im1, const uint8_t *im2, int w, int h, int pitch, float c1x, float c1y,
                                                float c2x, float c2y, int patch)
{
    float half = ((float)patch - 1.0f) / 2.0f;
    float s1x = c1x - half, s1y = c1y - half, s2x = c2x - half, s2y = c2y - half;
    int ip1x = (int)floorf(s1x), ip1y = (int)floorf(s1y), ip2x = (int)floorf(s2x), ip2y = (int)floorf(s2y);
    float x12 = s1x - (float)ip1x, y12 = s1y - (float)ip1y, x22 = s2x - (float)ip2x, y22 = s2y - (float)ip2y;
    float x11 = (float)(1.0 - (double)x12), y11 = (float)(1.0 - (double)y12);
    float x21 = (float)(1.0 - (double)x22), y21 = (float)(1.0 - (double)y22);
    float m11 = x11 * y11, m12 = x12 * y11, m13 = x11 * y12, m14 = x12 * y12;
    float m21 = x21 * y21, m22 = x22 * y21, m23 = x21 * y22, m24 = x22 * y22;
    float intensity = 0.f;
    if (ip1y >= 0 && ip1y + patch < h && ip2y >= 0 && ip2y + patch < h && ip1x >= 0 && ip1x + patch < w && ip2x >= 0 &&
        ip2x + patch < w) {
        for (int i = 0; i < patch; i++) {
            const uint8_t *s11 = im1 + (i + ip1y) * pitch + ip1x, *s12 = s11 + pitch;
            const uint8_t *s21 = im2 + (i + ip2y) * pitch + ip2x, *s22 = s21 + pitch;
            for (int j = 0; j < patch; j++) {
                float i1 = 0.f, i2 = 0.f;
                i1 += m11 * (float)s11[j]; i1 += m12 * (float)s11[j + 1]; i1 += m13 * (float)s12[j]; i1 += m14 * (float)s12[j + 1];
                i2 += m21 * (float)s21[j]; i2 += m22 * (float)s21[j + 1]; i2 += m23 * (float)s22[j]; i2 += m24 * (float)s22[j + 1];
                intensity += fabsf(i1 - i2);
            }
        }
    }
    return intensity;
}

// 6x6 SPD solve in double (LDL^T); returns false when H is not numerically positive definite
__device__ bool solve6(const double *Hu /*21 upper-tri row-major*/, const double *b, double *x)
{
    double A[6][6];
    int k = 0;
    for (int i = 0; i < 6; i++)
        for (int j = i; j < 6; j++) { A[i][j] = Hu[k]; A[j][i] = Hu[k]; k++; }
    double maxd = 0;
    for (int i = 0; i < 6; i++) maxd = fmax(maxd, A[i][i]);
    if (!(maxd > 0)) return false;
    double L[6][6], D[6];
    for (int j = 0; j < 6; j++) {
        double d = A[j][j];
        for (int q = 0; q < j; q++) d -= L[j][q] * L[j][q] * D[q];
        if (!(d > 1e-13 * maxd)) return false;
        D[j] = d;
        for (int i = j + 1; i < 6; i++) {
            double s = A[i][j];
            for (int q = 0; q < j; q++) s -= L[i][q] * L[j][q] * D[q];
            L[i][j] = s / d;
        }
    }
    double y[6];
    for (int i = 0; i < 6; i++) {
        double s = b[i];
        for (int q = 0; q < i; q++) s -= L[i][q] * y[q];
        y[i] = s;
    }
    for (int i = 5; i >= 0; i--) {
        double s = y[i] / D[i];
        for (int q = i + 1; q < 6; q++) s -= L[q][i] * x[q];
        x[i] = s;
    }
    return true;
}

__global__ void __launch_bounds__(ALIGN_THREADS, 1) sparse_align_kernel(AlignArgs a, int img_bytes_cap)
{
    extern __shared__ __align__(128) uint8_t smem_raw[];
    AlignHdr *hdr = reinterpret_cast<AlignHdr *>(smem_raw);
    double *red_scratch = reinterpret_cast<double *>(smem_raw + HDR_BYTES);
    uint8_t *img_area = smem_raw + HDR_BYTES + RED_BYTES;

    const int tid = threadIdx.x, nthr = blockDim.x;
    const DevCam cam = a.cam;
    if (tid == 0) {
        mbar_init(&hdr->bar, 1);
        hdr->n = min(*a.n_ptr, a.max_kps);
        for (int k = 0; k < 6; k++) hdr->x0[k] = a.pose_in[k];
        for (int k = 0; k < 16; k++) a.evals_out[k] = 0;
    }
    __syncthreads();
    const int n = hdr->n;
    uint32_t phase = 0;
    float *scr = a.scratch;
    const int SN = a.max_kps;

    const int lv_hi = (a.probe_level >= 0) ? a.probe_level : cam.max_levels - 1;
    const int lv_lo = (a.probe_level >= 0) ? a.probe_level : cam.min_level;

    for (int level = lv_hi; level >= lv_lo; level--) {
        const LevelDesc P = a.prev[level], C = a.cur[level];
        const int w = P.w, h = P.h;
        const uint8_t *pimg, *cimg;
        const uint32_t bytes16 = ((uint32_t)(w * h) + 15u) & ~15u;
        // ---- stage both level images in shared memory with TMA bulk copies
        if ((int)(2 * bytes16) <= img_bytes_cap) {
            __syncthreads();  // everyone finished reading the previous level's tiles
            if (tid == 0) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_expect_tx(&hdr->bar, 2 * bytes16);
                tma_load_1d(img_area, P.ptr, bytes16, &hdr->bar);
                tma_load_1d(img_area + bytes16, C.ptr, bytes16, &hdr->bar);
            }
            mbar_wait(&hdr->bar, phase);
            phase ^= 1;
            pimg = img_area; cimg = img_area + bytes16;
        } else {
            pimg = P.ptr; cimg = C.ptr;  // level too large for shared memory: read through L1/L2
        }
        const int pitch = w;  // halfSample levels are stored with pitch == width

        // setLevel (pose_estimator.cpp:541-562)
        const int divider = 1 << level;
        const float lfx = cam.fx / (float)divider, lfy = cam.fy / (float)divider;
        const float lcx = cam.cx / (float)divider, lcy = cam.cy / (float)divider;

        // ---- per-level cache of the pose-independent reference terms (calculate_hessian :346-395 image part,
        //      get_gradient :449-460 reference part)
        for (int i = tid; i < n; i += nthr) {
            if (a.flags && (a.flags[i] & SVO_F_IGN_TEMP)) continue;
            float bx = a.kps2d[2 * i], by = a.kps2d[2 * i + 1];
            if (level != 0) { bx = bx / (float)divider; by = by / (float)divider; }
            // Hessian loop coordinates
            float kx = bx - 2.f, ky = by - 2.f;
            // residual loop reference coordinates
            float rx = bx - 2.f, ry = by - 2.f;
            unsigned mask = 0;
            int e = 0;
            for (int r = 0; r < 4; r++) {
                for (int c = 0; c < 4; c++, e++) {
                    float g0 = 0.f, g1 = 0.f;
                    if (!(((double)kx - 2.0) < 0 || ((double)ky - 2.0) < 0 || ((double)kx + 3.0) >= w || ((double)ky + 3.0) >= h)) {
                        float i1 = patch_sum(pimg, pitch, kx + 1.f, ky), i2 = patch_sum(pimg, pitch, kx - 1.f, ky);
                        float i3 = patch_sum(pimg, pitch, kx, ky + 1.f), i4 = patch_sum(pimg, pitch, kx, ky - 1.f);
                        g0 = i1 - i2; g1 = i3 - i4;
                        mask |= 1u << e;
                    }
                    scr[(size_t)e * SN + i] = g0;
                    scr[(size_t)(16 + e) * SN + i] = g1;
                    kx += 1.f;
                    float sp = 0.f;
                    if (!(((double)rx - 1.0) < 0 || ((double)ry - 1.0) < 0 || ((double)rx + 2.0) > w || ((double)ry + 2.0) > h)) {
                        sp = patch_sum(pimg, pitch, rx, ry);
                        mask |= 1u << (16 + e);
                    }
                    scr[(size_t)(32 + e) * SN + i] = sp;
                    rx += 1.f;
                }
                kx -= 4.f; ky += 1.f;
                ry += 1.f; rx -= 4.f;
            }
            scr[(size_t)48 * SN + i] = __uint_as_float(mask);
        }
        // (each thread only ever reads back the scratch entries it wrote itself: no barrier needed)

        // ---- Gauss-Newton driver (estimate_pose_at_level :166-222).  ctrl: what to evaluate next.
        //  mode 0: cost at x0 (initial)   mode 1: gradient at x0   mode 2: cost at xt   mode 3: level done
        int mode = 0;
        int it = 0;            // the shared loop counter `i`
        float kstep = 1.f;
        int n_evals = 0, n_grads = 0;
        while (mode != 3) {
            const float *x = (mode == 2) ? hdr->xt : hdr->x0;
            // pose matrices for this evaluation (PoseManager::set_pose, pose_manager.cpp:9-17)
            if (tid == 0) {
                dev_rodrigues_d(-x[3], -x[4], -x[5], hdr->Rd);
                for (int k = 0; k < 9; k++) hdr->Rif[k] = (float)hdr->Rd[k];
                if (mode == 1) dev_rodrigues_f(x[3], x[4], x[5], hdr->Rf);
            }
            __syncthreads();
            const float tx = x[0], ty = x[1], tz = x[2];
            double acc[NRED];
#pragma unroll
            for (int k = 0; k < NRED; k++) acc[k] = 0.0;

            for (int i = tid; i < n; i += nthr) {
                if (a.flags && (a.flags[i] & SVO_F_IGN_TEMP)) continue;
                const float Px = a.kps3d[3 * i], Py = a.kps3d[3 * i + 1], Pz = a.kps3d[3 * i + 2];
                float bx = a.kps2d[2 * i], by = a.kps2d[2 * i + 1];
                if (level != 0) { bx = bx / (float)divider; by = by / (float)divider; }
                float u, v;
                dev_project(hdr->Rd, Px, Py, Pz, tx, ty, tz, lfx, lfy, lcx, lcy, cam.k1, cam.k2, cam.p1, cam.p2, cam.k3, u, v);
                if (mode != 1) {
                    // do_calc: bilinear SAD of the win x win patches (prev @ reference position, cur @ projection)
                    acc[27] += (double)intensity_diff(pimg, cimg, w, h, pitch, bx, by, u, v, cam.win_pose);
                } else {
                    // calculate_hessian: camera-frame point and 2x6 Jacobian
                    float X, Y, Z;
                    dev_m33v(hdr->Rif, Px - tx, Py - ty, Pz - tz, X, Y, Z);
                    float J[12];
                    J[0] = -lfx / Z; J[1] = 0.f; J[2] = lfx * X / (Z * Z); J[3] = lfx * X * Y / (Z * Z);
                    J[4] = -lfx * (1 + (X * X) / (Z * Z)); J[5] = lfx * Y / Z;
                    J[6] = 0.f; J[7] = -lfy / Z; J[8] = lfy * Y / (Z * Z); J[9] = lfy * (1 + (Y * Y) / (Z * Z));
                    J[10] = -lfy * X * Y / (Z * Z); J[11] = -lfy * X / Z;
                    const unsigned mask = __float_as_uint(scr[(size_t)48 * SN + i]);
                    float qx = u - 2.f, qy = v - 2.f;       // residual loop, current-image coordinates
                    float rx = bx - 2.f, ry = by - 2.f;     // reference coordinates (for the bounds test only)
                    int e = 0;
                    for (int r = 0; r < 4; r++) {
                        for (int c = 0; c < 4; c++, e++) {
                            float gj[6];
                            if (mask & (1u << e)) {
                                const float g0 = scr[(size_t)e * SN + i], g1 = scr[(size_t)(16 + e) * SN + i];
#pragma unroll
                                for (int k = 0; k < 6; k++) {
                                    float s = 0.f;
                                    s += g0 * J[k];
                                    s += g1 * J[6 + k];
                                    gj[k] = s;
                                }
                            } else {
#pragma unroll
                                for (int k = 0; k < 6; k++) gj[k] = 0.f;
                            }
                            float diff = 0.f;
                            if ((mask & (1u << (16 + e))) &&
                                !(((double)qx - 1.0) < 0 || ((double)qy - 1.0) < 0 || ((double)qx + 2.0) > w || ((double)qy + 2.0) > h)) {
                                diff = patch_sum(cimg, pitch, qx, qy) - scr[(size_t)(32 + e) * SN + i];
                            }
                            int hk = 0;
#pragma unroll
                            for (int p = 0; p < 6; p++)
#pragma unroll
                                for (int q = p; q < 6; q++) { acc[hk] += (double)(gj[p] * gj[q]); hk++; }
#pragma unroll
                            for (int p = 0; p < 6; p++) acc[21 + p] -= (double)(gj[p] * diff);
                            qx += 1.f; rx += 1.f;
                        }
                        qy += 1.f; ry += 1.f;
                        qx -= 4.f; rx -= 4.f;
                    }
                }
            }
            block_reduce_sum<NRED>(acc, red_scratch, hdr->red_out);

            // ---- serial part: thread 0 updates the solver state, everybody reads it after the barrier
            if (tid == 0) {
                if (mode == 0) {
                    hdr->prev_cost = (float)hdr->red_out[27];
                    n_evals++;
                    hdr->ctrl = (it < 50) ? 1 : 3;
                    if (a.probe_level >= 0) hdr->ctrl = 1;
                } else if (mode == 1) {
                    n_grads++;
                    double dx[6];
                    float delta[6], pg[6];
                    bool ok = solve6(hdr->red_out, hdr->red_out + 21, dx);
                    for (int k = 0; k < 6; k++) delta[k] = ok ? (float)dx[k] : 0.f;
                    dev_expmap(delta, pg);
                    dev_m33v(hdr->Rf, pg[0], pg[1], pg[2], hdr->grad[0], hdr->grad[1], hdr->grad[2]);
                    dev_m33v(hdr->Rf, pg[3], pg[4], pg[5], hdr->grad[3], hdr->grad[4], hdr->grad[5]);
                    kstep = 1.f;
                    for (int k = 0; k < 6; k++) hdr->xt[k] = hdr->x0[k] + (kstep * hdr->grad[k]);
                    hdr->ctrl = 2;
                    if (a.probe_level >= 0) {
                        for (int k = 0; k < 6; k++) a.probe_grad[k] = hdr->grad[k];
                        hdr->ctrl = 3;
                    }
                } else {  // mode 2: cost at xt
                    float new_cost = (float)hdr->red_out[27];
                    n_evals++;
                    if (new_cost < hdr->prev_cost) {
                        for (int k = 0; k < 6; k++) hdr->x0[k] = hdr->xt[k];
                        hdr->prev_cost = new_cost;
                        it++;  // outer loop increment after `break`
                        hdr->ctrl = (it < 50) ? 1 : 3;
                    } else if (fabsf(new_cost - hdr->prev_cost) < 1.0f) {
                        hdr->ctrl = 3;
                    } else {
                        kstep = kstep / 2;
                        it++;  // inner loop increment
                        if (it < 50) {
                            for (int k = 0; k < 6; k++) hdr->xt[k] = hdr->x0[k] + (kstep * hdr->grad[k]);
                            hdr->ctrl = 2;
                        } else {
                            hdr->ctrl = 3;
                        }
                    }
                }
            }
            __syncthreads();
            mode = hdr->ctrl;
        }
        if (tid == 0) {
            if (level < 8) { a.evals_out[2 * level] = n_evals; a.evals_out[2 * level + 1] = n_grads; }
        }
    }
    if (tid == 0) {
        for (int k = 0; k < 6; k++) a.pose_out[k] = hdr->x0[k];
        *a.cost_out = hdr->prev_cost;
    }
}

size_t align_smem_bytes(const AlignArgs &a)
{
    size_t need = 0;
    int hi = a.probe_level >= 0 ? a.probe_level : a.cam.max_levels - 1;
    int lo = a.probe_level >= 0 ? a.probe_level : a.cam.min_level;
    for (int l = lo; l <= hi; l++) {
        size_t b = (((size_t)a.prev[l].w * a.prev[l].h) + 15) & ~(size_t)15;
        if (2 * b <= 200 * 1024) need = need > 2 * b ? need : 2 * b;
    }
    return HDR_BYTES + RED_BYTES + need;
}

cudaError_t align_init_device()
{
    return cudaFuncSetAttribute(sparse_align_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024));
}

void launch_align(const AlignArgs &a, cudaStream_t st)
{
    size_t smem = align_smem_bytes(a);
    int cap = (int)(smem - HDR_BYTES - RED_BYTES);
    sparse_align_kernel<<<1, ALIGN_THREADS, smem, st>>>(a, cap);
}
