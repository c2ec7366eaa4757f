// Reprojection-error pose refinement — PoseRefiner::update_pose + PoseRefinerCallback
// (src/lib/pose_refinement.cpp:236-290, :321-412) as one persistent CTA: cost = sum |proj - kp| over
// keypoints with no ignore flag, Gauss-Newton step from H = sum J^T J, e = sum J^T d (|d| <= 3 px),
// exponential_map, additive update with step halving, shared 50-evaluation counter, stop threshold 1e-4.
// Plus project_keypoints (src/lib/transform_keypoints.cpp:11-47) as a stand-alone kernel.
#include "kernels.cuh"

#define REF_MAX_WARPS 16   // CTAs of 256 threads, or 512 for frames with more than 1024 keypoints
#define RNGRAD 27

// 6x6 SPD solve in double (LDL^T, one reciprocal per pivot)
__device__ bool refine_solve6(const double *Hu, const double *b, double *x)
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
    double L[6][6], D[6], Dinv[6];
    bool ok = true;
#pragma unroll
    for (int j = 0; j < 6; j++) {
        double d = A[j][j];
#pragma unroll
        for (int q = 0; q < j; q++) d -= L[j][q] * L[j][q] * D[q];
        if (!(d > 1e-13 * maxd)) { ok = false; d = 1.0; }
        D[j] = d;
        Dinv[j] = 1.0 / d;
#pragma unroll
        for (int i = j + 1; i < 6; i++) {
            double t = A[i][j];
#pragma unroll
            for (int q = 0; q < j; q++) t -= L[i][q] * L[j][q] * D[q];
            L[i][j] = t * Dinv[j];
        }
    }
    if (!ok) return false;
    double y[6];
#pragma unroll
    for (int i = 0; i < 6; i++) {
        double t = b[i];
#pragma unroll
        for (int q = 0; q < i; q++) t -= L[i][q] * y[q];
        y[i] = t;
    }
#pragma unroll
    for (int i = 5; i >= 0; i--) {
        double t = y[i] * Dinv[i];
#pragma unroll
        for (int q = i + 1; q < 6; q++) t -= L[q][i] * x[q];
        x[i] = t;
    }
    return true;
}

struct RefHdr {
    double Rd[18][9];   // two tables of 8 step sizes (x0 + 2^-j grad) + 2 spare slots, as in sparse_align_kernel
    double cost_part[2][REF_MAX_WARPS];
    float grad_part[RNGRAD][REF_MAX_WARPS];
    double red_out[RNGRAD];
    float grad[6];
};

// Solver state lives in registers, identical in every thread (decisions are recomputed redundantly from the
// broadcast partial sums: one barrier per cost evaluation).  The rotation of the NEXT trial pose (halved step)
// is computed speculatively by the last thread while the others evaluate the current one.
template <int REF_THREADS>
__global__ void __launch_bounds__(REF_THREADS, 1) reproj_refine_kernel(RefineArgs a, int max_kps)
{
    constexpr int REF_WARPS = REF_THREADS / 32;
    __shared__ RefHdr hdr;
    const int tid = threadIdx.x, nthr = REF_THREADS, lane = tid & 31, warp = tid >> 5;
    const int n = min(*a.n_ptr, max_kps);
    const DevCam cam = a.cam;
    const bool nodist = dev_cam_nodist(cam);   // 40 % fewer double-precision instructions, which bound this one-CTA kernel (64 FP64 results per clock per SM)
    float x0[6], xt[6], grad[6];
#pragma unroll
    for (int k = 0; k < 6; k++) { x0[k] = a.pose_in[k]; xt[k] = x0[k]; grad[k] = 0.f; }
    if (a.rd_in) { if (tid < 9) hdr.Rd[16][tid] = a.rd_in[tid]; }
    else if (tid == 0) dev_rodrigues_d(-x0[3], -x0[4], -x0[5], hdr.Rd[16]);
    __syncthreads();
    int mode = 0, it = 0, n_evals = 0, n_grads = 0, cbuf = 0, jstep = 0;
    int x0slot = 16, sp = 1, tb = 0, xtslot = 0;
    float kstep = 1.f, prev_cost = 0.f;
    while (mode != 3) {
        const float *x = (mode == 2) ? xt : x0;
        const double *Rd = hdr.Rd[(mode == 2) ? xtslot : x0slot];
        const float tx = x[0], ty = x[1], tz = x[2];
        if (mode != 1) {
            double part = 0.0;
            for (int i = tid; i < n; i += nthr) {
                if (a.flags[i] & (SVO_F_IGN_REFINE | SVO_F_IGN_COMPLETE | SVO_F_IGN_TEMP)) continue;
                float u, v;
                dev_project_nd(nodist, Rd, a.kps3d[3 * i], a.kps3d[3 * i + 1], a.kps3d[3 * i + 2], tx, ty, tz, cam.fx, cam.fy, cam.cx, cam.cy, cam, u, v);
                float d0 = fabsf(u - a.kps2d[2 * i]), d1 = fabsf(v - a.kps2d[2 * i + 1]);
                part += (double)(d0 + d1);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) part += __shfl_down_sync(0xffffffffu, part, o);
            if (lane == 0) hdr.cost_part[cbuf][warp] = part;
            __syncthreads();
            double tot = 0.0;
#pragma unroll
            for (int q = 0; q < REF_WARPS; q++) tot += hdr.cost_part[cbuf][q];
            cbuf ^= 1;
            const float cost = (float)tot;
            n_evals++;
            if (mode == 0) {
                prev_cost = cost;
                mode = 1;
            } else if (cost < prev_cost) {
#pragma unroll
                for (int k = 0; k < 6; k++) x0[k] = xt[k];
                prev_cost = cost;
                it++;
                mode = (it < 50) ? 1 : 3;
                x0slot = xtslot;
                if (xtslot >= 16) sp ^= 1; else tb ^= 1;
            } else if (fabs((double)(cost - prev_cost)) < 0.0001) {
                mode = 3;
            } else {
                kstep = kstep / 2;
                jstep++;
                it++;
                if (it < 50) {
#pragma unroll
                    for (int k = 0; k < 6; k++) xt[k] = x0[k] + (kstep * grad[k]);
                    if (jstep < 8) {
                        xtslot = tb * 8 + jstep;
                    } else {
                        xtslot = 16 + sp;
                        if (tid == 0) dev_rodrigues_d(-xt[3], -xt[4], -xt[5], hdr.Rd[xtslot]);
                        __syncthreads();
                    }
                    mode = 2;
                } else
                    mode = 3;
            }
        } else {
            float Rif[9];
#pragma unroll
            for (int k = 0; k < 9; k++) Rif[k] = (float)Rd[k];
            float acc[RNGRAD];
#pragma unroll
            for (int k = 0; k < RNGRAD; k++) acc[k] = 0.f;
            for (int i = tid; i < n; i += nthr) {
                if (a.flags[i] & (SVO_F_IGN_REFINE | SVO_F_IGN_COMPLETE | SVO_F_IGN_TEMP)) continue;
                const float Px = a.kps3d[3 * i], Py = a.kps3d[3 * i + 1], Pz = a.kps3d[3 * i + 2];
                float u, v;
                dev_project_nd(nodist, Rd, Px, Py, Pz, tx, ty, tz, cam.fx, cam.fy, cam.cx, cam.cy, cam, u, v);
                float d0 = a.kps2d[2 * i] - u, d1 = a.kps2d[2 * i + 1] - v;
                if ((fabs((double)d0) > 3.0) || (fabs((double)d1) > 3.0)) continue;
                float X, Y, Z;
                dev_m33v(Rif, Px - tx, Py - ty, Pz - tz, X, Y, Z);
                const float fx = cam.fx, fy = cam.fy;
                float J[12];
                J[0] = -fx / Z; J[1] = 0.f; J[2] = fx * X / (Z * Z); J[3] = fx * X * Y / (Z * Z);
                J[4] = -fx * (1 + (X * X) / (Z * Z)); J[5] = fx * Y / Z;
                J[6] = 0.f; J[7] = -fy / Z; J[8] = fy * Y / (Z * Z); J[9] = fy * (1 + (Y * Y) / (Z * Z));
                J[10] = -fy * X * Y / (Z * Z); J[11] = -fy * X / Z;
                int hk = 0;
#pragma unroll
                for (int p = 0; p < 6; p++)
#pragma unroll
                    for (int q = p; q < 6; q++) {
                        float t = 0.f;
                        t += J[p] * J[q];
                        t += J[6 + p] * J[6 + q];
                        acc[hk] += t;
                        hk++;
                    }
#pragma unroll
                for (int p = 0; p < 6; p++) {
                    float t = 0.f;
                    t += J[p] * d0;
                    t += J[6 + p] * d1;
                    acc[21 + p] += t;
                }
            }
            {
                const float t = warp_sum_scatter<RNGRAD>(acc, lane);   // lane k: warp sum of term k
                if (lane < RNGRAD) hdr.grad_part[lane][warp] = t;
            }
            __syncthreads();
            if (tid < RNGRAD) {
                double t = 0.0;
#pragma unroll
                for (int q = 0; q < REF_WARPS; q++) t += (double)hdr.grad_part[tid][q];
                hdr.red_out[tid] = t;
            }
            __syncthreads();
            if (tid == 0) {
                double dx[6];
                float tw[6], g[6];
                if (refine_solve6(hdr.red_out, hdr.red_out + 21, dx)) {
#pragma unroll
                    for (int k = 0; k < 6; k++) tw[k] = (float)dx[k];
                } else {
                    dev_pinv_step(hdr.red_out, hdr.red_out + 21, tw);   // rank-deficient H: the reference's pseudo-inverse step
                }
                dev_expmap(tw, g);  // used as is — not rotated to world (pose_refinement.cpp:401-411)
#pragma unroll
                for (int k = 0; k < 6; k++) hdr.grad[k] = g[k];
            }
            __syncthreads();
            n_grads++;
#pragma unroll
            for (int k = 0; k < 6; k++) grad[k] = hdr.grad[k];
            if (lane == 0 && warp < 8) {   // warp j prepares the matrix of x0 + 2^-j grad
                float kj = 1.f;
                for (int q = 0; q < warp; q++) kj = kj / 2;
                dev_rodrigues_d(-(x0[3] + (kj * grad[3])), -(x0[4] + (kj * grad[4])), -(x0[5] + (kj * grad[5])), hdr.Rd[tb * 8 + warp]);
            }
            __syncthreads();
            kstep = 1.f;
            jstep = 0;
#pragma unroll
            for (int k = 0; k < 6; k++) xt[k] = x0[k] + (kstep * grad[k]);
            xtslot = tb * 8;
            mode = 2;
        }
    }
    if (tid == 0) {
        for (int k = 0; k < 6; k++) a.pose_out[k] = x0[k];
        *a.cost_out = prev_cost;
        a.evals_out[0] = n_evals; a.evals_out[1] = n_grads;
    }
    if (a.rd_out && tid < 9) a.rd_out[tid] = hdr.Rd[x0slot][tid];   // the matrix of the refined pose, for the depth filter
}

// bucket: upper bound of the keypoint count known at launch (graph capture) time; it picks the CTA size, so that a frame gives
// the same bits whether it is replayed from a graph or launched directly
void launch_refine(const RefineArgs &a, int bucket, cudaStream_t st)
{
    if (bucket > 1024) reproj_refine_kernel<512><<<1, 512, 0, st>>>(a, 1 << 30);   // C4: 114 instead of 128 us; no gain for a few hundred keypoints
    else reproj_refine_kernel<256><<<1, 256, 0, st>>>(a, 1 << 30);
}

__global__ void __launch_bounds__(128) project_kernel(const float *pose, const float *kps3d, const int *n_ptr, int max_kps, DevCam cam,
                                                      float *kps2d)
{
    __shared__ double Rd[9];
    if (threadIdx.x == 0) dev_rodrigues_d(-pose[3], -pose[4], -pose[5], Rd);
    __syncthreads();
    const int n = min(*n_ptr, max_kps);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float u, v;
    dev_project(Rd, kps3d[3 * i], kps3d[3 * i + 1], kps3d[3 * i + 2], pose[0], pose[1], pose[2], cam.fx, cam.fy, cam.cx, cam.cy, cam.k1,
                cam.k2, cam.p1, cam.p2, cam.k3, u, v);
    kps2d[2 * i] = u; kps2d[2 * i + 1] = v;
}

void launch_project(const float *pose, const float *kps3d, const int *n_ptr, int max_kps, DevCam cam, float *kps2d, cudaStream_t st)
{
    if (max_kps <= 0) return;
    project_kernel<<<(max_kps + 127) / 128, 128, 0, st>>>(pose, kps3d, n_ptr, max_kps, cam, kps2d);
}
