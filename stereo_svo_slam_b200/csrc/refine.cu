// Reprojection-error pose refinement — PoseRefiner::update_pose + PoseRefinerCallback
// (src/lib/pose_refinement.cpp:236-290, :321-412) as one persistent CTA: cost = sum |proj - kp| over
// keypoints with no ignore flag, Gauss-Newton step from H = sum J^T J, e = sum J^T d (|d| <= 3 px),
// exponential_map, additive update with step halving, shared 50-evaluation counter, stop threshold 1e-4.
// Plus project_keypoints (src/lib/transform_keypoints.cpp:11-47) as a stand-alone kernel.
#include "kernels.cuh"

#define REF_THREADS 256
#define RNRED 28

__device__ bool refine_solve6(const double *Hu, const double *b, double *x)
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

struct RefHdr {
    double red_out[RNRED];
    double Rd[9];
    float Rif[9];
    float x0[6], xt[6], grad[6];
    float prev_cost;
    int ctrl;
};

__global__ void __launch_bounds__(REF_THREADS, 1) reproj_refine_kernel(RefineArgs a, int max_kps)
{
    __shared__ RefHdr hdr;
    __shared__ double red_scratch[RNRED * 32];
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int n = min(*a.n_ptr, max_kps);
    const DevCam cam = a.cam;
    if (tid == 0)
        for (int k = 0; k < 6; k++) hdr.x0[k] = a.pose_in[k];
    __syncthreads();
    int mode = 0, it = 0, n_evals = 0, n_grads = 0;
    float kstep = 1.f;
    while (mode != 3) {
        const float *x = (mode == 2) ? hdr.xt : hdr.x0;
        if (tid == 0) {
            dev_rodrigues_d(-x[3], -x[4], -x[5], hdr.Rd);
            for (int k = 0; k < 9; k++) hdr.Rif[k] = (float)hdr.Rd[k];
        }
        __syncthreads();
        const float tx = x[0], ty = x[1], tz = x[2];
        double acc[RNRED];
#pragma unroll
        for (int k = 0; k < RNRED; k++) acc[k] = 0.0;
        for (int i = tid; i < n; i += nthr) {
            if (a.flags[i] & (SVO_F_IGN_REFINE | SVO_F_IGN_COMPLETE | SVO_F_IGN_TEMP)) continue;
            const float Px = a.kps3d[3 * i], Py = a.kps3d[3 * i + 1], Pz = a.kps3d[3 * i + 2];
            const float kx = a.kps2d[2 * i], ky = a.kps2d[2 * i + 1];
            float u, v;
            dev_project(hdr.Rd, Px, Py, Pz, tx, ty, tz, cam.fx, cam.fy, cam.cx, cam.cy, cam.k1, cam.k2, cam.p1, cam.p2, cam.k3, u, v);
            if (mode != 1) {
                float d0 = fabsf(u - kx), d1 = fabsf(v - ky);
                acc[27] += (double)(d0 + d1);
            } else {
                float d0 = kx - u, d1 = ky - v;
                if ((fabs((double)d0) > 3.0) || (fabs((double)d1) > 3.0)) continue;
                float X, Y, Z;
                dev_m33v(hdr.Rif, Px - tx, Py - ty, Pz - tz, X, Y, Z);
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
                        float s = 0.f;
                        s += J[p] * J[q];
                        s += J[6 + p] * J[6 + q];
                        acc[hk] += (double)s;
                        hk++;
                    }
#pragma unroll
                for (int p = 0; p < 6; p++) {
                    float s = 0.f;
                    s += J[p] * d0;
                    s += J[6 + p] * d1;
                    acc[21 + p] += (double)s;
                }
            }
        }
        block_reduce_sum<RNRED>(acc, red_scratch, hdr.red_out);
        if (tid == 0) {
            if (mode == 0) {
                hdr.prev_cost = (float)hdr.red_out[27];
                n_evals++;
                hdr.ctrl = 1;
            } else if (mode == 1) {
                n_grads++;
                double dx[6];
                float tw[6];
                bool ok = refine_solve6(hdr.red_out, hdr.red_out + 21, dx);
                for (int k = 0; k < 6; k++) tw[k] = ok ? (float)dx[k] : 0.f;
                dev_expmap(tw, hdr.grad);  // used as is — not rotated to world (pose_refinement.cpp:401-411)
                kstep = 1.f;
                for (int k = 0; k < 6; k++) hdr.xt[k] = hdr.x0[k] + (kstep * hdr.grad[k]);
                hdr.ctrl = 2;
            } else {
                float new_cost = (float)hdr.red_out[27];
                n_evals++;
                if (new_cost < hdr.prev_cost) {
                    for (int k = 0; k < 6; k++) hdr.x0[k] = hdr.xt[k];
                    hdr.prev_cost = new_cost;
                    it++;
                    hdr.ctrl = (it < 50) ? 1 : 3;
                } else if (fabs((double)(new_cost - hdr.prev_cost)) < 0.0001) {
                    hdr.ctrl = 3;
                } else {
                    kstep = kstep / 2;
                    it++;
                    if (it < 50) {
                        for (int k = 0; k < 6; k++) hdr.xt[k] = hdr.x0[k] + (kstep * hdr.grad[k]);
                        hdr.ctrl = 2;
                    } else
                        hdr.ctrl = 3;
                }
            }
        }
        __syncthreads();
        mode = hdr.ctrl;
    }
    if (tid == 0) {
        for (int k = 0; k < 6; k++) a.pose_out[k] = hdr.x0[k];
        *a.cost_out = hdr.prev_cost;
        a.evals_out[0] = n_evals; a.evals_out[1] = n_grads;
    }
}

static int g_refine_max_kps = 0;
void launch_refine_n(const RefineArgs &a, int max_kps, cudaStream_t st) { reproj_refine_kernel<<<1, REF_THREADS, 0, st>>>(a, max_kps); }
void launch_refine(const RefineArgs &a, cudaStream_t st) { launch_refine_n(a, 1 << 30, st); }

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
