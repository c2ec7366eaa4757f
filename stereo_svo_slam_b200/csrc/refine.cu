// Reprojection-error pose refinement — PoseRefiner::update_pose + PoseRefinerCallback
// (src/lib/pose_refinement.cpp:236-290, :321-412) as one persistent CTA: cost = sum |proj - kp| over
// keypoints with no ignore flag, Gauss-Newton step from H = sum J^T J, e = sum J^T d (|d| <= 3 px),
// exponential_map, additive update with step halving, shared 50-evaluation counter, stop threshold 1e-4.
// Plus project_keypoints (src/lib/transform_keypoints.cpp:11-47) as a stand-alone kernel.
#include "kernels.cuh"

#define REF_MAX_WARPS 16   // CTAs of 256 threads, or 512 for frames with more than 1024 keypoints
#define RNGRAD 27

#define REF_DEPTH 24       // halving steps whose rotation matrices are prepared right after a gradient (x0 + 2^-j grad, j < 24)
#define REF_CACHED 2       // keypoints per thread held in registers for the whole solve (flags, 3-D point, 2-D observation)
struct RefHdr {
    unsigned long long xbar[2];              // cluster variant: cost-exchange barriers (alternating)
    double Rd[2 * REF_DEPTH + 2][9];         // two tables of REF_DEPTH step sizes + 2 spare slots (deeper halvings, computed on demand)
    double cost_part[2][REF_MAX_WARPS];
    double cl_cost[2][8];                    // cluster variant: the costs of the trials of one round, one per CTA (written through DSMEM)
    float grad_part[RNGRAD][REF_MAX_WARPS];
    double red_out[RNGRAD];
    float grad[6];
};

// Solver state lives in registers, identical in every thread (decisions are recomputed redundantly from the broadcast partial
// sums: one barrier per cost evaluation).
//
// The line search of a running sequence is a run of rejected-and-halved steps: the refinement of a C3 frame looks like
// "G a G r r r r r r r r r r a G r r r r r r r r r r r s" (G gradient, a accepted, r rejected, s stop: ~20 cost evaluations for ~4
// gradients, up to the 50 the reference allows).  Once the gradient is known, so are all trial poses x0 + 2^-j grad and their
// rotation matrices: 24 lanes prepare the matrices of j = 0 .. 23 at once (the first version computed j >= 8 on demand, one
// double-precision Rodrigues on one thread per evaluation).
//
// NCL == 8 (thread-block cluster, one sequence that has the GPU to itself): CTA c evaluates trial jstep + c over ALL keypoints,
// exactly like the single CTA would (same thread -> keypoint mapping, same reduction tree: the same bits), the eight costs are
// exchanged through distributed shared memory (st.async + mbarrier complete_tx) and every thread of every CTA replays the accept /
// stop / halve rule of pose_refinement.cpp:258-284 over them in order — a run of eight rejections costs one round instead of
// eight.  Gradients are evaluated redundantly by every CTA (no exchange).  Evaluation counts and decisions are the sequential ones.
template <int REF_THREADS, int NCL>
__global__ void __launch_bounds__(REF_THREADS, 1) reproj_refine_kernel(RefineArgs a, int max_kps)
{
    constexpr int REF_WARPS = REF_THREADS / 32;
    __shared__ RefHdr hdr;
    const int tid = threadIdx.x, nthr = REF_THREADS, lane = tid & 31, warp = tid >> 5;
    const int crank = NCL > 1 ? (int)cluster_ctarank() : 0;
    if (NCL > 1) {
        if (tid == 0) { mbar_init(&hdr.xbar[0], 1); mbar_init(&hdr.xbar[1], 1); }
        __syncthreads();
        cluster_sync_all();   // every CTA's barriers are live before anyone pushes into them
    }
    const int n = min(*a.n_ptr, max_kps);
    SolverTrace trc(a.trace, tid == 0 && crank == 0);
    trc.stamp(0);
    const DevCam cam = a.cam;
    const bool nodist = dev_cam_nodist(cam);   // 40 % fewer double-precision instructions, which bound this one-CTA kernel (64 FP64 results per clock per SM)
    float x0[6], xt[6], grad[6];
#pragma unroll
    for (int k = 0; k < 6; k++) { x0[k] = a.pose_in[k]; xt[k] = x0[k]; grad[k] = 0.f; }
    if (a.rd_in) { if (tid < 9) hdr.Rd[2 * REF_DEPTH][tid] = a.rd_in[tid]; }
    else if (tid == 0) dev_rodrigues_d(-x0[3], -x0[4], -x0[5], hdr.Rd[2 * REF_DEPTH]);
    // this thread's first keypoints stay in registers: the ~25 evaluations of a solve re-read nothing
    bool c_act[REF_CACHED];
    float c_P[REF_CACHED][3], c_kp[REF_CACHED][2];
#pragma unroll
    for (int q = 0; q < REF_CACHED; q++) {
        const int i = tid + q * nthr;
        c_act[q] = i < n && !(a.flags[i] & (SVO_F_IGN_REFINE | SVO_F_IGN_COMPLETE | SVO_F_IGN_TEMP));
        c_P[q][0] = c_P[q][1] = c_P[q][2] = 0.f; c_kp[q][0] = c_kp[q][1] = 0.f;
        if (c_act[q]) {
            c_P[q][0] = a.kps3d[3 * i]; c_P[q][1] = a.kps3d[3 * i + 1]; c_P[q][2] = a.kps3d[3 * i + 2];
            c_kp[q][0] = a.kps2d[2 * i]; c_kp[q][1] = a.kps2d[2 * i + 1];
        }
    }
    // keypoint q of this thread (global index i): from the registers above, or from memory beyond them; false = ignored keypoint
    auto fetch_kp = [&](int q, int i, float &Px, float &Py, float &Pz, float &k0, float &k1) -> bool {
        if (q < REF_CACHED) {
            bool act = c_act[0];
            Px = c_P[0][0]; Py = c_P[0][1]; Pz = c_P[0][2]; k0 = c_kp[0][0]; k1 = c_kp[0][1];
#pragma unroll
            for (int r = 1; r < REF_CACHED; r++)
                if (q == r) { act = c_act[r]; Px = c_P[r][0]; Py = c_P[r][1]; Pz = c_P[r][2]; k0 = c_kp[r][0]; k1 = c_kp[r][1]; }
            return act;
        }
        if (a.flags[i] & (SVO_F_IGN_REFINE | SVO_F_IGN_COMPLETE | SVO_F_IGN_TEMP)) return false;
        Px = a.kps3d[3 * i]; Py = a.kps3d[3 * i + 1]; Pz = a.kps3d[3 * i + 2];
        k0 = a.kps2d[2 * i]; k1 = a.kps2d[2 * i + 1];
        return true;
    };
    __syncthreads();
    trc.stamp(1);
    int mode = 0, it = 0, n_evals = 0, n_grads = 0, cbuf = 0, xb = 0, jstep = 0;
    uint32_t xphase[2] = {0, 0};
    int x0slot = 2 * REF_DEPTH, sp = 1, tb = 0, xtslot = 0;
    float kstep = 1.f, prev_cost = 0.f;
    while (mode != 3) {
        if (mode != 1) {
            // trials evaluated in this round: CTA c takes halving step jstep + c as long as its matrix is in the table (j < REF_DEPTH),
            // the iteration budget allows it (it + c < 50) and the pose is a trial at all (mode 2, table slot)
            int nvalid = 1;
            if (NCL > 1 && mode == 2 && xtslot < 2 * REF_DEPTH) { while (nvalid < NCL && jstep + nvalid < REF_DEPTH && it + nvalid < 50) nvalid++; }
            const bool exchange = nvalid > 1;
            const int g_mine = exchange ? crank : 0;
            const bool mine = g_mine < nvalid;
            float xg[6], kg = kstep;
            for (int q = 0; q < g_mine; q++) kg = kg / 2;
#pragma unroll
            for (int k = 0; k < 6; k++) xg[k] = (mode == 2) ? (g_mine == 0 ? xt[k] : x0[k] + (kg * grad[k])) : x0[k];
            const double *Rd = hdr.Rd[(mode == 2) ? (mine ? xtslot + g_mine : xtslot) : x0slot];
            const float tx = xg[0], ty = xg[1], tz = xg[2];
            if (mine) {
                double part = 0.0;
                // (one loop body for cached and uncached keypoints)
                for (int q = 0, i = tid; i < n; q++, i += nthr) {
                    float Px, Py, Pz, k0, k1;
                    if (!fetch_kp(q, i, Px, Py, Pz, k0, k1)) continue;
                    float u, v;
                    dev_project_nd(nodist, Rd, Px, Py, Pz, tx, ty, tz, cam.fx, cam.fy, cam.cx, cam.cy, cam, u, v);
                    float d0 = fabsf(u - k0), d1 = fabsf(v - k1);
                    part += (double)(d0 + d1);
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) part += __shfl_down_sync(0xffffffffu, part, o);
                if (lane == 0) hdr.cost_part[cbuf][warp] = part;
            }
            __syncthreads();
            const int cb = cbuf;
            cbuf ^= 1;
            const double *costs = nullptr;
            double tot_local = 0.0;
            if (NCL > 1 && exchange) {
                // thread t < NCL pushes this CTA's cost into CTA t's table and signals CTA t's barrier.  CTAs without a trial in
                // this round push too (an unused slot): EVERY peer's push of exchange e + 1 then tells that the peer is done reading
                // the table of exchange e, which is what makes the two alternating buffers safe
                if (tid < NCL) {
                    const double tot = mine ? tree_sum<REF_WARPS>(hdr.cost_part[cb]) : 0.0;
                    dsmem_push_f64(&hdr.cl_cost[xb][crank], &hdr.xbar[xb], (unsigned)tid, tot);
                }
                if (tid == 0) mbar_expect_tx(&hdr.xbar[xb], NCL * 8);
                mbar_wait(&hdr.xbar[xb], xphase[xb]);
                xphase[xb] ^= 1;
                costs = hdr.cl_cost[xb];
                xb ^= 1;   // alternates over the whole solve: a buffer is free again exactly two exchanges later (see sparse_align_kernel)
            } else {
                tot_local = tree_sum<REF_WARPS>(hdr.cost_part[cb]);
            }
            trc.stamp(4, nvalid);
            // replay of the sequential driver over the costs of this round
            for (int g = 0; g < nvalid && mode != 3 && mode != 1; g++) {
                const float cost = (float)(costs ? costs[g] : tot_local);
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
                    if (xtslot >= 2 * REF_DEPTH) sp ^= 1; else tb ^= 1;
                } else if (fabs((double)(cost - prev_cost)) < 0.0001) {
                    mode = 3;
                } else {
                    kstep = kstep / 2;
                    jstep++;
                    it++;
                    if (it < 50) {
#pragma unroll
                        for (int k = 0; k < 6; k++) xt[k] = x0[k] + (kstep * grad[k]);
                        if (jstep < REF_DEPTH) {
                            xtslot = tb * REF_DEPTH + jstep;
                        } else {   // more than REF_DEPTH - 1 halvings: on demand (g is the last trial of its round here)
                            xtslot = 2 * REF_DEPTH + sp;
                            if (tid == 0) dev_rodrigues_d(-xt[3], -xt[4], -xt[5], hdr.Rd[xtslot]);
                            __syncthreads();
                        }
                        mode = 2;
                    } else
                        mode = 3;
                }
            }
        } else {
            const double *Rd = hdr.Rd[x0slot];
            const float tx = x0[0], ty = x0[1], tz = x0[2];
            {
                float Rif[9];
#pragma unroll
                for (int k = 0; k < 9; k++) Rif[k] = (float)Rd[k];
                float acc[RNGRAD];
#pragma unroll
                for (int k = 0; k < RNGRAD; k++) acc[k] = 0.f;
                auto accumulate = [&](float Px, float Py, float Pz, float k0, float k1) {
                    float u, v;
                    dev_project_nd(nodist, Rd, Px, Py, Pz, tx, ty, tz, cam.fx, cam.fy, cam.cx, cam.cy, cam, u, v);
                    float d0 = k0 - u, d1 = k1 - v;
                    if ((fabs((double)d0) > 3.0) || (fabs((double)d1) > 3.0)) return;
                    float X, Y, Z;
                    dev_m33v(Rif, Px - tx, Py - ty, Pz - tz, X, Y, Z);
                    float J[12];
                    dev_jacobian(cam.fx, cam.fy, X, Y, Z, J);
                    int hk = 0;
#pragma unroll
                    for (int p = 0; p < 6; p++)
#pragma unroll
                        for (int q2 = p; q2 < 6; q2++) {
                            float t = 0.f;
                            t += J[p] * J[q2];
                            t += J[6 + p] * J[6 + q2];
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
                };
                for (int q = 0, i = tid; i < n; q++, i += nthr) {
                    float Px, Py, Pz, k0, k1;
                    if (fetch_kp(q, i, Px, Py, Pz, k0, k1)) accumulate(Px, Py, Pz, k0, k1);
                }
                trc.stamp(7);
                const float t = warp_sum_scatter<RNGRAD>(acc, lane);   // lane k: warp sum of term k
                if (lane < RNGRAD) hdr.grad_part[lane][warp] = t;
            }
            __syncthreads();
            if (tid < RNGRAD) {
                hdr.red_out[tid] = tree_sum<REF_WARPS>(hdr.grad_part[tid]);
            }
            __syncthreads();
            trc.stamp(8);
            n_grads++;
            // 24 lanes (3 in each of 8 warps) solve the same 6x6 system side by side and go straight on to the matrix of "their"
            // trial pose x0 + 2^-j grad, j = warp + 8 * lane: no block barrier between the solve and the Rodrigues formulas
            if (warp < 8 && lane < REF_DEPTH / 8) {
                double dx[6];
                float tw[6], g[6];
                if (dev_solve6(hdr.red_out, hdr.red_out + 21, dx)) {
#pragma unroll
                    for (int k = 0; k < 6; k++) tw[k] = (float)dx[k];
                } else {
                    dev_pinv_step(hdr.red_out, hdr.red_out + 21, tw);   // rank-deficient H: the reference's pseudo-inverse step
                }
                trc.stamp(9);
                dev_expmap(tw, g);  // used as is — not rotated to world (pose_refinement.cpp:401-411)
                if (tid == 0) {
#pragma unroll
                    for (int k = 0; k < 6; k++) hdr.grad[k] = g[k];
                }
                const int j = warp + 8 * lane;
                float kj = 1.f;
                for (int q = 0; q < j; q++) kj = kj / 2;   // the driver's own sequence of halvings
                dev_rodrigues_d(-(x0[3] + (kj * g[3])), -(x0[4] + (kj * g[4])), -(x0[5] + (kj * g[5])), hdr.Rd[tb * REF_DEPTH + j]);
                trc.stamp(11);
            }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < 6; k++) grad[k] = hdr.grad[k];
            kstep = 1.f;
            jstep = 0;
#pragma unroll
            for (int k = 0; k < 6; k++) xt[k] = x0[k] + (kstep * grad[k]);
            xtslot = tb * REF_DEPTH;
            mode = 2;
            trc.stamp(5);
        }
    }
    trc.stamp(6);
    trc.finish();
    if (tid == 0 && crank == 0) {
        for (int k = 0; k < 6; k++) a.pose_out[k] = x0[k];
        *a.cost_out = prev_cost;
        a.evals_out[0] = n_evals; a.evals_out[1] = n_grads;
    }
    if (a.rd_out && crank == 0 && tid < 9) a.rd_out[tid] = hdr.Rd[x0slot][tid];   // the matrix of the refined pose, for the depth filter
    if (NCL > 1) cluster_sync_all();   // no CTA exits while a peer may still push into its shared memory
}

template <int REF_THREADS>
static void launch_refine_cluster(const RefineArgs &a, cudaStream_t st)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(8, 1, 1);
    cfg.blockDim = dim3(REF_THREADS, 1, 1);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 8; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, reproj_refine_kernel<REF_THREADS, 8>, a, 1 << 30);
}

// bucket: upper bound of the keypoint count known at launch (graph capture) time; it picks the CTA size, so that a frame gives
// the same bits whether it is replayed from a graph or launched directly.  wide: the 8-CTA cluster variant (eight trial poses per
// round; same bits as the single CTA) — for a sequence that has the GPU to itself, where the SMs it occupies are idle anyway.
void launch_refine(const RefineArgs &a, int bucket, bool wide, cudaStream_t st)
{
    static const int big_from = getenv("SVO_REFINE_THREADS") && atoi(getenv("SVO_REFINE_THREADS")) == 512 ? 256 : 1024;   // developer A/B switch
    if (bucket > big_from) {   // C4: 114 instead of 128 us with 512 threads; no gain for a few hundred keypoints
        if (wide) launch_refine_cluster<512>(a, st);
        else reproj_refine_kernel<512, 1><<<1, 512, 0, st>>>(a, 1 << 30);
    } else {
        if (wide) launch_refine_cluster<256>(a, st);
        else reproj_refine_kernel<256, 1><<<1, 256, 0, st>>>(a, 1 << 30);
    }
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
