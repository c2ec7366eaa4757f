// Internal launcher interface between the C-ABI layer (context.cu) and the kernels.
#pragma once
#include "common.cuh"

struct LevelDesc {  // one image level in device memory
    uint8_t *ptr;   // first pixel of the logical image (for padded LK levels: points INSIDE the padded buffer)
    int w, h, pitch;
};

struct ImageSetDev {  // device view of one stereo frame ("slot")
    LevelDesc left[SVO_MAX_LEVELS];  // halfSample pyramid, pitch == w (contiguous rows)
    LevelDesc right0;
    LevelDesc lk[SVO_LK_LEVELS];     // LK pyramid, padded by SVO_LK_PAD px of REFLECT_101 border
    LevelDesc lkd[SVO_LK_LEVELS];    // Scharr derivatives of the LK levels: (Ix, Iy) int16 pairs, 4 bytes per pixel, pitch in BYTES,
                                     // SVO_LK_PAD px of zeros around (cv::buildOpticalFlowPyramid's deriv border is CONSTANT 0);
                                     // filled on demand (keyframes), see launch_lk_scharr
    int n_levels;
};


// ---- keypoint-block transfer (context.cu launches it as a kernel; the depth filter runs it as its epilogue)
// Between the page-locked host mirror and the device block, by the SMs instead of the copy engines: the block is laid out
// for the context's keypoint CAPACITY (C3: 115 KB in, 133 KB out), a DMA copy moves all of it and costs ~10-15 us of engine
// time per frame and direction (at 50 k frames/s the D2H engine would be ~75 % busy); this moves the n live entries of
// every array (C3, 430 keypoints: 21 KB in, 28 KB out) with 16-byte accesses.
struct IoCopyArgs {
    const uint8_t *src;
    uint8_t *dst;
    const int *n_ptr;        // live keypoints (in `src`'s header for the import, in device memory for the export)
    int max_n, narr;
    struct { unsigned off, elem, per_n; } arr[20];   // per_n: n * elem bytes, else elem bytes
};
#ifdef __CUDACC__
__device__ __forceinline__ void io_copy_block(const IoCopyArgs &a)   // whole CTA
{
    __shared__ unsigned io_first[21];   // first 16-byte chunk of every array in the flattened chunk list
    const int n = min(max(*a.n_ptr, 0), a.max_n);
    if (threadIdx.x == 0) {
        unsigned acc = 0;
        for (int k = 0; k < a.narr; k++) {
            io_first[k] = acc;
            acc += ((a.arr[k].per_n ? (unsigned)n * a.arr[k].elem : a.arr[k].elem) + 15) / 16;   // arrays are 64-byte aligned and padded
        }
        io_first[a.narr] = acc;
    }
    __syncthreads();
    const unsigned total = io_first[a.narr];
    // all loads of a batch are issued before the first store: when `src` is host memory the whole block costs about one
    // PCIe round trip after the one that fetched n
    for (unsigned base = threadIdx.x; base < total; base += blockDim.x * 8) {
        uint4 v[8];
        unsigned off[8];
#pragma unroll
        for (int q = 0; q < 8; q++) {
            const unsigned c = base + q * blockDim.x;
            off[q] = 0xffffffffu;
            if (c < total) {
                int k = 0;
                while (c >= io_first[k + 1]) k++;
                off[q] = a.arr[k].off + (c - io_first[k]) * 16;
                v[q] = *reinterpret_cast<const uint4 *>(a.src + off[q]);
            }
        }
#pragma unroll
        for (int q = 0; q < 8; q++)
            if (off[q] != 0xffffffffu) *reinterpret_cast<uint4 *>(a.dst + off[q]) = v[q];
    }
}
#endif

// ---- pyramid.cu
void launch_pyr_halfsample(const ImageSetDev &s, cudaStream_t st);
bool launch_lk_pyramid(const ImageSetDev &s, cudaStream_t st, const IoCopyArgs *io = nullptr);
bool lk_side_fusable(const ImageSetDev &s);   // pad + level 1 (+ keypoint import) in one launch
void launch_lk_scharr(const ImageSetDev &s, cudaStream_t st);   // derivative images of the three LK levels
int pyr_launch_count(const ImageSetDev &s);
struct IngestArgs {
    const uint8_t *src[2];   // left / right source (device pointers: device memory or mapped page-locked host memory)
    size_t spitch[2];
    uint8_t *dst[2];         // contiguous rows (pitch == w)
    int w, h;
    unsigned long long *t_start;          // device: %globaltimer when the frame's first kernel starts, or null
};
bool ingest_supported(const IngestArgs &a);
void launch_ingest(const IngestArgs &a, bool deep_queue, cudaStream_t st);
void launch_copy16(const void *src, void *dst, size_t bytes, int ctas, cudaStream_t st);

// ---- rectify.cu (EuRoC front end, euroc_input.cpp:48-49, :69-73)
struct RectifyMapArgs {
    double K[9], D[5], ir[9];  // camera matrix, (k1,k2,p1,p2,k3), inverse of P[:3,:3]*R
    int w, h;
    float *map1, *map2;        // w*h each
};
void launch_rectify_map(const RectifyMapArgs &a, cudaStream_t st);
void launch_rectify_pack(const float *map1, const float *map2, int n, int sw, int sh, uint32_t *packed, cudaStream_t st);
struct RemapArgs {
    const uint8_t *src[2];     // raw (distorted) images, left / right
    const uint32_t *map[2];    // packed maps; null = that image is not rectified
    uint8_t *dst[2];
    int w, h, dpitch;          // output size
    int sw, sh, spitch;        // source size
};
void launch_remap(const RemapArgs &a, cudaStream_t st);

// ---- align.cu
struct AlignArgs {
    LevelDesc prev[SVO_MAX_LEVELS], cur[SVO_MAX_LEVELS];
    const float *kps2d;    // n*2 previous-frame positions (level 0)
    const float *kps3d;    // n*3
    const uint8_t *flags;  // n or null
    const int *n_ptr;      // device pointer to n
    const float *pose_in;  // 6
    float *pose_out;       // 6
    float *cost_out;       // 1
    int *evals_out;        // 16
    float *scratch;        // cache of reference-patch gradients / sums: align_scratch_floats(max_kps) floats (global, L2 resident)
    int max_kps;
    DevCam cam;
    // probe mode: single level, single evaluation
    int probe_level;       // -1 = normal
    float *probe_grad;     // 6
    int cluster;           // CTAs (SMs) cooperating on one evaluation of the frame: 1, 2, 4, 8 or 16
    int groups;            // such clusters side by side, one trial pose of the line search each: 1, or 2 / 4 with cluster 8 / 4
    double *rd_out;        // 9: Rodrigues(-r) of pose_out in double, for the kernels that project with the aligned pose next (or null)
    int *dbg;              // developer aid (SVO_DEBUG_MARKS): where a stuck barrier wait was, else null
    unsigned long long *trace;   // developer aid (SVO_SOLVER_TRACE): [0] = entries, then (clock64 << 16 | extra << 8 | tag) per phase, else null
};
cudaError_t launch_align(const AlignArgs &a, cudaStream_t st);
size_t align_scratch_floats(int max_kps);
size_t align_smem_bytes(const AlignArgs &a);
cudaError_t align_init_device();  // once per device: opt in to 227 KB dynamic shared memory

// ---- klt.cu
#define KLT_TPL_CHUNKS 12                               // 16-byte chunks per lane: 3 planes x 32 int16
#define KLT_TPL_BYTES (KLT_TPL_CHUNKS * 32 * 16)        // per (keypoint, level)
struct KfTemplates {       // template cache of one keyframe: its keypoints with index first .. first + count - 1
    const uint4 *data;     // null: none (templates are built per frame)
    const float4 *hdr;
    int first, count;
};
struct KltTemplateArgs {
    LevelDesc lk[SVO_LK_LEVELS], lkd[SVO_LK_LEVELS];
    const float *kps2d;    // n*2 keyframe positions (device)
    int n;
    uint4 *data;
    float4 *hdr;
};
void launch_klt_templates(const KltTemplateArgs &a, cudaStream_t st);
struct KltArgs {
    const LevelDesc *kf_lk_table;  // device table: [keyframe_id][2 * SVO_LK_LEVELS]: image levels, then derivative levels
    const int *keyframe_ids;       // n (device) or null => use prev_fixed
    const KfTemplates *kf_tpl_table;  // device table [keyframe_id] or null
    const int *kp_index;           // n (device): index of the keypoint in its origin keyframe (template cache lookup) or null
    LevelDesc prev_fixed[SVO_LK_LEVELS];
    LevelDesc prev_fixed_deriv[SVO_LK_LEVELS];
    LevelDesc cur[SVO_LK_LEVELS];
    const float *prev_pts;         // n*2 (keyframe coordinates)
    const float *init_pts;         // n*2 or null => project kps3d with *pose
    const float *kps3d;            // n*3 (projection prologue)
    const float *pose;             // 6   (projection prologue)
    const double *pose_rd;         // 9: Rodrigues(-r) of *pose as the alignment kernel left it, or null (computed per warp)
    const int *n_ptr;
    float *next_pts;               // n*2
    uint8_t *status;               // n
    float *err;                    // n
    int *iters;                    // n or null: LK iterations summed over levels
    // gating epilogue (pose_refinement.cpp:125-150); null flags => no gating
    uint8_t *flags;                // n in/out
    float *kps2d_out;              // n*2: accepted tracked position, else the projected position
    int max_kps;
    DevCam cam;
};
void launch_klt(const KltArgs &a, bool wide, cudaStream_t st);   // wide: two warps per keypoint (a sequence alone on the device)

// ---- refine.cu
struct RefineArgs {
    const float *kps2d, *kps3d;
    const uint8_t *flags;
    const int *n_ptr;
    const float *pose_in;
    float *pose_out, *cost_out;
    int *evals_out;  // 2
    DevCam cam;
    const double *rd_in;   // 9: Rodrigues(-r) of pose_in as the alignment kernel left it (or null: computed here)
    double *rd_out;        // 9: Rodrigues(-r) of pose_out (or null)
    unsigned long long *trace;   // developer aid (SVO_SOLVER_TRACE), as in AlignArgs
};
void launch_refine(const RefineArgs &a, int bucket, bool wide, cudaStream_t st);
void launch_project(const float *pose, const float *kps3d, const int *n_ptr, int max_kps, DevCam cam, float *kps2d, cudaStream_t st);

// ---- stereo.cu
struct SsdArgs {
    LevelDesc left0, right0;
    const float *kps2d;
    const int *n_ptr;
    int mode;
    float *disparity;
    int max_kps;
    DevCam cam;
};
void launch_stereo_ssd(const SsdArgs &a, cudaStream_t st);

struct FilterArgs {
    const float *kf_pose_table;   // [keyframe_id][6]
    const int *keyframe_ids;
    const float *disparity;       // n
    const float *kps2d;           // n*2 positions used by the filter (KLT-refined or projected)
    const float *ref_kps2d;       // n*2 position in origin keyframe
    float *kps3d;                 // n*3 in/out
    uint8_t *flags;               // n in/out
    int *inlier, *outlier;        // n in/out
    float *kf_state;              // n*2 in/out
    const float *pose;            // 6 frame pose (refined)
    float *kps2d_out;             // n*2 re-projection of the updated points
    const int *n_ptr;
    int max_kps;
    DevCam cam;
    const double *rdn_in;         // 9: Rodrigues(-r) of `pose` as the refinement kernel left it (or null: computed here)
    int do_export;                // run io_copy_block(exp) when the update is done (results straight to the host mirror)
    IoCopyArgs exp;
    // completion record in the host mirror, written after the export: {frame sequence number, -, t_start, t_end (ns)} — the host
    // polls the sequence number instead of synchronising the stream (null: no record)
    unsigned long long *done_rec;        // 4 x u64 in mapped host memory
    const unsigned *seq_ptr;             // device copy of the frame header's sequence number
    const unsigned long long *t_start;   // device: stamp left by the ingest kernel
    unsigned long long *trace;           // developer aid (SVO_SOLVER_TRACE), as in AlignArgs, or null
};
void launch_depth_filter(const FilterArgs &a, cudaStream_t st);

// ---- detect.cu
struct DetectArgs {
    LevelDesc img;
    int grid_w, grid_h, level;
    int *score_map;      // w*h ints scratch (FAST scores, 0 = not a corner)
    float *cell_xy;      // cells*2
    float *cell_score;   // cells
    int *cell_type;      // cells
    int cells_x, cells_y;
};
void launch_detect(const DetectArgs &a, cudaStream_t st);
void launch_fast_list(const LevelDesc &img, int *score_map, int *xys, int max_out, int *count, cudaStream_t st);
