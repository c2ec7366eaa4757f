// Device layer of the C-ABI (include/svo_cuda.h, svo_* functions): context, device-resident image sets,
// per-stage entry points and the fused per-frame tracking sequence.  Host code only orchestrates: every
// arithmetic step of the hot path runs in the kernels of pyramid.cu / align.cu / klt.cu / refine.cu /
// stereo.cu / detect.cu.  There is no CPU fallback.
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <vector>
#include <atomic>
#include <chrono>

#include "kernels.cuh"

static char g_create_err[256] = "";
#include <chrono>
static const bool g_trace = getenv("SVO_TRACE_KF") != nullptr;   // developer trace of slow host-side calls
static inline double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

// developer aid (SVO_DEBUG_MARKS=1): single-thread kernels that stamp a progress id into mapped host memory between the
// stages of a frame, so a stuck stream can be attributed to a stage from the host (svo_debug_marks)
__global__ void mark_kernel(volatile int *p, int slot, int v) { p[slot] = v; p[0] = v; }

// Image-set arena: all contexts of one device and one image geometry carve their slots out of shared chunks, so that a
// keyframe (which needs a new image set) does not call cudaMalloc in steady state — an allocation under load stalls every
// stream of the process for milliseconds (tens to hundreds of ms with 32 sequences in flight).
#include <mutex>
struct SlotArena {
    int device; size_t slot_bytes; long long geom; int users;   // geom: image size and level count (same layout => zero frames stay valid)
    std::vector<uint8_t *> chunks, free_list;
};
static std::mutex g_arena_mu;
static std::vector<SlotArena *> g_arenas;
#define ARENA_CHUNK_SLOTS 32

static SlotArena *arena_get(int device, size_t slot_bytes, long long geom)
{
    std::lock_guard<std::mutex> lk(g_arena_mu);
    for (SlotArena *a : g_arenas)
        if (a->device == device && a->slot_bytes == slot_bytes && a->geom == geom) { a->users++; return a; }
    SlotArena *a = new SlotArena{device, slot_bytes, geom, 1, {}, {}};
    g_arenas.push_back(a);
    return a;
}
static cudaError_t arena_take(SlotArena *a, uint8_t **out)
{
    std::lock_guard<std::mutex> lk(g_arena_mu);
    if (a->free_list.empty()) {
        uint8_t *c = nullptr;
        cudaError_t e = cudaMalloc(&c, a->slot_bytes * ARENA_CHUNK_SLOTS);
        if (e != cudaSuccess) return e;
        e = cudaMemset(c, 0, a->slot_bytes * ARENA_CHUNK_SLOTS);
        if (e != cudaSuccess) { cudaFree(c); return e; }
        a->chunks.push_back(c);
        for (int k = ARENA_CHUNK_SLOTS - 1; k >= 0; k--) a->free_list.push_back(c + (size_t)k * a->slot_bytes);
    }
    *out = a->free_list.back();
    a->free_list.pop_back();
    return cudaSuccess;
}
static void arena_give(SlotArena *a, uint8_t *p)
{
    std::lock_guard<std::mutex> lk(g_arena_mu);
    a->free_list.push_back(p);
}
static void arena_release(SlotArena *a)
{
    std::lock_guard<std::mutex> lk(g_arena_mu);
    if (--a->users > 0) return;
    for (uint8_t *c : a->chunks) cudaFree(c);
    for (size_t k = 0; k < g_arenas.size(); k++)
        if (g_arenas[k] == a) { g_arenas.erase(g_arenas.begin() + k); break; }
    delete a;
}

__global__ void __launch_bounds__(256) io_copy_kernel(IoCopyArgs a) { io_copy_block(a); }

struct Slot {
    uint8_t *base = nullptr;
    ImageSetDev dev;
    int refcount = 0;
    bool derivs_valid = false;   // lkd[] holds the Scharr derivatives of lk[]
};

// layout of the keypoint I/O block (same offsets on device and in the pinned host mirror)
struct IoLayout {
    size_t n, pose_prior, prev_kps2d, ref_kps2d, kf_id, kp_index;  // in
    size_t kps3d, flags, inlier, outlier, kf_state;                // in/out
    size_t pose_aligned, pose_refined, rd_aligned, rd_refined, costs, evals, klt_pts, klt_err, klt_status, klt_iters, disparity, kps2d_ref_in, kps2d_out;  // out
    size_t in_end, inout_begin, total;
    size_t hdr_src, hdr_seq;   // inside the first 64-byte block, behind n: frame sources {left, right, pitch, pitch} (4 x u64), sequence number
    size_t stamps;             // device only: %globaltimer at the start of the frame's first kernel
    size_t done;               // host mirror: completion record {sequence number, -, t_start, t_end}
};

struct svo_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    svo_camera_settings cs;
    DevCam cam;
    int W = 0, H = 0, max_kps = 0, n_levels = 0;
    // slot layout
    size_t slot_bytes = 0;
    size_t off_left[SVO_MAX_LEVELS], off_right0, off_lk[SVO_LK_LEVELS], off_lkd[SVO_LK_LEVELS], image_bytes;
    int lkdpitch[SVO_LK_LEVELS];
    int lw[SVO_MAX_LEVELS], lh[SVO_MAX_LEVELS], lkw[SVO_LK_LEVELS], lkh[SVO_LK_LEVELS], lkpitch[SVO_LK_LEVELS];
    std::vector<Slot> slots;
    SlotArena *arena = nullptr;
    uint8_t *h_stage[2] = {nullptr, nullptr};  // pinned upload staging (double buffered)
    uint8_t *d_stage[2] = {nullptr, nullptr};  // ... as the device sees them (zero-copy ingest)
    const uint8_t *zc_left = nullptr, *zc_right = nullptr;   // device-visible addresses of the host images being uploaded
    int stage_idx = 0;
    // keyframe tables
    LevelDesc *d_kf_lk = nullptr;
    float *d_kf_pose = nullptr;
    KfTemplates *d_kf_tpl = nullptr;           // per keyframe: LK templates of the keypoints it introduced (svo_keyframe_set_templates)
    bool use_templates = true;
    std::vector<uint8_t *> tpl_chunks;         // bump-allocated, never freed before the context goes (keyframes live forever)
    size_t tpl_chunk_bytes = 0, tpl_chunk_used = 0, tpl_first_bytes = 0;
    float *d_tpl_kps = nullptr;                // cell_cap * 2 floats: keyframe positions handed to klt_template_kernel
    int kf_cap = 0, kf_count = 0;
    std::vector<int> kf_slot;
    // keypoint I/O
    IoLayout lay;
    uint8_t *d_io = nullptr, *h_io = nullptr;
    uint8_t *d_hio = nullptr;                 // h_io as the device sees it (SM-driven keypoint-block transfers), else null
    IoCopyArgs io_in, io_out;
    float *d_align_scratch = nullptr;
    uint8_t *d_detect_scratch = nullptr;  // 2*W*H bytes
    float *d_cell_xy = nullptr, *d_cell_score = nullptr;
    int *d_cell_type = nullptr;
    int cell_cap = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaStream_t stream2 = nullptr;            // side branch of the captured frame graph (LK pyramid || alignment, SSD || refinement)
    cudaEvent_t fev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    float last_ms = 0;
    int last_launches = 0;
    bool track_pending = false;
    long long launch_total = 0;
    bool profiling = false;
    cudaEvent_t sev[9] = {nullptr};  // stage events
    // CUDA-graph cache of the whole per-frame sequence (upload + pyramids + tracking), keyed by the fixed resources
    struct FrameGraph {
        int prev_slot, cur_slot, bucket, src_kind, stage_idx;
        bool slim;
        cudaGraphExec_t exec;
        cudaGraphNode_t left_copy, right_copy;
        unsigned long long last_use;
        int launches;
    };
    std::vector<FrameGraph> graphs;
    bool use_graphs = true;
    int solver_width = -1;    // svo_set_solver_width: -1 auto (wide while fewer than four sequences share the device), 0 sequential, 1 wide
    int align_groups4 = 4;    // wide line search with 4-CTA clusters: 4 or 2 trial poses per round (developer switch SVO_ALIGN_GROUPS)
    int align_cluster = 0;    // SMs per alignment solve (svo_set_align_cluster); 0 = by keypoint count: 8, or 16 above 1024 keypoints
    bool use_fork = true;     // two-branch frame graph (SVO_NO_FORK=1: linear chain)
    bool import_rode = false; // the keypoint import of the frame being captured is part of lk_side_kernel (no launch of its own)
    bool use_ingest = true;   // SM-driven frame ingest instead of copy-engine DMA (SVO_NO_INGEST=1 turns it off)
    // two driver calls per frame (ingest launch + graph launch) instead of nine: the frame's device time comes from %globaltimer
    // stamps instead of two event records and a query, completion from a sequence number the last kernel stores behind the results
    // in the host mirror instead of a stream synchronisation (SVO_NO_SLIM=1: off)
    bool use_slim = true;
    bool capturing_slim = false, slim_pending = false;
    unsigned frame_seq = 0;
    unsigned long long graph_clock = 0;
    long long graph_launches = 0, graph_captures = 0, graph_updates = 0;
    float stage_ms[8] = {0};
    // EuRoC rectification in front of the pyramid build (euroc_input.cpp:48-49, :69-73); [0] = left, [1] = right input
    float *d_rect_map[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};
    uint32_t *d_rect_packed[2] = {nullptr, nullptr};
    uint8_t *d_raw[2] = {nullptr, nullptr};   // raw (distorted) images of the frame being uploaded
    int *h_marks = nullptr, *d_marks = nullptr;   // SVO_DEBUG_MARKS
    unsigned long long *h_trace = nullptr, *d_trace = nullptr;   // SVO_SOLVER_TRACE: 2 x 1024 words (alignment, refinement), mapped host memory
    int mark_seq = 0;
    char err[256];
};

// diagnostic (tools/ only): SVO_DIAG_DUP=<stage>:<count> repeats an idempotent stage of the frame sequence so that its
// marginal GPU cost at saturation can be read off the throughput (stages: pyr, align, klt, refine, ssd)
static int diag_dup(const char *stage)
{
    static const char *env = getenv("SVO_DIAG_DUP");
    if (!env) return 0;
    const size_t l = strlen(stage);
    if (strncmp(env, stage, l) == 0 && env[l] == ':') return atoi(env + l + 1);
    return 0;
}

static void mark(svo_ctx *ctx, int id)
{
    if (ctx->d_marks) mark_kernel<<<1, 1, 0, ctx->stream>>>(ctx->d_marks, id, ++ctx->mark_seq);
}

static void destroy_graph(svo_ctx::FrameGraph &g);

template <class T> static T *io_ptr(uint8_t *base, size_t off) { return reinterpret_cast<T *>(base + off); }

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static void make_layout(IoLayout &L, int M)
{
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t r = o; o = align_up(o + bytes, 64); return r; };
    L.n = take(16);
    L.pose_prior = take(6 * 4);
    L.prev_kps2d = take((size_t)M * 8);
    L.ref_kps2d = take((size_t)M * 8);
    L.kf_id = take((size_t)M * 4);
    L.kp_index = take((size_t)M * 4);
    L.in_end = o;
    L.inout_begin = o;
    L.kps3d = take((size_t)M * 12);
    L.flags = take((size_t)M);
    L.inlier = take((size_t)M * 4);
    L.outlier = take((size_t)M * 4);
    L.kf_state = take((size_t)M * 8);
    L.pose_aligned = take(6 * 4);
    L.pose_refined = take(6 * 4);
    L.rd_aligned = take(9 * 8);     // Rodrigues(-r) of the two poses in double, handed from kernel to kernel (device only)
    L.rd_refined = take(9 * 8);
    L.costs = take(2 * 4);
    L.evals = take(18 * 4);
    L.klt_pts = take((size_t)M * 8);
    L.klt_err = take((size_t)M * 4);
    L.klt_status = take((size_t)M);
    L.klt_iters = take((size_t)M * 4);
    L.disparity = take((size_t)M * 4);
    L.kps2d_ref_in = take((size_t)M * 8);
    L.kps2d_out = take((size_t)M * 8);
    L.stamps = take(16);
    L.done = take(32);
    L.total = o;
    L.hdr_src = L.n + 16;
    L.hdr_seq = L.n + 48;
}

#define CK(expr) SVO_CUDA_CHECK(ctx->err, expr)

static inline void svo_cpu_relax()
{
#if defined(__x86_64__) || defined(__i386__)
    __builtin_ia32_pause();
#elif defined(__aarch64__)
    asm volatile("yield");
#endif
}

extern "C" int svo_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

extern "C" const char *svo_last_error(svo_ctx *ctx) { return ctx ? ctx->err : g_create_err; }

static int alloc_slot(svo_ctx *ctx, int *slot_out);
extern "C" int svo_ctx_destroy(svo_ctx *ctx);
static int fail_create(svo_ctx *ctx, int code, const char *msg)
{
    snprintf(g_create_err, sizeof(g_create_err), "%s", msg);
    if (ctx) svo_ctx_destroy(ctx);   // null-safe per member: streams, events, pinned and device buffers, arena share
    return code;
}

extern "C" int svo_ctx_create(const svo_camera_settings *s, int device, int width, int height, int max_keypoints, svo_ctx **out)
{
    if (!s || !out || width < 16 || height < 16) return fail_create(nullptr, SVO_ERR_INVALID, "bad arguments");
    if (s->max_pyramid_levels < 1 || s->max_pyramid_levels > 7 || s->min_pyramid_level_pose_estimation < 0 ||
        s->min_pyramid_level_pose_estimation > s->max_pyramid_levels)
        return fail_create(nullptr, SVO_ERR_INVALID, "max_pyramid_levels must be in 1..7 and min level in 0..max");
    if (s->window_size_opt_flow < 3 || s->window_size_opt_flow > 31) return fail_create(nullptr, SVO_ERR_INVALID, "window_size_opt_flow must be in 3..31");
    if (s->window_size_depth_calculator < 1 || s->window_size_depth_calculator > 63 || s->search_x < 0 || s->search_y < 0)
        return fail_create(nullptr, SVO_ERR_INVALID, "bad depth calculator window / search range");
    if (s->window_size_pose_estimator < 1 || s->window_size_pose_estimator > 16) return fail_create(nullptr, SVO_ERR_INVALID, "window_size_pose_estimator must be in 1..16");
    if (s->grid_width < 2 || s->grid_height < 2) return fail_create(nullptr, SVO_ERR_INVALID, "grid cell must be at least 2x2");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail_create(nullptr, SVO_ERR_NO_DEVICE, "no CUDA device (there is no CPU fallback)");
    if (device < 0 || device >= ndev) return fail_create(nullptr, SVO_ERR_INVALID, "device index out of range");
    svo_ctx *ctx = new svo_ctx();
    ctx->err[0] = 0;
    ctx->device = device;
    ctx->cs = *s;
    ctx->cam = make_devcam(*s);
    ctx->W = width; ctx->H = height;
    ctx->n_levels = s->max_pyramid_levels;
    ctx->use_graphs = getenv("SVO_NO_GRAPHS") == nullptr;
    ctx->use_ingest = getenv("SVO_NO_INGEST") == nullptr;
    ctx->use_fork = getenv("SVO_NO_FORK") == nullptr;
    ctx->use_slim = getenv("SVO_NO_SLIM") == nullptr;
    if (getenv("SVO_INGEST_MIX")) {   // developer experiment: every second context uploads by copy-engine DMA
        static int counter = 0;
        ctx->use_ingest = (counter++ & 1) == 0;
    }
    if (getenv("SVO_SOLVER_WIDTH")) {
        const int w = atoi(getenv("SVO_SOLVER_WIDTH"));
        if (w >= -1 && w <= 1) ctx->solver_width = w;
    }
    if (getenv("SVO_ALIGN_GROUPS") && atoi(getenv("SVO_ALIGN_GROUPS")) == 2) ctx->align_groups4 = 2;
    if (getenv("SVO_ALIGN_CLUSTER")) {
        const int c = atoi(getenv("SVO_ALIGN_CLUSTER"));
        if (c == 1 || c == 2 || c == 4 || c == 8 || c == 16) ctx->align_cluster = c;
    }
    if (getenv("SVO_SOLVER_TRACE")) {
        if (cudaHostAlloc(&ctx->h_trace, 2048 * sizeof(unsigned long long), cudaHostAllocMapped) == cudaSuccess) {
            memset(ctx->h_trace, 0, 2048 * sizeof(unsigned long long));
            cudaHostGetDevicePointer((void **)&ctx->d_trace, ctx->h_trace, 0);
        }
    }
    if (getenv("SVO_DEBUG_MARKS")) {
        if (cudaHostAlloc(&ctx->h_marks, 64 * sizeof(int), cudaHostAllocMapped) == cudaSuccess) {
            memset(ctx->h_marks, 0, 64 * sizeof(int));
            cudaHostGetDevicePointer((void **)&ctx->d_marks, ctx->h_marks, 0);
        }
    }
    if (max_keypoints <= 0) {
        // one keypoint per grid cell per keyframe, surviving ones from older keyframes on top: 4x cells is ample
        max_keypoints = 4 * (width / s->grid_width + 1) * (height / s->grid_height + 1);
    }
    ctx->max_kps = (int)align_up((size_t)max_keypoints, 32);
#define CKC(expr)                                                                                      \
    do {                                                                                               \
        cudaError_t _e = (expr);                                                                       \
        if (_e != cudaSuccess) {                                                                       \
            char m[256];                                                                               \
            snprintf(m, sizeof(m), "%s failed: %s", #expr, cudaGetErrorString(_e));                    \
            return fail_create(ctx, SVO_ERR_CUDA, m);                                                  \
        }                                                                                              \
    } while (0)
    CKC(cudaSetDevice(device));
    CKC(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    CKC(cudaStreamCreateWithFlags(&ctx->stream2, cudaStreamNonBlocking));
    for (int k = 0; k < 5; k++) CKC(cudaEventCreateWithFlags(&ctx->fev[k], cudaEventDisableTiming));
    CKC(cudaEventCreate(&ctx->ev0));
    CKC(cudaEventCreate(&ctx->ev1));
    for (int k = 0; k < 9; k++) CKC(cudaEventCreate(&ctx->sev[k]));
    CKC(align_init_device());
    // ---- slot layout
    size_t o = 0;
    int w = width, h = height;
    for (int l = 0; l < ctx->n_levels; l++) {
        ctx->lw[l] = w; ctx->lh[l] = h;
        ctx->off_left[l] = o;
        o = align_up(o + (size_t)w * h + 16, 256);
        w /= 2; h /= 2;
        if (w < 1 || h < 1) return fail_create(ctx, SVO_ERR_INVALID, "image too small for max_pyramid_levels");
    }
    ctx->off_right0 = o;
    o = align_up(o + (size_t)width * height + 16, 256);
    w = width; h = height;
    for (int l = 0; l < SVO_LK_LEVELS; l++) {
        ctx->lkw[l] = w; ctx->lkh[l] = h;
        ctx->lkpitch[l] = (int)align_up((size_t)w + 2 * SVO_LK_PAD, 16);
        ctx->off_lk[l] = o;
        o = align_up(o + (size_t)ctx->lkpitch[l] * (h + 2 * SVO_LK_PAD) + 16, 256);
        w = (w + 1) / 2; h = (h + 1) / 2;
    }
    ctx->image_bytes = o;   // everything a keyframe copies; the derivative levels below are rebuilt from the copy
    for (int l = 0; l < SVO_LK_LEVELS; l++) {
        ctx->lkdpitch[l] = (int)align_up(((size_t)ctx->lkw[l] + 2 * SVO_LK_PAD) * 4, 16);
        ctx->off_lkd[l] = o;
        o = align_up(o + (size_t)ctx->lkdpitch[l] * (ctx->lkh[l] + 2 * SVO_LK_PAD) + 16, 256);
    }
    ctx->slot_bytes = o;
    ctx->arena = arena_get(device, ctx->slot_bytes, ((long long)width << 40) | ((long long)height << 16) | ctx->n_levels);
    {   // two frame slots + the first keyframes up front: steady-state tracking never allocates
        int ids[6];
        for (int k = 0; k < 6; k++) { int rc0 = alloc_slot(ctx, &ids[k]); if (rc0) return fail_create(ctx, rc0, ctx->err); }
        for (int k = 0; k < 6; k++) ctx->slots[ids[k]].refcount = 0;
    }
    for (int k = 0; k < 2; k++) {
        CKC(cudaMallocHost(&ctx->h_stage[k], (size_t)2 * width * height));
        if (cudaHostGetDevicePointer((void **)&ctx->d_stage[k], ctx->h_stage[k], 0) != cudaSuccess) { ctx->d_stage[k] = nullptr; cudaGetLastError(); }
    }
    make_layout(ctx->lay, ctx->max_kps);
    CKC(cudaMalloc(&ctx->d_io, ctx->lay.total));
    CKC(cudaMemset(ctx->d_io, 0, ctx->lay.total));
    CKC(cudaMallocHost(&ctx->h_io, ctx->lay.total));
    memset(ctx->h_io, 0, ctx->lay.total);
    if (getenv("SVO_IO_DMA") || cudaHostGetDevicePointer((void **)&ctx->d_hio, ctx->h_io, 0) != cudaSuccess) { ctx->d_hio = nullptr; cudaGetLastError(); }
    {
        const IoLayout &L = ctx->lay;
        IoCopyArgs &in = ctx->io_in, &out = ctx->io_out;
        in.src = ctx->d_hio; in.dst = ctx->d_io; in.n_ptr = reinterpret_cast<const int *>(ctx->d_hio + L.n); in.max_n = ctx->max_kps;
        out.src = ctx->d_io; out.dst = ctx->d_hio; out.n_ptr = reinterpret_cast<const int *>(ctx->d_io + L.n); out.max_n = ctx->max_kps;
        int k = 0;
        auto add = [&](IoCopyArgs &a, size_t off, unsigned elem, unsigned per_n) { a.arr[k].off = (unsigned)off; a.arr[k].elem = elem; a.arr[k].per_n = per_n; k++; };
        add(in, L.n, 64, 0); /* the 64-byte frame header: n, ..., sequence number */ add(in, L.pose_prior, 24, 0); add(in, L.prev_kps2d, 8, 1); add(in, L.ref_kps2d, 8, 1); add(in, L.kf_id, 4, 1); add(in, L.kp_index, 4, 1);
        add(in, L.kps3d, 12, 1); add(in, L.flags, 1, 1); add(in, L.inlier, 4, 1); add(in, L.outlier, 4, 1); add(in, L.kf_state, 8, 1);
        in.narr = k; k = 0;
        add(out, L.kps3d, 12, 1); add(out, L.flags, 1, 1); add(out, L.inlier, 4, 1); add(out, L.outlier, 4, 1); add(out, L.kf_state, 8, 1);
        add(out, L.pose_aligned, 24, 0); add(out, L.pose_refined, 24, 0); add(out, L.costs, 8, 0); add(out, L.evals, 72, 0);
        add(out, L.klt_pts, 8, 1); add(out, L.klt_err, 4, 1); add(out, L.klt_status, 1, 1); add(out, L.klt_iters, 4, 1);
        add(out, L.disparity, 4, 1); add(out, L.kps2d_ref_in, 8, 1); add(out, L.kps2d_out, 8, 1);
        out.narr = k;
    }
    CKC(cudaMalloc(&ctx->d_align_scratch, align_scratch_floats(ctx->max_kps) * sizeof(float)));
    CKC(cudaMalloc(&ctx->d_detect_scratch, (size_t)2 * width * height + 64));
    ctx->cell_cap = (width / 2 + 1) * (height / 2 + 1) / 1 + 16;
    ctx->cell_cap = (width / s->grid_width + 2) * (height / s->grid_height + 2) * 4 + 64;
    CKC(cudaMalloc(&ctx->d_cell_xy, (size_t)ctx->cell_cap * 8));
    CKC(cudaMalloc(&ctx->d_cell_score, (size_t)ctx->cell_cap * 4));
    CKC(cudaMalloc(&ctx->d_cell_type, (size_t)ctx->cell_cap * 4));
    ctx->kf_cap = 64;
    CKC(cudaMalloc(&ctx->d_kf_lk, (size_t)ctx->kf_cap * 2 * SVO_LK_LEVELS * sizeof(LevelDesc)));
    CKC(cudaMalloc(&ctx->d_kf_pose, (size_t)ctx->kf_cap * 24 * sizeof(float)));
    CKC(cudaMalloc(&ctx->d_kf_tpl, (size_t)ctx->kf_cap * sizeof(KfTemplates)));
    CKC(cudaMemset(ctx->d_kf_tpl, 0, (size_t)ctx->kf_cap * sizeof(KfTemplates)));
    CKC(cudaMalloc(&ctx->d_tpl_kps, (size_t)ctx->cell_cap * 8));
    ctx->use_templates = getenv("SVO_NO_TEMPLATES") == nullptr && s->window_size_opt_flow == 31;
    if (ctx->use_templates) {   // first chunk of the template cache up front: the first keyframes do not allocate
        const size_t cells = (size_t)(width / s->grid_width + 1) * (height / s->grid_height + 1);   // a first keyframe fills every cell
        const size_t first = std::max((size_t)32 << 20, align_up(cells * (SVO_LK_LEVELS * (KLT_TPL_BYTES + sizeof(float4))) * 5 / 4, 1 << 20));
        uint8_t *c = nullptr;
        CKC(cudaMalloc(&c, first));
        ctx->tpl_chunks.push_back(c);
        ctx->tpl_chunk_bytes = first; ctx->tpl_chunk_used = 0; ctx->tpl_first_bytes = first;
    }
#undef CKC
    *out = ctx;
    return SVO_OK;
}

extern "C" int svo_ctx_destroy(svo_ctx *ctx)
{
    if (!ctx) return SVO_ERR_INVALID;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    for (auto &g : ctx->graphs) destroy_graph(g);
    if (ctx->arena) {
        for (auto &s : ctx->slots) if (s.base) arena_give(ctx->arena, s.base);
        arena_release(ctx->arena);
    }
    for (int k = 0; k < 2; k++) if (ctx->h_stage[k]) cudaFreeHost(ctx->h_stage[k]);
    if (ctx->d_io) cudaFree(ctx->d_io);
    if (ctx->h_io) cudaFreeHost(ctx->h_io);
    if (ctx->h_marks) cudaFreeHost(ctx->h_marks);
    if (ctx->h_trace) cudaFreeHost(ctx->h_trace);
    if (ctx->d_align_scratch) cudaFree(ctx->d_align_scratch);
    if (ctx->d_detect_scratch) cudaFree(ctx->d_detect_scratch);
    if (ctx->d_cell_xy) cudaFree(ctx->d_cell_xy);
    if (ctx->d_cell_score) cudaFree(ctx->d_cell_score);
    if (ctx->d_cell_type) cudaFree(ctx->d_cell_type);
    if (ctx->d_kf_lk) cudaFree(ctx->d_kf_lk);
    if (ctx->d_kf_pose) cudaFree(ctx->d_kf_pose);
    if (ctx->d_kf_tpl) cudaFree(ctx->d_kf_tpl);
    if (ctx->d_tpl_kps) cudaFree(ctx->d_tpl_kps);
    for (uint8_t *c : ctx->tpl_chunks) cudaFree(c);
    for (int k = 0; k < 2; k++) {
        if (ctx->d_rect_map[k][0]) cudaFree(ctx->d_rect_map[k][0]);
        if (ctx->d_rect_map[k][1]) cudaFree(ctx->d_rect_map[k][1]);
        if (ctx->d_rect_packed[k]) cudaFree(ctx->d_rect_packed[k]);
        if (ctx->d_raw[k]) cudaFree(ctx->d_raw[k]);
    }
    for (int k = 0; k < 5; k++) if (ctx->fev[k]) cudaEventDestroy(ctx->fev[k]);
    if (ctx->stream2) cudaStreamDestroy(ctx->stream2);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    for (int k = 0; k < 9; k++) if (ctx->sev[k]) cudaEventDestroy(ctx->sev[k]);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return SVO_OK;
}

// Start over with an empty sequence: every image set and keyframe is given back, the keyframe tables and the template cache
// are rewound.  All device memory, the streams and the captured frame graphs stay (they refer to slots and tables by address).
extern "C" int svo_ctx_reset(svo_ctx *ctx)
{
    if (!ctx) return SVO_ERR_INVALID;
    ctx->err[0] = 0;
    if (ctx->track_pending) { snprintf(ctx->err, sizeof(ctx->err), "svo_ctx_reset while a frame is in flight"); return SVO_ERR_STATE; }
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    for (auto &s : ctx->slots) { s.refcount = 0; s.derivs_valid = false; }
    ctx->kf_count = 0;
    ctx->kf_slot.clear();
    if (ctx->d_kf_tpl) CK(cudaMemsetAsync(ctx->d_kf_tpl, 0, (size_t)ctx->kf_cap * sizeof(KfTemplates), ctx->stream));
    while (ctx->tpl_chunks.size() > 1) { cudaFree(ctx->tpl_chunks.back()); ctx->tpl_chunks.pop_back(); }
    ctx->tpl_chunk_used = 0;
    if (!ctx->tpl_chunks.empty()) ctx->tpl_chunk_bytes = ctx->tpl_first_bytes;
    return SVO_OK;
}

extern "C" int svo_sync(svo_ctx *ctx)
{
    if (!ctx) return SVO_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    return SVO_OK;
}

static int slot_ok(svo_ctx *ctx, int slot) { return slot >= 0 && slot < (int)ctx->slots.size() && ctx->slots[slot].refcount > 0; }

static int alloc_slot(svo_ctx *ctx, int *slot_out)
{
    int id = -1;
    for (size_t i = 0; i < ctx->slots.size(); i++)
        if (ctx->slots[i].refcount == 0) { id = (int)i; break; }
    if (id < 0) {
        Slot s;
        CK(arena_take(ctx->arena, &s.base));
        ImageSetDev &d = s.dev;
        d.n_levels = ctx->n_levels;
        for (int l = 0; l < SVO_MAX_LEVELS; l++) d.left[l] = LevelDesc{nullptr, 0, 0, 0};
        for (int l = 0; l < ctx->n_levels; l++) d.left[l] = LevelDesc{s.base + ctx->off_left[l], ctx->lw[l], ctx->lh[l], ctx->lw[l]};
        d.right0 = LevelDesc{s.base + ctx->off_right0, ctx->W, ctx->H, ctx->W};
        for (int l = 0; l < SVO_LK_LEVELS; l++)
            d.lk[l] = LevelDesc{s.base + ctx->off_lk[l] + (size_t)SVO_LK_PAD * ctx->lkpitch[l] + SVO_LK_PAD, ctx->lkw[l], ctx->lkh[l], ctx->lkpitch[l]};
        for (int l = 0; l < SVO_LK_LEVELS; l++)   // the zero frame comes from the arena's one-off memset; only the interior is ever written
            d.lkd[l] = LevelDesc{s.base + ctx->off_lkd[l] + (size_t)SVO_LK_PAD * ctx->lkdpitch[l] + (size_t)SVO_LK_PAD * 4, ctx->lkw[l], ctx->lkh[l], ctx->lkdpitch[l]};
        ctx->slots.push_back(s);
        id = (int)ctx->slots.size() - 1;
    }
    ctx->slots[id].refcount = 1;
    ctx->slots[id].derivs_valid = false;
    *slot_out = id;
    return SVO_OK;
}

// Scharr derivative levels of a slot (cv::buildOpticalFlowPyramid stores them next to every image level; here they
// are only built for image sets that serve as the REFERENCE of an optical-flow call, i.e. keyframes)
static int ensure_derivs(svo_ctx *ctx, int slot)
{
    Slot &s = ctx->slots[slot];
    if (s.derivs_valid) return SVO_OK;
    launch_lk_scharr(s.dev, ctx->stream);
    ctx->launch_total += SVO_LK_LEVELS;
    CK(cudaGetLastError());
    s.derivs_valid = true;
    return SVO_OK;
}

// src_kind: 0 = pageable host memory (staged through the context's pinned buffer), 1 = page-locked host memory
// (direct DMA), 2 = device memory
static int classify_source(const uint8_t *left, const uint8_t *right, const uint8_t **dev_left = nullptr, const uint8_t **dev_right = nullptr)
{
    static const bool skip_query = getenv("SVO_ASSUME_PAGEABLE") != nullptr;   // diagnostic switches
    static const bool assume_pinned = getenv("SVO_ASSUME_PINNED") != nullptr;  // (measures the cost of the two queries)
    if (dev_left) *dev_left = nullptr;
    if (dev_right) *dev_right = nullptr;
    if (skip_query) return 0;
    if (assume_pinned) {
        if (dev_left) *dev_left = left;
        if (dev_right) *dev_right = right;
        return 1;
    }
    cudaPointerAttributes al, ar;
    bool ok = cudaPointerGetAttributes(&al, left) == cudaSuccess && cudaPointerGetAttributes(&ar, right) == cudaSuccess;
    cudaGetLastError();  // clear a possible "invalid value" from querying unregistered memory on old drivers
    if (!ok) return 0;
    if (al.type == cudaMemoryTypeHost && ar.type == cudaMemoryTypeHost) {
        // the same query yields the address at which the device sees the page-locked buffer (zero-copy ingest)
        if (dev_left) *dev_left = (const uint8_t *)al.devicePointer;
        if (dev_right) *dev_right = (const uint8_t *)ar.devicePointer;
        return 1;
    }
    if ((al.type == cudaMemoryTypeDevice || al.type == cudaMemoryTypeManaged) && (ar.type == cudaMemoryTypeDevice || ar.type == cudaMemoryTypeManaged)) return 2;
    return 0;
}

// host part of a staged upload: copy the caller's (possibly strided) rows into the pinned staging buffer
static uint8_t *stage_images(svo_ctx *ctx, const uint8_t *left, size_t ls, const uint8_t *right, size_t rs, int stage_idx)
{
    const size_t img = (size_t)ctx->W * ctx->H;
    uint8_t *stage = ctx->h_stage[stage_idx];
    if (ls == (size_t)ctx->W && rs == (size_t)ctx->W) {
        memcpy(stage, left, img);
        memcpy(stage + img, right, img);
    } else {
        for (int y = 0; y < ctx->H; y++) {
            memcpy(stage + (size_t)y * ctx->W, left + (size_t)y * ls, ctx->W);
            memcpy(stage + img + (size_t)y * ctx->W, right + (size_t)y * rs, ctx->W);
        }
    }
    return stage;
}

// stream part of an upload: copies + pyramid kernels (stereo_slam.cpp:135-139).  For src_kind 0 `left`/`right`
// are the two halves of the staging buffer.
static int enqueue_pyramids(svo_ctx *ctx, Slot &s, bool forked = false, bool with_import = false);
static int enqueue_upload(svo_ctx *ctx, Slot &s, const uint8_t *left, size_t ls, const uint8_t *right, size_t rs, int src_kind,
                          bool copies_only = false)
{
    const size_t img = (size_t)ctx->W * ctx->H;
    const cudaMemcpyKind kind = src_kind == 2 ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    // a rectified input lands in the context's raw buffer; the remap kernel in front of the pyramids writes level 0
    uint8_t *dl = ctx->d_rect_packed[0] ? ctx->d_raw[0] : s.dev.left[0].ptr;
    uint8_t *dr = ctx->d_rect_packed[1] ? ctx->d_raw[1] : s.dev.right0.ptr;
    // preferred path: the SMs fetch the pair themselves (zero-copy from page-locked host memory, or device memory)
    if (ctx->use_ingest) {
        IngestArgs ia;
        ia.t_start = nullptr;
        ia.dst[0] = dl; ia.dst[1] = dr; ia.w = ctx->W; ia.h = ctx->H; ia.spitch[0] = ls; ia.spitch[1] = rs;
        bool ok = true;
        if (src_kind == 2) { ia.src[0] = left; ia.src[1] = right; }
        else { ia.src[0] = ctx->zc_left; ia.src[1] = ctx->zc_right; ok = ctx->zc_left && ctx->zc_right; }
        if (ok && ingest_supported(ia)) {
            mark(ctx, 1);
            // fewer than four live sequences of this geometry on the device: latency matters more than the other streams' traffic
            launch_ingest(ia, ctx->arena->users < 4, ctx->stream);
            mark(ctx, 2);
            ctx->launch_total += 1;
            CK(cudaGetLastError());
            if (copies_only) return SVO_OK;
            return enqueue_pyramids(ctx, s);
        }
    }
    // contiguous rows: one linear DMA (a 2-D copy of 752-byte rows costs one descriptor per row)
    if (ls == (size_t)ctx->W) CK(cudaMemcpyAsync(dl, left, img, kind, ctx->stream));
    else CK(cudaMemcpy2DAsync(dl, ctx->W, left, ls, ctx->W, ctx->H, kind, ctx->stream));
    if (rs == (size_t)ctx->W) CK(cudaMemcpyAsync(dr, right, img, kind, ctx->stream));
    else CK(cudaMemcpy2DAsync(dr, ctx->W, right, rs, ctx->W, ctx->H, kind, ctx->stream));
    if (copies_only) return SVO_OK;
    return enqueue_pyramids(ctx, s);
}

static bool rectifying(const svo_ctx *ctx) { return ctx->d_rect_packed[0] || ctx->d_rect_packed[1]; }
static int frame_pyr_launches(const svo_ctx *ctx, const Slot &s) { return pyr_launch_count(s.dev) + (rectifying(ctx) ? 1 : 0); }

// forked (only while capturing the frame graph): the LK pyramid goes to a side branch that rejoins in front of the KLT
// kernel, so it runs beside the half-sample pyramid, the keypoint upload and the alignment solve instead of before them
// with_import (forked only): the keypoint block is fetched at the head of the side branch, beside the half-sample pyramid
static int enqueue_pyramids(svo_ctx *ctx, Slot &s, bool forked, bool with_import)
{
    if (rectifying(ctx)) {
        RemapArgs ra;
        ra.src[0] = ctx->d_raw[0]; ra.src[1] = ctx->d_raw[1];
        ra.map[0] = ctx->d_rect_packed[0]; ra.map[1] = ctx->d_rect_packed[1];
        ra.dst[0] = s.dev.left[0].ptr; ra.dst[1] = s.dev.right0.ptr;
        ra.w = ra.sw = ctx->W; ra.h = ra.sh = ctx->H; ra.dpitch = ra.spitch = ctx->W;
        launch_remap(ra, ctx->stream);
    }
    cudaStream_t lk_stream = ctx->stream;
    if (forked) {
        CK(cudaEventRecord(ctx->fev[0], ctx->stream));
        CK(cudaStreamWaitEvent(ctx->stream2, ctx->fev[0], 0));
        lk_stream = ctx->stream2;
    }
    for (int k = 0; k <= diag_dup("pyr"); k++) {
        launch_pyr_halfsample(s.dev, ctx->stream);
        if (k == 0 && forked && with_import) {
            // the keypoint import rides in the first kernel of the side branch (lk_side_kernel); if that kernel cannot be used
            // for this geometry it goes ahead of the pyramid as a kernel of its own
            const bool fusable = lk_side_fusable(s.dev);
            if (!fusable) io_copy_kernel<<<1, 256, 0, ctx->stream2>>>(ctx->io_in);
            launch_lk_pyramid(s.dev, lk_stream, fusable ? &ctx->io_in : nullptr);
            ctx->import_rode = fusable;
            CK(cudaEventRecord(ctx->fev[4], ctx->stream2));
        } else
            launch_lk_pyramid(s.dev, lk_stream);
    }
    if (forked) CK(cudaEventRecord(ctx->fev[1], ctx->stream2));
    mark(ctx, 3);
    CK(cudaGetLastError());
    return SVO_OK;
}

static int upload_common(svo_ctx *ctx, const uint8_t *left, size_t ls, const uint8_t *right, size_t rs, int src_kind, int *slot_out)
{
    CK(cudaSetDevice(ctx->device));
    int id;
    int rc = alloc_slot(ctx, &id);
    if (rc) return rc;
    Slot &s = ctx->slots[id];
    const size_t img = (size_t)ctx->W * ctx->H;
    if (src_kind == 0) {
        // pageable memory goes through the context's own pinned staging buffer (double buffered; uploads on one
        // stream are ordered and the caller synchronises once per frame)
        uint8_t *stage = stage_images(ctx, left, ls, right, rs, ctx->stage_idx);
        ctx->zc_left = ctx->d_stage[ctx->stage_idx];
        ctx->zc_right = ctx->zc_left ? ctx->zc_left + img : nullptr;
        ctx->stage_idx ^= 1;
        left = stage; right = stage + img; ls = rs = (size_t)ctx->W;
    }
    if (ctx->profiling) CK(cudaEventRecord(ctx->sev[0], ctx->stream));
    if ((rc = enqueue_upload(ctx, s, left, ls, right, rs, src_kind))) { s.refcount = 0; return rc; }
    ctx->launch_total += frame_pyr_launches(ctx, s);
    if (ctx->profiling) CK(cudaEventRecord(ctx->sev[1], ctx->stream));
    *slot_out = id;
    return SVO_OK;
}

extern "C" int svo_upload_stereo(svo_ctx *ctx, const uint8_t *left, size_t ls, const uint8_t *right, size_t rs, int *slot_out)
{
    if (ctx) ctx->err[0] = 0;   // a failure below reports its own message, never a stale one
    if (!ctx || !left || !right || !slot_out || ls < (size_t)ctx->W || rs < (size_t)ctx->W) return SVO_ERR_INVALID;
    int kind = classify_source(left, right, &ctx->zc_left, &ctx->zc_right);
    if (kind == 2) {  // device memory must come through svo_upload_stereo_device: the host path would memcpy() from it
        snprintf(ctx->err, sizeof(ctx->err), "svo_upload_stereo: device pointer passed to the host entry point");
        return SVO_ERR_INVALID;
    }
    return upload_common(ctx, left, ls, right, rs, kind, slot_out);
}

extern "C" int svo_upload_stereo_device(svo_ctx *ctx, const uint8_t *d_left, size_t ls, const uint8_t *d_right, size_t rs, int *slot_out)
{
    if (ctx) ctx->err[0] = 0;   // a failure below reports its own message, never a stale one
    if (!ctx || !d_left || !d_right || !slot_out || ls < (size_t)ctx->W || rs < (size_t)ctx->W) return SVO_ERR_INVALID;
    return upload_common(ctx, d_left, ls, d_right, rs, 2, slot_out);
}

extern "C" int svo_keypoint_capacity(svo_ctx *ctx, int *max_keypoints)
{
    if (!ctx || !max_keypoints) return SVO_ERR_INVALID;
    *max_keypoints = ctx->max_kps;
    return SVO_OK;
}

extern "C" int svo_launch_count(svo_ctx *ctx, long long *launches)
{
    if (!ctx || !launches) return SVO_ERR_INVALID;
    *launches = ctx->launch_total;
    return SVO_OK;
}

// (P[:3,:3] * R)^-1 in double by the closed form cv::Matx33d::inv uses (euroc_input.cpp:48-49 -> initUndistortRectifyMap)
static bool inv_pr(const double P[9], const double R[9], double ir[9])
{
    double M[9];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) M[i * 3 + j] = P[i * 3 + 0] * R[0 * 3 + j] + P[i * 3 + 1] * R[1 * 3 + j] + P[i * 3 + 2] * R[2 * 3 + j];
    const double det = M[0] * (M[4] * M[8] - M[5] * M[7]) - M[1] * (M[3] * M[8] - M[5] * M[6]) + M[2] * (M[3] * M[7] - M[4] * M[6]);
    if (det == 0) return false;
    const double id = 1.0 / det;
    ir[0] = (M[4] * M[8] - M[5] * M[7]) * id; ir[1] = (M[2] * M[7] - M[1] * M[8]) * id; ir[2] = (M[1] * M[5] - M[2] * M[4]) * id;
    ir[3] = (M[5] * M[6] - M[3] * M[8]) * id; ir[4] = (M[0] * M[8] - M[2] * M[6]) * id; ir[5] = (M[2] * M[3] - M[0] * M[5]) * id;
    ir[6] = (M[3] * M[7] - M[4] * M[6]) * id; ir[7] = (M[1] * M[6] - M[0] * M[7]) * id; ir[8] = (M[0] * M[4] - M[1] * M[3]) * id;
    return true;
}

extern "C" int svo_set_rectification(svo_ctx *ctx, int which, const double K[9], const double D[5], const double R[9], const double P[9])
{
    if (!ctx || (which != 0 && which != 1) || !K || !D || !R || !P) return SVO_ERR_INVALID;
    if (ctx->W > 2046 || ctx->H > 2046) { snprintf(ctx->err, sizeof(ctx->err), "rectification supports images up to 2046x2046"); return SVO_ERR_INVALID; }
    if (ctx->track_pending) { snprintf(ctx->err, sizeof(ctx->err), "svo_set_rectification while a frame is in flight"); return SVO_ERR_STATE; }
    RectifyMapArgs a;
    if (!inv_pr(P, R, a.ir)) { snprintf(ctx->err, sizeof(ctx->err), "P*R is singular"); return SVO_ERR_INVALID; }
    CK(cudaSetDevice(ctx->device));
    const size_t n = (size_t)ctx->W * ctx->H;
    for (int k = 0; k < 2; k++)
        if (!ctx->d_rect_map[which][k]) CK(cudaMalloc(&ctx->d_rect_map[which][k], n * sizeof(float)));
    uint32_t *packed = ctx->d_rect_packed[which];
    if (!packed) CK(cudaMalloc(&packed, n * sizeof(uint32_t)));
    if (!ctx->d_raw[which]) CK(cudaMalloc(&ctx->d_raw[which], n + 16));
    for (int k = 0; k < 9; k++) a.K[k] = K[k];
    for (int k = 0; k < 5; k++) a.D[k] = D[k];
    a.w = ctx->W; a.h = ctx->H; a.map1 = ctx->d_rect_map[which][0]; a.map2 = ctx->d_rect_map[which][1];
    launch_rectify_map(a, ctx->stream);
    launch_rectify_pack(a.map1, a.map2, (int)n, ctx->W, ctx->H, packed, ctx->stream);
    ctx->launch_total += 2;
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->d_rect_packed[which] = packed;
    for (auto &g : ctx->graphs) destroy_graph(g);   // captured frame sequences do not contain the remap node
    ctx->graphs.clear();
    return SVO_OK;
}

extern "C" int svo_clear_rectification(svo_ctx *ctx)
{
    if (!ctx) return SVO_ERR_INVALID;
    if (ctx->track_pending) { snprintf(ctx->err, sizeof(ctx->err), "svo_clear_rectification while a frame is in flight"); return SVO_ERR_STATE; }
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    for (int k = 0; k < 2; k++) {
        if (ctx->d_rect_packed[k]) cudaFree(ctx->d_rect_packed[k]);
        ctx->d_rect_packed[k] = nullptr;
    }
    for (auto &g : ctx->graphs) destroy_graph(g);
    ctx->graphs.clear();
    return SVO_OK;
}

extern "C" int svo_rectification_maps(svo_ctx *ctx, int which, float *map1, float *map2)
{
    if (!ctx || (which != 0 && which != 1) || !map1 || !map2) return SVO_ERR_INVALID;
    if (!ctx->d_rect_packed[which]) { snprintf(ctx->err, sizeof(ctx->err), "no rectification set for input %d", which); return SVO_ERR_STATE; }
    CK(cudaSetDevice(ctx->device));
    const size_t n = (size_t)ctx->W * ctx->H * sizeof(float);
    CK(cudaMemcpyAsync(map1, ctx->d_rect_map[which][0], n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(map2, ctx->d_rect_map[which][1], n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return SVO_OK;
}

extern "C" int svo_debug_zero_copy_bandwidth(svo_ctx *ctx, const void *pinned_host, size_t bytes, int ctas, int reps, float *gb_per_s)
{
    if (!ctx || !pinned_host || !gb_per_s || bytes < 16 || ctas < 1 || reps < 1) return SVO_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    void *dp = nullptr, *dst = nullptr;
    CK(cudaHostGetDevicePointer(&dp, const_cast<void *>(pinned_host), 0));
    CK(cudaMalloc(&dst, bytes));
    launch_copy16(dp, dst, bytes, ctas, ctx->stream);
    CK(cudaEventRecord(ctx->ev0, ctx->stream));
    for (int k = 0; k < reps; k++) launch_copy16(dp, dst, bytes, ctas, ctx->stream);
    CK(cudaEventRecord(ctx->ev1, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    cudaFree(dst);
    *gb_per_s = (float)((double)bytes * reps / (ms * 1e-3) / 1e9);
    return SVO_OK;
}

extern "C" int svo_debug_solver_trace(svo_ctx *ctx, int which, unsigned long long *out, int cap)
{
    if (!ctx || !out || which < 0 || which > 2 || cap < 1) return SVO_ERR_INVALID;
    if (!ctx->h_trace) { out[0] = 0; return SVO_OK; }
    const volatile unsigned long long *src = ctx->h_trace + (which == 2 ? 512 : 1024 * which);
    const unsigned long long n0 = src[0]; const int n = n0 < 1000 ? (int)n0 : 1000;
    for (int k = 0; k <= n && k < cap; k++) out[k] = src[k];
    return SVO_OK;
}

extern "C" int svo_debug_marks(svo_ctx *ctx, int *out64)
{
    if (!ctx || !out64) return SVO_ERR_INVALID;
    for (int k = 0; k < 64; k++) out64[k] = ctx->h_marks ? ((volatile int *)ctx->h_marks)[k] : -1;
    out64[15] = ctx->mark_seq;
    return SVO_OK;
}

// wide line searches (svo_set_solver_width) for this launch / capture?
static bool wide_solvers(const svo_ctx *ctx) { return ctx->solver_width < 0 ? ctx->arena->users < 4 : ctx->solver_width != 0; }

extern "C" int svo_set_align_cluster(svo_ctx *ctx, int ctas)
{
    if (!ctx || (ctas != 0 && ctas != 1 && ctas != 2 && ctas != 4 && ctas != 8 && ctas != 16)) return SVO_ERR_INVALID;
    if (ctx->track_pending) { snprintf(ctx->err, sizeof(ctx->err), "svo_set_align_cluster while a frame is in flight"); return SVO_ERR_STATE; }
    if (ctas != ctx->align_cluster) {
        ctx->align_cluster = ctas;
        for (auto &g : ctx->graphs) destroy_graph(g);   // captured sequences hold the old launch shape
        ctx->graphs.clear();
    }
    return SVO_OK;
}

extern "C" int svo_set_solver_width(svo_ctx *ctx, int wide)
{
    if (!ctx || wide < -1 || wide > 1) return SVO_ERR_INVALID;
    if (ctx->track_pending) { snprintf(ctx->err, sizeof(ctx->err), "svo_set_solver_width while a frame is in flight"); return SVO_ERR_STATE; }
    if (wide != ctx->solver_width) {
        ctx->solver_width = wide;
        for (auto &g : ctx->graphs) destroy_graph(g);   // captured sequences hold the old launch shape
        ctx->graphs.clear();
    }
    return SVO_OK;
}

extern "C" int svo_set_profiling(svo_ctx *ctx, int on)
{
    if (!ctx) return SVO_ERR_INVALID;
    ctx->profiling = on != 0;
    return SVO_OK;
}

extern "C" int svo_last_stage_ms(svo_ctx *ctx, float *stage_ms8)
{
    if (!ctx || !stage_ms8) return SVO_ERR_INVALID;
    for (int k = 0; k < 8; k++) stage_ms8[k] = ctx->stage_ms[k];
    return SVO_OK;
}

extern "C" int svo_slot_retain(svo_ctx *ctx, int slot)
{
    if (!ctx || !slot_ok(ctx, slot)) return SVO_ERR_INVALID;
    ctx->slots[slot].refcount++;
    return SVO_OK;
}

extern "C" int svo_slot_release(svo_ctx *ctx, int slot)
{
    if (!ctx || !slot_ok(ctx, slot)) return SVO_ERR_INVALID;
    ctx->slots[slot].refcount--;
    return SVO_OK;
}

static bool level_desc(svo_ctx *ctx, int slot, int kind, int level, LevelDesc &d)
{
    const ImageSetDev &s = ctx->slots[slot].dev;
    if (kind == 0 && level >= 0 && level < ctx->n_levels) { d = s.left[level]; return true; }
    if (kind == 1 && level == 0) { d = s.right0; return true; }
    if (kind == 2 && level >= 0 && level < SVO_LK_LEVELS) { d = s.lk[level]; return true; }
    return false;
}

extern "C" int svo_slot_level_size(svo_ctx *ctx, int kind, int level, int *width, int *height)
{
    if (!ctx || !width || !height) return SVO_ERR_INVALID;
    if (kind == 0 && level >= 0 && level < ctx->n_levels) { *width = ctx->lw[level]; *height = ctx->lh[level]; return SVO_OK; }
    if (kind == 1 && level == 0) { *width = ctx->W; *height = ctx->H; return SVO_OK; }
    if ((kind == 2 || kind == 3) && level >= 0 && level < SVO_LK_LEVELS) { *width = ctx->lkw[level]; *height = ctx->lkh[level]; return SVO_OK; }
    if ((kind == 4 || kind == 5) && level >= 0 && level < SVO_LK_LEVELS) { *width = ctx->lkw[level] + 2 * SVO_LK_PAD; *height = ctx->lkh[level] + 2 * SVO_LK_PAD; return SVO_OK; }
    return SVO_ERR_INVALID;
}

extern "C" int svo_download_level(svo_ctx *ctx, int slot, int kind, int level, uint8_t *out, size_t out_stride)
{
    if (!ctx || !out || !slot_ok(ctx, slot)) return SVO_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    if (kind >= 3 && kind <= 5) {
        // 3: Scharr level (Ix, Iy int16 pairs, 4 bytes per pixel); 4 / 5: LK image level / Scharr level WITH the SVO_LK_PAD frame
        // cv::buildOpticalFlowPyramid keeps around them (BORDER_REFLECT_101 pixels / BORDER_CONSTANT zeros), stereo_slam.cpp:139
        if (level < 0 || level >= SVO_LK_LEVELS) return SVO_ERR_INVALID;
        if (kind != 4) { int rc = ensure_derivs(ctx, slot); if (rc) return rc; }
        const ImageSetDev &s = ctx->slots[slot].dev;
        const LevelDesc d = kind == 4 ? s.lk[level] : s.lkd[level];
        const size_t esz = kind == 4 ? 1 : 4;
        const int pad = kind == 3 ? 0 : SVO_LK_PAD;
        const size_t row = (size_t)(d.w + 2 * pad) * esz;
        if (out_stride < row) return SVO_ERR_INVALID;
        const uint8_t *src = d.ptr - (size_t)pad * d.pitch - (size_t)pad * esz;
        CK(cudaMemcpy2DAsync(out, out_stride, src, d.pitch, row, d.h + 2 * pad, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        return SVO_OK;
    }
    LevelDesc d;
    if (!level_desc(ctx, slot, kind, level, d) || out_stride < (size_t)d.w) return SVO_ERR_INVALID;
    CK(cudaMemcpy2DAsync(out, out_stride, d.ptr, d.pitch, d.w, d.h, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return SVO_OK;
}

// ------------------------------------------------------------------------------------------------ helpers
static int up(svo_ctx *ctx, size_t off, const void *src, size_t bytes)
{
    if (bytes == 0) return SVO_OK;
    CK(cudaMemcpyAsync(ctx->d_io + off, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return SVO_OK;
}
static int down(svo_ctx *ctx, void *dst, size_t off, size_t bytes)
{
    if (bytes == 0) return SVO_OK;
    CK(cudaMemcpyAsync(dst, ctx->d_io + off, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    return SVO_OK;
}
static int set_n(svo_ctx *ctx, int n)
{
    int v[4] = {n, 0, 0, 0};
    CK(cudaMemcpyAsync(ctx->d_io + ctx->lay.n, v, sizeof(v), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));  // v is on the stack
    return SVO_OK;
}
#define DP(T, field) io_ptr<T>(ctx->d_io, ctx->lay.field)

static int check_n(svo_ctx *ctx, int n)
{
    if (n < 0) return SVO_ERR_INVALID;
    if (n > ctx->max_kps) { snprintf(ctx->err, sizeof(ctx->err), "%d keypoints exceed the context capacity %d", n, ctx->max_kps); return SVO_ERR_CAPACITY; }
    return SVO_OK;
}

// ------------------------------------------------------------------------------------------------ stages
extern "C" int svo_detect_keypoints(svo_ctx *ctx, int slot, int level, int grid_w, int grid_h, int max_out, float *xy, float *score,
                                    int *type, int *n_out)
{
    if (ctx) ctx->err[0] = 0;   // a failure below reports its own message, never a stale one
    if (!ctx || !slot_ok(ctx, slot) || level < 0 || level >= ctx->n_levels || grid_w < 1 || grid_h < 1 || !n_out) return SVO_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    DetectArgs a;
    a.img = ctx->slots[slot].dev.left[level];
    a.grid_w = grid_w; a.grid_h = grid_h; a.level = level;
    a.cells_x = a.img.w / grid_w;
    a.cells_y = a.img.h / grid_h;
    if (a.cells_y < 1) a.cells_y = 1;  // the reference always emits the first cell row (corner_detector.cpp:27-31, :73-76)
    int cells = a.cells_x * a.cells_y;
    if (cells > ctx->cell_cap) { snprintf(ctx->err, sizeof(ctx->err), "%d grid cells exceed capacity %d", cells, ctx->cell_cap); return SVO_ERR_CAPACITY; }
    a.score_map = reinterpret_cast<int *>(ctx->d_detect_scratch);
    a.cell_xy = ctx->d_cell_xy; a.cell_score = ctx->d_cell_score; a.cell_type = ctx->d_cell_type;
    launch_detect(a, ctx->stream);
    ctx->launch_total += 2;
    CK(cudaGetLastError());
    int m = cells < max_out ? cells : max_out;
    if (m > 0) {
        CK(cudaMemcpyAsync(xy, ctx->d_cell_xy, (size_t)m * 8, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(score, ctx->d_cell_score, (size_t)m * 4, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(type, ctx->d_cell_type, (size_t)m * 4, cudaMemcpyDeviceToHost, ctx->stream));
    }
    CK(cudaStreamSynchronize(ctx->stream));
    *n_out = cells;
    return SVO_OK;
}

extern "C" int svo_fast_corners(svo_ctx *ctx, int slot, int level, int max_out, int *xys, int *n_out)
{
    if (!ctx || !slot_ok(ctx, slot) || level < 0 || level >= ctx->n_levels || !n_out) return SVO_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    LevelDesc img = ctx->slots[slot].dev.left[level];
    launch_fast_list(img, reinterpret_cast<int *>(ctx->d_detect_scratch), nullptr, 0, nullptr, ctx->stream);
    ctx->launch_total += 2;
    CK(cudaGetLastError());
    std::vector<uint8_t> nms((size_t)img.w * img.h);
    CK(cudaMemcpyAsync(nms.data(), ctx->d_detect_scratch + (size_t)img.w * img.h, nms.size(), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    int n = 0;
    for (int y = 0; y < img.h; y++)
        for (int x = 0; x < img.w; x++) {
            int s = nms[(size_t)y * img.w + x];
            if (!s) continue;
            if (n < max_out) { xys[3 * n] = x; xys[3 * n + 1] = y; xys[3 * n + 2] = s; }
            n++;
        }
    *n_out = n;
    return SVO_OK;
}

extern "C" int svo_stereo_match(svo_ctx *ctx, int slot, const float *kps2d, int n, int mode, float *disparity)
{
    if (ctx) ctx->err[0] = 0;   // a failure below reports its own message, never a stale one
    if (!ctx || !slot_ok(ctx, slot) || !kps2d || !disparity || (mode != 0 && mode != 1)) return SVO_ERR_INVALID;
    int rc = check_n(ctx, n);
    if (rc) return rc;
    if (n == 0) return SVO_OK;
    CK(cudaSetDevice(ctx->device));
    if ((rc = set_n(ctx, n))) return rc;
    if ((rc = up(ctx, ctx->lay.kps2d_ref_in, kps2d, (size_t)n * 8))) return rc;
    SsdArgs a;
    a.left0 = ctx->slots[slot].dev.left[0]; a.right0 = ctx->slots[slot].dev.right0;
    a.kps2d = DP(float, kps2d_ref_in); a.n_ptr = DP(int, n); a.mode = mode; a.disparity = DP(float, disparity);
    a.max_kps = n; a.cam = ctx->cam;
    launch_stereo_ssd(a, ctx->stream);
    ctx->launch_total += 1;
    CK(cudaGetLastError());
    if ((rc = down(ctx, disparity, ctx->lay.disparity, (size_t)n * 4))) return rc;
    CK(cudaStreamSynchronize(ctx->stream));
    return SVO_OK;
}

static void fill_align_args(svo_ctx *ctx, int prev_slot, int cur_slot, AlignArgs &a, int n, bool use_flags)
{
    for (int l = 0; l < SVO_MAX_LEVELS; l++) { a.prev[l] = ctx->slots[prev_slot].dev.left[l]; a.cur[l] = ctx->slots[cur_slot].dev.left[l]; }
    a.kps2d = DP(float, prev_kps2d); a.kps3d = DP(float, kps3d);
    a.flags = use_flags ? DP(uint8_t, flags) : nullptr;
    a.n_ptr = DP(int, n);
    a.pose_in = DP(float, pose_prior); a.pose_out = DP(float, pose_aligned);
    a.cost_out = DP(float, costs); a.evals_out = DP(int, evals);
    a.scratch = ctx->d_align_scratch; a.max_kps = ctx->max_kps; a.cam = ctx->cam;
    a.probe_level = -1; a.probe_grad = nullptr;
    a.dbg = ctx->d_marks ? ctx->d_marks + 16 : nullptr;
    a.trace = ctx->d_trace;
    a.cluster = ctx->align_cluster ? ctx->align_cluster : (n > 1024 ? 16 : 8);
    // wide line search: two (cluster 8) or up to four (cluster 4) trial poses per round; a frame of more than 1024 keypoints keeps
    // all 16 SMs of the largest cluster on one evaluation
    a.groups = 1;
    if (wide_solvers(ctx)) {
        if (a.cluster == 8) a.groups = 2;
        else if (a.cluster == 4) a.groups = ctx->align_groups4;
    }
    a.rd_out = nullptr;
    (void)n;
}

extern "C" int svo_align(svo_ctx *ctx, int prev_slot, int cur_slot, const float *kps2d, const float *kps3d, const uint8_t *flags, int n,
                         const float pose_in[6], float pose_out[6], float *cost, int *evals16)
{
    if (!ctx || !slot_ok(ctx, prev_slot) || !slot_ok(ctx, cur_slot) || !kps2d || !kps3d || !pose_in || !pose_out) return SVO_ERR_INVALID;
    int rc = check_n(ctx, n);
    if (rc) return rc;
    CK(cudaSetDevice(ctx->device));
    if ((rc = set_n(ctx, n))) return rc;
    if ((rc = up(ctx, ctx->lay.prev_kps2d, kps2d, (size_t)n * 8))) return rc;
    if ((rc = up(ctx, ctx->lay.kps3d, kps3d, (size_t)n * 12))) return rc;
    if (flags && (rc = up(ctx, ctx->lay.flags, flags, (size_t)n))) return rc;
    if ((rc = up(ctx, ctx->lay.pose_prior, pose_in, 24))) return rc;
    AlignArgs a;
    fill_align_args(ctx, prev_slot, cur_slot, a, n, flags != nullptr);
    CK(launch_align(a, ctx->stream));
    ctx->launch_total += 1;
    CK(cudaGetLastError());
    if ((rc = down(ctx, pose_out, ctx->lay.pose_aligned, 24))) return rc;
    if (cost && (rc = down(ctx, cost, ctx->lay.costs, 4))) return rc;
    if (evals16 && (rc = down(ctx, evals16, ctx->lay.evals, 64))) return rc;
    CK(cudaStreamSynchronize(ctx->stream));
    return SVO_OK;
}

extern "C" int svo_align_probe(svo_ctx *ctx, int prev_slot, int cur_slot, const float *kps2d, const float *kps3d, int n, int level,
                               const float pose[6], float *cost, float grad[6])
{
    if (!ctx || !slot_ok(ctx, prev_slot) || !slot_ok(ctx, cur_slot) || !kps2d || !kps3d || !pose || level < 0 || level >= ctx->n_levels)
        return SVO_ERR_INVALID;
    int rc = check_n(ctx, n);
    if (rc) return rc;
    CK(cudaSetDevice(ctx->device));
    if ((rc = set_n(ctx, n))) return rc;
    if ((rc = up(ctx, ctx->lay.prev_kps2d, kps2d, (size_t)n * 8))) return rc;
    if ((rc = up(ctx, ctx->lay.kps3d, kps3d, (size_t)n * 12))) return rc;
    if ((rc = up(ctx, ctx->lay.pose_prior, pose, 24))) return rc;
    AlignArgs a;
    fill_align_args(ctx, prev_slot, cur_slot, a, n, false);
    a.probe_level = level;
    a.probe_grad = DP(float, pose_refined);
    CK(launch_align(a, ctx->stream));
    ctx->launch_total += 1;
    CK(cudaGetLastError());
    if (cost && (rc = down(ctx, cost, ctx->lay.costs, 4))) return rc;
    if (grad && (rc = down(ctx, grad, ctx->lay.pose_refined, 24))) return rc;
    CK(cudaStreamSynchronize(ctx->stream));
    return SVO_OK;
}

static int klt_common(svo_ctx *ctx, const int *keyframe_ids, int prev_slot, int cur_slot, const float *prev_pts, const float *init_pts, int n,
                      float *next_pts, uint8_t *status, float *err)
{
    if (!prev_pts || !init_pts || !next_pts || !status || !err) return SVO_ERR_INVALID;
    int rc = check_n(ctx, n);
    if (rc) return rc;
    if (n == 0) return SVO_OK;
    CK(cudaSetDevice(ctx->device));
    if ((rc = set_n(ctx, n))) return rc;
    if ((rc = up(ctx, ctx->lay.ref_kps2d, prev_pts, (size_t)n * 8))) return rc;
    if ((rc = up(ctx, ctx->lay.kps2d_ref_in, init_pts, (size_t)n * 8))) return rc;
    KltArgs a;
    memset(&a, 0, sizeof(a));
    if (keyframe_ids) {
        for (int i = 0; i < n; i++)
            if (keyframe_ids[i] < 0 || keyframe_ids[i] >= ctx->kf_count) return SVO_ERR_INVALID;
        if ((rc = up(ctx, ctx->lay.kf_id, keyframe_ids, (size_t)n * 4))) return rc;
        a.kf_lk_table = ctx->d_kf_lk;
        a.keyframe_ids = DP(int, kf_id);
    } else {
        if ((rc = ensure_derivs(ctx, prev_slot))) return rc;
        for (int l = 0; l < SVO_LK_LEVELS; l++) { a.prev_fixed[l] = ctx->slots[prev_slot].dev.lk[l]; a.prev_fixed_deriv[l] = ctx->slots[prev_slot].dev.lkd[l]; }
    }
    for (int l = 0; l < SVO_LK_LEVELS; l++) a.cur[l] = ctx->slots[cur_slot].dev.lk[l];
    a.prev_pts = DP(float, ref_kps2d); a.init_pts = DP(float, kps2d_ref_in);
    a.n_ptr = DP(int, n);
    a.next_pts = DP(float, klt_pts); a.status = DP(uint8_t, klt_status); a.err = DP(float, klt_err);
    a.flags = nullptr; a.kps2d_out = nullptr; a.iters = nullptr;
    a.max_kps = n; a.cam = ctx->cam;
    launch_klt(a, wide_solvers(ctx), ctx->stream);
    ctx->launch_total += 1;
    CK(cudaGetLastError());
    if ((rc = down(ctx, next_pts, ctx->lay.klt_pts, (size_t)n * 8))) return rc;
    if ((rc = down(ctx, status, ctx->lay.klt_status, (size_t)n))) return rc;
    if ((rc = down(ctx, err, ctx->lay.klt_err, (size_t)n * 4))) return rc;
    CK(cudaStreamSynchronize(ctx->stream));
    return SVO_OK;
}

extern "C" int svo_klt(svo_ctx *ctx, const int *keyframe_ids, int cur_slot, const float *prev_pts, const float *init_pts, int n,
                       float *next_pts, uint8_t *status, float *err)
{
    if (!ctx || !keyframe_ids || !slot_ok(ctx, cur_slot)) return SVO_ERR_INVALID;
    return klt_common(ctx, keyframe_ids, -1, cur_slot, prev_pts, init_pts, n, next_pts, status, err);
}

extern "C" int svo_klt_slots(svo_ctx *ctx, int prev_slot, int cur_slot, const float *prev_pts, const float *init_pts, int n, float *next_pts,
                             uint8_t *status, float *err)
{
    if (!ctx || !slot_ok(ctx, prev_slot) || !slot_ok(ctx, cur_slot)) return SVO_ERR_INVALID;
    return klt_common(ctx, nullptr, prev_slot, cur_slot, prev_pts, init_pts, n, next_pts, status, err);
}

extern "C" int svo_reproj_refine(svo_ctx *ctx, const float *kps2d, const float *kps3d, const uint8_t *flags, int n, const float pose_in[6],
                                 float pose_out[6], float *cost, int *evals2)
{
    if (!ctx || !kps2d || !kps3d || !flags || !pose_in || !pose_out) return SVO_ERR_INVALID;
    int rc = check_n(ctx, n);
    if (rc) return rc;
    CK(cudaSetDevice(ctx->device));
    if ((rc = set_n(ctx, n))) return rc;
    if ((rc = up(ctx, ctx->lay.kps2d_ref_in, kps2d, (size_t)n * 8))) return rc;
    if ((rc = up(ctx, ctx->lay.kps3d, kps3d, (size_t)n * 12))) return rc;
    if ((rc = up(ctx, ctx->lay.flags, flags, (size_t)n))) return rc;
    if ((rc = up(ctx, ctx->lay.pose_aligned, pose_in, 24))) return rc;
    RefineArgs a;
    a.kps2d = DP(float, kps2d_ref_in); a.kps3d = DP(float, kps3d); a.flags = DP(uint8_t, flags); a.n_ptr = DP(int, n);
    a.pose_in = DP(float, pose_aligned); a.pose_out = DP(float, pose_refined);
    a.cost_out = DP(float, costs) + 1; a.evals_out = DP(int, evals) + 16; a.cam = ctx->cam;
    a.rd_in = nullptr; a.rd_out = nullptr; a.trace = ctx->d_trace ? ctx->d_trace + 1024 : nullptr;
    launch_refine(a, (n + 127) / 128 * 128, wide_solvers(ctx), ctx->stream);
    ctx->launch_total += 1;
    CK(cudaGetLastError());
    if ((rc = down(ctx, pose_out, ctx->lay.pose_refined, 24))) return rc;
    if (cost && (rc = down(ctx, cost, ctx->lay.costs + 4, 4))) return rc;
    if (evals2 && (rc = down(ctx, evals2, ctx->lay.evals + 64, 8))) return rc;
    CK(cudaStreamSynchronize(ctx->stream));
    return SVO_OK;
}

extern "C" int svo_depth_filter_update(svo_ctx *ctx, int slot, int n, const float *kps2d, const float *ref_kps2d, const int *keyframe_ids,
                                       float *kps3d, uint8_t *flags, int *inlier_count, int *outlier_count, float *kf_state,
                                       const float pose[6], float *disparity, float *kps2d_out)
{
    if (!ctx || !slot_ok(ctx, slot) || !pose) return SVO_ERR_INVALID;
    int rc = check_n(ctx, n);
    if (rc) return rc;
    if (n == 0) return SVO_OK;
    if (!kps2d || !ref_kps2d || !keyframe_ids || !kps3d || !flags || !inlier_count || !outlier_count || !kf_state) return SVO_ERR_INVALID;
    for (int i = 0; i < n; i++)
        if (keyframe_ids[i] < 0 || keyframe_ids[i] >= ctx->kf_count) { snprintf(ctx->err, sizeof(ctx->err), "keypoint %d: unknown keyframe id %d", i, keyframe_ids[i]); return SVO_ERR_INVALID; }
    CK(cudaSetDevice(ctx->device));
    if ((rc = set_n(ctx, n))) return rc;
    if ((rc = up(ctx, ctx->lay.kps2d_ref_in, kps2d, (size_t)n * 8))) return rc;
    if ((rc = up(ctx, ctx->lay.ref_kps2d, ref_kps2d, (size_t)n * 8))) return rc;
    if ((rc = up(ctx, ctx->lay.kf_id, keyframe_ids, (size_t)n * 4))) return rc;
    if ((rc = up(ctx, ctx->lay.kps3d, kps3d, (size_t)n * 12))) return rc;
    if ((rc = up(ctx, ctx->lay.flags, flags, (size_t)n))) return rc;
    if ((rc = up(ctx, ctx->lay.inlier, inlier_count, (size_t)n * 4))) return rc;
    if ((rc = up(ctx, ctx->lay.outlier, outlier_count, (size_t)n * 4))) return rc;
    if ((rc = up(ctx, ctx->lay.kf_state, kf_state, (size_t)n * 8))) return rc;
    if ((rc = up(ctx, ctx->lay.pose_refined, pose, 24))) return rc;
    SsdArgs sa;
    sa.left0 = ctx->slots[slot].dev.left[0]; sa.right0 = ctx->slots[slot].dev.right0;
    sa.kps2d = DP(float, kps2d_ref_in); sa.n_ptr = DP(int, n); sa.mode = 1; sa.disparity = DP(float, disparity);
    sa.max_kps = n; sa.cam = ctx->cam;
    launch_stereo_ssd(sa, ctx->stream);
    FilterArgs fa;
    fa.kf_pose_table = ctx->d_kf_pose; fa.keyframe_ids = DP(int, kf_id); fa.disparity = DP(float, disparity);
    fa.kps2d = DP(float, kps2d_ref_in); fa.ref_kps2d = DP(float, ref_kps2d); fa.kps3d = DP(float, kps3d);
    fa.flags = DP(uint8_t, flags); fa.inlier = DP(int, inlier); fa.outlier = DP(int, outlier); fa.kf_state = DP(float, kf_state);
    fa.pose = DP(float, pose_refined); fa.kps2d_out = DP(float, kps2d_out); fa.n_ptr = DP(int, n); fa.max_kps = n; fa.cam = ctx->cam;
    fa.do_export = 0; fa.rdn_in = nullptr; fa.done_rec = nullptr; fa.seq_ptr = nullptr; fa.t_start = nullptr; fa.trace = nullptr;
    launch_depth_filter(fa, ctx->stream);
    ctx->launch_total += 2;
    CK(cudaGetLastError());
    if ((rc = down(ctx, kps3d, ctx->lay.kps3d, (size_t)n * 12))) return rc;
    if ((rc = down(ctx, flags, ctx->lay.flags, (size_t)n))) return rc;
    if ((rc = down(ctx, inlier_count, ctx->lay.inlier, (size_t)n * 4))) return rc;
    if ((rc = down(ctx, outlier_count, ctx->lay.outlier, (size_t)n * 4))) return rc;
    if ((rc = down(ctx, kf_state, ctx->lay.kf_state, (size_t)n * 8))) return rc;
    if (disparity && (rc = down(ctx, disparity, ctx->lay.disparity, (size_t)n * 4))) return rc;
    if (kps2d_out && (rc = down(ctx, kps2d_out, ctx->lay.kps2d_out, (size_t)n * 8))) return rc;
    CK(cudaStreamSynchronize(ctx->stream));
    return SVO_OK;
}

extern "C" int svo_project(svo_ctx *ctx, const float pose[6], const float *kps3d, int n, float *kps2d)
{
    if (!ctx || !pose || !kps3d || !kps2d) return SVO_ERR_INVALID;
    int rc = check_n(ctx, n);
    if (rc) return rc;
    if (n == 0) return SVO_OK;
    CK(cudaSetDevice(ctx->device));
    if ((rc = set_n(ctx, n))) return rc;
    if ((rc = up(ctx, ctx->lay.kps3d, kps3d, (size_t)n * 12))) return rc;
    if ((rc = up(ctx, ctx->lay.pose_refined, pose, 24))) return rc;
    launch_project(DP(float, pose_refined), DP(float, kps3d), DP(int, n), n, ctx->cam, DP(float, kps2d_out), ctx->stream);
    ctx->launch_total += 1;
    CK(cudaGetLastError());
    if ((rc = down(ctx, kps2d, ctx->lay.kps2d_out, (size_t)n * 8))) return rc;
    CK(cudaStreamSynchronize(ctx->stream));
    return SVO_OK;
}

// host Rodrigues (same expression order as cv::Rodrigues) for the keyframe pose table
static void host_rodrigues_f(const float r[3], float R[9])
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

extern "C" int svo_keyframe_commit(svo_ctx *ctx, int slot, const float pose[6], int *keyframe_id_out)
{
    if (ctx) ctx->err[0] = 0;   // a failure below reports its own message, never a stale one
    if (!ctx || !slot_ok(ctx, slot) || !pose || !keyframe_id_out) return SVO_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    if (ctx->kf_count == ctx->kf_cap) {
        int ncap = ctx->kf_cap * 2;
        LevelDesc *nl;
        float *np;
        KfTemplates *nt;
        CK(cudaMalloc(&nl, (size_t)ncap * 2 * SVO_LK_LEVELS * sizeof(LevelDesc)));
        CK(cudaMalloc(&np, (size_t)ncap * 24 * sizeof(float)));
        CK(cudaMalloc(&nt, (size_t)ncap * sizeof(KfTemplates)));
        CK(cudaStreamSynchronize(ctx->stream));
        CK(cudaMemset(nt, 0, (size_t)ncap * sizeof(KfTemplates)));
        CK(cudaMemcpy(nt, ctx->d_kf_tpl, (size_t)ctx->kf_cap * sizeof(KfTemplates), cudaMemcpyDeviceToDevice));
        cudaFree(ctx->d_kf_tpl);
        ctx->d_kf_tpl = nt;
        CK(cudaMemcpy(nl, ctx->d_kf_lk, (size_t)ctx->kf_cap * 2 * SVO_LK_LEVELS * sizeof(LevelDesc), cudaMemcpyDeviceToDevice));
        CK(cudaMemcpy(np, ctx->d_kf_pose, (size_t)ctx->kf_cap * 24 * sizeof(float), cudaMemcpyDeviceToDevice));
        cudaFree(ctx->d_kf_lk); cudaFree(ctx->d_kf_pose);
        ctx->d_kf_lk = nl; ctx->d_kf_pose = np; ctx->kf_cap = ncap;
        for (auto &g : ctx->graphs) destroy_graph(g);   // captured kernels hold the old table pointers
        ctx->graphs.clear();
    }
    // the keyframe gets its OWN copy of the image set (device-to-device, 1.5 MB): frame slots then keep alternating
    // between two fixed buffers, which is what lets the per-frame sequence be replayed as a CUDA graph
    int kslot;
    {
        int rc2 = alloc_slot(ctx, &kslot);
        if (rc2) return rc2;
        CK(cudaMemcpyAsync(ctx->slots[kslot].base, ctx->slots[slot].base, ctx->image_bytes, cudaMemcpyDeviceToDevice, ctx->stream));
        if ((rc2 = ensure_derivs(ctx, kslot))) return rc2;
    }
    int id = ctx->kf_count;
    float rec[24];
    for (int k = 0; k < 6; k++) rec[k] = pose[k];
    float rp[3] = {pose[3], pose[4], pose[5]}, rn[3] = {-pose[3], -pose[4], -pose[5]};
    host_rodrigues_f(rp, rec + 6);
    host_rodrigues_f(rn, rec + 15);
    CK(cudaMemcpyAsync(ctx->d_kf_pose + (size_t)id * 24, rec, sizeof(rec), cudaMemcpyHostToDevice, ctx->stream));
    LevelDesc kfd[2 * SVO_LK_LEVELS];
    for (int l = 0; l < SVO_LK_LEVELS; l++) { kfd[l] = ctx->slots[kslot].dev.lk[l]; kfd[SVO_LK_LEVELS + l] = ctx->slots[kslot].dev.lkd[l]; }
    CK(cudaMemcpyAsync(ctx->d_kf_lk + (size_t)id * 2 * SVO_LK_LEVELS, kfd, sizeof(kfd), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));  // rec is on the stack
    slot = kslot;
    ctx->kf_slot.push_back(slot);
    ctx->kf_count++;
    *keyframe_id_out = id;
    return SVO_OK;
}

extern "C" int svo_keyframe_slot(svo_ctx *ctx, int keyframe_id, int *slot_out)
{
    if (ctx) ctx->err[0] = 0;   // a failure below reports its own message, never a stale one
    if (!ctx || !slot_out || keyframe_id < 0 || keyframe_id >= ctx->kf_count) return SVO_ERR_INVALID;
    *slot_out = ctx->kf_slot[keyframe_id];
    return SVO_OK;
}

// LK templates of the keypoints a keyframe introduced (indices first .. first + count - 1 of its keypoint list, positions
// kps2d in keyframe coordinates).  calcOpticalFlowPyrLK rebuilds the 31x31 template (Iw, Ix, Iy and the structure tensor) of
// every keypoint on every level for every frame from the keyframe's pyramid (optical_flow.cpp:41-44); position and pyramid
// never change after the keyframe is created, so the templates are built once here and the tracking kernel fetches them.
extern "C" int svo_keyframe_set_templates(svo_ctx *ctx, int keyframe_id, const float *kps2d, int first, int count)
{
    if (ctx) ctx->err[0] = 0;   // a failure below reports its own message, never a stale one
    if (!ctx || keyframe_id < 0 || keyframe_id >= ctx->kf_count || first < 0 || count < 0 || (count > 0 && !kps2d)) return SVO_ERR_INVALID;
    if (!ctx->use_templates || count == 0) return SVO_OK;
    if (count > ctx->cell_cap) { snprintf(ctx->err, sizeof(ctx->err), "%d templates exceed the cell capacity %d", count, ctx->cell_cap); return SVO_ERR_CAPACITY; }
    CK(cudaSetDevice(ctx->device));
    const size_t data_bytes = (size_t)count * SVO_LK_LEVELS * KLT_TPL_BYTES, hdr_bytes = align_up((size_t)count * SVO_LK_LEVELS * sizeof(float4), 256);
    const size_t need = data_bytes + hdr_bytes;
    if (ctx->tpl_chunks.empty() || ctx->tpl_chunk_used + need > ctx->tpl_chunk_bytes) {
        const size_t chunk = std::max(need, (size_t)32 << 20);
        uint8_t *c = nullptr;
        CK(cudaMalloc(&c, chunk));
        ctx->tpl_chunks.push_back(c);
        ctx->tpl_chunk_bytes = chunk; ctx->tpl_chunk_used = 0;
    }
    uint8_t *base = ctx->tpl_chunks.back() + ctx->tpl_chunk_used;
    ctx->tpl_chunk_used += need;
    CK(cudaMemcpyAsync(ctx->d_tpl_kps, kps2d, (size_t)count * 8, cudaMemcpyHostToDevice, ctx->stream));
    const Slot &ks = ctx->slots[ctx->kf_slot[keyframe_id]];
    KltTemplateArgs ta;
    for (int l = 0; l < SVO_LK_LEVELS; l++) { ta.lk[l] = ks.dev.lk[l]; ta.lkd[l] = ks.dev.lkd[l]; }
    ta.kps2d = ctx->d_tpl_kps; ta.n = count;
    ta.data = reinterpret_cast<uint4 *>(base); ta.hdr = reinterpret_cast<float4 *>(base + data_bytes);
    launch_klt_templates(ta, ctx->stream);
    ctx->launch_total += 1;
    CK(cudaGetLastError());
    KfTemplates rec;
    rec.data = ta.data; rec.hdr = ta.hdr; rec.first = first; rec.count = count;
    CK(cudaMemcpyAsync(ctx->d_kf_tpl + keyframe_id, &rec, sizeof(rec), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));   // rec is on the stack, kps2d is the caller's
    return SVO_OK;
}

// ------------------------------------------------------------------------------------------------ fused frame
static int validate_and_pack(svo_ctx *ctx, svo_track_io *io)
{
    if (ctx->track_pending) { snprintf(ctx->err, sizeof(ctx->err), "tracking frame begun twice without _end"); return SVO_ERR_STATE; }
    const int n = io->n;
    int rc = check_n(ctx, n);
    if (rc) return rc;
    if (n > 0 && (!io->prev_kps2d || !io->kps3d || !io->ref_kps2d || !io->keyframe_id || !io->flags || !io->inlier_count ||
                  !io->outlier_count || !io->kf_state || !io->kps2d))
        return SVO_ERR_INVALID;
    for (int i = 0; i < n; i++)
        if (io->keyframe_id[i] < 0 || io->keyframe_id[i] >= ctx->kf_count) { snprintf(ctx->err, sizeof(ctx->err), "keypoint %d: unknown keyframe id %d", i, io->keyframe_id[i]); return SVO_ERR_INVALID; }
    const IoLayout &L = ctx->lay;
    uint8_t *h = ctx->h_io;
    *io_ptr<int>(h, L.n) = n;
    memcpy(h + L.pose_prior, io->pose_prior, 24);
    memcpy(h + L.prev_kps2d, io->prev_kps2d, (size_t)n * 8);
    memcpy(h + L.ref_kps2d, io->ref_kps2d, (size_t)n * 8);
    memcpy(h + L.kf_id, io->keyframe_id, (size_t)n * 4);
    if (io->keypoint_index) memcpy(h + L.kp_index, io->keypoint_index, (size_t)n * 4);
    else memset(h + L.kp_index, 0xff, (size_t)n * 4);   // -1: no template lookup
    memcpy(h + L.kps3d, io->kps3d, (size_t)n * 12);
    memcpy(h + L.flags, io->flags, (size_t)n);
    memcpy(h + L.inlier, io->inlier_count, (size_t)n * 4);
    memcpy(h + L.outlier, io->outlier_count, (size_t)n * 4);
    memcpy(h + L.kf_state, io->kf_state, (size_t)n * 8);
    return SVO_OK;
}

// stream part of a tracking frame; grid_n = number of per-keypoint CTAs to launch (>= n; the kernels read n from
// device memory), prof = record the per-stage events
static int enqueue_track(svo_ctx *ctx, int prev_slot, int cur_slot, int n, int grid_n, bool prof, int *launches_out, bool forked = false,
                         bool import_on_branch = false)
{
    const IoLayout &L = ctx->lay;
    uint8_t *h = ctx->h_io;
    int launches = 0;
    // all inputs: the live entries by an SM-driven copy from the page-locked mirror, else one DMA of the whole capacity
    // (in + in/out regions are contiguous)
    if (import_on_branch) { CK(cudaStreamWaitEvent(ctx->stream, ctx->fev[4], 0)); if (!ctx->import_rode) launches++; }
    else if (ctx->d_hio) { io_copy_kernel<<<1, 256, 0, ctx->stream>>>(ctx->io_in); launches++; }
    else CK(cudaMemcpyAsync(ctx->d_io, h, L.pose_aligned, cudaMemcpyHostToDevice, ctx->stream));
    // 1. sparse image alignment (stereo_slam.cpp:60-67)
    AlignArgs aa;
    fill_align_args(ctx, prev_slot, cur_slot, aa, grid_n, true);   // grid_n: the keypoint bucket (decides the cluster size with the graph)
    aa.rd_out = DP(double, rd_aligned);
    if (prof) CK(cudaEventRecord(ctx->sev[2], ctx->stream));
    mark(ctx, 4);
    CK(launch_align(aa, ctx->stream)); launches++;
    for (int k = 0; k < diag_dup("align"); k++) CK(launch_align(aa, ctx->stream));
    mark(ctx, 5);
    if (prof) CK(cudaEventRecord(ctx->sev[3], ctx->stream));
    if (forked) CK(cudaStreamWaitEvent(ctx->stream, ctx->fev[1], 0));   // LK pyramid branch rejoins
    if (n > 0) {
        // 2. projection with the aligned pose + KLT against the origin keyframes + gating (stereo_slam.cpp:71-83)
        KltArgs ka;
        memset(&ka, 0, sizeof(ka));
        ka.kf_lk_table = ctx->d_kf_lk; ka.keyframe_ids = DP(int, kf_id);
        if (ctx->use_templates) { ka.kf_tpl_table = ctx->d_kf_tpl; ka.kp_index = DP(int, kp_index); }
        for (int l = 0; l < SVO_LK_LEVELS; l++) ka.cur[l] = ctx->slots[cur_slot].dev.lk[l];
        ka.prev_pts = DP(float, ref_kps2d); ka.init_pts = nullptr; ka.kps3d = DP(float, kps3d); ka.pose = DP(float, pose_aligned);
        ka.pose_rd = DP(double, rd_aligned);
        ka.n_ptr = DP(int, n); ka.next_pts = DP(float, klt_pts); ka.status = DP(uint8_t, klt_status); ka.err = DP(float, klt_err);
        ka.flags = DP(uint8_t, flags); ka.kps2d_out = DP(float, kps2d_ref_in); ka.max_kps = grid_n; ka.cam = ctx->cam;
        ka.iters = DP(int, klt_iters);
        launch_klt(ka, wide_solvers(ctx), ctx->stream); launches++;
        for (int k = 0; k < diag_dup("klt"); k++) launch_klt(ka, wide_solvers(ctx), ctx->stream);
        mark(ctx, 6);
    }
    // the stereo SSD of the depth filter only needs the positions the KLT stage left: with `forked` it runs beside the
    // reprojection refinement and rejoins in front of the filter update
    SsdArgs sa;
    if (n > 0) {
        sa.left0 = ctx->slots[cur_slot].dev.left[0]; sa.right0 = ctx->slots[cur_slot].dev.right0;
        sa.kps2d = DP(float, kps2d_ref_in); sa.n_ptr = DP(int, n); sa.mode = 1; sa.disparity = DP(float, disparity);
        sa.max_kps = grid_n; sa.cam = ctx->cam;
        if (forked) {
            CK(cudaEventRecord(ctx->fev[2], ctx->stream));
            CK(cudaStreamWaitEvent(ctx->stream2, ctx->fev[2], 0));
            launch_stereo_ssd(sa, ctx->stream2); launches++;
            for (int k = 0; k < diag_dup("ssd"); k++) { launch_stereo_ssd(sa, ctx->stream2); launches++; }
            CK(cudaEventRecord(ctx->fev[3], ctx->stream2));
        }
    }
    if (prof) CK(cudaEventRecord(ctx->sev[4], ctx->stream));
    // 3. reprojection Gauss-Newton (pose_refinement.cpp:175-177)
    RefineArgs ra;
    ra.kps2d = DP(float, kps2d_ref_in); ra.kps3d = DP(float, kps3d); ra.flags = DP(uint8_t, flags); ra.n_ptr = DP(int, n);
    ra.pose_in = DP(float, pose_aligned); ra.pose_out = DP(float, pose_refined);
    ra.cost_out = DP(float, costs) + 1; ra.evals_out = DP(int, evals) + 16; ra.cam = ctx->cam;
    ra.rd_in = DP(double, rd_aligned); ra.rd_out = DP(double, rd_refined); ra.trace = ctx->d_trace ? ctx->d_trace + 1024 : nullptr;
    const int ref_bucket = std::min(ctx->max_kps, (grid_n + 127) / 128 * 128);   // == the graph's bucket
    launch_refine(ra, ref_bucket, wide_solvers(ctx), ctx->stream); launches++;
    for (int k = 0; k < diag_dup("refine"); k++) launch_refine(ra, ref_bucket, wide_solvers(ctx), ctx->stream);
    mark(ctx, 7);
    if (prof) CK(cudaEventRecord(ctx->sev[5], ctx->stream));
    if (n > 0) {
        // 4. depth filter: disparities on the current stereo pair, then vote / triangulate / Kalman / flags / re-project
        if (forked) {
            CK(cudaStreamWaitEvent(ctx->stream, ctx->fev[3], 0));
        } else {
            launch_stereo_ssd(sa, ctx->stream); launches++;
            for (int k = 0; k < diag_dup("ssd"); k++) { launch_stereo_ssd(sa, ctx->stream); launches++; }
        }
        if (prof) CK(cudaEventRecord(ctx->sev[6], ctx->stream));
        FilterArgs fa;
        fa.kf_pose_table = ctx->d_kf_pose; fa.keyframe_ids = DP(int, kf_id); fa.disparity = DP(float, disparity);
        fa.kps2d = DP(float, kps2d_ref_in); fa.ref_kps2d = DP(float, ref_kps2d); fa.kps3d = DP(float, kps3d);
        fa.flags = DP(uint8_t, flags); fa.inlier = DP(int, inlier); fa.outlier = DP(int, outlier); fa.kf_state = DP(float, kf_state);
        fa.pose = DP(float, pose_refined); fa.kps2d_out = DP(float, kps2d_out); fa.n_ptr = DP(int, n); fa.max_kps = grid_n; fa.cam = ctx->cam;
        fa.do_export = ctx->d_hio != nullptr; fa.exp = ctx->io_out;   // results go to the host mirror from the same kernel
        fa.rdn_in = DP(double, rd_refined);
        fa.done_rec = nullptr; fa.seq_ptr = nullptr; fa.t_start = nullptr;
        fa.trace = ctx->d_trace ? ctx->d_trace + 512 : nullptr;   // third trace region (the alignment kernel uses < 100 of its 1024 words)
        if (ctx->capturing_slim && ctx->d_hio && !getenv("SVO_SLIM_NOREC")) {
            fa.done_rec = reinterpret_cast<unsigned long long *>(ctx->d_hio + L.done);
            fa.seq_ptr = reinterpret_cast<const unsigned *>(ctx->d_io + L.hdr_seq);
            fa.t_start = reinterpret_cast<const unsigned long long *>(ctx->d_io + L.stamps);
        }
        mark(ctx, 8);
        launch_depth_filter(fa, ctx->stream); launches++;
        mark(ctx, 9);
    } else if (prof) {
        CK(cudaEventRecord(ctx->sev[6], ctx->stream));
    }
    if (prof) CK(cudaEventRecord(ctx->sev[7], ctx->stream));
    CK(cudaGetLastError());
    // all outputs back to the page-locked mirror (in/out + out regions are contiguous)
    if (ctx->d_hio) { if (n <= 0) { io_copy_kernel<<<1, 256, 0, ctx->stream>>>(ctx->io_out); launches++; } }
    else CK(cudaMemcpyAsync(h + L.inout_begin, ctx->d_io + L.inout_begin, L.total - L.inout_begin, cudaMemcpyDeviceToHost, ctx->stream));
    mark(ctx, 10);
    *launches_out = launches;
    return SVO_OK;
}

extern "C" int svo_track_frame_begin(svo_ctx *ctx, int prev_slot, int cur_slot, svo_track_io *io)
{
    if (ctx) ctx->err[0] = 0;   // a failure below reports its own message, never a stale one
    if (!ctx || !io || !slot_ok(ctx, prev_slot) || !slot_ok(ctx, cur_slot)) return SVO_ERR_INVALID;
    int rc = validate_and_pack(ctx, io);
    if (rc) return rc;
    CK(cudaSetDevice(ctx->device));
    const bool prof = ctx->profiling;
    CK(cudaEventRecord(ctx->ev0, ctx->stream));
    int launches = 0;
    if ((rc = enqueue_track(ctx, prev_slot, cur_slot, io->n, io->n, prof, &launches))) return rc;
    CK(cudaEventRecord(ctx->ev1, ctx->stream));
    if (prof) CK(cudaEventRecord(ctx->sev[8], ctx->stream));
    ctx->last_launches = launches;
    ctx->launch_total += launches;
    ctx->track_pending = true;
    return SVO_OK;
}

// ---- upload + tracking as one replayable CUDA graph ------------------------------------------------------------
static void destroy_graph(svo_ctx::FrameGraph &g)
{
    if (g.exec) cudaGraphExecDestroy(g.exec);
    g.exec = nullptr;
}

static int capture_frame_graph(svo_ctx *ctx, svo_ctx::FrameGraph &g, int n, cudaGraph_t *graph_out)
{
    cudaGraph_t graph = nullptr;
    Slot &s = ctx->slots[g.cur_slot];
    CK(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
    int launches = 0;
    const bool forked = ctx->use_fork, branch_import = forked && ctx->d_hio != nullptr;
    ctx->capturing_slim = g.slim;
    int rc = enqueue_pyramids(ctx, s, forked, branch_import);
    if (!rc) rc = enqueue_track(ctx, g.prev_slot, g.cur_slot, n, g.bucket, false, &launches, forked, branch_import);
    ctx->capturing_slim = false;
    cudaError_t e = cudaStreamEndCapture(ctx->stream, &graph);
    if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
    CK(e);
    g.launches = launches + frame_pyr_launches(ctx, s);
    *graph_out = graph;
    return SVO_OK;
}

/* upload + pyramids + tracking in one call (what StereoSlam::new_image does for a tracking frame). */
extern "C" int svo_frame_begin(svo_ctx *ctx, const uint8_t *left, size_t ls, const uint8_t *right, size_t rs, int on_device, int prev_slot,
                               svo_track_io *io, int *cur_slot_out)
{
    if (ctx) ctx->err[0] = 0;   // a failure below reports its own message, never a stale one
    if (!ctx) return SVO_ERR_INVALID;
    if (!left || !right || !io || !cur_slot_out || ls < (size_t)ctx->W || rs < (size_t)ctx->W || !slot_ok(ctx, prev_slot)) {
        snprintf(ctx->err, sizeof(ctx->err), "svo_frame_begin: null argument, row stride below the image width, or previous slot %d not live", prev_slot);
        return SVO_ERR_INVALID;
    }
    const double tt0 = g_trace ? now_ms() : 0;
    int rc = validate_and_pack(ctx, io);
    if (rc) return rc;
    CK(cudaSetDevice(ctx->device));
    int src_kind = on_device ? 2 : classify_source(left, right, &ctx->zc_left, &ctx->zc_right);
    if (!on_device && src_kind == 2) {
        snprintf(ctx->err, sizeof(ctx->err), "svo_frame_begin: device pointer passed with on_device == 0");
        return SVO_ERR_INVALID;
    }
    // a frame that fails after its image set was taken gives the set back
    struct SlotGuard {
        svo_ctx *c; int id; bool keep;
        ~SlotGuard() { if (!keep && id >= 0) c->slots[id].refcount = 0; }
    } guard{ctx, -1, false};
    const int n = io->n;
    const bool contiguous = ls == (size_t)ctx->W && rs == (size_t)ctx->W;
    const bool graph_ok = ctx->use_graphs && !ctx->profiling && n > 0;
    (void)contiguous;
    int cur;
    const size_t img = (size_t)ctx->W * ctx->H;
    if (!graph_ok) {
        CK(cudaEventRecord(ctx->ev0, ctx->stream));
        if ((rc = upload_common(ctx, left, ls, right, rs, src_kind, &cur))) return rc;
        guard.id = cur;
        *cur_slot_out = cur;
        const bool prof = ctx->profiling;
        int launches = 0;
        if ((rc = enqueue_track(ctx, prev_slot, cur, n, n, prof, &launches))) return rc;
        CK(cudaEventRecord(ctx->ev1, ctx->stream));
        if (prof) CK(cudaEventRecord(ctx->sev[8], ctx->stream));
        ctx->last_launches = launches;
        ctx->launch_total += launches;
        ctx->track_pending = true;
        guard.keep = true;
        return SVO_OK;
    }
    if ((rc = alloc_slot(ctx, &cur))) return rc;
    guard.id = cur;
    *cur_slot_out = cur;
    int stage_idx = 0;
    if (src_kind == 0) {
        stage_idx = ctx->stage_idx;
        uint8_t *stage = stage_images(ctx, left, ls, right, rs, stage_idx);
        ctx->zc_left = ctx->d_stage[stage_idx];
        ctx->zc_right = ctx->zc_left ? ctx->zc_left + img : nullptr;
        ctx->stage_idx ^= 1;
        left = stage; right = stage + img; ls = rs = (size_t)ctx->W;
    }
    // Slim form (one driver call per frame): the SMs can fetch the pair themselves and the keypoint block is host-mapped, so the
    // ingest kernel is the first node of the replayed graph; it reads this frame's sources from the frame header.
    const uint8_t *vs_left = src_kind == 2 ? left : ctx->zc_left, *vs_right = src_kind == 2 ? right : ctx->zc_right;
    const bool slim = ctx->use_slim && ctx->use_ingest && ctx->d_hio && vs_left && vs_right && !(ctx->W & 15) &&
                      !((reinterpret_cast<uintptr_t>(vs_left) | reinterpret_cast<uintptr_t>(vs_right) | ls | rs) & 15);
    const double tt1 = g_trace ? now_ms() : 0;
    if (slim) {
        *reinterpret_cast<unsigned *>(ctx->h_io + ctx->lay.hdr_seq) = ++ctx->frame_seq;
        // the ingest stays a launch of its own: as a graph node it would have to fetch this frame's sources from the frame
        // header first — one more dependent PCIe round trip in front of every frame (measured: +0.25 ms of latency under load)
        IngestArgs ia;
        ia.src[0] = vs_left; ia.src[1] = vs_right; ia.spitch[0] = ls; ia.spitch[1] = rs;
        ia.dst[0] = ctx->d_rect_packed[0] ? ctx->d_raw[0] : ctx->slots[cur].dev.left[0].ptr;
        ia.dst[1] = ctx->d_rect_packed[1] ? ctx->d_raw[1] : ctx->slots[cur].dev.right0.ptr;
        ia.w = ctx->W; ia.h = ctx->H;
        ia.t_start = reinterpret_cast<unsigned long long *>(ctx->d_io + ctx->lay.stamps);
        launch_ingest(ia, ctx->arena->users < 4, ctx->stream);
        ctx->launch_total += 1;
    } else {
        // the two image copies go on the stream directly (their source changes every frame); pyramids + tracking replay
        CK(cudaEventRecord(ctx->ev0, ctx->stream));   // the frame's device time starts with its ingest
        if ((rc = enqueue_upload(ctx, ctx->slots[cur], left, ls, right, rs, src_kind, true))) return rc;
    }
    const double tt2 = g_trace ? now_ms() : 0;
    const int bucket = std::min(ctx->max_kps, (n + 127) / 128 * 128);
    // one executable graph per (previous, current) image-set pair; when the keypoint count moves to another grid-size
    // bucket the sequence is re-captured (cheap) and the executable is UPDATED in place (cudaGraphExecUpdate: same
    // topology, new grid sizes) — no instantiation, hence no device allocation, after the first two tracking frames
    svo_ctx::FrameGraph *g = nullptr;
    for (auto &c : ctx->graphs)
        if (c.exec && c.prev_slot == prev_slot && c.cur_slot == cur && c.slim == slim) { g = &c; break; }
    if (g && g->bucket != bucket) {
        svo_ctx::FrameGraph ng = *g;
        ng.bucket = bucket; ng.exec = nullptr;
        cudaGraph_t graph = nullptr;
        if ((rc = capture_frame_graph(ctx, ng, n, &graph))) return rc;
        cudaGraphExecUpdateResultInfo info;
        cudaError_t ue = cudaGraphExecUpdate(g->exec, graph, &info);
        if (ue == cudaSuccess) {
            g->bucket = bucket; g->launches = ng.launches;
            ctx->graph_updates++;
        } else {
            cudaGetLastError();
            destroy_graph(*g);
            CK(cudaGraphInstantiate(&g->exec, graph, 0));
            g->bucket = bucket; g->launches = ng.launches;
            ctx->graph_captures++;
        }
        cudaGraphDestroy(graph);
        if (g_trace) fprintf(stderr, "[graph] update prev %d cur %d bucket %d (%s)\n", prev_slot, cur, bucket, ue == cudaSuccess ? "in place" : "re-instantiated");
    }
    if (!g) {
        if (ctx->graphs.size() >= 12) {  // evict the least recently used
            size_t v = 0;
            for (size_t k = 1; k < ctx->graphs.size(); k++) if (ctx->graphs[k].last_use < ctx->graphs[v].last_use) v = k;
            destroy_graph(ctx->graphs[v]);
            ctx->graphs.erase(ctx->graphs.begin() + v);
        }
        svo_ctx::FrameGraph ng;
        ng.prev_slot = prev_slot; ng.cur_slot = cur; ng.bucket = bucket; ng.src_kind = src_kind; ng.stage_idx = stage_idx; ng.slim = slim;
        ng.exec = nullptr; ng.left_copy = ng.right_copy = nullptr; ng.last_use = 0; ng.launches = 0;
        cudaGraph_t graph = nullptr;
        if ((rc = capture_frame_graph(ctx, ng, n, &graph))) return rc;
        cudaError_t ie = cudaGraphInstantiate(&ng.exec, graph, 0);
        cudaGraphDestroy(graph);
        CK(ie);
        ctx->graph_captures++;
        if (g_trace) fprintf(stderr, "[graph] capture prev %d cur %d bucket %d (cache %zu)\n", prev_slot, cur, bucket, ctx->graphs.size());
        ctx->graphs.push_back(ng);
        g = &ctx->graphs.back();
    }
    g->last_use = ++ctx->graph_clock;
    const double tt3 = g_trace ? now_ms() : 0;
    CK(cudaGraphLaunch(g->exec, ctx->stream));
    if (!slim) CK(cudaEventRecord(ctx->ev1, ctx->stream));
    ctx->slim_pending = slim;
    if (g_trace && now_ms() - tt0 > 3.0)
        fprintf(stderr, "[slow frame_begin] pack+alloc %.2f upload %.2f graph lookup/capture %.2f launch %.2f ms\n", tt1 - tt0, tt2 - tt1, tt3 - tt2, now_ms() - tt3);
    ctx->graph_launches++;
    ctx->last_launches = g->launches;
    ctx->launch_total += g->launches;
    ctx->track_pending = true;
    guard.keep = true;
    return SVO_OK;
}

extern "C" int svo_set_graphs(svo_ctx *ctx, int on)
{
    if (!ctx) return SVO_ERR_INVALID;
    ctx->use_graphs = on != 0;
    return SVO_OK;
}

extern "C" int svo_graph_stats(svo_ctx *ctx, long long *graph_launches, long long *graph_captures)
{
    if (!ctx) return SVO_ERR_INVALID;
    if (graph_launches) *graph_launches = ctx->graph_launches;
    if (graph_captures) *graph_captures = ctx->graph_captures;
    return SVO_OK;
}

extern "C" int svo_track_frame_end(svo_ctx *ctx, svo_track_io *io)
{
    if (ctx) ctx->err[0] = 0;   // a failure below reports its own message, never a stale one
    if (!ctx || !io) return SVO_ERR_INVALID;
    if (!ctx->track_pending) { snprintf(ctx->err, sizeof(ctx->err), "svo_track_frame_end without _begin"); return SVO_ERR_STATE; }
    ctx->track_pending = false;
    const double ts0 = g_trace ? now_ms() : 0;
    bool polled = false;
    static const bool slim_sync = getenv("SVO_SLIM_SYNC") != nullptr;   // developer switch: slim graph, but wait with cudaStreamSynchronize
    if (ctx->slim_pending && slim_sync) ctx->slim_pending = false;
    if (ctx->slim_pending) {
        // the frame's last kernel stores the sequence number behind its results (system-scope fence): poll it, no driver call
        ctx->slim_pending = false;
        volatile unsigned *flag = reinterpret_cast<volatile unsigned *>(ctx->h_io + ctx->lay.done);
        const unsigned want = ctx->frame_seq;
        const auto t_poll = std::chrono::steady_clock::now();
        for (unsigned spins = 0; !polled; spins++) {
            if (*flag == want) { polled = true; break; }
            svo_cpu_relax();
            if ((spins & 0xfff) == 0xfff && std::chrono::steady_clock::now() - t_poll > std::chrono::seconds(2)) break;   // a fault never raises the flag
        }
        std::atomic_thread_fence(std::memory_order_acquire);
        if (polled) {
            const unsigned long long *rec = reinterpret_cast<const unsigned long long *>(ctx->h_io + ctx->lay.done);
            ctx->last_ms = (float)((double)(rec[3] - rec[2]) * 1e-6);
        }
    }
    if (!polled) {
        CK(cudaSetDevice(ctx->device));
        CK(cudaStreamSynchronize(ctx->stream));
        if (g_trace && now_ms() - ts0 > 20.0) fprintf(stderr, "[slow frame_end] stream sync %.2f ms\n", now_ms() - ts0);
        if (cudaEventElapsedTime(&ctx->last_ms, ctx->ev0, ctx->ev1) != cudaSuccess) { cudaGetLastError(); ctx->last_ms = 0.f; }
    }
    if (ctx->profiling) {
        // sev: 0 upload start, 1 pyramids done, 2 inputs on device, 3 align, 4 klt, 5 refine, 6 ssd, 7 filter, 8 d2h
        float t;
        if (cudaEventElapsedTime(&t, ctx->sev[0], ctx->sev[1]) == cudaSuccess) ctx->stage_ms[0] = t;
        for (int k = 1; k <= 6; k++) { CK(cudaEventElapsedTime(&t, ctx->sev[k + 1], ctx->sev[k + 2])); ctx->stage_ms[k] = t; }
        if (cudaEventElapsedTime(&t, ctx->sev[0], ctx->sev[8]) == cudaSuccess) ctx->stage_ms[7] = t;
    }
    const IoLayout &L = ctx->lay;
    const uint8_t *h = ctx->h_io;
    const int n = io->n;
    memcpy(io->kps3d, h + L.kps3d, (size_t)n * 12);
    memcpy(io->flags, h + L.flags, (size_t)n);
    memcpy(io->inlier_count, h + L.inlier, (size_t)n * 4);
    memcpy(io->outlier_count, h + L.outlier, (size_t)n * 4);
    memcpy(io->kf_state, h + L.kf_state, (size_t)n * 8);
    memcpy(io->kps2d, h + L.kps2d_out, (size_t)n * 8);
    memcpy(io->pose_aligned, h + L.pose_aligned, 24);
    memcpy(io->pose_refined, h + L.pose_refined, 24);
    io->align_cost = io_ptr<float>(const_cast<uint8_t *>(h), L.costs)[0];
    io->refine_cost = io_ptr<float>(const_cast<uint8_t *>(h), L.costs)[1];
    memcpy(io->align_evals, h + L.evals, 64);
    memcpy(io->refine_evals, h + L.evals + 64, 8);
    if (io->klt_pts) memcpy(io->klt_pts, h + L.klt_pts, (size_t)n * 8);
    if (io->klt_err) memcpy(io->klt_err, h + L.klt_err, (size_t)n * 4);
    if (io->klt_status) memcpy(io->klt_status, h + L.klt_status, (size_t)n);
    if (io->disparity) memcpy(io->disparity, h + L.disparity, (size_t)n * 4);
    if (io->kps2d_refine_in) memcpy(io->kps2d_refine_in, h + L.kps2d_ref_in, (size_t)n * 8);
    if (io->klt_iters) memcpy(io->klt_iters, h + L.klt_iters, (size_t)n * 4);
    return SVO_OK;
}

extern "C" int svo_track_frame(svo_ctx *ctx, int prev_slot, int cur_slot, svo_track_io *io)
{
    int rc = svo_track_frame_begin(ctx, prev_slot, cur_slot, io);
    if (rc) return rc;
    return svo_track_frame_end(ctx, io);
}

extern "C" int svo_last_track_timing(svo_ctx *ctx, float *gpu_ms, int *launches)
{
    if (!ctx) return SVO_ERR_INVALID;
    if (gpu_ms) *gpu_ms = ctx->last_ms;
    if (launches) *launches = ctx->last_launches;
    return SVO_OK;
}
