// Image pyramids of one stereo frame (stereo_slam.cpp:135-139).
//
// pyr_halfsample_kernel : createImgPyramid/halfSample (stereo_slam.cpp:93-121).  ONE launch produces every level:
//   a CTA stages a 64x64 tile of level 0 in shared memory and reduces it level by level (truncating
//   integer mean of 2x2, exactly the reference's arithmetic), so level 0 is read from HBM once and each
//   coarser level is written once.  Pure streaming kernel -> HBM roofline (algorithmic bytes: SURVEY §8d).
// lk_pad_level0_kernel / lk_pyrdown_kernel : cv::buildOpticalFlowPyramid(left, win, 2) image levels:
//   level 0 copy, levels 1..2 = cv::pyrDown ([1 4 6 4 1]^2, (sum+128)>>8, even samples,
//   size (w+1)/2 x (h+1)/2).  Every level is stored with a SVO_LK_PAD-pixel BORDER_REFLECT_101 frame, which
//   is both pyrDown's border rule and the padding calcOpticalFlowPyrLK expects around its windows.
// lk_scharr_kernel : the interleaved int16 Scharr derivative levels OpenCV stores next to them (zero border), built only for
//   image sets that become keyframes — the only ones optical flow ever uses as its reference (optical_flow.cpp:41-44).
#include <cstdlib>
#include "kernels.cuh"

#define TILE 64

__global__ void __launch_bounds__(256) pyr_halfsample_kernel(ImageSetDev s)
{
    __shared__ __align__(16) uint8_t t0[TILE * TILE];           // level 0 tile
    __shared__ __align__(16) uint8_t tl[TILE / 2 * TILE / 2 + TILE / 4 * TILE / 4 + 256];  // coarser tiles, ping-pong free (sizes shrink)

    const int tx0 = blockIdx.x * TILE, ty0 = blockIdx.y * TILE;
    const LevelDesc L0 = s.left[0];
    const int tid = threadIdx.x;

    // ---- stage level 0 (16-byte vector loads when the row is aligned, zero fill outside the image)
    const bool vec_ok = ((L0.pitch & 15) == 0) && ((reinterpret_cast<uintptr_t>(L0.ptr) & 15) == 0);
    for (int i = tid; i < TILE * TILE / 16; i += blockDim.x) {
        int r = i / (TILE / 16), c16 = (i % (TILE / 16)) * 16;
        int gy = ty0 + r, gx = tx0 + c16;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (gy < L0.h) {
            if (vec_ok && gx + 16 <= L0.w) {
                v = __ldg(reinterpret_cast<const uint4 *>(L0.ptr + (size_t)gy * L0.pitch + gx));
            } else {
                uint8_t b[16];
#pragma unroll
                for (int k = 0; k < 16; k++) b[k] = (gx + k < L0.w) ? __ldg(L0.ptr + (size_t)gy * L0.pitch + gx + k) : 0;
                v = *reinterpret_cast<uint4 *>(b);
            }
        }
        *reinterpret_cast<uint4 *>(&t0[r * TILE + c16]) = v;
    }
    __syncthreads();

    // ---- reduce level by level inside the tile
    const uint8_t *src = t0;
    int sw = TILE;  // source tile width/height
    uint8_t *dst = tl;
    for (int l = 1; l < s.n_levels; l++) {
        const int dw = sw >> 1;
        if (dw == 0) break;
        const LevelDesc D = s.left[l];
        const int gx0 = tx0 >> l, gy0 = ty0 >> l;
        for (int i = tid; i < dw * dw; i += blockDim.x) {
            int r = i / dw, c = i % dw;
            const uint8_t *p = src + (2 * r) * sw + 2 * c;
            uint8_t v = (uint8_t)((p[0] + p[1] + p[sw] + p[sw + 1]) / 4);
            dst[r * dw + c] = v;
            int gx = gx0 + c, gy = gy0 + r;
            if (gx < D.w && gy < D.h) D.ptr[(size_t)gy * D.pitch + gx] = v;
        }
        __syncthreads();
        src = dst;
        dst = dst + dw * dw;
        sw = dw;
    }
}

void launch_pyr_halfsample(const ImageSetDev &s, cudaStream_t st)
{
    if (s.n_levels <= 1) return;
    dim3 grid((s.left[0].w + TILE - 1) / TILE, (s.left[0].h + TILE - 1) / TILE);
    pyr_halfsample_kernel<<<grid, 256, 0, st>>>(s);
}

// level 0 of the LK pyramid: copy of `left` with a REFLECT_101 frame of SVO_LK_PAD px
__global__ void __launch_bounds__(256) lk_pad_level0_kernel(LevelDesc src, LevelDesc dst)
{
    const int PW = dst.w + 2 * SVO_LK_PAD, PH = dst.h + 2 * SVO_LK_PAD;
    int xp = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    int yp = blockIdx.y;
    if (xp >= PW || yp >= PH) return;
    int y = dev_reflect101(yp - SVO_LK_PAD, dst.h);
    const uint8_t *srow = src.ptr + (size_t)y * src.pitch;
    uint8_t *drow = dst.ptr + ((ptrdiff_t)(yp - SVO_LK_PAD)) * dst.pitch - SVO_LK_PAD;
    uint8_t b[4];
#pragma unroll
    for (int k = 0; k < 4; k++) b[k] = srow[dev_reflect101(xp + k - SVO_LK_PAD, dst.w)];
    if (xp + 4 <= PW) *reinterpret_cast<uchar4 *>(drow + xp) = make_uchar4(b[0], b[1], b[2], b[3]);
    else
        for (int k = 0; k < 4 && xp + k < PW; k++) drow[xp + k] = b[k];
}

// cv::pyrDown of a padded level into the next padded level (interior AND border: border pixels are the
// REFLECT_101 image of interior ones, recomputed instead of copied)
__global__ void __launch_bounds__(256) lk_pyrdown_kernel(LevelDesc src, LevelDesc dst)
{
    const int PW = dst.w + 2 * SVO_LK_PAD, PH = dst.h + 2 * SVO_LK_PAD;
    int xp = blockIdx.x * blockDim.x + threadIdx.x;
    int yp = blockIdx.y;
    if (xp >= PW || yp >= PH) return;
    int x = dev_reflect101(xp - SVO_LK_PAD, dst.w), y = dev_reflect101(yp - SVO_LK_PAD, dst.h);
    // taps 2x-2..2x+2 reach at most 2 px outside the source level: inside its reflected frame
    const uint8_t *p = src.ptr + (ptrdiff_t)(2 * y - 2) * src.pitch + (2 * x - 2);
    int acc = 0;
    const int wgt[5] = {1, 4, 6, 4, 1};
#pragma unroll
    for (int r = 0; r < 5; r++) {
        const uint8_t *q = p + (ptrdiff_t)r * src.pitch;
        int row = q[0] + q[4] + 4 * (q[1] + q[3]) + 6 * q[2];
        acc += wgt[r] * row;
    }
    dst.ptr[(ptrdiff_t)(yp - SVO_LK_PAD) * dst.pitch + (xp - SVO_LK_PAD)] = (uint8_t)((acc + 128) >> 8);
}

// 16 bytes per thread.  Blocks with blockIdx.y < padded height copy the interior chunks of one row (one aligned 16-byte load
// each: the 32-pixel frame is two chunks wide, so chunk c of the padded row is source bytes 16(c-2) ..); the blocks after
// them assemble the four frame chunks of every row byte by byte through the reflection — kept in warps of their own so
// that the copy warps never run the slow path.  Needs width, source pitch and both base addresses to be multiples of 16
// (the per-byte kernel covers the rest).
__device__ __forceinline__ void lk_pad_rows_job(const LevelDesc &src, const LevelDesc &dst, int job, int nthreads, int tid);
__global__ void __launch_bounds__(64) lk_pad_level0_vec_kernel(LevelDesc src, LevelDesc dst)
{
    lk_pad_rows_job(src, dst, (int)blockIdx.y, 64, threadIdx.x);
}

// ---- device bodies shared by the stand-alone kernels above and the fused side-branch kernel below
__device__ __forceinline__ void lk_pad_rows_job(const LevelDesc &src, const LevelDesc &dst, int job, int nthreads, int tid)
{
    const int PH = dst.h + 2 * SVO_LK_PAD;
    const int chunks = (dst.w + 2 * SVO_LK_PAD) >> 4;
    if (job < PH) {
        const int yp = job;
        const int y = dev_reflect101(yp - SVO_LK_PAD, dst.h);
        const uint8_t *srow = src.ptr + (size_t)y * src.pitch;
        uint8_t *drow = dst.ptr + ((ptrdiff_t)(yp - SVO_LK_PAD)) * dst.pitch - SVO_LK_PAD;
        for (int c = 2 + tid; c < chunks - 2; c += nthreads)
            *reinterpret_cast<uint4 *>(drow + c * 16) = __ldg(reinterpret_cast<const uint4 *>(srow + (c - 2) * 16));
        return;
    }
    const int item = (job - PH) * nthreads + tid;   // (row, one of the four frame chunks)
    const int yp = item >> 2, which = item & 3;
    if (yp >= PH) return;
    const int c = which < 2 ? which : chunks - 4 + which;
    const int y = dev_reflect101(yp - SVO_LK_PAD, dst.h);
    const uint8_t *srow = src.ptr + (size_t)y * src.pitch;
    uint8_t *drow = dst.ptr + ((ptrdiff_t)(yp - SVO_LK_PAD)) * dst.pitch - SVO_LK_PAD;
    const int x0 = c * 16 - SVO_LK_PAD;
    uint8_t b[16];
#pragma unroll
    for (int k = 0; k < 16; k++) b[k] = srow[dev_reflect101(x0 + k, dst.w)];
    *reinterpret_cast<uint4 *>(drow + c * 16) = *reinterpret_cast<uint4 *>(b);
}

// cv::pyrDown of the UNPADDED source image into padded LK level 1: the same taps as lk_pyrdown_kernel on padded level 0, whose
// frame is the REFLECT_101 image of this source — read through the reflection here, so level 1 does not wait for level 0
__device__ __forceinline__ void lk_pyrdown_from_source(const LevelDesc &src, const LevelDesc &dst, int xp, int yp)
{
    const int PW = dst.w + 2 * SVO_LK_PAD, PH = dst.h + 2 * SVO_LK_PAD;
    if (xp >= PW || yp >= PH) return;
    const int x = dev_reflect101(xp - SVO_LK_PAD, dst.w), y = dev_reflect101(yp - SVO_LK_PAD, dst.h);
    int acc = 0;
    const int wgt[5] = {1, 4, 6, 4, 1};
    if (2 * x - 2 >= 0 && 2 * x + 2 < src.w && 2 * y - 2 >= 0 && 2 * y + 2 < src.h) {   // all 25 taps inside the image: no reflection
        const uint8_t *p = src.ptr + (size_t)(2 * y - 2) * src.pitch + (2 * x - 2);
#pragma unroll
        for (int r = 0; r < 5; r++) {
            const uint8_t *q = p + (size_t)r * src.pitch;
            int row = q[0] + q[4] + 4 * (q[1] + q[3]) + 6 * q[2];
            acc += wgt[r] * row;
        }
    } else {
        int cx[5];
#pragma unroll
        for (int k = 0; k < 5; k++) cx[k] = dev_reflect101(2 * x - 2 + k, src.w);
#pragma unroll
        for (int r = 0; r < 5; r++) {
            const uint8_t *q = src.ptr + (size_t)dev_reflect101(2 * y - 2 + r, src.h) * src.pitch;
            int row = q[cx[0]] + q[cx[4]] + 4 * (q[cx[1]] + q[cx[3]]) + 6 * q[cx[2]];
            acc += wgt[r] * row;
        }
    }
    dst.ptr[(ptrdiff_t)(yp - SVO_LK_PAD) * dst.pitch + (xp - SVO_LK_PAD)] = (uint8_t)((acc + 128) >> 8);
}

// The side branch of a tracking frame in ONE launch (a kernel launch costs ~0.27 us of GPU front end at saturation, whatever
// its size): CTAs [0, n_pad) pad level 0, [n_pad, n_pad + n_down) make padded level 1 straight from the source image, and the
// last CTA (if any) imports the keypoint block.  Nothing in here depends on anything else in here.
__global__ void __launch_bounds__(256) lk_side_kernel(LevelDesc l0, LevelDesc lk0, LevelDesc lk1, int n_pad, int n_down, int down_bx, IoCopyArgs io,
                                                      int with_import)
{
    const int b = blockIdx.x;
    if (b < n_pad) { lk_pad_rows_job(l0, lk0, b, 256, threadIdx.x); return; }
    if (b < n_pad + n_down) {
        const int j = b - n_pad;
        lk_pyrdown_from_source(l0, lk1, (j % down_bx) * 256 + threadIdx.x, j / down_bx);
        return;
    }
    if (with_import) io_copy_block(io);
}

static bool lk_pad_vec_ok(const ImageSetDev &s)
{
    const LevelDesc &d = s.lk[0];
    const LevelDesc &l0 = s.left[0];
    const uint8_t *dorigin = d.ptr - (ptrdiff_t)SVO_LK_PAD * d.pitch - SVO_LK_PAD;
    return (d.w % 16 == 0) && (l0.pitch % 16 == 0) && (d.pitch % 16 == 0) && ((reinterpret_cast<uintptr_t>(l0.ptr) & 15) == 0) &&
           ((reinterpret_cast<uintptr_t>(dorigin) & 15) == 0);
}

bool lk_side_fusable(const ImageSetDev &s)
{
    static const bool no_fuse = getenv("SVO_NO_SIDE_FUSION") != nullptr;
    return lk_pad_vec_ok(s) && SVO_LK_LEVELS >= 2 && !no_fuse;
}

// io != null: the keypoint import rides along if the fused kernel is used; returns whether it did
bool launch_lk_pyramid(const ImageSetDev &s, cudaStream_t st, const IoCopyArgs *io)
{
    bool imported = false;
    int first = 1;
    if (lk_side_fusable(s)) {
        const LevelDesc &d0 = s.lk[0], &d1 = s.lk[1];
        const int PH0 = d0.h + 2 * SVO_LK_PAD;
        const int n_pad = PH0 + (PH0 * 4 + 255) / 256;
        const int PW1 = d1.w + 2 * SVO_LK_PAD, PH1 = d1.h + 2 * SVO_LK_PAD;
        const int down_bx = (PW1 + 255) / 256, n_down = down_bx * PH1;
        IoCopyArgs none;
        memset(&none, 0, sizeof(none));
        lk_side_kernel<<<n_pad + n_down + (io ? 1 : 0), 256, 0, st>>>(s.left[0], d0, d1, n_pad, n_down, down_bx, io ? *io : none, io ? 1 : 0);
        imported = io != nullptr;
        first = 2;
    } else {
        const LevelDesc &d = s.lk[0];
        const int PW = d.w + 2 * SVO_LK_PAD, PH = d.h + 2 * SVO_LK_PAD;
        const LevelDesc &l0 = s.left[0];
        if (lk_pad_vec_ok(s)) {
            dim3 grid(1, PH + (PH * 4 + 63) / 64);
            lk_pad_level0_vec_kernel<<<grid, 64, 0, st>>>(l0, d);
        } else {
            dim3 grid((PW / 4 + 1 + 255) / 256, PH);
            lk_pad_level0_kernel<<<grid, 256, 0, st>>>(l0, d);
        }
    }
    for (int l = first; l < SVO_LK_LEVELS; l++) {
        const LevelDesc &d = s.lk[l];
        int PW = d.w + 2 * SVO_LK_PAD, PH = d.h + 2 * SVO_LK_PAD;
        dim3 grid((PW + 255) / 256, PH);
        lk_pyrdown_kernel<<<grid, 256, 0, st>>>(s.lk[l - 1], d);
    }
    return imported;
}

// Frame ingest: both images of a stereo pair, from ANY device-visible memory — page-locked host memory read by the SMs
// over PCIe (zero-copy; a 361 KB cudaMemcpyAsync costs ~25 us of copy-engine time, of which ~18 us is fixed overhead,
// so two copies per frame cap one engine at ~20 k frames/s) or device memory — into level 0 of the image set (or into
// the raw buffers of the rectifier).  One 16-byte load per thread, all 722 KB of a C3 pair in flight at once.
#define INGEST_CH 24      // cap on 16-byte chunks in flight per thread (128 threads x 24 x 16 B = 48 KB of shared memory per CTA)
// The bytes in flight over PCIe are parked in SHARED memory by cp.async (LDGSTS), not in registers, and there are FEW of
// them: 4 persistent CTAs per image x 128 threads x 3 chunks = 49 KB per stereo pair.  Measured with 32 sequences per GPU
// (end-to-end frames/s): one 16-byte load per thread, 722 KB in flight 37.2 k; 4 CTAs of 8 register-held loads 42.4 k;
// this kernel 41.4-42.9 k for 25-100 KB in flight, 38.9 k at 290 KB.  More outstanding system-memory reads than the link
// needs (bandwidth x ~2 us) only queue in front of the other kernels' memory traffic.  A shallow queue costs a sequence
// that runs alone ~15 us per frame, so contexts that have the GPU (almost) to themselves use a deep one (deep_queue: every
// chunk of the pair in flight at once).  Every thread reads back only the chunks it fetched itself, so there is no
// barrier.  SVO_INGEST_CTAS / SVO_INGEST_DEPTH override the two numbers of the shallow configuration.
__global__ void __launch_bounds__(128) ingest_kernel(IngestArgs a, int depth)
{
    extern __shared__ __align__(16) uint4 ingest_buf[];
    const int z = blockIdx.y;
    const uint8_t *__restrict__ src = z ? a.src[1] : a.src[0];
    uint8_t *__restrict__ dst = z ? a.dst[1] : a.dst[0];
    const size_t sp = z ? a.spitch[1] : a.spitch[0];
    if (a.t_start && blockIdx.x == 0 && z == 0 && threadIdx.x == 0) {   // device-side time stamp of the frame's first kernel
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        *a.t_start = t;
    }
    const int c16 = a.w >> 4;                        // 16-byte chunks per row
    const int total = c16 * a.h;
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(ingest_buf) + threadIdx.x * 16;
    for (int base = blockIdx.x * (128 * depth) + threadIdx.x; base < total; base += gridDim.x * 128 * depth) {
        // chunk k of this thread = base + k * 128: every instruction of a warp covers 512 contiguous bytes
#pragma unroll 4
        for (int k = 0; k < depth; k++) {
            const int idx = base + k * 128;
            if (idx < total) {
                const int row = idx / c16, col = (idx - row * c16) << 4;
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sbase + k * 128 * 16), "l"(src + (size_t)row * sp + col) : "memory");
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
#pragma unroll 4
        for (int k = 0; k < depth; k++) {
            const int idx = base + k * 128;
            if (idx < total) {
                const int row = idx / c16, col = (idx - row * c16) << 4;
                *reinterpret_cast<uint4 *>(dst + (size_t)row * a.w + col) = ingest_buf[k * 128 + threadIdx.x];
            }
        }
    }
}

bool ingest_supported(const IngestArgs &a)
{
    if (a.w & 15) return false;
    for (int z = 0; z < 2; z++)
        if ((reinterpret_cast<uintptr_t>(a.src[z]) & 15) || (a.spitch[z] & 15) || (reinterpret_cast<uintptr_t>(a.dst[z]) & 15)) return false;
    return true;
}

void launch_ingest(const IngestArgs &a, bool deep_queue, cudaStream_t st)
{
    const int total = (a.w >> 4) * a.h;
    static const int max_ctas = getenv("SVO_INGEST_CTAS") ? atoi(getenv("SVO_INGEST_CTAS")) : 4;   // per image
    static const int depth_env = getenv("SVO_INGEST_DEPTH") ? atoi(getenv("SVO_INGEST_DEPTH")) : 3;
    int depth = depth_env < 1 ? 1 : (depth_env > INGEST_CH ? INGEST_CH : depth_env);
    int ctas = max_ctas < 1 ? 1 : max_ctas;
    if (deep_queue) { depth = 8; ctas = 1 << 20; }   // a sequence running alone: everything in flight at once (lowest latency)
    const int need = (total + 128 * depth - 1) / (128 * depth);
    dim3 grid(need < ctas ? need : ctas, 2);
    ingest_kernel<<<grid, 128, 128 * depth * 16, st>>>(a, depth);
}

// developer probe: zero-copy read bandwidth of the SMs (same access pattern as ingest_kernel) over a large buffer
__global__ void __launch_bounds__(256) copy16_kernel(const uint4 *__restrict__ src, uint4 *__restrict__ dst, size_t n16)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) dst[i] = __ldcs(src + i);
}
// A linear copy driven by the TMA unit: cp.async.bulk in (system or device memory -> shared), cp.async.bulk out (shared -> global),
// ONE thread, a 4-stage ring with two loads in flight.  Chunks first, first + stride, ... of `chunk` bytes.
__device__ __forceinline__ void bulk_copy_thread(const uint8_t *src, uint8_t *dst, size_t bytes, int chunk, long long first, long long stride,
                                                 uint8_t *ring, unsigned long long *bar)
{
    for (int s = 0; s < 4; s++) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&bar[s])), "r"(1) : "memory");
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const long long nchunks = (long long)((bytes + chunk - 1) / chunk);
    const long long mine = first < nchunks ? (nchunks - first + stride - 1) / stride : 0;
    auto issue = [&](long long k) {
        const size_t off = (size_t)(first + k * stride) * chunk;
        const uint32_t n = (uint32_t)((bytes - off < (size_t)chunk) ? (bytes - off) : (size_t)chunk);
        const int s = (int)(k & 3);
        const uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar[s]), d = (uint32_t)__cvta_generic_to_shared(ring + (size_t)s * chunk);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(n) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(d), "l"(src + off), "r"(n), "r"(b) : "memory");
    };
    for (long long k = 0; k < mine && k < 2; k++) issue(k);
    for (long long k = 0; k < mine; k++) {
        const int s = (int)(k & 3);
        const uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar[s]), parity = (uint32_t)((k >> 2) & 1);
        uint32_t ok = 0;
        while (!ok)
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(b), "r"(parity) : "memory");
        const size_t off = (size_t)(first + k * stride) * chunk;
        const uint32_t n = (uint32_t)((bytes - off < (size_t)chunk) ? (bytes - off) : (size_t)chunk);
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + off), "r"((uint32_t)__cvta_generic_to_shared(ring + (size_t)s * chunk)), "r"(n) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 2;" ::: "memory");   // the store of chunk k - 2 has read its stage: reload it
        if (k + 2 < mine) issue(k + 2);
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// developer probe (SVO_ZC_BULK=<chunk bytes>): does the bulk-copy engine read system memory faster than the SMs' 16-byte loads?
// (No: both reach the link's 51 GB/s; the TMA gets there with 4 single-thread CTAs instead of 32 CTAs of 256 threads.)
__global__ void __launch_bounds__(32) copy_bulk_kernel(const uint8_t *src, uint8_t *dst, size_t bytes, int chunk)
{
    extern __shared__ __align__(128) uint8_t bulk_buf[];
    __shared__ unsigned long long bar[4];
    if (threadIdx.x != 0) return;
    bulk_copy_thread(src, dst, bytes, chunk, blockIdx.x, gridDim.x, bulk_buf, bar);
}

void launch_copy16(const void *src, void *dst, size_t bytes, int ctas, cudaStream_t st)
{
    static const int bulk = getenv("SVO_ZC_BULK") ? atoi(getenv("SVO_ZC_BULK")) : 0;
    if (bulk >= 16 && bulk <= 48 * 1024 && bulk % 16 == 0) {
        static bool once = false;
        if (!once) { cudaFuncSetAttribute(copy_bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); once = true; }
        copy_bulk_kernel<<<ctas, 32, 4 * (size_t)bulk, st>>>((const uint8_t *)src, (uint8_t *)dst, bytes, bulk);
        return;
    }
    copy16_kernel<<<ctas, 256, 0, st>>>((const uint4 *)src, (uint4 *)dst, bytes / 16);
}

// Scharr derivatives of one padded LK level (calcScharrDeriv inside cv::buildOpticalFlowPyramid): (Ix, Iy) as an int16
// pair per pixel, taps through the level's REFLECT_101 frame, only the interior is written (the frame of the derivative
// level stays 0 = BORDER_CONSTANT).  Runs once per KEYFRAME, not per frame: a thread makes 4 pixels (one 16-byte store).
__global__ void __launch_bounds__(128) lk_scharr_kernel(LevelDesc src, LevelDesc dst)
{
    const int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4, y = blockIdx.y;
    if (x0 >= src.w) return;
    const uint8_t *r0 = src.ptr + (ptrdiff_t)(y - 1) * src.pitch + x0 - 1, *r1 = r0 + src.pitch, *r2 = r1 + src.pitch;
    int t[6], m[6], b[6];
#pragma unroll
    for (int k = 0; k < 6; k++) { t[k] = r0[k]; m[k] = r1[k]; b[k] = r2[k]; }   // up to 2 px past the row end: inside the frame
    uint32_t out[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int gx = 3 * (t[k + 2] + b[k + 2]) + 10 * m[k + 2] - (3 * (t[k] + b[k]) + 10 * m[k]);
        const int gy = 3 * ((b[k] - t[k]) + (b[k + 2] - t[k + 2])) + 10 * (b[k + 1] - t[k + 1]);
        out[k] = ((uint32_t)gx & 0xffffu) | ((uint32_t)gy << 16);
    }
    uint32_t *d = reinterpret_cast<uint32_t *>(dst.ptr + (size_t)y * dst.pitch) + x0;
    if (x0 + 4 <= src.w) *reinterpret_cast<uint4 *>(d) = make_uint4(out[0], out[1], out[2], out[3]);
    else
        for (int k = 0; k < 4 && x0 + k < src.w; k++) d[k] = out[k];
}

void launch_lk_scharr(const ImageSetDev &s, cudaStream_t st)
{
    for (int l = 0; l < SVO_LK_LEVELS; l++) {
        dim3 grid(((s.lk[l].w + 3) / 4 + 127) / 128, s.lk[l].h);
        lk_scharr_kernel<<<grid, 128, 0, st>>>(s.lk[l], s.lkd[l]);
    }
}

int pyr_launch_count(const ImageSetDev &s) { return (s.n_levels > 1 ? 1 : 0) + SVO_LK_LEVELS - (lk_side_fusable(s) ? 1 : 0); }
