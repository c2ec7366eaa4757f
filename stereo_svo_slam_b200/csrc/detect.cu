// Grid keypoint detection — CornerDetector::detect_keypoints (src/lib/corner_detector.cpp:13-79):
// cv::FastFeatureDetector::create(6) (FAST-9/16, threshold 6, non-max suppression, response = OpenCV's
// cornerScore) -> best response per grid cell (first maximum in raster order); cells without a corner fall
// back to the arg-max of cv::Sobel(image, -1, 1, 0) (8-bit saturated x-derivative, REFLECT_101) scanned
// column-major.  Exactly one keypoint per cell, cells in row-major order.  Integer work: bit-exact.
//
//   fast_score_kernel : one thread per pixel, 16-pixel Bresenham ring from L1/L2, 9-contiguous test on two
//                       16-bit masks, score by the min/max arc recurrence.  Streaming (HBM/L2 bound).
//   cell_select_kernel: one CTA per grid cell; NMS against the 8 neighbours in the score map, block arg-max;
//                       Sobel fallback computed on the fly (never materialised).
#include "kernels.cuh"

__constant__ int c_fast_dx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
__constant__ int c_fast_dy[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};

__device__ __forceinline__ bool has9(unsigned m)
{
    unsigned mm = m | (m << 16);
    unsigned r = mm;
#pragma unroll
    for (int i = 1; i < 9; i++) r &= (mm >> i);
    return (r & 0xFFFFu) != 0;
}

__global__ void __launch_bounds__(256) fast_score_kernel(LevelDesc img, uint8_t *score, int thr)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= img.w || y >= img.h) return;
    int out = 0;
    if (x >= 3 && x < img.w - 3 && y >= 3 && y < img.h - 3) {
        const uint8_t *p = img.ptr + (size_t)y * img.pitch + x;
        const int v = p[0];
        int d[25];
        unsigned br = 0, dk = 0;
#pragma unroll
        for (int k = 0; k < 16; k++) {
            int q = p[c_fast_dy[k] * img.pitch + c_fast_dx[k]];
            d[k] = v - q;
            if (q > v + thr) br |= 1u << k;
            if (q < v - thr) dk |= 1u << k;
        }
        if (has9(br) || has9(dk)) {
#pragma unroll
            for (int k = 16; k < 25; k++) d[k] = d[k - 16];
            int a0 = thr;
#pragma unroll
            for (int k = 0; k < 16; k += 2) {
                int a = min(d[k + 1], d[k + 2]);
                a = min(a, d[k + 3]);
                if (a <= a0) continue;
                a = min(a, d[k + 4]); a = min(a, d[k + 5]); a = min(a, d[k + 6]); a = min(a, d[k + 7]); a = min(a, d[k + 8]);
                a0 = max(a0, min(a, d[k]));
                a0 = max(a0, min(a, d[k + 9]));
            }
            int b0 = -a0;
#pragma unroll
            for (int k = 0; k < 16; k += 2) {
                int b = max(d[k + 1], d[k + 2]);
                b = max(b, d[k + 3]); b = max(b, d[k + 4]); b = max(b, d[k + 5]);
                if (b >= b0) continue;
                b = max(b, d[k + 6]); b = max(b, d[k + 7]); b = max(b, d[k + 8]);
                b0 = min(b0, max(b, d[k]));
                b0 = min(b0, max(b, d[k + 9]));
            }
            out = -b0 - 1;
        }
    }
    score[(size_t)y * img.w + x] = (uint8_t)out;
}

__device__ __forceinline__ int nms_score(const uint8_t *score, int w, int h, int x, int y)
{
    if (x < 3 || y < 3 || x >= w - 3 || y >= h - 3) return 0;
    const uint8_t *c = score + (size_t)y * w + x;
    const int s = c[0];
    if (s == 0) return 0;
    if (s > c[-1] && s > c[1] && s > c[-w - 1] && s > c[-w] && s > c[-w + 1] && s > c[w - 1] && s > c[w] && s > c[w + 1]) return s;
    return 0;
}

__device__ __forceinline__ int sobel_x_u8(const LevelDesc &img, int x, int y)
{
    const int xm = dev_reflect101(x - 1, img.w), xp = dev_reflect101(x + 1, img.w);
    const uint8_t *r0 = img.ptr + (size_t)dev_reflect101(y - 1, img.h) * img.pitch;
    const uint8_t *r1 = img.ptr + (size_t)y * img.pitch;
    const uint8_t *r2 = img.ptr + (size_t)dev_reflect101(y + 1, img.h) * img.pitch;
    int v = (r0[xp] + 2 * r1[xp] + r2[xp]) - (r0[xm] + 2 * r1[xm] + r2[xm]);
    return min(255, max(0, v));
}

__global__ void __launch_bounds__(128) cell_select_kernel(DetectArgs a, const uint8_t *score)
{
    __shared__ unsigned long long red[4];
    const int cx = blockIdx.x, cy = blockIdx.y;
    const int left = cx * a.grid_w, top = cy * a.grid_h;
    const int gw = a.grid_w, gh = a.grid_h;
    const int tid = threadIdx.x;
    const int w = a.img.w, h = a.img.h;
    // ---- FAST: max response, first in raster order.  key = score << 32 | (0xffffffff - raster index)
    unsigned long long best = 0;
    for (int k = tid; k < gw * gh; k += blockDim.x) {
        int ly = k / gw, lx = k - ly * gw;
        int x = left + lx, y = top + ly;
        if (x >= w || y >= h) continue;
        int s = nms_score(score, w, h, x, y);
        if (s > 0) {
            unsigned long long key = ((unsigned long long)s << 32) | (unsigned)(0xffffffffu - (unsigned)k);
            best = key > best ? key : best;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        unsigned long long other = __shfl_down_sync(0xffffffffu, best, o);
        best = other > best ? other : best;
    }
    if ((tid & 31) == 0) red[tid >> 5] = best;
    __syncthreads();
    best = red[0];
    for (int q = 1; q < 4; q++) best = red[q] > best ? red[q] : best;
    __syncthreads();
    const int cell = cy * a.cells_x + cx;
    if (best != 0) {
        if (tid == 0) {
            int k = (int)(0xffffffffu - (unsigned)(best & 0xffffffffu));
            int ly = k / gw, lx = k - ly * gw;
            a.cell_xy[2 * cell] = (float)(left + lx); a.cell_xy[2 * cell + 1] = (float)(top + ly);
            a.cell_score[cell] = (float)(int)(best >> 32);
            a.cell_type[cell] = SVO_KP_FAST;
        }
        return;
    }
    // ---- edgelet fallback: column-major scan (x outer, y inner), strictly greater wins, seed score -1
    best = 0;
    bool any = false;
    for (int k = tid; k < gw * gh; k += blockDim.x) {
        int lx = k / gh, ly = k - lx * gh;  // column-major order index
        int x = left + lx, y = top + ly;
        if (x >= w) continue;
        int s = (y < h) ? sobel_x_u8(a.img, x, y) : 0;
        unsigned long long key = ((unsigned long long)(s + 1) << 32) | (unsigned)(0xffffffffu - (unsigned)k);
        best = key > best ? key : best;
        any = true;
    }
    (void)any;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        unsigned long long other = __shfl_down_sync(0xffffffffu, best, o);
        best = other > best ? other : best;
    }
    if ((tid & 31) == 0) red[tid >> 5] = best;
    __syncthreads();
    if (tid == 0) {
        best = red[0];
        for (int q = 1; q < 4; q++) best = red[q] > best ? red[q] : best;
        int k = (int)(0xffffffffu - (unsigned)(best & 0xffffffffu));
        int lx = k / gh, ly = k - lx * gh;
        a.cell_xy[2 * cell] = (float)(left + lx); a.cell_xy[2 * cell + 1] = (float)(top + ly);
        a.cell_score[cell] = (float)((int)(best >> 32) - 1);
        a.cell_type[cell] = SVO_KP_EDGELET;
    }
}

void launch_detect(const DetectArgs &a, cudaStream_t st)
{
    uint8_t *score = reinterpret_cast<uint8_t *>(a.score_map);
    dim3 g1((a.img.w + 255) / 256, a.img.h);
    fast_score_kernel<<<g1, 256, 0, st>>>(a.img, score, 6);
    if (a.cells_x > 0 && a.cells_y > 0) {
        dim3 g2(a.cells_x, a.cells_y);
        cell_select_kernel<<<g2, 128, 0, st>>>(a, score);
    }
}

// diagnostic: NMS'd score map (0 = suppressed / not a corner), extracted in raster order on the host
__global__ void __launch_bounds__(256) fast_nms_kernel(LevelDesc img, const uint8_t *score, uint8_t *out)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= img.w || y >= img.h) return;
    out[(size_t)y * img.w + x] = (uint8_t)nms_score(score, img.w, img.h, x, y);
}

void launch_fast_list(const LevelDesc &img, int *score_map, int *xys, int max_out, int *count, cudaStream_t st)
{
    // score_map: scratch of >= 2*w*h bytes; first w*h = scores, second w*h = NMS'd scores (downloaded by the caller)
    (void)xys; (void)max_out; (void)count;
    uint8_t *score = reinterpret_cast<uint8_t *>(score_map);
    dim3 g1((img.w + 255) / 256, img.h);
    fast_score_kernel<<<g1, 256, 0, st>>>(img, score, 6);
    fast_nms_kernel<<<g1, 256, 0, st>>>(img, score, score + (size_t)img.w * img.h);
}
