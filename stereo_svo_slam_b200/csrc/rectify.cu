// EuRoC front end in front of the pyramid build (src/app/euroc_input.cpp:48-49, :69-73):
//
// rectify_map_kernel  : cv::initUndistortRectifyMap(K, D(k1,k2,p1,p2,k3), R, P[:3,:3], size, CV_32F) — once per
//   context and camera.  OpenCV walks every row with running sums (_x += ir[0] ...) in double; the value of a
//   pixel therefore depends on the 751 additions before it, so one thread owns one row and repeats exactly that
//   sequence (no FMA: -fmad=false).  One-off cost, not on the per-frame path.
// rectify_pack_kernel : the per-pixel constants cv::remap(INTER_LINEAR) derives from the float maps — integer
//   source position and the 5-bit fractions of cvRound(map * 32) — packed into one 32-bit word per pixel, so the
//   per-frame kernel reads 4 B/px of map instead of 8 and does no float work at all.
// remap_kernel        : cv::remap(src, dst, map1, map2, INTER_LINEAR, BORDER_CONSTANT 0) for both images of a
//   stereo pair in one launch, written straight into level 0 of the image set: weights (32-fx)(32-fy)*32 ...
//   (INTER_REMAP_COEF_SCALE 2^15), (sum + 2^14) >> 15.  Streaming kernel: 4 B map + 1 B source + 1 B output per
//   pixel, coalesced; the source gather is near-sequential because rectification is a smooth warp.
#include "kernels.cuh"

__global__ void __launch_bounds__(64) rectify_map_kernel(RectifyMapArgs a)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.h) return;
    const double *ir = a.ir;
    const double fx = a.K[0], fy = a.K[4], u0 = a.K[2], v0 = a.K[5];
    const double k1 = a.D[0], k2 = a.D[1], p1 = a.D[2], p2 = a.D[3], k3 = a.D[4];
    double _x = i * ir[1] + ir[2], _y = i * ir[4] + ir[5], _w = i * ir[7] + ir[8];
    float *m1 = a.map1 + (size_t)i * a.w, *m2 = a.map2 + (size_t)i * a.w;
    for (int j = 0; j < a.w; j++, _x += ir[0], _y += ir[3], _w += ir[6]) {
        double wi = 1. / _w, x = _x * wi, y = _y * wi;
        double x2 = x * x, y2 = y * y;
        double r2 = x2 + y2, _2xy = 2 * x * y;
        double kr = (1 + ((k3 * r2 + k2) * r2 + k1) * r2);  // rational part k4..k6 = 0: denominator is exactly 1
        double xd = (x * kr + p1 * _2xy + p2 * (r2 + 2 * x2));
        double yd = (y * kr + p1 * (r2 + 2 * y2) + p2 * _2xy);
        m1[j] = (float)(fx * xd + u0);
        m2[j] = (float)(fy * yd + v0);
    }
}

void launch_rectify_map(const RectifyMapArgs &a, cudaStream_t st)
{
    rectify_map_kernel<<<(a.h + 63) / 64, 64, 0, st>>>(a);
}

// packed word: [31:27] fy, [26:22] fx, [21:11] iy + 1, [10:0] ix + 1; positions whose two taps along an axis are
// both outside the source collapse to the sentinel (size + 1), so the per-frame kernel needs two unsigned compares
__global__ void __launch_bounds__(256) rectify_pack_kernel(const float *map1, const float *map2, int n, int sw, int sh, uint32_t *packed)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int sx = __float2int_rn(map1[i] * 32.f), sy = __float2int_rn(map2[i] * 32.f);  // cvRound(map * INTER_TAB_SIZE)
    const int fx = sx & 31, fy = sy & 31;
    int ix = max(-32768, min(32767, sx >> 5)), iy = max(-32768, min(32767, sy >> 5));     // saturate_cast<short>
    const unsigned ixc = (ix < -1 || ix > sw - 1) ? (unsigned)(sw + 1) : (unsigned)(ix + 1);
    const unsigned iyc = (iy < -1 || iy > sh - 1) ? (unsigned)(sh + 1) : (unsigned)(iy + 1);
    packed[i] = ((unsigned)fy << 27) | ((unsigned)fx << 22) | (iyc << 11) | ixc;
}

void launch_rectify_pack(const float *map1, const float *map2, int n, int sw, int sh, uint32_t *packed, cudaStream_t st)
{
    rectify_pack_kernel<<<(n + 255) / 256, 256, 0, st>>>(map1, map2, n, sw, sh, packed);
}

__device__ __forceinline__ uint32_t remap_px(const uint8_t *__restrict__ src, int sw, int sh, int spitch, uint32_t m)
{
    const unsigned x1 = m & 2047u, y1 = (m >> 11) & 2047u;
    const unsigned x0 = x1 - 1u, y0 = y1 - 1u;
    const int fx = (m >> 22) & 31, fy = m >> 27;
    const bool vx0 = x0 < (unsigned)sw, vx1 = x1 < (unsigned)sw, vy0 = y0 < (unsigned)sh, vy1 = y1 < (unsigned)sh;
    const uint8_t *r0 = src + (size_t)y0 * spitch, *r1 = src + (size_t)y1 * spitch;
    const int p00 = (vx0 && vy0) ? __ldg(r0 + x0) : 0, p01 = (vx1 && vy0) ? __ldg(r0 + x1) : 0;
    const int p10 = (vx0 && vy1) ? __ldg(r1 + x0) : 0, p11 = (vx1 && vy1) ? __ldg(r1 + x1) : 0;
    const int w00 = (32 - fx) * (32 - fy) * 32, w01 = fx * (32 - fy) * 32, w10 = (32 - fx) * fy * 32, w11 = fx * fy * 32;
    return (uint32_t)((p00 * w00 + p01 * w01 + p10 * w10 + p11 * w11 + (1 << 14)) >> 15);
}

// blockIdx.z = image of the pair (0 left, 1 right); a thread produces 4 consecutive pixels of one row
__global__ void __launch_bounds__(256) remap_kernel(RemapArgs a)
{
    const int z = blockIdx.z;
    const uint32_t *__restrict__ map = z ? a.map[1] : a.map[0];
    if (!map) return;
    const uint8_t *__restrict__ src = z ? a.src[1] : a.src[0];
    uint8_t *__restrict__ dst = z ? a.dst[1] : a.dst[0];
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    const int x = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (x >= a.w || y >= a.h) return;
    const uint32_t *mrow = map + (size_t)y * a.w + x;
    uint8_t *drow = dst + (size_t)y * a.dpitch + x;
    if (x + 4 <= a.w && ((a.w | a.dpitch) & 3) == 0) {
        const uint4 m = __ldg(reinterpret_cast<const uint4 *>(mrow));
        const uint32_t v0 = remap_px(src, a.sw, a.sh, a.spitch, m.x), v1 = remap_px(src, a.sw, a.sh, a.spitch, m.y);
        const uint32_t v2 = remap_px(src, a.sw, a.sh, a.spitch, m.z), v3 = remap_px(src, a.sw, a.sh, a.spitch, m.w);
        *reinterpret_cast<uint32_t *>(drow) = v0 | (v1 << 8) | (v2 << 16) | (v3 << 24);
    } else {
        for (int k = 0; k < 4 && x + k < a.w; k++) drow[k] = (uint8_t)remap_px(src, a.sw, a.sh, a.spitch, __ldg(mrow + k));
    }
}

void launch_remap(const RemapArgs &a, cudaStream_t st)
{
    dim3 block(64, 4), grid(((a.w + 3) / 4 + 63) / 64, (a.h + 3) / 4, 2);
    remap_kernel<<<grid, block, 0, st>>>(a);
}
