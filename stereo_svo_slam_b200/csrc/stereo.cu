// Stereo SSD patch matching and the per-point depth filter.
//
// stereo_ssd_kernel replaces the template match in DepthCalculator::calculate_depth
// (src/lib/depth_calculator.cpp:200-239, mode 0) and DepthFilter::calculate_disparities
// (src/lib/depth_filter.cpp:259-327, mode 1): cv::matchTemplate(TM_SQDIFF) + cv::minMaxLoc + the
// "mean column of the tied minima right/below the first minimum" rule.  The SSD is computed as exact
// uint32 (the mathematical definition of TM_SQDIFF; OpenCV's float map carries +-40 units of DFT noise,
// SURVEY.md Appendix B.5), so the result is bit-identical to the integer oracle.
//   One CTA per keypoint; template and search ROI staged in shared memory; each thread owns result-map
//   positions; arg-min by a (value, raster index) lexicographic block reduction = "first minimum in raster order".
//   Bound: integer ALU / shared-memory bandwidth (~640 k multiply-adds per keypoint), not HBM.
//
// depth_filter_kernel replaces DepthFilter::outlier_check (:52-128), DepthFilter::update_kps3d (:130-257),
// the flag post-processing in StereoSlam::new_image (src/lib/stereo_slam.cpp:205-226) and the final
// re-projection (:228-229).  One thread per keypoint, reference float operation order.
#include <cstdlib>
#include "kernels.cuh"

#define SSD_THREADS 256

struct SsdGeom { int x11, x12, y11, y12, x21, x22, y21, y22; };

// window clipping rules of depth_calculator.cpp:205-214 / depth_filter.cpp:281-301; returns false => disparity -1
__device__ __forceinline__ bool ssd_geometry(float kx, float ky, int W, int H, int win, int sx, int sy, int mode, SsdGeom &g)
{
    const int wb = win / 2, wa = (win + 1) / 2;
    const int x = (int)kx, y = (int)ky;
    g.x11 = max(0, x - wb); g.x12 = min(W - 1, x + wa);
    g.y11 = max(0, y - wb); g.y12 = min(H, y + wa);
    if (mode == 1 && (g.x12 <= 0 || g.y12 <= 0 || g.x11 >= (W - 1) || g.y11 >= (H - 1))) return false;
    g.x21 = max(0, x - wb); g.x22 = min(W - 1, x + wa + sx);
    g.y21 = max(0, y - wb - sy); g.y22 = min(H - 1, y + wa + sy);
    if (mode == 1 && (g.x22 <= 0 || g.y22 <= 0 || g.x21 >= (W - 1) || g.y21 >= (H - 1))) return false;
    const int tw = g.x12 - g.x11, th = g.y12 - g.y11, rw = g.x22 - g.x21, rh = g.y22 - g.y21;
    if (tw <= 0 || th <= 0 || rw < tw || rh < th) return false;  // cv::matchTemplate would throw
    return true;
}

// tpl / roi rows are stored as 32-bit words (row pitch in words); bytes beyond the window are zero.
// SSD(j,k) = sum a^2 - 2 sum a*b + sum b^2 with the two dot products on the 4-way byte dot-product unit
// (IDP4A); unaligned ROI words are assembled from two aligned shared-memory words with a funnel shift.
__global__ void __launch_bounds__(SSD_THREADS) stereo_ssd_kernel(SsdArgs a, int tpw, int rpw, int roi_rows, int map_cap)
{
    extern __shared__ __align__(16) uint32_t ssd_smem32[];
    const int win = a.cam.win_depth;
    uint32_t *tpl = ssd_smem32;                 // win rows x tpw words
    uint32_t *roi = tpl + win * tpw;            // roi_rows x rpw words
    uint32_t *map = roi + roi_rows * rpw;       // map_cap entries
    __shared__ unsigned long long red[SSD_THREADS / 32];
    __shared__ unsigned long long best_s;
    __shared__ int cnt_s[SSD_THREADS / 32], sum_s[SSD_THREADS / 32];
    __shared__ unsigned saa_s[SSD_THREADS / 32];

    const int i = blockIdx.x;
    const int n = min(*a.n_ptr, a.max_kps);
    if (i >= n) return;
    const int tid = threadIdx.x;
    const LevelDesc L = a.left0, R = a.right0;
    SsdGeom g;
    if (!ssd_geometry(a.kps2d[2 * i], a.kps2d[2 * i + 1], L.w, L.h, win, a.cam.search_x, a.cam.search_y, a.mode, g)) {
        if (tid == 0) a.disparity[i] = -1.f;
        return;
    }
    const int tw = g.x12 - g.x11, th = g.y12 - g.y11, rw = g.x22 - g.x21, rh = g.y22 - g.y21;
    const int mw = rw - tw + 1, mh = rh - th + 1;
    const int twords = (tw + 3) >> 2;
    const uint32_t last_mask = (tw & 3) ? ((1u << ((tw & 3) * 8)) - 1u) : 0xffffffffu;

    // ---- stage (zero padded) and accumulate sum a^2
    uint8_t *tpl8 = reinterpret_cast<uint8_t *>(tpl), *roi8 = reinterpret_cast<uint8_t *>(roi);
    for (int k = tid; k < th * tpw; k += SSD_THREADS) tpl[k] = 0;
    for (int k = tid; k < (rh + 1) * rpw && k < roi_rows * rpw; k += SSD_THREADS) roi[k] = 0;
    __syncthreads();
    unsigned saa = 0;
    for (int k = tid; k < th * tw; k += SSD_THREADS) {
        int r = k / tw, c = k - r * tw;
        unsigned v = L.ptr[(size_t)(g.y11 + r) * L.pitch + g.x11 + c];
        tpl8[r * tpw * 4 + c] = (uint8_t)v;
        saa += v * v;
    }
    for (int k = tid; k < rh * rw; k += SSD_THREADS) {
        int r = k / rw, c = k - r * rw;
        roi8[r * rpw * 4 + c] = R.ptr[(size_t)(g.y21 + r) * R.pitch + g.x21 + c];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) saa += __shfl_down_sync(0xffffffffu, saa, o);
    if ((tid & 31) == 0) saa_s[tid >> 5] = saa;
    __syncthreads();
    saa = 0;
#pragma unroll
    for (int q = 0; q < SSD_THREADS / 32; q++) saa += saa_s[q];

    // ---- SSD map + running (value, index) minimum per thread
    unsigned long long best = ~0ull;
    for (int p = tid; p < mw * mh; p += SSD_THREADS) {
        const int k = p / mw, j = p - k * mw;
        const int jw = j >> 2, sh = (j & 3) * 8;
        unsigned sab = 0, sbb = 0;
        for (int y = 0; y < th; y++) {
            const uint32_t *rr = roi + (k + y) * rpw + jw;
            const uint32_t *tt = tpl + y * tpw;
            uint32_t w0 = rr[0];
            for (int x = 0; x < twords; x++) {
                const uint32_t w1 = rr[x + 1];
                uint32_t b = __funnelshift_r(w0, w1, sh);
                if (x == twords - 1) b &= last_mask;
                const uint32_t av = tt[x];
                sab = __dp4a(av, b, sab);
                sbb = __dp4a(b, b, sbb);
                w0 = w1;
            }
        }
        const uint32_t sv = saa + sbb - 2u * sab;
        map[p] = sv;
        unsigned long long key = ((unsigned long long)sv << 32) | (unsigned)p;
        best = key < best ? key : best;
    }
    // block arg-min (smallest value, then smallest raster index == cv::minMaxLoc's first minimum)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        unsigned long long other = __shfl_down_sync(0xffffffffu, best, o);
        best = other < best ? other : best;
    }
    if ((tid & 31) == 0) red[tid >> 5] = best;
    __syncthreads();
    if (tid == 0) {
        unsigned long long b = red[0];
        for (int w = 1; w < SSD_THREADS / 32; w++) b = red[w] < b ? red[w] : b;
        best_s = b;
    }
    __syncthreads();
    const uint32_t minv = (uint32_t)(best_s >> 32);
    const int mp = (int)(best_s & 0xffffffffu);
    const int mly = mp / mw, mlx = mp - mly * mw;
    // ---- tie rule (depth_calculator.cpp:226-237): mean column index of entries <= min right/below the first minimum
    int cnt = 0, sum = 0;
    for (int p = tid; p < mw * mh; p += SSD_THREADS) {
        const int k = p / mw, j = p - k * mw;
        if (j >= mlx && k >= mly && map[p] <= minv) { cnt++; sum += j; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { cnt += __shfl_down_sync(0xffffffffu, cnt, o); sum += __shfl_down_sync(0xffffffffu, sum, o); }
    if ((tid & 31) == 0) { cnt_s[tid >> 5] = cnt; sum_s[tid >> 5] = sum; }
    __syncthreads();
    if (tid == 0) {
        int c = 0, sm = 0;
        for (int w = 0; w < SSD_THREADS / 32; w++) { c += cnt_s[w]; sm += sum_s[w]; }
        float minPos = (float)sm / (float)c;  // float sum of small integers is exact
        a.disparity[i] = (a.mode == 1) ? fmaxf(0.5f, minPos) : minPos;
    }
}

// ---------------------------------------------------------------------------------------------------------
// Column kernel (windows up to 32 px wide, up to 16 vertical search positions — every shipped configuration):
// one thread owns one horizontal search position j and ALL vertical positions k of it.  Each ROI row is read and
// funnel-shifted once and then dotted against the (up to 13) template rows it meets, so the shared-memory traffic
// and shift work per SSD drop ~3x and the template rows arrive as broadcast 128-bit loads.  64 threads per keypoint.
// ---------------------------------------------------------------------------------------------------------
#define SSDC_THREADS 64
#define SSDC_MAXK 16

__global__ void __launch_bounds__(SSDC_THREADS) stereo_ssd_col_kernel(SsdArgs a, int rpw, int roi_rows, int map_cap)
{
    extern __shared__ __align__(16) uint32_t ssd_smem32[];
    const int win = a.cam.win_depth;
    uint32_t *tpl = ssd_smem32;                 // 32 rows x 8 words (16-byte aligned rows)
    uint32_t *roi = tpl + 32 * 8;               // roi_rows x rpw words
    uint32_t *map = roi + roi_rows * rpw;       // map_cap entries
    __shared__ unsigned long long red[SSDC_THREADS / 32];
    __shared__ unsigned long long best_s;
    __shared__ int cnt_s[SSDC_THREADS / 32], sum_s[SSDC_THREADS / 32];
    __shared__ unsigned saa_s[SSDC_THREADS / 32];

    const int i = blockIdx.x;
    const int n = min(*a.n_ptr, a.max_kps);
    if (i >= n) return;
    const int tid = threadIdx.x;
    const LevelDesc L = a.left0, R = a.right0;
    SsdGeom g;
    if (!ssd_geometry(a.kps2d[2 * i], a.kps2d[2 * i + 1], L.w, L.h, win, a.cam.search_x, a.cam.search_y, a.mode, g)) {
        if (tid == 0) a.disparity[i] = -1.f;
        return;
    }
    const int tw = g.x12 - g.x11, th = g.y12 - g.y11, rw = g.x22 - g.x21, rh = g.y22 - g.y21;
    const int mw = rw - tw + 1, mh = rh - th + 1;
    const uint32_t last_mask = (tw & 3) ? ((1u << ((tw & 3) * 8)) - 1u) : 0xffffffffu;
    const int twords = (tw + 3) >> 2;

    // ---- stage template and ROI row-wise: every thread fetches whole rows as independent aligned word loads and
    //      realigns them with funnel shifts (bytes past the window are masked where they are consumed)
    unsigned saa = 0;
    if (tid < th) {
        const uint8_t *rowp = L.ptr + (size_t)(g.y11 + tid) * L.pitch + g.x11;
        const uintptr_t addr = reinterpret_cast<uintptr_t>(rowp);
        const uint32_t *base = reinterpret_cast<const uint32_t *>(addr & ~(uintptr_t)3);
        const int sh = (int)(addr & 3) * 8;
        uint32_t wv[9];
#pragma unroll
        for (int k = 0; k < 9; k++) wv[k] = base[k];
#pragma unroll
        for (int k = 0; k < 8; k++) {
            uint32_t v = __funnelshift_r(wv[k], wv[k + 1], sh);
            if (k == twords - 1) v &= last_mask;
            if (k >= twords) v = 0;
            tpl[tid * 8 + k] = v;
            saa = __dp4a(v, v, saa);
        }
    }
    const int rwords = min(rpw - 1, ((rw + 3) >> 2) + 1);   // words of a ROI row that any search position can touch
    for (int r = tid; r < rh; r += SSDC_THREADS) {
        const uint8_t *rowp = R.ptr + (size_t)(g.y21 + r) * R.pitch + g.x21;
        const uintptr_t addr = reinterpret_cast<uintptr_t>(rowp);
        const uint32_t *base = reinterpret_cast<const uint32_t *>(addr & ~(uintptr_t)3);
        const int sh = (int)(addr & 3) * 8;
        uint32_t w0 = base[0];
        for (int k = 0; k < rwords; k += 4) {   // 4 independent loads per trip
            uint32_t w1 = base[k + 1], w2 = base[k + 2], w3 = base[k + 3], w4 = base[k + 4];
            roi[r * rpw + k] = __funnelshift_r(w0, w1, sh);
            if (k + 1 < rpw) roi[r * rpw + k + 1] = __funnelshift_r(w1, w2, sh);
            if (k + 2 < rpw) roi[r * rpw + k + 2] = __funnelshift_r(w2, w3, sh);
            if (k + 3 < rpw) roi[r * rpw + k + 3] = __funnelshift_r(w3, w4, sh);
            w0 = w4;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) saa += __shfl_down_sync(0xffffffffu, saa, o);
    if ((tid & 31) == 0) saa_s[tid >> 5] = saa;
    __syncthreads();
    saa = 0;
#pragma unroll
    for (int q = 0; q < SSDC_THREADS / 32; q++) saa += saa_s[q];

    unsigned long long best = ~0ull;
    for (int j = tid; j < mw; j += SSDC_THREADS) {
        const int jw = j >> 2, sh = (j & 3) * 8;
        unsigned sab[SSDC_MAXK], sbb[SSDC_MAXK];
#pragma unroll
        for (int k = 0; k < SSDC_MAXK; k++) { sab[k] = 0; sbb[k] = 0; }
        for (int r = 0; r < rh; r++) {
            const uint32_t *rr = roi + r * rpw + jw;
            uint32_t b[8];
            uint32_t w0 = rr[0];
#pragma unroll
            for (int x = 0; x < 8; x++) {
                const uint32_t w1 = rr[x + 1];
                uint32_t v = __funnelshift_r(w0, w1, sh);
                if (x == twords - 1) v &= last_mask;
                if (x >= twords) v = 0;
                b[x] = v;
                w0 = w1;
            }
            unsigned bb = 0;
#pragma unroll
            for (int x = 0; x < 8; x++) bb = __dp4a(b[x], b[x], bb);
#pragma unroll
            for (int k = 0; k < SSDC_MAXK; k++) {
                const int y = r - k;          // template row meeting ROI row r at vertical position k (uniform in the CTA)
                if (k < mh && y >= 0 && y < th) {
                    const uint4 t0 = *reinterpret_cast<const uint4 *>(tpl + y * 8);
                    const uint4 t1 = *reinterpret_cast<const uint4 *>(tpl + y * 8 + 4);
                    unsigned acc = sab[k];
                    acc = __dp4a(t0.x, b[0], acc); acc = __dp4a(t0.y, b[1], acc); acc = __dp4a(t0.z, b[2], acc); acc = __dp4a(t0.w, b[3], acc);
                    acc = __dp4a(t1.x, b[4], acc); acc = __dp4a(t1.y, b[5], acc); acc = __dp4a(t1.z, b[6], acc); acc = __dp4a(t1.w, b[7], acc);
                    sab[k] = acc;
                    sbb[k] += bb;
                }
            }
        }
#pragma unroll
        for (int k = 0; k < SSDC_MAXK; k++) {
            if (k < mh) {
                const uint32_t sv = saa + sbb[k] - 2u * sab[k];
                const int p = k * mw + j;
                map[p] = sv;
                unsigned long long key = ((unsigned long long)sv << 32) | (unsigned)p;
                best = key < best ? key : best;
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        unsigned long long other = __shfl_down_sync(0xffffffffu, best, o);
        best = other < best ? other : best;
    }
    if ((tid & 31) == 0) red[tid >> 5] = best;
    __syncthreads();
    if (tid == 0) {
        unsigned long long b = red[0];
        for (int w = 1; w < SSDC_THREADS / 32; w++) b = red[w] < b ? red[w] : b;
        best_s = b;
    }
    __syncthreads();
    const uint32_t minv = (uint32_t)(best_s >> 32);
    const int mp = (int)(best_s & 0xffffffffu);
    const int mly = mp / mw, mlx = mp - mly * mw;
    int cnt = 0, sum = 0;
    for (int p = tid; p < mw * mh; p += SSDC_THREADS) {
        const int k = p / mw, j = p - k * mw;
        if (j >= mlx && k >= mly && map[p] <= minv) { cnt++; sum += j; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { cnt += __shfl_down_sync(0xffffffffu, cnt, o); sum += __shfl_down_sync(0xffffffffu, sum, o); }
    if ((tid & 31) == 0) { cnt_s[tid >> 5] = cnt; sum_s[tid >> 5] = sum; }
    __syncthreads();
    if (tid == 0) {
        int c = 0, sm = 0;
        for (int w = 0; w < SSDC_THREADS / 32; w++) { c += cnt_s[w]; sm += sum_s[w]; }
        float minPos = (float)sm / (float)c;
        a.disparity[i] = (a.mode == 1) ? fmaxf(0.5f, minPos) : minPos;
    }
}


// ---------------------------------------------------------------------------------------------------------
// Tensor-core kernel (windows up to 32 px, up to 16 vertical and 64 horizontal search positions — every shipped
// configuration).  The cross term of the SSD is the one dense contraction on the tracking path: for a keypoint,
//     C[k][j] = sum_r sum_x Tz[r - k][x] * R[r][x + j]         (Tz = template, zero outside its th rows / tw columns)
// is, for every ROI row r, a 16 x 32 (vertical position k, x) times 32 x 8 (x, horizontal position j) product
// accumulated over r — exact in the integer tensor core (mma.sync m16n8k32 u8*u8 -> s32; sums stay < 2^26).
// Stages (the kernel below gives one keypoint to four warps):
//   * staging: lane = image row; a lane reads its row as aligned words, realigns it with funnel shifts and writes the
//     ROI four times, shifted by 0..3 bytes, so that every B fragment (4 consecutive bytes at ANY byte offset x + j)
//     is one aligned, conflict-free LDS.32.  The same pass leaves the running sum of squares of the row in shared
//     memory (P[r][i] = sum_{i' < i} R[r][i']^2).
//   * A fragments are template rows r - k: the template is stored once with 15 zero rows above and zero rows below,
//     so moving to the next ROI row just moves the read pointer by one row (a Toeplitz operand never materialised).
//   * sum b^2 of a window = vertical sliding sum of P[r][j + tw] - P[r][j]; lane = horizontal position.
//   * arg-min (value, raster index) and the tie rule run on the accumulator fragments in registers.
// ~2.8 k warp instructions per keypoint instead of ~25 k for the byte-dot-product kernel below.
// ---------------------------------------------------------------------------------------------------------
#define SSDM_TP 12    // template row pitch (words): rows r-g, g = 0..7 land in distinct banks
#define SSDM_CP 24    // pitch of one shifted ROI copy (words); 4 copies + 2 pad words per ROI row
#define SSDM_RPW (4 * SSDM_CP + 2)
#define SSDM_TPAD 15  // zero rows above the template (k up to 15)
#define SSDM_THREADS 128

__device__ __forceinline__ void mma_u8_16832(int (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1)
{
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// Four warps own one keypoint (the ~38 KB of staged data is what limits how many keypoints an SM holds, so the time a
// keypoint keeps it is what counts): staging is split over (ROI row, row half) pairs, the sum-b^2 pass runs in warps
// 0-1 while warps 2-3 already issue their MMA tiles, the tiles of 8 horizontal positions are dealt out two per warp.
template <int NT>
__global__ void __launch_bounds__(SSDM_THREADS) stereo_ssd_mma_kernel(SsdArgs a, int roi_rows)
{
    extern __shared__ __align__(16) uint32_t ssd_smem32[];
    constexpr int PP = 8 * NT + 33;               // pitch of the prefix-of-squares rows
    constexpr int NW = 2 * NT + 8;                // words of a shifted copy any fragment can touch: (g>>2) + 2nt + tig + 4 <= 2NT + 6
    constexpr int HW = NW / 2;                    // ... per row half
    constexpr int MT = (NT + 3) / 4;              // MMA tiles per warp
    const int win = a.cam.win_depth;
    uint32_t *tpl = ssd_smem32;                               // (SSDM_TPAD + roi_rows + 1) x SSDM_TP
    uint32_t *roi = tpl + (SSDM_TPAD + roi_rows + 1) * SSDM_TP;  // roi_rows x SSDM_RPW
    uint32_t *psq = roi + roi_rows * SSDM_RPW;                // roi_rows x PP
    uint32_t *sbb = psq + roi_rows * PP;                      // 16 x 8NT
    __shared__ unsigned s_saa;
    __shared__ unsigned long long s_best[SSDM_THREADS / 32];
    __shared__ int s_cnt[SSDM_THREADS / 32], s_sum[SSDM_THREADS / 32];

    const int i = blockIdx.x;
    const int n = min(*a.n_ptr, a.max_kps);
    if (i >= n) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const LevelDesc L = a.left0, R = a.right0;
    SsdGeom g;
    if (!ssd_geometry(a.kps2d[2 * i], a.kps2d[2 * i + 1], L.w, L.h, win, a.cam.search_x, a.cam.search_y, a.mode, g)) {   // CTA uniform
        if (tid == 0) a.disparity[i] = -1.f;
        return;
    }
    const int tw = g.x12 - g.x11, th = g.y12 - g.y11, rw = g.x22 - g.x21, rh = g.y22 - g.y21;
    const int mw = rw - tw + 1, mh = rh - th + 1;
    const uint32_t last_mask = (tw & 3) ? ((1u << ((tw & 3) * 8)) - 1u) : 0xffffffffu;
    const int twords = (tw + 3) >> 2;

    // ---- zero frame of the template
    for (int k = tid; k < (SSDM_TPAD + rh + 1) * SSDM_TP; k += SSDM_THREADS) tpl[k] = 0;
    __syncthreads();
    if (warp == 3) {
        // ---- template: lane = template row (th <= 32); this warp has no ROI rows to stage
        unsigned saa = 0;
        if (lane < th) {
            const uint8_t *rowp = L.ptr + (size_t)(g.y11 + lane) * L.pitch + g.x11;
            const uintptr_t addr = reinterpret_cast<uintptr_t>(rowp);
            const uint32_t *base = reinterpret_cast<const uint32_t *>(addr & ~(uintptr_t)3);
            const int sh = (int)(addr & 3) * 8;
            uint32_t wv[9];
#pragma unroll
            for (int k = 0; k < 9; k++) wv[k] = base[k];
#pragma unroll
            for (int k = 0; k < 8; k++) {
                uint32_t v = __funnelshift_r(wv[k], wv[k + 1], sh);
                if (k == twords - 1) v &= last_mask;
                if (k >= twords) v = 0;
                tpl[(SSDM_TPAD + lane) * SSDM_TP + k] = v;
                saa = __dp4a(v, v, saa);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) saa += __shfl_xor_sync(0xffffffffu, saa, o);
        if (lane == 0) s_saa = saa;
    } else {
        // ---- ROI: thread = (row, half of the row); rows beyond 48 (never with the supported settings) loop
        for (int it0 = 0; it0 < 2 * rh; it0 += 96) {   // warp-uniform trip count: the partner shuffle below needs every lane
            const int item = it0 + tid;
            const bool active = item < 2 * rh;
            const int r = active ? (item >> 1) : 0, hf = item & 1;
            const uint8_t *rowp = R.ptr + (size_t)(g.y21 + r) * R.pitch + g.x21;
            const uintptr_t addr = reinterpret_cast<uintptr_t>(rowp);
            const uint32_t *base = reinterpret_cast<const uint32_t *>(addr & ~(uintptr_t)3);
            const int sh = (int)(addr & 3) * 8;
            const int wlast = ((int)(addr & 3) + rw - 1) >> 2;   // last aligned word holding a byte of the ROI row
            const int k0 = hf * HW;
            uint32_t v0[HW + 1];                                 // bytes 4(k0+k) .. of the row, realigned
            {
                uint32_t w0 = base[min(k0, wlast)];
#pragma unroll
                for (int k = 0; k < HW + 1; k++) {
                    const uint32_t w1 = base[min(k0 + k + 1, wlast)];
                    v0[k] = __funnelshift_r(w0, w1, sh);
                    w0 = w1;
                }
            }
            uint32_t *dst = roi + r * SSDM_RPW + k0;
#pragma unroll
            for (int k = 0; k < HW; k++) {
                if (!active) break;
                dst[k] = v0[k];
                dst[SSDM_CP + k] = __funnelshift_r(v0[k], v0[k + 1], 8);
                dst[2 * SSDM_CP + k] = __funnelshift_r(v0[k], v0[k + 1], 16);
                dst[3 * SSDM_CP + k] = __funnelshift_r(v0[k], v0[k + 1], 24);
            }
            // exclusive prefix of squares along the row: psq[r][i] = sum_{i' < i} R[r][i']^2, i = 0 .. 4NW; the second half
            // starts from the first half's total (its partner is the neighbouring lane)
            unsigned tot = 0;
#pragma unroll
            for (int k = 0; k < HW; k++) tot = __dp4a(v0[k], v0[k], tot);
            const unsigned left = __shfl_up_sync(0xffffffffu, tot, 1);
            unsigned acc = hf ? left : 0u;
            uint32_t *pp = psq + r * PP + 4 * k0;
            if (!active) continue;
#pragma unroll
            for (int k = 0; k < HW; k++) {
#pragma unroll
                for (int b = 0; b < 4; b++) {
                    pp[4 * k + b] = acc;
                    const unsigned px = (v0[k] >> (8 * b)) & 255u;
                    acc += px * px;
                }
            }
            if (hf) pp[4 * HW] = acc;
        }
    }
    __syncthreads();

    // ---- sum b^2 per search position (warps 0-1: thread = horizontal position j, vertical sliding window over th rows)
    if (tid < 8 * NT && tid < mw) {
        const int j = tid;
        unsigned v = 0;
        for (int r = 0; r < th; r++) v += psq[r * PP + j + tw] - psq[r * PP + j];
        sbb[j] = v;
        for (int k = 1; k < mh; k++) {
            v += psq[(k + th - 1) * PP + j + tw] - psq[(k + th - 1) * PP + j];
            v -= psq[(k - 1) * PP + j + tw] - psq[(k - 1) * PP + j];
            sbb[k * 8 * NT + j] = v;
        }
    }

    // ---- cross term on the tensor core: warp w' = (warp + 2) & 3 owns tiles w' * MT .. w' * MT + MT - 1
    const int gid = lane >> 2, tig = lane & 3;
    const int t0 = ((warp + 2) & 3) * MT;
    int acc[MT][4];
#pragma unroll
    for (int t = 0; t < MT; t++) { acc[t][0] = 0; acc[t][1] = 0; acc[t][2] = 0; acc[t][3] = 0; }
    const int ntiles = min(NT, (mw + 7) >> 3);   // tiles holding at least one valid position (CTA uniform)
    if (t0 < ntiles) {
        const uint32_t *ta = tpl + (SSDM_TPAD - gid) * SSDM_TP + tig;
        const uint32_t *rb = roi + (gid & 3) * SSDM_CP + (gid >> 2) + tig + 2 * t0;
        for (int r = 0; r < rh; r++) {
            const uint32_t a0 = ta[0], a2 = ta[4], a1 = ta[-8 * SSDM_TP], a3 = ta[-8 * SSDM_TP + 4];
#pragma unroll
            for (int t = 0; t < MT; t++) {
                if (t0 + t < ntiles) mma_u8_16832(acc[t], a0, a1, a2, a3, rb[2 * t], rb[2 * t + 4]);
            }
            ta += SSDM_TP;
            rb += SSDM_RPW;
        }
    }
    __syncthreads();

    // ---- SSD = sum a^2 + sum b^2 - 2 sum ab; arg-min by (value, raster index) == cv::minMaxLoc's first minimum
    const unsigned saa = s_saa;
    unsigned long long best = ~0ull;
    uint32_t sv[MT][4];
#pragma unroll
    for (int t = 0; t < MT; t++) {
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int k = gid + (q >> 1) * 8, j = 8 * (t0 + t) + 2 * tig + (q & 1);
            sv[t][q] = 0xffffffffu;
            if (k < mh && j < mw) {
                const uint32_t v = saa + sbb[k * 8 * NT + j] - 2u * (uint32_t)acc[t][q];
                sv[t][q] = v;
                const unsigned long long key = ((unsigned long long)v << 32) | (unsigned)(k * mw + j);
                best = key < best ? key : best;
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
        best = other < best ? other : best;
    }
    if (lane == 0) s_best[warp] = best;
    __syncthreads();
#pragma unroll
    for (int w = 0; w < SSDM_THREADS / 32; w++) best = s_best[w] < best ? s_best[w] : best;
    const uint32_t minv = (uint32_t)(best >> 32);
    const int mp = (int)(best & 0xffffffffu);
    const int mly = mp / mw, mlx = mp - mly * mw;
    // ---- tie rule (depth_calculator.cpp:226-237): mean column index of entries <= min right/below the first minimum
    int cnt = 0, sum = 0;
#pragma unroll
    for (int t = 0; t < MT; t++) {
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int k = gid + (q >> 1) * 8, j = 8 * (t0 + t) + 2 * tig + (q & 1);
            if (k < mh && j < mw && j >= mlx && k >= mly && sv[t][q] <= minv) { cnt++; sum += j; }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { cnt += __shfl_xor_sync(0xffffffffu, cnt, o); sum += __shfl_xor_sync(0xffffffffu, sum, o); }
    if (lane == 0) { s_cnt[warp] = cnt; s_sum[warp] = sum; }
    __syncthreads();
    if (tid == 0) {
        int c = 0, sm = 0;
#pragma unroll
        for (int w = 0; w < SSDM_THREADS / 32; w++) { c += s_cnt[w]; sm += s_sum[w]; }
        const float minPos = (float)sm / (float)c;  // float sum of small integers is exact
        a.disparity[i] = (a.mode == 1) ? fmaxf(0.5f, minPos) : minPos;
    }
}

template <int NT>
static void launch_ssd_mma(const SsdArgs &a, int roi_rows, cudaStream_t st)
{
    const size_t smem = ((size_t)(SSDM_TPAD + roi_rows + 1) * SSDM_TP + (size_t)roi_rows * SSDM_RPW + (size_t)roi_rows * (8 * NT + 33) +
                         (size_t)16 * 8 * NT) * 4;
    if (smem > 48 * 1024) cudaFuncSetAttribute(stereo_ssd_mma_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    stereo_ssd_mma_kernel<NT><<<a.max_kps, SSDM_THREADS, smem, st>>>(a, roi_rows);
}

void launch_stereo_ssd(const SsdArgs &a, cudaStream_t st)
{
    if (a.max_kps <= 0) return;
    const int win = a.cam.win_depth;
    const int map_cap = (a.cam.search_x + 1) * (2 * a.cam.search_y + 1);
    static const bool no_mma = getenv("SVO_SSD_NO_MMA") != nullptr;   // A/B switch for profiling
    if (!no_mma && win <= 32 && 2 * a.cam.search_y + 1 <= 16 && a.cam.search_x + 1 <= 64) {
        const int roi_rows = win + 2 * a.cam.search_y + 1;
        switch ((a.cam.search_x + 8) / 8) {   // ceil((search_x + 1) / 8) tiles of 8 horizontal positions
        case 1: launch_ssd_mma<1>(a, roi_rows, st); break;
        case 2: launch_ssd_mma<2>(a, roi_rows, st); break;
        case 3: launch_ssd_mma<3>(a, roi_rows, st); break;
        case 4: launch_ssd_mma<4>(a, roi_rows, st); break;
        case 5: launch_ssd_mma<5>(a, roi_rows, st); break;
        case 6: launch_ssd_mma<6>(a, roi_rows, st); break;
        case 7: launch_ssd_mma<7>(a, roi_rows, st); break;
        default: launch_ssd_mma<8>(a, roi_rows, st); break;
        }
        return;
    }
    if (win <= 32 && 2 * a.cam.search_y + 1 <= SSDC_MAXK) {
        const int rpw = (win + a.cam.search_x + 3) / 4 + 10;      // 9 words are read from the last search column
        const int roi_rows = win + 2 * a.cam.search_y + 1;
        size_t smem = ((size_t)32 * 8 + (size_t)roi_rows * rpw + (size_t)map_cap) * 4 + 16;
        stereo_ssd_col_kernel<<<a.max_kps, SSDC_THREADS, smem, st>>>(a, rpw, roi_rows, map_cap);
        return;
    }
    const int tpw = (win + 3) / 4 + 1;                           // words per template row (+1: never read past)
    const int rpw = (win + a.cam.search_x + 3) / 4 + 2;          // words per ROI row (+1 word read by the funnel shift)
    const int roi_rows = win + 2 * a.cam.search_y + 1;
    size_t smem = ((size_t)win * tpw + (size_t)roi_rows * rpw + (size_t)map_cap) * 4 + 16;
    stereo_ssd_kernel<<<a.max_kps, SSD_THREADS, smem, st>>>(a, tpw, rpw, roi_rows, map_cap);
}

// ---------------------------------------------------------------------------------------------------------
// depth filter
// ---------------------------------------------------------------------------------------------------------
struct FrameMats { float R[9]; double Rdn[9]; };

// One CTA walks all keypoints (a thread per keypoint, grid-stride), so that the kernel can finish with the export of the
// frame's results to the host mirror (io_copy_block) without a grid-wide barrier — one launch less per frame.
#define DF_THREADS 512
__global__ void __launch_bounds__(DF_THREADS) depth_filter_kernel(FilterArgs a)
{
    __shared__ FrameMats fm;
    const int n = min(*a.n_ptr, a.max_kps);
    SolverTrace trc(a.trace, threadIdx.x == 0);
    trc.stamp(0);
    if (a.rdn_in) {
        // Rodrigues(-r) in double as the refinement kernel left it; float(Rodrigues(r)) is its transpose bit for bit
        if (threadIdx.x < 9) {
            const double v = a.rdn_in[threadIdx.x];
            fm.Rdn[threadIdx.x] = v;
            fm.R[(threadIdx.x % 3) * 3 + threadIdx.x / 3] = (float)v;
        }
    } else if (threadIdx.x == 0) {
        const float *p = a.pose;
        dev_rodrigues_f(p[3], p[4], p[5], fm.R);
        dev_rodrigues_d(-p[3], -p[4], -p[5], fm.Rdn);
    }
    __syncthreads();
    trc.stamp(1);
    for (int i = threadIdx.x; i < n; i += DF_THREADS) {
        const DevCam cam = a.cam;
        const float fx = cam.fx, fy = cam.fy, cx = cam.cx, cy = cam.cy, baseline = cam.baseline;
        const float *c2 = a.pose;  // frame translation
        const float *kfp = a.kf_pose_table + (size_t)a.keyframe_ids[i] * 24;  // pose(6) R(9) Ri(9)
        const float *c1 = kfp, *Rk = kfp + 6, *Rik = kfp + 15;
        const float u = a.kps2d[2 * i], v = a.kps2d[2 * i + 1];
        uint8_t flags = a.flags[i];
        int inl = a.inlier[i], outl = a.outlier[i];
        float P0 = a.kps3d[3 * i], P1 = a.kps3d[3 * i + 1], P2 = a.kps3d[3 * i + 2];

        // ---- outlier_check (depth_filter.cpp:52-128)
        {
            const float d = a.disparity[i];
            const float _z = baseline / fmaxf(d, 0.5f);
            const float _x = (u - cx) / fx * _z;
            const float _y = (v - cy) / fy * _z;
            float w0, w1, w2;
            dev_m33v(fm.R, _x, _y, _z, w0, w1, w2);
            w0 += c2[0]; w1 += c2[1]; w2 += c2[2];
            float a0, a1, a2, b0, b1, b2;
            dev_m33v(Rik, w0 - c1[0], w1 - c1[1], w2 - c1[2], a0, a1, a2);
            dev_m33v(Rik, P0 - c1[0], P1 - c1[1], P2 - c1[2], b0, b1, b2);
            const float disp_ref = baseline / b2;
            const float disp = baseline / a2;
            const float pixel_distance = disp - disp_ref;
            if (fabsf(pixel_distance) > 5 * 0.5f) outl++;
            else inl++;
        }
        // ---- update_kps3d (depth_filter.cpp:130-257)
        float kx = a.kf_state[2 * i], kP = a.kf_state[2 * i + 1];
        {
            float d0, d1, d2;
            dev_m33v(Rik, fabsf(c1[0] - c2[0]), fabsf(c1[1] - c2[1]), fabsf(c1[2] - c2[2]), d0, d1, d2);
            if (flags & (SVO_F_IGN_COMPLETE | SVO_F_IGN_REFINE)) {
                outl++;
            } else if (!((double)d0 < 0.1 && (double)d1 < 0.1)) {
                const float rfx = a.ref_kps2d[2 * i], rfy = a.ref_kps2d[2 * i + 1];
                float p10, p11, p12, p20, p21, p22;
                dev_m33v(Rk, rfx - cx, rfy - cy, fx, p10, p11, p12);
                dev_m33v(fm.R, u - cx, v - cy, fx, p20, p21, p22);
                // least squares [p1 -p2] l = c2 - c1 (cv::solve(DECOMP_SVD) in the reference; normal equations in double here)
                const double y0 = (double)(c2[0] - c1[0]), y1 = (double)(c2[1] - c1[1]), y2 = (double)(c2[2] - c1[2]);
                const double a00 = (double)p10 * p10 + (double)p11 * p11 + (double)p12 * p12;
                const double a01 = -((double)p10 * p20 + (double)p11 * p21 + (double)p12 * p22);
                const double a11 = (double)p20 * p20 + (double)p21 * p21 + (double)p22 * p22;
                const double r0 = (double)p10 * y0 + (double)p11 * y1 + (double)p12 * y2;
                const double r1 = -((double)p20 * y0 + (double)p21 * y1 + (double)p22 * y2);
                const double det = a00 * a11 - a01 * a01;
                const float l0 = (float)((a11 * r0 - a01 * r1) / det);
                // 1x1 cv::KalmanFilter (A = H = 1, Q = 1e-4): predict + correct, OpenCV's float/double mix
                const float dev = (float)(0.5 / (double)sqrtf(d0 * d0 + d1 * d1));
                const float Rn = dev * dev;
                const float xpre = kx;
                const float Ppre = (float)((double)kP + (double)1e-4f);
                // Vec3f new_p = inv_rotation_kf*l(0)*(p1-c1)  (matrix scaled first)
                float Ms[9];
    #pragma unroll
                for (int q = 0; q < 9; q++) Ms[q] = Rik[q] * l0;
                float np0, np1, np2;
                dev_m33v(Ms, p10 - c1[0], p11 - c1[1], p12 - c1[2], np0, np1, np2);
                const float meas = 1 / np2;
                const float t3 = (float)((double)Ppre + (double)Rn);
                // gain through OpenCV's SVD solve of a 1x1 system
                const double wd = sqrt((double)t3 * (double)t3);
                const float uu = t3 * (float)(wd > 1.1754943508222875e-38 ? 1. / wd : 0.);
                const float wf = (float)wd;
                float K = 0.f;
                if (fabs((double)wf) > (double)wf * (double)(float)(2.220446049250313e-16 * 2)) {
                    double s = (double)uu * (double)Ppre;
                    s *= 1. / (double)wf;
                    K = (float)(0.0 + s * 1.0);
                }
                const float t5 = (float)(-((double)xpre) + (double)meas);
                kx = (float)((double)K * (double)t5 + (double)xpre);
                kP = (float)(-((double)K * (double)Ppre) + (double)Ppre);
                const float _z = (float)(1.0 / (double)kx);
                const float _x = (rfx - cx) / fx * _z;
                const float _y = (rfy - cy) / fy * _z;
                float o0, o1, o2;
                dev_m33v(Rk, _x, _y, _z, o0, o1, o2);
                P0 = c1[0] + o0; P1 = c1[1] + o1; P2 = c1[2] + o2;
            }
        }
        // ---- post-processing (stereo_slam.cpp:205-226)
        if (outl > inl) flags |= SVO_F_IGN_COMPLETE;
        if (inl > outl) flags &= (uint8_t)~SVO_F_IGN_TEMP;
        a.kps3d[3 * i] = P0; a.kps3d[3 * i + 1] = P1; a.kps3d[3 * i + 2] = P2;
        a.flags[i] = flags; a.inlier[i] = inl; a.outlier[i] = outl;
        a.kf_state[2 * i] = kx; a.kf_state[2 * i + 1] = kP;
        // ---- re-projection of the updated point (stereo_slam.cpp:228-229)
        float ou, ov;
        dev_project(fm.Rdn, P0, P1, P2, c2[0], c2[1], c2[2], fx, fy, cx, cy, cam.k1, cam.k2, cam.p1, cam.p2, cam.k3, ou, ov);
        a.kps2d_out[2 * i] = ou; a.kps2d_out[2 * i + 1] = ov;
    }
    trc.stamp(2);
    if (a.do_export) {
        __syncthreads();   // this CTA wrote the last results of the frame; everything earlier in the stream is complete
        trc.stamp(3);
        io_copy_block(a.exp);
        trc.stamp(5);
        if (a.done_rec) {
            // completion record: the CTA barrier orders every thread's result stores before thread 0's system-scope fence
            // (fences are cumulative), the fence orders them before the sequence number the host polls for — ONE fence per frame
            __syncthreads();
            trc.stamp(6);
            if (threadIdx.x == 0) {
                unsigned long long t;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
                a.done_rec[2] = *a.t_start;
                a.done_rec[3] = t;
                __threadfence_system();
                trc.stamp(7);
                *reinterpret_cast<volatile unsigned *>(a.done_rec) = *a.seq_ptr;
            }
        }
    }
    trc.stamp(8);
    trc.finish();
}

void launch_depth_filter(const FilterArgs &a, cudaStream_t st)
{
    if (a.max_kps <= 0 && !a.do_export) return;
    depth_filter_kernel<<<1, DF_THREADS, 0, st>>>(a);
}
