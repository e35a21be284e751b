// csrc/seam.cu -- MapperTemplate::create_masks() without images on the GPU (modules/octvr/src/template.cpp:155-204 +
// cv::detail::DistanceSeamFinder, modules/stitching/src/seam_finders.cpp:86-133):
//   masks resized (cv::resize INTER_LINEAR, 8UC1) to a working width of <= 960 px, per camera the 3x3 chamfer distance
//   transform cv::distanceTransform(mask, DIST_L2, 3) (imgproc/src/distransform.cpp:69-139; full-width masks on a 3x tiled
//   copy so that the distance wraps around the panorama), every working pixel given to the camera with the largest
//   distance (ties: the lower camera index), the losers' masks zeroed there, and the masks resized back to the ROI size.
// Bit-identical to prep.cpp distance_seam_masks (and through it to the reference: tests/golden/tmpl_*.npz).
//
// The chamfer transform is a raster recurrence, t(x) = min(a(x), t(x-1) + HV) along a row with a(x) taken from the row
// above.  Along a row that is a min-plus prefix scan -- t(x) = HV x + min_{k <= x}(a(k) - HV k) -- so one CTA walks the
// rows of an image and scans each row in parallel (integers: exact in any order).  The backward pass is the mirror image.
#include "prep.h"
#include <cuda_runtime.h>
#include <memory>

namespace ob {

namespace {
constexpr int CH_THREADS = 1024, CH_ITEMS = 8;               // rows of up to 8192 px
constexpr int CH_HV = 62587, CH_DIAG = 89738, CH_FAR = 0x7FFFFFFF >> 2;   // cvRound(0.955 * 65536), cvRound(1.3693 * 65536), INT_MAX >> 2

struct ResizeU8 {
    const uint8_t* src; int sw, sh; uint8_t* dst; int dw, dh;
    const int* xofs; const short2* xa; const int* yofs; const short2* yb;
};
// cv::resize(8UC1, INTER_LINEAR) (imgwarp.cpp:3224-3500,1387-1500): 11-bit coefficients, the arithmetic of prep.cpp resize_linear
__global__ void __launch_bounds__(256) k_resize_u8(const ResizeU8 p)
{
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    if (x >= p.dw || y >= p.dh) return;
    const int sx = __ldg(p.xofs + x), sx1 = min(sx + 1, p.sw - 1), sy = __ldg(p.yofs + y);
    const short2 a = __ldg(p.xa + x), b = __ldg(p.yb + y);
    const uint8_t* s0 = p.src + (size_t)min(max(sy, 0), p.sh - 1) * p.sw;
    const uint8_t* s1 = p.src + (size_t)min(max(sy + 1, 0), p.sh - 1) * p.sw;
    const int top = s0[sx] * a.x + s0[sx1] * a.y, bot = s1[sx] * a.x + s1[sx1] * a.y;
    p.dst[(size_t)y * p.dw + x] = (uint8_t)((((b.x * (top >> 4)) >> 16) + ((b.y * (bot >> 4)) >> 16) + 2) >> 2);
}

struct ChamferJob { const uint8_t* mask; int w, h, tiles; int* tmp; float* dist; };   // tiles = 3: wrap-around (virtual width 3 w, middle copy kept)

// inclusive min-scan over the CTA of v[0..CH_ITEMS) per thread (thread t owns elements t*CH_ITEMS ..); reverse = suffix scan
template <bool REVERSE>
__device__ __forceinline__ void block_min_scan(int (&v)[CH_ITEMS], int* s_warp)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (!REVERSE) { for (int k = 1; k < CH_ITEMS; k++) v[k] = min(v[k], v[k - 1]); }
    else { for (int k = CH_ITEMS - 2; k >= 0; k--) v[k] = min(v[k], v[k + 1]); }
    int tot = REVERSE ? v[0] : v[CH_ITEMS - 1];               // this thread's total, scanned across the warp
    #pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int n = REVERSE ? __shfl_down_sync(0xffffffffu, tot, o) : __shfl_up_sync(0xffffffffu, tot, o);
        if (REVERSE ? lane + o < 32 : lane >= o) tot = min(tot, n);
    }
    if (lane == (REVERSE ? 0 : 31)) s_warp[warp] = tot;
    __syncthreads();
    if (warp == 0) {
        int w = s_warp[lane];
        #pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int n = REVERSE ? __shfl_down_sync(0xffffffffu, w, o) : __shfl_up_sync(0xffffffffu, w, o);
            if (REVERSE ? lane + o < 32 : lane >= o) w = min(w, n);
        }
        s_warp[32 + lane] = w;
    }
    __syncthreads();
    // exclusive prefix of the preceding threads: previous lanes of this warp and previous warps
    int ex = REVERSE ? __shfl_down_sync(0xffffffffu, tot, 1) : __shfl_up_sync(0xffffffffu, tot, 1);
    if (REVERSE ? lane == 31 : lane == 0) ex = 0x7FFFFFFF;
    const int pw = REVERSE ? (warp < 31 ? s_warp[32 + warp + 1] : 0x7FFFFFFF) : (warp > 0 ? s_warp[32 + warp - 1] : 0x7FFFFFFF);
    ex = min(ex, pw);
    #pragma unroll
    for (int k = 0; k < CH_ITEMS; k++) v[k] = min(v[k], ex);
    __syncthreads();
}

__global__ void __launch_bounds__(CH_THREADS) k_chamfer(const ChamferJob* jobs)
{
    extern __shared__ int s_rows[];                            // two rows of (VW + 2) ints + 64 scan slots
    const ChamferJob j = jobs[blockIdx.x];
    const int VW = j.w * j.tiles;                              // virtual width
    int* rowA = s_rows + 1, *rowB = s_rows + (VW + 2) + 1, *s_warp = s_rows + 2 * (VW + 2);
    const int x0 = threadIdx.x * CH_ITEMS;
    for (int x = threadIdx.x - 1; x <= VW; x += CH_THREADS) if (x >= -1) { rowA[x] = CH_FAR; rowB[x] = CH_FAR; }
    if (threadIdx.x == 0) { rowA[-1] = rowB[-1] = CH_FAR; rowA[VW] = rowB[VW] = CH_FAR; }
    __syncthreads();
    int* prev = rowA, *cur = rowB;
    // forward raster: NW, N, NE from the row above, W by the scan
    for (int y = 0; y < j.h; y++) {
        const uint8_t* m = j.mask + (size_t)y * j.w;
        int v[CH_ITEMS];
        #pragma unroll
        for (int k = 0; k < CH_ITEMS; k++) {
            const int x = x0 + k;
            int a = 0x7FFFFFFF;
            if (x < VW) {
                a = 0;
                if (m[x % j.w]) a = min(min(prev[x - 1] + CH_DIAG, prev[x] + CH_HV), prev[x + 1] + CH_DIAG);
                a -= CH_HV * x;
            }
            v[k] = a;
        }
        if (threadIdx.x == 0) v[0] = min(v[0], CH_FAR + CH_HV);                  // the border element left of x = 0
        block_min_scan<false>(v, s_warp);
        #pragma unroll
        for (int k = 0; k < CH_ITEMS; k++) {
            const int x = x0 + k;
            if (x < VW) { const int t = v[k] + CH_HV * x; cur[x] = t; j.tmp[(size_t)y * VW + x] = t; }
        }
        __syncthreads();
        int* sw = prev; prev = cur; cur = sw;
    }
    // backward raster: SE, S, SW from the row below, E by the reverse scan
    for (int x = threadIdx.x; x < VW; x += CH_THREADS) prev[x] = CH_FAR;
    __syncthreads();
    const float scale = 1.f / 65536;
    for (int y = j.h - 1; y >= 0; y--) {
        int v[CH_ITEMS];
        #pragma unroll
        for (int k = 0; k < CH_ITEMS; k++) {
            const int x = x0 + k;
            int c = 0x7FFFFFFF;
            if (x < VW) {
                c = j.tmp[(size_t)y * VW + x];
                c = min(min(min(c, prev[x + 1] + CH_DIAG), prev[x] + CH_HV), prev[x - 1] + CH_DIAG);
                c += CH_HV * x;
            }
            v[k] = c;
        }
        if (x0 <= VW - 1 && VW - 1 < x0 + CH_ITEMS) v[VW - 1 - x0] = min(v[VW - 1 - x0], CH_FAR + CH_HV + CH_HV * (VW - 1));   // border right of the last column
        block_min_scan<true>(v, s_warp);
        #pragma unroll
        for (int k = 0; k < CH_ITEMS; k++) {
            const int x = x0 + k;
            if (x < VW) {
                const int t = v[k] - CH_HV * x;
                cur[x] = t;
                if (j.tiles == 1) j.dist[(size_t)y * j.w + x] = (float)t * scale;
                else if (x >= j.w && x < 2 * j.w) j.dist[(size_t)y * j.w + (x - j.w)] = (float)t * scale;
            }
        }
        __syncthreads();
        int* sw = prev; prev = cur; cur = sw;
    }
}

struct SeamCam { int x, y, w, h; const float* dist; uint8_t* um; };
struct WinnerParams { SeamCam cam[16]; int n, rx, ry, rw, rh; };
// seam_finders.cpp:113-131: per pixel the candidates are ordered by distance, descending (the reference's std::sort over
// <= 16 entries keeps the lower index first among equals); every camera but the first loses the pixel
__global__ void __launch_bounds__(256) k_seam_winner(const __grid_constant__ WinnerParams p)
{
    const int x = p.rx + blockIdx.x * 32 + threadIdx.x, y = p.ry + blockIdx.y * 8 + threadIdx.y;
    if (x >= p.rx + p.rw || y >= p.ry + p.rh) return;
    int win = -1; float wd = 0.f;
    for (int k = 0; k < p.n; k++) {
        const SeamCam& c = p.cam[k];
        const int lx = x - c.x, ly = y - c.y;
        const float d = (lx >= 0 && ly >= 0 && lx < c.w && ly < c.h) ? c.dist[(size_t)ly * c.w + lx] : -1.f;
        if (win < 0 || d > wd) { win = k; wd = d; }
    }
    for (int k = 0; k < p.n; k++) {
        if (k == win) continue;
        const SeamCam& c = p.cam[k];
        const int lx = x - c.x, ly = y - c.y;
        if (lx >= 0 && ly >= 0 && lx < c.w && ly < c.h) c.um[(size_t)ly * c.w + lx] = 0;
    }
}

template <class T> struct Dev {
    T* p = nullptr;
    explicit Dev(size_t n) { OB_CUDA(cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T))); }
    Dev(const T* h, size_t n) : Dev(n) { if (n) OB_CUDA(cudaMemcpy(p, h, n * sizeof(T), cudaMemcpyHostToDevice)); }
    ~Dev() { cudaFree(p); }
    Dev(const Dev&) = delete;
};
void resize_u8(const uint8_t* d_src, int sw, int sh, uint8_t* d_dst, int dw, int dh)
{
    std::vector<int> xofs, yofs;
    std::vector<short> xa, yb;
    resize_linear_tables(sw, dw, true, xofs, xa);
    resize_linear_tables(sh, dh, false, yofs, yb);
    Dev<int> dx(xofs.data(), xofs.size()), dy(yofs.data(), yofs.size());
    Dev<short> da(xa.data(), xa.size()), db(yb.data(), yb.size());
    ResizeU8 p{ d_src, sw, sh, d_dst, dw, dh, dx.p, reinterpret_cast<const short2*>(da.p), dy.p, reinterpret_cast<const short2*>(db.p) };
    k_resize_u8<<<dim3((dw + 31) / 32, (dh + 7) / 8), dim3(32, 8)>>>(p);
    OB_CUDA(cudaGetLastError());
    OB_CUDA(cudaDeviceSynchronize());                         // the tables are freed on return
}
}  // namespace

// d_dist[i] = cv::distanceTransform(d_mask[i], DIST_L2, 3) for n device images (one CTA each, widths <= 8192)
void chamfer_l2_gpu_batch(const uint8_t* const* d_mask, const int* w, const int* h, float* const* d_dist, int n)
{
    std::vector<ChamferJob> jobs(n);
    std::vector<std::unique_ptr<Dev<int>>> tmp(n);
    int max_w = 0;
    for (int i = 0; i < n; i++) {
        OB_CHECK(w[i] > 0 && h[i] > 0 && w[i] <= CH_THREADS * CH_ITEMS, "chamfer: image width");
        tmp[i].reset(new Dev<int>((size_t)w[i] * h[i]));
        jobs[i] = ChamferJob{ d_mask[i], w[i], h[i], 1, tmp[i]->p, d_dist[i] };
        max_w = std::max(max_w, w[i]);
    }
    Dev<ChamferJob> d_jobs(jobs.data(), jobs.size());
    const size_t smem = (size_t)(2 * (max_w + 2) + 64) * sizeof(int);
    OB_CUDA(cudaFuncSetAttribute(k_chamfer, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_chamfer<<<n, CH_THREADS, smem>>>(d_jobs.p);
    OB_CUDA(cudaGetLastError());
    OB_CUDA(cudaDeviceSynchronize());                         // tmp and the job list are freed on return
}

bool distance_seam_masks_gpu(const std::vector<TInput>& in, int out_w, int device, std::vector<Img<uint8_t>>& out)
{
    const int n = (int)in.size();
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) { cudaGetLastError(); return false; }
    if (device < 0 || device >= count) device = 0;
    if (n < 1 || n > 16) return false;
    OB_CUDA(cudaSetDevice(device));
    const double scale = std::min(1.0, 960.0 / out_w);
    std::vector<Rect> sr(n);
    std::vector<std::unique_ptr<Dev<uint8_t>>> d_mask(n), d_um(n);
    std::vector<std::unique_ptr<Dev<float>>> d_dist(n);
    std::vector<std::unique_ptr<Dev<int>>> d_tmp(n);
    for (int i = 0; i < n; i++) {
        const Rect& r = in[i].roi;
        sr[i] = Rect{ (int)(r.x * scale), (int)(r.y * scale), (int)(r.w * scale), (int)(r.h * scale) };
        if (sr[i].w <= 0 || sr[i].h <= 0) return false;
        d_mask[i].reset(new Dev<uint8_t>(in[i].mask.d.data(), in[i].mask.d.size()));
        d_um[i].reset(new Dev<uint8_t>((size_t)sr[i].w * sr[i].h));
        resize_u8(d_mask[i]->p, r.w, r.h, d_um[i]->p, sr[i].w, sr[i].h);
    }
    Rect R = sr[0];
    for (int i = 1; i < n; i++) R = rect_union(R, sr[i]);
    std::vector<ChamferJob> jobs(n);
    int max_vw = 0;
    for (int i = 0; i < n; i++) {
        const int tiles = (sr[i].x == 0 && sr[i].w == R.w) ? 3 : 1;                // full-width mask: wrap-around DT on a 3x tiled copy
        const int vw = sr[i].w * tiles;
        if (vw > CH_THREADS * CH_ITEMS) return false;
        max_vw = std::max(max_vw, vw);
        d_dist[i].reset(new Dev<float>((size_t)sr[i].w * sr[i].h));
        d_tmp[i].reset(new Dev<int>((size_t)vw * sr[i].h));
        jobs[i] = ChamferJob{ d_um[i]->p, sr[i].w, sr[i].h, tiles, d_tmp[i]->p, d_dist[i]->p };
    }
    Dev<ChamferJob> d_jobs(jobs.data(), jobs.size());
    const size_t smem = (size_t)(2 * (max_vw + 2) + 64) * sizeof(int);
    OB_CUDA(cudaFuncSetAttribute(k_chamfer, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_chamfer<<<n, CH_THREADS, smem>>>(d_jobs.p);
    OB_CUDA(cudaGetLastError());
    WinnerParams wp;
    wp.n = n; wp.rx = R.x; wp.ry = R.y; wp.rw = R.w; wp.rh = R.h;
    for (int i = 0; i < n; i++) wp.cam[i] = SeamCam{ sr[i].x, sr[i].y, sr[i].w, sr[i].h, d_dist[i]->p, d_um[i]->p };
    k_seam_winner<<<dim3((R.w + 31) / 32, (R.h + 7) / 8), dim3(32, 8)>>>(wp);
    OB_CUDA(cudaGetLastError());
    OB_CUDA(cudaDeviceSynchronize());
    out.assign(n, Img<uint8_t>());
    for (int i = 0; i < n; i++) {
        const Rect& r = in[i].roi;
        Dev<uint8_t> d_full((size_t)r.w * r.h);
        resize_u8(d_um[i]->p, sr[i].w, sr[i].h, d_full.p, r.w, r.h);
        out[i] = Img<uint8_t>(r.w, r.h);
        OB_CUDA(cudaMemcpy(out[i].d.data(), d_full.p, (size_t)r.w * r.h, cudaMemcpyDeviceToHost));
    }
    return true;
}

}  // namespace ob
