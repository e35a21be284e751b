// csrc/mapper.cpp -- vr::Mapper (modules/octvr/src/mapper.{hpp,cpp}) behind the C ABI:
// construction packs the template into tile-compacted device tables, stitch() launches the
// per-frame kernels on the caller's stream.  No CPU fallback: a missing device is an error.
#include "mapper.h"
#include "prep.h"
#include <cuda.h>
#include <map>
#include <algorithm>
#include <cmath>
#include <climits>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <functional>
#include <queue>

using namespace ob;

namespace ob {

template <class T> static T* dev_upload(const T* h, size_t n)
{
    T* d = nullptr;
    OB_CUDA(cudaMalloc(&d, std::max<size_t>(n, 1) * sizeof(T)));
    if (n) OB_CUDA(cudaMemcpy(d, h, n * sizeof(T), cudaMemcpyHostToDevice));
    return d;
}
template <class T> static T* dev_alloc(size_t n, bool zero = false)
{
    T* d = nullptr;
    OB_CUDA(cudaMalloc(&d, std::max<size_t>(n, 1) * sizeof(T)));
    if (zero) OB_CUDA(cudaMemset(d, 0, std::max<size_t>(n, 1) * sizeof(T)));
    return d;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no -lcuda link dependency)
typedef CUresult (*TensorMapEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                      const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                      CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static TensorMapEncodeFn tensor_map_encoder()
{
    static TensorMapEncodeFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p)
            fail(OCTVR_ERR_CUDA, "cuTensorMapEncodeTiled is not available in this driver");
        fn = (TensorMapEncodeFn)p;
    }
    return fn;
}

// one 128-byte CUtensorMap over an RGBX plane (u32 elements, no swizzle, zero fill outside the plane) for w x h boxes
void encode_rgbx_tensor_map(void* out128, const uint32_t* plane, int plane_w, int plane_h, int box_w, int box_h)
{
    static_assert(sizeof(CUtensorMap) == 128, "CUtensorMap is 128 bytes");
    const cuuint64_t gdim[2] = { (cuuint64_t)plane_w, (cuuint64_t)plane_h };
    const cuuint64_t gstride[1] = { (cuuint64_t)plane_w * 4 };
    const cuuint32_t box[2] = { (cuuint32_t)box_w, (cuuint32_t)box_h };
    const cuuint32_t estr[2] = { 1, 1 };
    CUresult r = tensor_map_encoder()(reinterpret_cast<CUtensorMap*>(out128), CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, const_cast<uint32_t*>(plane), gdim, gstride,
                                      box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) fail(OCTVR_ERR_CUDA, "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
}

// table entry for one ROI pixel: fixed-point source position -> tap offset, fractions, border bits
// Entries that do not contribute are {offset 0, no flags}: the blend kernel runs them branch-free with
// weight 0 (reads source pixel 0, adds floor(v*0) = 0); the gain kernel checks C_VALID.
static inline uint2 make_entry(int32_t sx, int32_t sy, int src_w, int src_h, bool valid)
{
    if (!valid) return make_uint2(0u, 0u);
    int ix = sx >> 5, iy = sy >> 5;
    ix = std::min(32767, std::max(-32768, ix));           // saturate_cast<short> (imgwarp.cpp:4394-4395)
    iy = std::min(32767, std::max(-32768, iy));
    const uint32_t fx = (uint32_t)(sx & 31), fy = (uint32_t)(sy & 31);
    const bool x0 = ix >= 0 && ix < src_w, x1 = ix + 1 >= 0 && ix + 1 < src_w;
    const bool y0 = iy >= 0 && iy < src_h, y1 = iy + 1 >= 0 && iy + 1 < src_h;
    const uint32_t taps = (uint32_t)(x0 && y0) | ((uint32_t)(x1 && y0) << 1) | ((uint32_t)(x0 && y1) << 2) | ((uint32_t)(x1 && y1) << 3);
    if (taps == 0) return make_uint2(0u, 0u);             // every tap outside: contributes 0, same as skipping
    uint32_t flags = C_VALID;
    if (taps != 15u) flags |= C_BORDER | (taps << C_TAP_SHIFT);
    const int off = iy * src_w + ix;                      // may be "negative" for border entries; only inside taps are read
    return make_uint2((uint32_t)off, fx | (fy << 5) | flags);
}

}  // namespace ob

octvr_mapper::~octvr_mapper()
{
    cudaSetDevice(device);
    for (auto p : d_rgbx) cudaFree(p);
    for (auto p : d_vig) cudaFree(p);
    for (auto p : d_ov_coords) cudaFree(p);
    cudaFree(d_rgb_scaled); delete scale_plan; delete preview_plan;
    cudaFree(d_tile_job_start); cudaFree(d_job_cam); cudaFree(d_coords); cudaFree(d_weights); cudaFree(d_jobs); cudaFree(d_entries); cudaFree(d_tmaps); cudaFree(d_fjobs); cudaFree(d_fbins); cudaFree(d_fitems);
    cudaFree(d_rjobs); cudaFree(d_rentries); cudaFree(d_ring_counter); cudaFree(d_dbg_ring);
    cudaFree(d_smask); cudaFree(d_gcoord); cudaFree(d_partial); cudaFree(d_ticket); cudaFree(d_gsamples); cudaFree(d_gchunks); cudaFree(d_gtotals);
    cudaFree(d_gains); cudaFree(d_gain_f32); cudaFree(d_gain_flag); cudaFree(d_gain_lut); cudaFree(d_rgb); cudaFree(d_dbg);
    if (h_gains) cudaFreeHost(h_gains);
    for (auto& e : ev) if (e) cudaEventDestroy(e);
    ob::multiband_destroy(mb);
}

// ---- fused layout (K_stitch_fused): 32x32 tiles, one job per (tile, camera[, row range]).  A job carries the bounding
// box of its bilinear taps in the camera's source plane, the list of 8 px x 2 row blocks of that box which some tap
// touches (the units of the in-kernel colour conversion; everything else in the box is never read) and 8-byte entries
// {byte offset of the top-left tap inside the box | fy << 16 | fx << 24, weight}.  Tiles are distributed over the
// persistent CTAs by longest-processing-time-first bin packing and the job records are laid out in the order in
// which each CTA consumes them, every record followed by the item list of the CTA's next job.
static bool build_fused(octvr_mapper& m, const octvr_template& t, const std::vector<Img<int32_t>>& sx, const std::vector<Img<int32_t>>& sy,
                        const std::vector<Img<float>>& W)
{
    const int n = m.n;
    const int tiles_x = (t.out_w + FT_W - 1) / FT_W, tiles_y = (t.out_h + FT_H - 1) / FT_H, ntiles = tiles_x * tiles_y;
    struct Job { int cam, r0, r1, bh; FJob rec; std::vector<uint16_t> items; };
    std::vector<std::vector<Job>> tile_jobs(ntiles);
    // pixel of a tile -> (valid, tap position, weight) for camera i
    auto sample = [&](int tl, int px, int py, int i, int& ix, int& iy, int32_t& fsx, int32_t& fsy, float& w) {
        const TInput& in = t.inputs[i];
        const int gx = (tl % tiles_x) * FT_W + px, gy = (tl / tiles_x) * FT_H + py;
        const int lx = gx - in.roi.x, ly = gy - in.roi.y;
        if (gx >= t.out_w || gy >= t.out_h || lx < 0 || ly < 0 || lx >= in.roi.w || ly >= in.roi.h) return false;
        w = W[i].row(ly)[lx];
        fsx = sx[i].row(ly)[lx]; fsy = sy[i].row(ly)[lx];
        const uint2 e = make_entry(fsx, fsy, m.in_w[i], m.in_h[i], in.mask.row(ly)[lx] != 0 && w != 0.f);
        if (!(e.y & C_VALID)) return false;
        ix = std::min(32767, std::max(-32768, fsx >> 5)); iy = std::min(32767, std::max(-32768, fsy >> 5));
        return true;
    };
    auto floor_div = [](int a, int b) { return a >= 0 ? a / b : -((-a + b - 1) / b); };
    // job for the rows [r0, r1) of tile tl and camera i: 0 = nothing contributes, 1 = ok, -1 = does not fit the stage
    auto make_job = [&](int tl, int i, int r0, int r1, Job& jb) {
        int xmin = INT32_MAX, xmax = INT32_MIN, ymin = INT32_MAX, ymax = INT32_MIN;
        for (int py = r0; py < r1; py++)
            for (int px = 0; px < FT_W; px++) {
                int ix, iy; int32_t fsx, fsy; float w;
                if (!sample(tl, px, py, i, ix, iy, fsx, fsy, w)) continue;
                xmin = std::min(xmin, ix); xmax = std::max(xmax, ix + 1); ymin = std::min(ymin, iy); ymax = std::max(ymax, iy + 1);
            }
        if (xmin > xmax) return 0;
        FJob& r = jb.rec;
        memset(&r, 0, sizeof(r));
        r.cam = i;
        r.bx0 = floor_div(xmin, 8) * 8; r.by0 = floor_div(ymin, 2) * 2;
        r.bw = (xmax - r.bx0 + 8) / 8 * 8;
        const int bh = (ymax - r.by0 + 2) / 2 * 2, groups = r.bw / 4, rps = bh / 2;
        if ((int64_t)r.bw * bh > FUSED_CAP || groups > 128 || rps > 128) return -1;
        std::vector<uint8_t> touched((size_t)groups * rps, 0);
        for (int py = r0; py < r1; py++)
            for (int px = 0; px < FT_W; px++) {
                int ix, iy; int32_t fsx, fsy; float w;
                if (!sample(tl, px, py, i, ix, iy, fsx, fsy, w)) continue;
                for (int dy = 0; dy < 2; dy++)
                    for (int dx = 0; dx < 2; dx++) touched[(size_t)((iy + dy - r.by0) >> 1) * groups + ((ix + dx - r.bx0) >> 2)] = 1;
            }
        jb.items.clear();
        for (int rp = 0; rp < rps; rp++)
            for (int g = 0; g < groups; g++) {
                if (!touched[(size_t)rp * groups + g]) continue;
                const int x0 = r.bx0 + 4 * g, y0 = r.by0 + 2 * rp;
                int cls = FITEM_SLOW;
                if (x0 + 4 <= 0 || x0 >= m.in_w[i] || y0 + 2 <= 0 || y0 >= m.in_h[i]) cls = FITEM_ZERO;
                else if (x0 >= 0 && x0 + 4 <= m.in_w[i] && y0 >= 0 && y0 + 2 <= m.in_h[i]) cls = FITEM_FAST;
                jb.items.push_back((uint16_t)(rp | (g << 7) | (cls << 14)));
            }
        if ((int)jb.items.size() > FUSED_MAXITEMS) return -1;
        r.nitems = (int)jb.items.size();
        jb.cam = i; jb.r0 = r0; jb.r1 = r1; jb.bh = bh;
        return 1;
    };
    bool ok = true;
    std::function<void(int, int, int, int)> add_jobs = [&](int tl, int i, int r0, int r1) {
        Job jb;
        const int st = make_job(tl, i, r0, r1, jb);
        if (st == 0) return;
        if (st == 1) { tile_jobs[tl].push_back(std::move(jb)); return; }
        if (r1 - r0 <= 1) { ok = false; return; }          // a single output row needs a box larger than the stage
        const int mid = (r0 + r1) / 2;
        add_jobs(tl, i, r0, mid); add_jobs(tl, i, mid, r1);
    };
    auto in_band = [&](int tl) { const int y = (tl / tiles_x) * FT_H; return y >= m.band_y0 && y < m.band_y1; };
    for (int tl = 0; tl < ntiles && ok; tl++) {
        if (!in_band(tl)) continue;                        // another rank's rows (row-band mode)
        for (int i = 0; i < n; i++) add_jobs(tl, i, 0, FT_H);
        if (tile_jobs[tl].empty()) {                       // nobody covers this tile: one job with no items and zero weights
            Job jb; memset(&jb.rec, 0, sizeof(jb.rec));
            jb.cam = -1; jb.r0 = jb.r1 = 0; jb.bh = 2; jb.rec.bw = 8;
            tile_jobs[tl].push_back(jb);
        }
    }
    if (!ok) return false;                                 // the caller falls back to the two-kernel path

    // ---- schedule: LPT bin packing of tiles onto the persistent CTAs ----
    int sms = 0;
    OB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, m.device));
    const int per_sm = fused_ctas_per_sm();
    if (per_sm <= 0) fail(OCTVR_ERR_CUDA, "k_stitch_fused cannot be resident on this device");
    int band_tiles = 0;
    for (int i = 0; i < ntiles; i++) band_tiles += in_band(i);
    const int grid = std::max(1, std::min(band_tiles, sms * per_sm));
    std::vector<int64_t> cost(ntiles);
    for (int tl = 0; tl < ntiles; tl++) {
        // per-thread instruction estimates: ~190 for a job's four pixels, ~150 per conversion round, ~110 for the epilogue
        int64_t c = 110;
        for (const Job& jb : tile_jobs[tl]) c += 190 + 150 * (int64_t)((jb.items.size() + FT_THREADS - 1) / FT_THREADS) + 60;
        cost[tl] = c;
    }
    std::vector<int> order;
    for (int i = 0; i < ntiles; i++) if (in_band(i)) order.push_back(i);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return cost[a] > cost[b]; });
    std::vector<std::vector<int>> bin_tiles(grid);
    {
        typedef std::pair<int64_t, int> Load;              // (load, bin): min-heap
        std::priority_queue<Load, std::vector<Load>, std::greater<Load>> heap;
        for (int b = 0; b < grid; b++) heap.push(Load(0, b));
        for (int tl : order) {
            Load l = heap.top(); heap.pop();
            bin_tiles[l.second].push_back(tl);
            heap.push(Load(l.first + cost[tl], l.second));
        }
    }

    // ---- tables in consumption order ----
    std::vector<FJob> jobs;
    std::vector<FBin> bins(grid);
    std::vector<uint16_t> items;
    size_t njobs = 0;
    for (auto& v : tile_jobs) njobs += v.size();
    OB_CHECK(njobs * FT_PX < ((size_t)1 << 31), "table too large");
    std::vector<uint2> entries(njobs * FT_PX, make_uint2(0u, 0u));
    jobs.reserve(njobs);
    for (int b = 0; b < grid; b++) {
        bins[b].start = (int)jobs.size();
        int prev_off = 0, prev_area = 0;                    // stage region of the CTA's previous job
        for (int tl : bin_tiles[b])
            for (size_t k = 0; k < tile_jobs[tl].size(); k++) {
                const Job& jb = tile_jobs[tl][k];
                FJob r = jb.rec;
                r.cam = std::max(jb.cam, 0) | (k + 1 == tile_jobs[tl].size() ? FJOB_LAST : 0);
                // place the box where the previous job's box is not (its gather may still be running when this job is
                // converted); if both do not fit side by side the kernel separates them with a barrier
                const int area = (jb.bh * r.bw + 7) / 8 * 8;
                if (prev_off + prev_area + area <= FUSED_CAP) r.stage_off = prev_off + prev_area;
                else if (area <= prev_off) r.stage_off = 0;
                else { r.stage_off = 0; if ((int)jobs.size() > bins[b].start) r.cam |= FJOB_SYNC; }
                prev_off = r.stage_off; prev_area = area;
                r.tile_xy = (uint32_t)(tl % tiles_x) | ((uint32_t)(tl / tiles_x) << 16);
                // descriptors at a fixed stride per job, so their address does not depend on the job record
                r.items_off = (uint32_t)items.size();
                items.insert(items.end(), jb.items.begin(), jb.items.end());
                items.resize(r.items_off + FUSED_MAXITEMS, (uint16_t)0);
                if (jb.cam >= 0) {
                    uint2* ent = entries.data() + jobs.size() * FT_PX;
                    for (int tid = 0; tid < FT_THREADS; tid++)
                        for (int q = 0; q < FT_PPT; q++) {
                            const int px = tid & 31, py = (tid >> 5) + 8 * q;
                            if (py < jb.r0 || py >= jb.r1) continue;
                            int ix, iy; int32_t fsx, fsy; float w;
                            if (!sample(tl, px, py, jb.cam, ix, iy, fsx, fsy, w)) continue;
                            const uint32_t off = (uint32_t)((r.stage_off + (iy - r.by0) * r.bw + (ix - r.bx0)) * 4);
                            uint32_t wbits; memcpy(&wbits, &w, 4);
                            ent[tid * FT_PPT + q] = make_uint2(off | ((uint32_t)(fsy & 31) << 16) | ((uint32_t)(fsx & 31) << 24), wbits);
                        }
                }
                jobs.push_back(r);
            }
        bins[b].end = (int)jobs.size();
    }
    m.fused = true;
    m.fused_grid = grid;
    m.njobs = njobs;
    m.d_fjobs = dev_upload(jobs.data(), jobs.size());
    m.d_fbins = dev_upload(bins.data(), bins.size());
    m.d_fitems = dev_upload(items.data(), items.size());
    m.d_entries = dev_upload(entries.data(), entries.size());
    m.table_bytes = (int64_t)(entries.size() * sizeof(uint2) + jobs.size() * sizeof(FJob) + items.size() * 2);
    return true;
}

static void build_mapper(octvr_mapper& m, const octvr_template& t, const int* in_sizes, int n_in,
                         int blend, bool enable_gain, int scale_w, int scale_h, int band_y0, int band_y1, int band_x0 = 0, int band_x1 = 0)
{
    InitTrace tr("build_mapper");
    const int n = (int)t.inputs.size();
    const int n_ov = (int)t.overlays.size();
    OB_CHECK(n >= 1 && n + n_ov <= MAX_CAMS, "1..16 inputs (including overlays) supported");
    OB_CHECK(n_in == n + n_ov, "in_sizes must list every input and overlay");
    if (n == 1) { enable_gain = false; blend = 0; }        // mapper.cpp:78-82
    m.n = n; m.n_ov = n_ov; m.out_w = t.out_w; m.out_h = t.out_h; m.blend = blend; m.gain = enable_gain;
    // Pixel-coordinate convention.  Default: source pixel index = u * W, as cv::remap on map * W (template.cpp:174-176) and
    // every CPU path of the reference.  OCTVR_TEXEL_CENTER=1: u * W - 0.5, where the reference's CUDA Mapper samples (tex2D with
    // normalised coordinates and linear filtering) -- for templates calibrated against that path.
    if (const char* e = getenv("OCTVR_TEXEL_CENTER")) m.texel_shift = atoi(e) != 0 ? 0.5f : 0.f;
    // scaled_output_size = scale_output.area() == 0 ? mt.out_size : scale_output (mapper.cpp:68)
    const bool scaled = scale_w > 0 && scale_h > 0 && (scale_w != t.out_w || scale_h != t.out_h);
    m.scaled_w = scaled ? scale_w : t.out_w; m.scaled_h = scaled ? scale_h : t.out_h;
    OB_CHECK(m.scaled_w % 2 == 0 && m.scaled_h % 2 == 0, "scale_output must be even (4:2:0)");
    if (band_y0 == 0 && band_y1 == 0) band_y1 = t.out_h;
    if ((band_y0 != 0 || band_y1 != t.out_h) && (n_ov > 0 || scaled))
        fail(OCTVR_ERR_UNSUPPORTED, "row-band mappers do not take overlays or scale_output");
    OB_CHECK(band_y0 >= 0 && band_y0 < band_y1 && band_y1 <= t.out_h, "row band must lie inside the output");
    OB_CHECK(band_y0 % 32 == 0 && (band_y1 % 32 == 0 || band_y1 == t.out_h), "row bands must be aligned to 32 output rows");
    m.band_y0 = band_y0; m.band_y1 = band_y1;
    if (band_x0 == 0 && band_x1 == 0) band_x1 = t.out_w;
    OB_CHECK(band_x0 >= 0 && band_x0 < band_x1 && band_x1 <= t.out_w, "column band must lie inside the output");
    OB_CHECK(band_x0 % 32 == 0 && (band_x1 % 32 == 0 || band_x1 == t.out_w), "column bands must be aligned to 32 output columns");
    if (band_x0 != 0 || band_x1 != t.out_w) {
        if (n_ov > 0 || scaled) fail(OCTVR_ERR_UNSUPPORTED, "band mappers do not take overlays or scale_output");
        if (blend <= 0 || n == 1) fail(OCTVR_ERR_UNSUPPORTED, "column bands are a multiband (blend > 0) partition; feather / no-blend mappers split by rows");
    }
    m.band_x0 = band_x0; m.band_x1 = band_x1;
    OB_CHECK(t.out_w % 2 == 0 && t.out_h % 2 == 0, "output size must be even (4:2:0)");
    for (int i = 0; i < n + n_ov; i++) {
        int w = in_sizes[2 * i], h = in_sizes[2 * i + 1];
        OB_CHECK(w > 0 && h > 0 && w % 2 == 0 && h % 2 == 0, "input sizes must be positive and even (async.cpp:44-46)");
        OB_CHECK((int64_t)w * h < (int64_t)1 << 30, "input too large");
        m.in_w.push_back(w); m.in_h.push_back(h);
    }

    // ---- per-camera source planes ----
    for (int i = 0; i < n + n_ov; i++) {
        float* dv = nullptr;
        const TInput& ti = i < n ? t.inputs[i] : t.overlays[i - n];
        if (!ti.vignette.empty()) {                        // mapper.cpp:108-112, 121-126
            Img<float> v = resize_linear(ti.vignette, m.in_w[i], m.in_h[i]);
            dv = dev_upload(v.d.data(), v.d.size());
        }
        m.d_vig.push_back(dv);
    }

    // ---- fixed-point coordinates and blend weights per ROI pixel ----
    m.pairs = 0; m.roi_area = 0;
    for (int i = 0; i < n; i++) {
        m.roi_area += (int64_t)t.inputs[i].roi.w * t.inputs[i].roi.h;
        for (uint8_t v : t.inputs[i].mask.d) m.pairs += v != 0;
    }
    // OCTVR_BLEND=fused selects the single-kernel path (K_stitch_fused: no RGBX image in HBM, half the DRAM traffic, but
    // measured slower per frame than convert + blend because the conversion no longer hides the latency of the gain
    // launch -- DESIGN.md section 4); default: K_convert(+gain) then K_blend_ring
    const char* mode = getenv("OCTVR_BLEND");
    const bool want_fused = blend <= 0 && mode && std::string(mode) == "fused" && n_ov == 0;   // overlays need the RGBX planes
    // RGBX planes written by K_convert: only the two-kernel and multiband paths need them
    if (!want_fused)
        for (int i = 0; i < n + n_ov; i++) m.d_rgbx.push_back(dev_alloc<uint32_t>((size_t)m.in_w[i] * m.in_h[i] + 4));
    m.src_row0.assign(n + n_ov, 0);
    m.src_col0.assign(n + n_ov, 0);
    m.win_col0.assign(n + n_ov, 0);
    for (int i = 0; i < n + n_ov; i++) { m.src_row1.push_back(m.in_h[i]); m.src_col1.push_back(m.in_w[i]); m.win_w.push_back(m.in_w[i]); }
    m.gain_col0.assign(n, INT32_MAX); m.gain_col1.assign(n, INT32_MIN);
    tr.lap("vignette + pairs + planes");
    // default layout (K_blend_ring): quantisation, feather weights, job list, boxes and entries by CUDA kernels (pack.cu)
    const bool is_band = m.band_y0 != 0 || m.band_y1 != t.out_h;
    const bool gpu_packed = blend <= 0 && !want_fused && pack_ring_gpu(m, t, blend);
    // host copies of the fixed-point coordinates: only where host code still walks them (multiband set-up, the source rows of
    // a row band, the non-default layouts); the gain tables sample a few thousand positions through fixed_at()
    std::vector<Img<int32_t>> sx(n), sy(n);
    bool host_xy = false;
    auto ensure_host_xy = [&]() {
        if (host_xy) return;
        for (int i = 0; i < n; i++) quantise_map(t.inputs[i].map1, t.inputs[i].map2, m.in_w[i], m.in_h[i], sx[i], sy[i], m.texel_shift);
        host_xy = true;
        tr.lap("quantise_map (host)");
    };
    if (blend <= 0 && (!gpu_packed || is_band)) ensure_host_xy();
    auto fixed_at = [&](int i, int lx, int ly, int32_t& fsx, int32_t& fsy) {
        if (host_xy && !sx[i].empty()) { fsx = sx[i].row(ly)[lx]; fsy = sy[i].row(ly)[lx]; return; }
        const float fw = (float)(double)m.in_w[i], fh = (float)(double)m.in_h[i];      // quantise_map, one pixel
        float px = t.inputs[i].map1.row(ly)[lx] * fw + 0.f, py = t.inputs[i].map2.row(ly)[lx] * fh + 0.f;
        if (m.texel_shift != 0.f) { px = px - m.texel_shift; py = py - m.texel_shift; }
        fsx = (int32_t)lrintf(px * 32.f); fsy = (int32_t)lrintf(py * 32.f);
    };

    // source rows a row-band mapper touches (feather / no blend: the taps of its own output rows; the multiband set-up
    // widens this to its row window)
    if (is_band && blend <= 0)
        for (int i = 0; i < n; i++) {
            const TInput& in = t.inputs[i];
            int lo = INT32_MAX, hi = INT32_MIN;
            for (int y = std::max(0, m.band_y0 - in.roi.y); y < std::min(in.roi.h, m.band_y1 - in.roi.y); y++)
                for (int x = 0; x < in.roi.w; x++)
                    if (in.mask.row(y)[x]) { const int iy = sy[i].row(y)[x] >> 5; lo = std::min(lo, iy); hi = std::max(hi, iy + 1); }
            if (lo > hi) { m.src_row0[i] = m.src_row1[i] = 0; continue; }
            m.src_row0[i] = std::max(0, lo) & ~1; m.src_row1[i] = std::min(m.in_h[i], hi + 1);
        }

    std::vector<Img<float>> W;
    if (blend <= 0 && !gpu_packed) {
        W = blend < 0 ? feather_weights(t.inputs, -blend) : overwrite_weights(t.inputs);
        tr.lap("feather / overwrite weights (host)");
        m.inv_n = blend < 0 ? (float)(1.0 / n) : 1.f;
        if (want_fused) build_fused(m, t, sx, sy, W);
        if (want_fused && !m.fused)      // the fused layout did not apply: the two-kernel path needs its planes after all
            for (int i = 0; i < n + n_ov; i++) m.d_rgbx.push_back(dev_alloc<uint32_t>((size_t)m.in_w[i] * m.in_h[i] + 4));
    }
    // overlay inputs: one table entry per roi pixel (cv::remap fixed point; valid <=> mask), mapper.cpp:116-127
    for (int k = 0; k < n_ov; k++) {
        const TInput& in = t.overlays[k];
        Img<int32_t> ox, oy;
        quantise_map(in.map1, in.map2, m.in_w[n + k], m.in_h[n + k], ox, oy, m.texel_shift);
        std::vector<uint2> ce((size_t)in.roi.w * in.roi.h);
        for (int y = 0; y < in.roi.h; y++)
            for (int x = 0; x < in.roi.w; x++)
                ce[(size_t)y * in.roi.w + x] = make_entry(ox.row(y)[x], oy.row(y)[x], m.in_w[n + k], m.in_h[n + k], in.mask.row(y)[x] != 0);
        m.d_ov_coords.push_back(dev_upload(ce.data(), ce.size()));
        m.ov_roi.push_back(in.roi);
    }
    if (scaled) {
        m.scale_plan = ob::resize_plan_create(t.out_w, t.out_h, m.scaled_w, m.scaled_h);
        m.d_rgb_scaled = dev_alloc<uint8_t>((size_t)m.scaled_w * m.scaled_h * 3, true);
    }

    tr.lap("planes + overlays");
    if (blend > 0) {
        // (the host set-up takes the coordinates over: nothing after it walks them)
        m.mb = ob::multiband_create(m, t, [&](std::vector<Img<int32_t>>& ox, std::vector<Img<int32_t>>& oy) {
            ensure_host_xy();
            ox = std::move(sx); oy = std::move(sy);
            sx.assign(n, Img<int32_t>()); sy.assign(n, Img<int32_t>()); host_xy = false;
        });
        tr.lap("multiband_create");
    } else if (!gpu_packed) {
        if (!m.fused) {
        // ---- tile-compacted tables ----
        const int tiles_x = (t.out_w + TILE_W - 1) / TILE_W, tiles_y = (t.out_h + TILE_H - 1) / TILE_H;
        const int ntiles = tiles_x * tiles_y;
        m.tiles_x = tiles_x; m.tiles_y = tiles_y;
        std::vector<uint8_t> used((size_t)ntiles * n, 0);  // [tile][cam]
        for (int i = 0; i < n; i++) {
            const TInput& in = t.inputs[i];
            for (int y = 0; y < in.roi.h; y++) {
                const uint8_t* mk = in.mask.row(y);
                const float* wr = W[i].row(y);
                if (in.roi.y + y < m.band_y0 || in.roi.y + y >= m.band_y1) continue;      // another rank's rows
                const int ty = (in.roi.y + y) / TILE_H;
                for (int x = 0; x < in.roi.w; x++)
                    if (mk[x] && wr[x] != 0.f) used[((size_t)ty * tiles_x + (in.roi.x + x) / TILE_W) * n + i] = 1;
            }
        }
        std::vector<uint32_t> job_start(ntiles + 1, 0);
        std::vector<uint8_t> job_cam;
        for (int tl = 0; tl < ntiles; tl++) {
            job_start[tl] = (uint32_t)job_cam.size();
            for (int i = 0; i < n; i++) if (used[(size_t)tl * n + i]) job_cam.push_back((uint8_t)i);
        }
        job_start[ntiles] = (uint32_t)job_cam.size();
        const size_t njobs = job_cam.size();
        OB_CHECK(njobs * TILE_PX < ((size_t)1 << 32), "table too large");
        m.njobs = njobs;
        m.d_tile_job_start = dev_upload(job_start.data(), job_start.size());

        // ---- staged layout: per-job source footprint + 8-byte entries (K_blend_staged) ----
        // eligible when every source width is a multiple of 4 px (16-byte cp.async chunks) and every job's
        // footprint fits the shared-memory stage; otherwise the direct-gather kernel and its 12-byte tables are used
        bool staged = true;
        for (int i = 0; i < n; i++) staged = staged && m.in_w[i] % 4 == 0;
        if (const char* e = getenv("OCTVR_BLEND")) { if (std::string(e) == "direct") staged = false; }
        std::vector<JobMeta> meta(njobs);
        for (auto& jm : meta) memset(&jm, 0, sizeof(jm));
        std::map<uint64_t, int> tmap_index;       // (camera, box w, box h) -> descriptor slot
        auto pixel_of = [&](int tl, int p, int i, int& lx, int& ly) {
            const int tx = tl % tiles_x, ty = tl / tiles_x;
            const int gx = tx * TILE_W + (p & (TILE_W - 1)), gy = ty * TILE_H + (p / TILE_W);
            lx = gx - t.inputs[i].roi.x; ly = gy - t.inputs[i].roi.y;
            return lx >= 0 && ly >= 0 && lx < t.inputs[i].roi.w && ly < t.inputs[i].roi.h && gx < t.out_w && gy < t.out_h;
        };
        if (staged) {
            for (int tl = 0; tl < ntiles && staged; tl++)
                for (uint32_t j = job_start[tl]; j < job_start[tl + 1]; j++) {
                    const int i = job_cam[j];
                    const TInput& in = t.inputs[i];
                    int xmin = INT32_MAX, xmax = INT32_MIN, ymin = INT32_MAX, ymax = INT32_MIN;
                    for (int p = 0; p < TILE_PX; p++) {
                        int lx, ly;
                        if (!pixel_of(tl, p, i, lx, ly)) continue;
                        const float w = W[i].row(ly)[lx];
                        const uint2 e = make_entry(sx[i].row(ly)[lx], sy[i].row(ly)[lx], m.in_w[i], m.in_h[i], in.mask.row(ly)[lx] != 0 && w != 0.f);
                        if (!(e.y & C_VALID)) continue;
                        const int ix = std::min(32767, std::max(-32768, sx[i].row(ly)[lx] >> 5)), iy = std::min(32767, std::max(-32768, sy[i].row(ly)[lx] >> 5));
                        xmin = std::min(xmin, ix); xmax = std::max(xmax, ix + 1); ymin = std::min(ymin, iy); ymax = std::max(ymax, iy + 1);
                    }
                    JobMeta& jm = meta[j];
                    jm.cam = i;
                    if (xmin > xmax) { xmin = xmax = ymin = ymax = 0; }
                    // TMA box: footprint size rounded up to a size class (one tensor map per camera and class in use)
                    auto size_class = [](int v) { return v <= 64 ? (v + 7) / 8 * 8 : v <= 128 ? (v + 15) / 16 * 16 : (v + 31) / 32 * 32; };
                    // measured on B200: the innermost TMA start coordinate must be 16-byte aligned (4 px here), else the
                    // copy raises "illegal instruction"; negative and out-of-range coordinates are fine (zero filled)
                    jm.bx0 = (int)std::floor(xmin / 4.0) * 4; jm.by0 = ymin;
                    jm.bw = size_class(xmax - jm.bx0 + 1); jm.bh = size_class(ymax - ymin + 1);
                    if (jm.bw > 256 || jm.bh > 256 || (int64_t)jm.bw * jm.bh > STAGE_CAP) { staged = false; break; }
                    const uint64_t key = ((uint64_t)i << 32) | ((uint64_t)jm.bw << 16) | (uint64_t)jm.bh;
                    auto it = tmap_index.find(key);
                    if (it == tmap_index.end()) it = tmap_index.emplace(key, (int)tmap_index.size()).first;
                    jm.tmap = it->second;
                }
            // pack each tile's jobs into stage-sized groups (almost always one group)
            for (int tl = 0; tl < ntiles && staged; tl++) {
                int grp = 0, used_px = 0;
                for (uint32_t j = job_start[tl]; j < job_start[tl + 1]; j++) {
                    const int area = meta[j].bw * meta[j].bh;
                    if (used_px + area > STAGE_CAP) { grp++; used_px = 0; }
                    meta[j].soff = used_px; meta[j].grp_nj = (grp << 8) | (int)(job_start[tl + 1] - job_start[tl]); meta[j].j0 = (int)job_start[tl];
                    used_px += area;
                }
            }
        }
        m.staged = staged;
        tr.lap("tile jobs + boxes");
        // K_blend_ring (default): the same boxes and tensor maps, entries regrouped for four pixels per thread, every
        // job's [entries | box] must fit the shared-memory ring.  OCTVR_BLEND=staged keeps the one-tile-per-CTA kernel.
        bool ring = staged;
        if (const char* e = getenv("OCTVR_BLEND")) { if (std::string(e) == "staged") ring = false; }
        for (size_t j = 0; j < njobs && ring; j++)
            ring = (((size_t)meta[j].bw * meta[j].bh * 4 + 127) & ~(size_t)127) <= (size_t)RING_BYTES && meta[j].bw < 4096;
        ring = ring && njobs < ((size_t)1 << 20);
        if (ring) {
            m.ring_ctas = ring_ctas_per_sm();
            OB_CUDA(cudaDeviceGetAttribute(&m.sm_count, cudaDevAttrMultiProcessorCount, m.device));
            if (m.ring_ctas <= 0) ring = false;
        }
        m.ring = ring;
        if (ring) {
            std::vector<uint4> entries(njobs * (TILE_PX / 2), make_uint4(0u, 0u, 0u, 0u));
            std::vector<uint4> recs((size_t)ntiles * n, make_uint4(0u, 0u, 0u, 0u));
            for (int tl = 0; tl < ntiles; tl++) {
                for (uint32_t j = job_start[tl]; j < job_start[tl + 1]; j++) {
                    const int i = job_cam[j];
                    const TInput& in = t.inputs[i];
                    const JobMeta& jm = meta[j];
                    recs[(size_t)tl * n + (j - job_start[tl])] = make_uint4(((uint32_t)jm.bx0 & 0xFFFFu) | ((uint32_t)jm.by0 << 16), (uint32_t)jm.tmap | ((uint32_t)i << 16),
                                                                            (uint32_t)(jm.bw * jm.bh * 4), (uint32_t)jm.bw | (j << 12));
                    OB_CHECK(jm.bx0 >= -32768 && jm.bx0 < 32768 && jm.by0 >= -32768 && jm.by0 < 32768 && jm.tmap < 65536, "box origin out of range");
                    uint32_t* ent = reinterpret_cast<uint32_t*>(entries.data() + (size_t)j * (TILE_PX / 2));
                    for (int p = 0; p < TILE_PX; p++) {
                        int lx, ly;
                        if (!pixel_of(tl, p, i, lx, ly)) continue;
                        const float w = W[i].row(ly)[lx];
                        const int32_t fsx = sx[i].row(ly)[lx], fsy = sy[i].row(ly)[lx];
                        const uint2 e = make_entry(fsx, fsy, m.in_w[i], m.in_h[i], in.mask.row(ly)[lx] != 0 && w != 0.f);
                        if (!(e.y & C_VALID)) continue;
                        const int ix = std::min(32767, std::max(-32768, fsx >> 5)), iy = std::min(32767, std::max(-32768, fsy >> 5));
                        const uint32_t off = (uint32_t)(((iy - jm.by0) * jm.bw + (ix - jm.bx0)) * 4);       // bytes inside the box
                        uint32_t wbits; memcpy(&wbits, &w, 4);
                        // pixel (col, row) of the tile -> thread (row & 7) >> 1 << 5 | col, slot q = (row & 1) | (row >> 3) << 1
                        const int col = p & (TILE_W - 1), row = p / TILE_W;
                        const int tid = (((row & 7) >> 1) << 5) | col, q = (row & 1) | ((row >> 3) << 1);
                        uint32_t* u4 = ent + ((size_t)(q >> 1) * 128 + tid) * 4;
                        u4[q & 1] = (off << 16) | ((uint32_t)(fsy & 31) << 8) | (uint32_t)(fsx & 31);
                        u4[2 + (q & 1)] = wbits;
                    }
                }
            }
            std::vector<CUtensorMap> tmaps(tmap_index.size());
            for (auto& kv : tmap_index) {
                const int cam = (int)(kv.first >> 32), bw = (int)((kv.first >> 16) & 0xFFFF), bh = (int)(kv.first & 0xFFFF);
                encode_rgbx_tensor_map(&tmaps[kv.second], m.d_rgbx[cam], m.in_w[cam], m.in_h[cam], bw, bh);
            }
            m.d_tmaps = (void*)dev_upload(tmaps.data(), tmaps.size());
            m.n_tmaps = (int)tmaps.size();
            m.d_rjobs = dev_upload(recs.data(), recs.size());
            m.d_rentries = dev_upload(entries.data(), entries.size());
            m.d_ring_counter = dev_alloc<unsigned int>(1, true);
            m.d_dbg_ring = dev_alloc<unsigned long long>(8, true);
            m.table_bytes = (int64_t)(entries.size() * sizeof(uint4) + recs.size() * 16);
        } else if (staged) {
            std::vector<uint2> entries(njobs * TILE_PX, make_uint2(0u, 0u));
            for (int tl = 0; tl < ntiles; tl++)
                for (uint32_t j = job_start[tl]; j < job_start[tl + 1]; j++) {
                    const int i = job_cam[j];
                    const TInput& in = t.inputs[i];
                    const JobMeta& jm = meta[j];
                    for (int p = 0; p < TILE_PX; p++) {
                        int lx, ly;
                        if (!pixel_of(tl, p, i, lx, ly)) continue;
                        const float w = W[i].row(ly)[lx];
                        const int32_t fsx = sx[i].row(ly)[lx], fsy = sy[i].row(ly)[lx];
                        const uint2 e = make_entry(fsx, fsy, m.in_w[i], m.in_h[i], in.mask.row(ly)[lx] != 0 && w != 0.f);
                        if (!(e.y & C_VALID)) continue;
                        const int ix = std::min(32767, std::max(-32768, fsx >> 5)), iy = std::min(32767, std::max(-32768, fsy >> 5));
                        const uint32_t off = (uint32_t)(jm.soff + (iy - jm.by0) * jm.bw + (ix - jm.bx0));
                        uint32_t wbits; memcpy(&wbits, &w, 4);
                        entries[(size_t)j * TILE_PX + p] = make_uint2((off << 2) | ((uint32_t)(fsy & 31) << 16) | ((uint32_t)(fsx & 31) << 24), wbits);   // byte offset | fy | fx
                    }
                }
            // tensor maps over the mapper-owned RGBX planes (static addresses): u32 elements, no swizzle, zero OOB fill
            std::vector<CUtensorMap> tmaps(tmap_index.size());
            for (auto& kv : tmap_index) {
                const int cam = (int)(kv.first >> 32), bw = (int)((kv.first >> 16) & 0xFFFF), bh = (int)(kv.first & 0xFFFF);
                const cuuint64_t gdim[2] = { (cuuint64_t)m.in_w[cam], (cuuint64_t)m.in_h[cam] };
                const cuuint64_t gstride[1] = { (cuuint64_t)m.in_w[cam] * 4 };
                const cuuint32_t box[2] = { (cuuint32_t)bw, (cuuint32_t)bh };
                const cuuint32_t estr[2] = { 1, 1 };
                CUresult r = tensor_map_encoder()(&tmaps[kv.second], CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, m.d_rgbx[cam], gdim, gstride, box, estr,
                                                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                if (r != CUDA_SUCCESS) fail(OCTVR_ERR_CUDA, "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
            }
            m.d_tmaps = (void*)dev_upload(tmaps.data(), tmaps.size());
            std::vector<JobMeta> recs((size_t)ntiles * MAX_CAMS);
            for (auto& r : recs) memset(&r, 0, sizeof(r));
            for (int tl = 0; tl < ntiles; tl++)
                for (uint32_t j = job_start[tl]; j < job_start[tl + 1]; j++) recs[(size_t)tl * MAX_CAMS + (j - job_start[tl])] = meta[j];
            m.d_jobs = dev_upload(recs.data(), recs.size());
            m.d_entries = dev_upload(entries.data(), entries.size());
            m.table_bytes = (int64_t)(entries.size() * sizeof(uint2) + (size_t)ntiles * 64 * 3);   // ~3 records read per tile
            m.n_tmaps = (int)tmaps.size();
        } else {
            std::vector<uint2> coords(njobs * TILE_PX);
            std::vector<float> weights(njobs * TILE_PX);
            for (int tl = 0; tl < ntiles; tl++)
                for (uint32_t j = job_start[tl]; j < job_start[tl + 1]; j++) {
                    const int i = job_cam[j];
                    const TInput& in = t.inputs[i];
                    for (int p = 0; p < TILE_PX; p++) {
                        int lx, ly;
                        uint2 e = make_uint2(0u, 0u);
                        float w = 0.f;
                        if (pixel_of(tl, p, i, lx, ly)) {
                            w = W[i].row(ly)[lx];
                            e = make_entry(sx[i].row(ly)[lx], sy[i].row(ly)[lx], m.in_w[i], m.in_h[i], in.mask.row(ly)[lx] != 0 && w != 0.f);
                        }
                        if (!(e.y & C_VALID)) w = 0.f;
                        coords[(size_t)j * TILE_PX + p] = e;
                        weights[(size_t)j * TILE_PX + p] = w;
                    }
                }
            m.d_job_cam = dev_upload(job_cam.data(), job_cam.size());
            m.d_coords = dev_upload(coords.data(), coords.size());
            m.d_weights = dev_upload(weights.data(), weights.size());
            m.table_bytes = (int64_t)(coords.size() * sizeof(uint2) + weights.size() * sizeof(float) + job_cam.size() + job_start.size() * 4);
        }
        }   // !fused
        tr.lap("entries + upload");
    }

    // ---- gain compensation tables (mapper.cpp:94-99,113-114,235-237) ----
    if (enable_gain) {
        const double ws = std::min(1.0, std::sqrt(0.1 * 1e6 / ((double)t.out_w * t.out_h)));
        GainParams& g = m.gp;
        memset(&g, 0, sizeof(g));
        g.n = n;
        std::vector<uint8_t> smask;
        std::vector<uint2> gcoord;
        uint32_t off = 0;
        for (int i = 0; i < n; i++) {
            const TInput& in = t.inputs[i];
            GainCam& c = g.cam[i];
            c.sx = (int)(in.roi.x * ws); c.sy = (int)(in.roi.y * ws); c.sw = (int)(in.roi.w * ws); c.sh = (int)(in.roi.h * ws);
            OB_CHECK(c.sw > 0 && c.sh > 0, "working-scale ROI is empty");
            c.off = off;
            Img<uint8_t> sm = resize_linear(in.mask, c.sw, c.sh);
            smask.insert(smask.end(), sm.d.begin(), sm.d.end());
            const double ifx = 1. / ((double)c.sw / in.roi.w), ify = 1. / ((double)c.sh / in.roi.h);   // resizeNN
            for (int dy = 0; dy < c.sh; dy++) {
                const int ly = std::min((int)std::floor(dy * ify), in.roi.h - 1);
                for (int dx = 0; dx < c.sw; dx++) {
                    const int lx = std::min((int)std::floor(dx * ifx), in.roi.w - 1);
                    int32_t fsx, fsy;
                    fixed_at(i, lx, ly, fsx, fsy);
                    uint2 e = make_entry(fsx, fsy, m.in_w[i], m.in_h[i], in.mask.row(ly)[lx] != 0);
                    if (e.y & C_VALID) {        // the gain kernel reads the input planes directly: keep (ix, iy), not a plane offset
                        const int ix = std::min(32767, std::max(-32768, fsx >> 5)), iy = std::min(32767, std::max(-32768, fsy >> 5));
                        e.x = (uint32_t)(ix + 1) | ((uint32_t)(iy + 1) << 16);
                        if (sm.row(dy)[dx] == 255) { m.gain_col0[i] = std::min(m.gain_col0[i], std::max(0, ix)); m.gain_col1[i] = std::max(m.gain_col1[i], std::min(m.in_w[i], ix + 2)); }
                    }
                    if (sm.row(dy)[dx] != 255) e = make_uint2(0xFFFFFFFFu, 0u);     // CPU compensator's intersect rule: mask == 255
                    gcoord.push_back(e);
                }
            }
            off += (uint32_t)c.sw * c.sh;
        }
        g.total = off;
        g.n_pairs = n * (n + 1) / 2;
        int x0 = g.cam[0].sx, y0 = g.cam[0].sy, x1 = x0 + g.cam[0].sw, y1 = y0 + g.cam[0].sh;
        for (int i = 1; i < n; i++) {
            x0 = std::min(x0, g.cam[i].sx); y0 = std::min(y0, g.cam[i].sy);
            x1 = std::max(x1, g.cam[i].sx + g.cam[i].sw); y1 = std::max(y1, g.cam[i].sy + g.cam[i].sh);
        }
        g.cx0 = x0; g.cy0 = y0; g.cw = x1 - x0; g.ch = y1 - y0;
        // samples regrouped per canvas chunk (<= 256 pixels, <= 512 samples): one CTA per chunk, two samples per thread;
        // every chunk owns 512 sample slots (empty ones all ones), so a thread finds its samples without a chunk table
        constexpr int CH_PX = 256, CH_SAMPLES = 512;       // kernels.cu: GAIN_PX, GAIN_BLK * GAIN_SPT
        std::vector<uint4> samples;
        {
            const uint4 empty = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
            size_t first = 0; int npx = 0, cnt = 0;
            // a chunk's samples are stored camera by camera (pixels in canvas order within a camera): the 32 lanes of a warp then
            // gather from one plane along one curve, i.e. from 2-3 cache lines per load instead of ~10 (the gather phase of the
            // statistics is bound by L1 wavefronts: 24 scattered byte loads per thread)
            auto close = [&]() {
                if (cnt) {
                    std::stable_sort(samples.begin() + first, samples.end(), [](const uint4& a, const uint4& b) { return (a.z & 255u) < (b.z & 255u); });
                    samples.resize(first + CH_SAMPLES, empty); first = samples.size();
                }
                npx = 0; cnt = 0;
            };
            std::vector<uint4> here;
            for (int pix = 0; pix < g.cw * g.ch; pix++) {
                const int X = g.cx0 + pix % g.cw, Y = g.cy0 + pix / g.cw;
                here.clear();
                for (int i = 0; i < n; i++) {
                    const GainCam& c = g.cam[i];
                    const int lx = X - c.sx, ly = Y - c.sy;
                    if (lx < 0 || ly < 0 || lx >= c.sw || ly >= c.sh) continue;
                    const uint2 e = gcoord[c.off + (size_t)ly * c.sw + lx];
                    if (e.x == 0xFFFFFFFFu) continue;            // working-scale mask != 255: not in any intersection
                    here.push_back(make_uint4(e.x, e.y, (uint32_t)i, 0u));
                }
                if (here.empty()) continue;
                if (npx == CH_PX || cnt + (int)here.size() > CH_SAMPLES) close();
                for (uint4 s : here) { s.z |= (uint32_t)npx << 8; samples.push_back(s); cnt++; }
                npx++;
            }
            close();
            if (samples.empty()) samples.resize(CH_SAMPLES, empty);
        }
        std::vector<int2> chunks(samples.size() / CH_SAMPLES, make_int2(0, 0));
        g.grid = (int)chunks.size();
        m.d_smask = dev_upload(smask.data(), smask.size());
        m.d_gcoord = dev_upload(gcoord.data(), gcoord.size());
        m.d_gsamples = dev_upload(samples.data(), samples.size());
        m.d_gchunks = dev_upload(chunks.data(), chunks.size());
        m.d_gtotals = dev_alloc<unsigned long long>((size_t)g.n_pairs * 5, true);
        g.samples = m.d_gsamples; g.chunks = m.d_gchunks; g.totals = m.d_gtotals;
        m.d_partial = dev_alloc<double>((size_t)g.n_pairs * 3, true);
        m.d_ticket = dev_alloc<unsigned int>(1, true);
        m.d_gains = dev_alloc<double>(MAX_CAMS, true);
        m.d_gain_f32 = dev_alloc<float>(MAX_CAMS, true);
        m.d_gain_flag = dev_alloc<int>(MAX_CAMS, true);
        m.d_gain_lut = dev_alloc<uint8_t>(MAX_CAMS * 256, true);
        OB_CUDA(cudaMallocHost(&m.h_gains, sizeof(double) * MAX_CAMS));
        g.smask = m.d_smask; g.gcoord = m.d_gcoord; g.partial = m.d_partial; g.ticket = m.d_ticket;
        m.d_dbg = dev_alloc<unsigned long long>(8 + 2 * 4096, true); g.dbg = m.d_dbg;
        g.dbg_trace = getenv("OCTVR_GAIN_TRACE") ? m.d_dbg : nullptr;
        g.gains = m.d_gains; g.gain_f32 = m.d_gain_f32; g.gain_flag = m.d_gain_flag; g.gain_lut = m.d_gain_lut;
        m.table_bytes += (int64_t)(samples.size() * sizeof(uint4) + chunks.size() * sizeof(int2));
    }
    tr.lap("gain tables");
    for (auto& e : m.ev) OB_CUDA(cudaEventCreate(&e));

}

static void check_frame(const octvr_frame& f, int w, int h, const char* what, bool allow_rgb = false)
{
    if (f.uv_pixel_stride == OCTVR_FMT_RGB24 || f.uv_pixel_stride == OCTVR_FMT_BGR24) {       // packed 8UC3 in the y plane
        OB_CHECK(allow_rgb, "packed RGB / BGR frames are an input format only");
        OB_CHECK(f.y && f.y_pitch >= (size_t)w * 3, what);
        return;
    }
    OB_CHECK(f.y && f.u && f.v, what);
    OB_CHECK(f.y_pitch >= (size_t)w && f.uv_pixel_stride >= 1 && f.uv_pixel_stride <= 2, what);
    OB_CHECK(f.u_pitch >= (size_t)(w / 2) * f.uv_pixel_stride && f.v_pitch >= (size_t)(w / 2) * f.uv_pixel_stride, what);
    (void)h;
}

namespace ob {
void mapper_stitch_internal(octvr_mapper& m, const octvr_frame* in, int n_in, const octvr_frame* out,
                            const double* gains, int n_gains, const double* d_gains_src, cudaStream_t s,
                            uint8_t* d_preview, size_t preview_pitch, int preview_w, int preview_h);
}
static void do_stitch(octvr_mapper& m, const octvr_frame* in, int n_in, const octvr_frame* out,
                      const double* gains, int n_gains, cudaStream_t s,
                      uint8_t* d_preview = nullptr, size_t preview_pitch = 0, int preview_w = 0, int preview_h = 0)
{
    ob::mapper_stitch_internal(m, in, n_in, out, gains, n_gains, nullptr, s, d_preview, preview_pitch, preview_w, preview_h);
}

// gains: host array of predefined gains, or d_gains_src: another mapper's device gains (gain sharing between
// output regions, async.cpp:75-86), or neither: computed from this frame.
void ob::mapper_stitch_internal(octvr_mapper& m, const octvr_frame* in, int n_in, const octvr_frame* user_out,
                                const double* gains, int n_gains, const double* d_gains_src, cudaStream_t s,
                                uint8_t* d_preview, size_t preview_pitch, int preview_w, int preview_h)
{
    const int n_all = m.n + m.n_ov;
    OB_CHECK(n_in == n_all, "wrong number of input frames");         // mapper.cpp:208
    OB_CUDA(cudaSetDevice(m.device));
    for (int i = 0; i < n_all; i++) check_frame(in[i], m.win_w[i], m.in_h[i], "bad input frame", true);
    if (user_out) check_frame(*user_out, m.scaled_w, m.scaled_h, "bad output frame");
    OB_CHECK(user_out || m.keep_rgb || d_preview, "no output requested");
    if (d_preview) OB_CHECK(preview_w > 0 && preview_h > 0 && preview_pitch >= (size_t)preview_w * 3, "bad preview buffer");
    // The blend kernels write the 4:2:0 output themselves unless something still has to happen to the RGB result
    // (overlays copied over it, resize to scale_output: mapper.cpp:279-306); those go through post.cu.
    const bool scaled = m.scale_plan != nullptr;
    const bool post_yuv = m.n_ov > 0 || scaled;
    const octvr_frame* out = post_yuv ? nullptr : user_out;
    m.rgb_this_frame = m.keep_rgb || post_yuv || d_preview != nullptr;
    if (m.profiling) OB_CUDA(cudaEventRecord(m.ev[0], s));

    ConvertParams cp;                 // also describes the input planes for the fused kernel
    memset(&cp, 0, sizeof(cp));
    cp.n = n_all;
    for (int i = 0; i < n_all; i++) {
        CamSrc& c = cp.cam[i];
        c.y = in[i].y; c.u = in[i].u; c.v = in[i].v;
        c.y_pitch = (uint32_t)in[i].y_pitch; c.u_pitch = (uint32_t)in[i].u_pitch; c.v_pitch = (uint32_t)in[i].v_pitch;
        c.uv_step = in[i].uv_pixel_stride; c.w = m.in_w[i]; c.h = m.in_h[i];
        if (c.uv_step == OCTVR_FMT_RGB24 || c.uv_step == OCTVR_FMT_BGR24) {     // packed 8UC3 (cv::Mat CV_8UC3 as the tools read it: BGR)
            OB_CHECK(!m.fused, "the single-kernel path (OCTVR_BLEND=fused) takes 4:2:0 input only");
            const bool bgr = c.uv_step == OCTVR_FMT_BGR24;
            c.y = in[i].y + (bgr ? 2 : 0); c.u = in[i].y + 1; c.v = in[i].y + (bgr ? 0 : 2);
            c.u_pitch = c.v_pitch = c.y_pitch; c.uv_step = 3; c.rgb = 1;
        }
        if (m.win_col0[i] != 0) {        // the frame holds source columns [win_col0, win_col0 + win_w): address column 0 virtually
            const int c0 = m.win_col0[i];
            if (c.rgb) { c.y -= 3 * (size_t)c0; c.u -= 3 * (size_t)c0; c.v -= 3 * (size_t)c0; }
            else { c.y -= c0; c.u -= (size_t)(c0 / 2) * c.uv_step; c.v -= (size_t)(c0 / 2) * c.uv_step; }
        }
        c.rgbx = m.fused ? nullptr : m.d_rgbx[i]; c.vignette = m.d_vig[i];
        const bool chroma_ok = c.uv_step == 1
            ? ((uintptr_t)c.u % 4 == 0 && (uintptr_t)c.v % 4 == 0 && c.u_pitch % 4 == 0 && c.v_pitch % 4 == 0)
            : ((uintptr_t)c.u % 8 == 0 && c.u_pitch % 8 == 0 && c.v == c.u + 1);
        c.aligned4 = ((uintptr_t)c.y % 8 == 0) && (c.y_pitch % 8 == 0) && (c.w % 4 == 0) && chroma_ok;
        c.row0 = m.src_row0[i]; c.row1 = m.src_row1[i];
        c.col0 = m.src_col0[i]; c.col1 = m.src_col1[i];
        c.xlo = m.win_col0[i]; c.xhi = m.win_col0[i] + m.win_w[i] - 1;
        cp.grid_x = std::max(cp.grid_x, (c.col1 - c.col0 + 255) / 256);
        cp.grid_y = std::max(cp.grid_y, (c.row1 - c.row0 + 15) / 16);
    }
    // one launch: gain statistics + solve (reading the input planes directly) in the first CTAs, conversion in the rest
    // (fused path: no conversion pass at all, the launch only carries the gain CTAs)
    const bool compute_gains = m.gain && !d_gains_src && !gains;
    if (m.fused) { cp.grid_x = cp.grid_y = 0; }
    if (compute_gains) {
        GainParams gp = m.gp;
        for (int i = 0; i < m.n; i++) gp.src[i] = cp.cam[i];
        launch_convert_gain(cp, &gp, s);
    } else if (!m.fused)
        launch_convert_gain(cp, nullptr, s);
    if (m.profiling) OB_CUDA(cudaEventRecord(m.ev[1], s));

    if (m.gain) {
        if (d_gains_src) {
            OB_CUDA(cudaMemcpyAsync(m.d_gains, d_gains_src, sizeof(double) * m.n, cudaMemcpyDeviceToDevice, s));
            launch_gain_finalize(m.gp, s);
        } else if (gains) {
            OB_CHECK(n_gains == m.n, "gains size must equal the number of inputs");
            // the previous frame may still be reading h_gains: wait for it before overwriting
            if (m.last_stream_valid) OB_CUDA(cudaStreamSynchronize(m.last_stream));
            memcpy(m.h_gains, gains, sizeof(double) * m.n);
            OB_CUDA(cudaMemcpyAsync(m.d_gains, m.h_gains, sizeof(double) * m.n, cudaMemcpyHostToDevice, s));
            launch_gain_finalize(m.gp, s);
        }
    }
    if (m.profiling) OB_CUDA(cudaEventRecord(m.ev[2], s));

    if (m.rgb_this_frame && !m.d_rgb) m.d_rgb = dev_alloc<uint8_t>((size_t)m.out_w * m.out_h * 3, true);
    if (m.mb) {
        ob::multiband_stitch(m, out, s);
    } else if (m.fused) {
        FusedParams fp;
        memset(&fp, 0, sizeof(fp));
        for (int i = 0; i < m.n; i++) fp.cam[i] = cp.cam[i];
        fp.n = m.n;
        fp.jobs = m.d_fjobs; fp.bins = m.d_fbins; fp.items = m.d_fitems; fp.entries = m.d_entries;
        fp.out_w = m.out_w; fp.out_h = m.out_h;
        if (out) {
            fp.oy = out->y; fp.ou = out->u; fp.ov = out->v;
            fp.oy_pitch = (uint32_t)out->y_pitch; fp.ou_pitch = (uint32_t)out->u_pitch; fp.ov_pitch = (uint32_t)out->v_pitch;
            fp.uv_step = out->uv_pixel_stride;
        }
        fp.rgb_out = m.rgb_this_frame ? m.d_rgb : nullptr; fp.rgb_pitch = (uint32_t)m.out_w * 3;
        fp.gain_f32 = m.d_gain_f32; fp.gain_flag = m.d_gain_flag; fp.gain_lut = m.d_gain_lut;
        fp.use_gain = m.gain ? 1 : 0;
        fp.inv_n = m.inv_n;
        launch_stitch_fused(fp, m.fused_grid, s);
    } else if (m.ring) {
        RingParams rp;
        memset(&rp, 0, sizeof(rp));
        rp.jobs = m.d_rjobs; rp.nslot = m.n; rp.entries = m.d_rentries; rp.tmaps = m.d_tmaps; rp.counter = m.d_ring_counter; rp.dbg = m.d_dbg_ring;
        rp.tiles_x = m.tiles_x; rp.out_w = m.out_w; rp.out_h = m.out_h;
        const int ty0 = m.band_y0 / TILE_H, ty1 = (m.band_y1 + TILE_H - 1) / TILE_H;
        rp.tile0 = ty0 * m.tiles_x; rp.ntiles = (ty1 - ty0) * m.tiles_x;
        if (out) {
            rp.oy = out->y; rp.ou = out->u; rp.ov = out->v;
            rp.oy_pitch = (uint32_t)out->y_pitch; rp.ou_pitch = (uint32_t)out->u_pitch; rp.ov_pitch = (uint32_t)out->v_pitch;
            rp.uv_step = out->uv_pixel_stride;
        }
        rp.rgb_out = m.rgb_this_frame ? m.d_rgb : nullptr; rp.rgb_pitch = (uint32_t)m.out_w * 3;
        rp.fast_store = out && !rp.rgb_out && rp.uv_step == 1 && (uintptr_t)rp.oy % 16 == 0 && (uintptr_t)rp.ou % 16 == 0 && (uintptr_t)rp.ov % 16 == 0 &&
                        rp.oy_pitch % 16 == 0 && rp.ou_pitch % 16 == 0 && rp.ov_pitch % 16 == 0;
        rp.gain_f32 = m.d_gain_f32; rp.gain_flag = m.d_gain_flag; rp.gain_lut = m.d_gain_lut;
        rp.use_gain = m.gain ? 1 : 0;
        rp.inv_n = m.inv_n;
        launch_blend_ring(rp, std::max(1, std::min(rp.ntiles, m.sm_count * m.ring_ctas)), s);
    } else if (m.staged) {
        StagedParams sp;
        memset(&sp, 0, sizeof(sp));
        for (int i = 0; i < m.n; i++) { sp.rgbx[i] = m.d_rgbx[i]; sp.src_w[i] = m.in_w[i]; sp.src_h[i] = m.in_h[i]; }
        sp.jobs = m.d_jobs; sp.entries = m.d_entries; sp.tmaps = m.d_tmaps;
        sp.tiles_x = m.tiles_x; sp.tiles_y = m.tiles_y; sp.out_w = m.out_w; sp.out_h = m.out_h;
        sp.tile_y0 = m.band_y0 / TILE_H; sp.tiles_y_run = (m.band_y1 + TILE_H - 1) / TILE_H - sp.tile_y0;
        if (out) {
            sp.oy = out->y; sp.ou = out->u; sp.ov = out->v;
            sp.oy_pitch = (uint32_t)out->y_pitch; sp.ou_pitch = (uint32_t)out->u_pitch; sp.ov_pitch = (uint32_t)out->v_pitch;
            sp.uv_step = out->uv_pixel_stride;
        }
        sp.rgb_out = m.rgb_this_frame ? m.d_rgb : nullptr; sp.rgb_pitch = (uint32_t)m.out_w * 3;
        sp.gain_f32 = m.d_gain_f32; sp.gain_flag = m.d_gain_flag; sp.gain_lut = m.d_gain_lut;
        sp.use_gain = m.gain ? 1 : 0;
        sp.inv_n = m.inv_n;
        launch_blend_staged(sp, s);
    } else {
        BlendParams bp;
        memset(&bp, 0, sizeof(bp));
        for (int i = 0; i < m.n; i++) { bp.rgbx[i] = m.d_rgbx[i]; bp.src_pitch[i] = m.in_w[i]; }
        bp.tile_job_start = m.d_tile_job_start; bp.job_cam = m.d_job_cam; bp.coords = m.d_coords; bp.weights = m.d_weights;
        bp.tiles_x = m.tiles_x; bp.tiles_y = m.tiles_y; bp.out_w = m.out_w; bp.out_h = m.out_h;
        bp.tile_y0 = m.band_y0 / TILE_H; bp.tiles_y_run = (m.band_y1 + TILE_H - 1) / TILE_H - bp.tile_y0;
        if (out) {
            bp.oy = out->y; bp.ou = out->u; bp.ov = out->v;
            bp.oy_pitch = (uint32_t)out->y_pitch; bp.ou_pitch = (uint32_t)out->u_pitch; bp.ov_pitch = (uint32_t)out->v_pitch;
            bp.uv_step = out->uv_pixel_stride;
        }
        bp.rgb_out = m.rgb_this_frame ? m.d_rgb : nullptr; bp.rgb_pitch = (uint32_t)m.out_w * 3;
        bp.gain_f32 = m.d_gain_f32; bp.gain_flag = m.d_gain_flag; bp.gain_lut = m.d_gain_lut;
        bp.use_gain = m.gain ? 1 : 0;
        bp.inv_n = m.inv_n;
        launch_blend(bp, s);
    }
    // ---- after the blender: overlays, scale_output, preview (mapper.cpp:279-312) ----
    for (int k = 0; k < m.n_ov; k++)
        launch_overlay(m.d_rgbx[m.n + k], m.in_w[m.n + k], m.d_ov_coords[k], m.ov_roi[k], m.d_rgb, (size_t)m.out_w * 3, s);
    if (post_yuv && user_out) {
        const uint8_t* src = m.d_rgb;
        if (scaled) { launch_resize_rgb(*m.scale_plan, m.d_rgb, (size_t)m.out_w * 3, m.d_rgb_scaled, (size_t)m.scaled_w * 3, s); src = m.d_rgb_scaled; }
        launch_rgb_to_yuv420(src, (size_t)m.scaled_w * 3, m.scaled_w, m.scaled_h, *user_out, s);
    }
    if (d_preview) {                  // cv::cuda::resize(result, preview_output, preview_output.size(), INTER_LINEAR), mapper.cpp:308-312
        if (!m.preview_plan || m.preview_plan->dw != preview_w || m.preview_plan->dh != preview_h) {
            if (m.preview_plan && m.last_stream_valid) OB_CUDA(cudaStreamSynchronize(m.last_stream));   // tables may still be in use
            delete m.preview_plan; m.preview_plan = nullptr;
            m.preview_plan = ob::resize_plan_create(m.out_w, m.out_h, preview_w, preview_h);
        }
        launch_resize_rgb(*m.preview_plan, m.d_rgb, (size_t)m.out_w * 3, d_preview, preview_pitch, s);
    }
    if (m.profiling) OB_CUDA(cudaEventRecord(m.ev[3], s));
    OB_CUDA(cudaGetLastError());
    m.last_stream = s; m.last_stream_valid = true;
}

extern "C" {

octvr_status octvr_mapper_create(const octvr_template* t, const int* in_sizes_wh, int n_in, int blend,
                                 int enable_gain, int scale_w, int scale_h, int device, octvr_mapper** out)
{
    return guard([&] {
        OB_CHECK(t && in_sizes_wh && out, "null argument");
        int count = 0;
        cudaError_t e = cudaGetDeviceCount(&count);
        if (e != cudaSuccess || count <= 0 || device < 0 || device >= count)
            fail(OCTVR_ERR_CUDA, "no usable CUDA device (the stitch path has no CPU fallback)");
        OB_CUDA(cudaSetDevice(device));
        std::unique_ptr<octvr_mapper> m(new octvr_mapper);
        m->device = device;
        build_mapper(*m, *t, in_sizes_wh, n_in, blend, enable_gain != 0, scale_w, scale_h, 0, 0);
        *out = m.release();
    });
}

octvr_status octvr_mapper_create_band(const octvr_template* t, const int* in_sizes_wh, int n_in, int blend,
                                      int enable_gain, int band_y0, int band_y1, int device, octvr_mapper** out)
{
    return guard([&] {
        OB_CHECK(t && in_sizes_wh && out, "null argument");
        int count = 0;
        cudaError_t e = cudaGetDeviceCount(&count);
        if (e != cudaSuccess || count <= 0 || device < 0 || device >= count)
            fail(OCTVR_ERR_CUDA, "no usable CUDA device (the stitch path has no CPU fallback)");
        OB_CUDA(cudaSetDevice(device));
        std::unique_ptr<octvr_mapper> m(new octvr_mapper);
        m->device = device;
        build_mapper(*m, *t, in_sizes_wh, n_in, blend, enable_gain != 0, 0, 0, band_y0, band_y1);
        *out = m.release();
    });
}

octvr_status octvr_mapper_create_window(const octvr_template* t, const int* in_sizes_wh, int n_in, int blend, int enable_gain,
                                        int x0, int x1, int y0, int y1, int device, octvr_mapper** out)
{
    return guard([&] {
        OB_CHECK(t && in_sizes_wh && out, "null argument");
        int count = 0;
        cudaError_t e = cudaGetDeviceCount(&count);
        if (e != cudaSuccess || count <= 0 || device < 0 || device >= count)
            fail(OCTVR_ERR_CUDA, "no usable CUDA device (the stitch path has no CPU fallback)");
        OB_CUDA(cudaSetDevice(device));
        std::unique_ptr<octvr_mapper> m(new octvr_mapper);
        m->device = device;
        build_mapper(*m, *t, in_sizes_wh, n_in, blend, enable_gain != 0, 0, 0, y0, y1, x0, x1);
        *out = m.release();
    });
}

octvr_status octvr_mapper_stitch(octvr_mapper* m, const octvr_frame* d_inputs, int n_inputs, const octvr_frame* d_output,
                                 uint8_t* d_preview_rgb, size_t preview_pitch, int preview_w, int preview_h,
                                 const double* gains, int n_gains, void* stream)
{
    return guard([&] {
        OB_CHECK(m && d_inputs, "null argument");
        do_stitch(*m, d_inputs, n_inputs, d_output, gains, n_gains, (cudaStream_t)stream, d_preview_rgb, preview_pitch, preview_w, preview_h);
    });
}

octvr_status octvr_mapper_stitch_packed(octvr_mapper* m, const uint8_t* const* d_inputs, const size_t* in_pitch, int n_inputs,
                                        uint8_t* d_output, size_t out_pitch, const double* gains, int n_gains, void* stream)
{
    return guard([&] {
        OB_CHECK(m && d_inputs && in_pitch && d_output, "null argument");
        OB_CHECK(n_inputs == m->n + m->n_ov, "wrong number of input frames");
        std::vector<octvr_frame> f(n_inputs);
        for (int i = 0; i < n_inputs; i++) {              // mapper.cpp:222-226
            uint8_t* b = const_cast<uint8_t*>(d_inputs[i]);
            f[i].y = b; f[i].u = b + (size_t)m->in_h[i] * in_pitch[i]; f[i].v = f[i].u + m->win_w[i] / 2;
            f[i].y_pitch = f[i].u_pitch = f[i].v_pitch = in_pitch[i]; f[i].uv_pixel_stride = 1;
        }
        octvr_frame o;                                     // mapper.cpp:296-302
        o.y = d_output; o.u = d_output + (size_t)m->scaled_h * out_pitch; o.v = o.u + m->scaled_w / 2;
        o.y_pitch = o.u_pitch = o.v_pitch = out_pitch; o.uv_pixel_stride = 1;
        do_stitch(*m, f.data(), n_inputs, &o, gains, n_gains, (cudaStream_t)stream);
    });
}

octvr_status octvr_mapper_set_keep_rgb(octvr_mapper* m, int on)
{
    return guard([&] { OB_CHECK(m, "null argument"); m->keep_rgb = on != 0; });
}

octvr_status octvr_mapper_result_rgb(octvr_mapper* m, uint8_t* h_rgb, size_t pitch)
{
    return guard([&] {
        OB_CHECK(m && h_rgb && pitch >= (size_t)m->out_w * 3, "bad argument");
        OB_CHECK(m->d_rgb, "no RGB result: call octvr_mapper_set_keep_rgb(m, 1) before stitching");
        OB_CUDA(cudaSetDevice(m->device));
        if (m->last_stream_valid) OB_CUDA(cudaStreamSynchronize(m->last_stream));
        OB_CUDA(cudaMemcpy2D(h_rgb, pitch, m->d_rgb, (size_t)m->out_w * 3, (size_t)m->out_w * 3, m->out_h, cudaMemcpyDeviceToHost));
    });
}

octvr_status octvr_mapper_gains(octvr_mapper* m, double* out, int n)
{
    return guard([&] {
        OB_CHECK(m && out, "null argument");
        if (!m->gain) { for (int i = 0; i < n; i++) out[i] = 1.0; return; }   // mapper.hpp:85-87 returns {} ; we report unity
        OB_CHECK(n == m->n, "gains size must equal the number of inputs");
        OB_CUDA(cudaSetDevice(m->device));
        if (m->last_stream_valid) OB_CUDA(cudaStreamSynchronize(m->last_stream));
        OB_CUDA(cudaMemcpy(out, m->d_gains, sizeof(double) * n, cudaMemcpyDeviceToHost));
    });
}

octvr_status octvr_mapper_stats(const octvr_mapper* m, int64_t* pairs, int64_t* roi_area, int64_t* table_bytes, int* launches)
{
    return guard([&] {
        OB_CHECK(m, "null argument");
        if (pairs) *pairs = m->pairs;
        if (roi_area) *roi_area = m->roi_area;
        if (table_bytes) *table_bytes = m->table_bytes;
        if (launches) *launches = m->mb ? ob::multiband_launches(*m) + 1 : m->fused ? 1 + (m->gain ? 1 : 0) : 2;
    });
}

octvr_status octvr_mapper_source_rows(const octvr_mapper* m, int* rows_lo_hi, int n)
{
    return guard([&] {
        OB_CHECK(m && rows_lo_hi && n == m->n, "bad argument");
        for (int i = 0; i < n; i++) { rows_lo_hi[2 * i] = m->src_row0[i]; rows_lo_hi[2 * i + 1] = m->src_row1[i]; }
    });
}

octvr_status octvr_mapper_set_profiling(octvr_mapper* m, int on)
{
    return guard([&] { OB_CHECK(m, "null argument"); m->profiling = on != 0; });
}

octvr_status octvr_mapper_source_cols(const octvr_mapper* m, int* cols_lo_hi, int n)
{
    return guard([&] {
        OB_CHECK(m && cols_lo_hi && n == m->n, "bad argument");
        for (int i = 0; i < n; i++) {      // what the tables read (that is converted) and what the gain samples read
            int lo = m->src_col0[i], hi = m->src_col1[i];
            if (m->gain && m->gain_col0[i] <= m->gain_col1[i]) {
                if (lo >= hi) { lo = m->gain_col0[i]; hi = m->gain_col1[i]; }
                else { lo = std::min(lo, m->gain_col0[i]); hi = std::max(hi, m->gain_col1[i]); }
            }
            cols_lo_hi[2 * i] = lo; cols_lo_hi[2 * i + 1] = hi;
        }
    });
}

octvr_status octvr_mapper_set_input_window(octvr_mapper* m, int cam, int col0, int width)
{
    return guard([&] {
        OB_CHECK(m && cam >= 0 && cam < m->n, "bad argument");
        OB_CHECK(!m->fused, "input windows are not available on the single-kernel path (OCTVR_BLEND=fused)");
        OB_CHECK(col0 >= 0 && width > 0 && col0 + width <= m->in_w[cam] && col0 % 8 == 0 && width % 8 == 0, "window must lie inside the frame, multiples of 8");
        int need[2 * MAX_CAMS];
        octvr_mapper_source_cols(m, need, m->n);
        if (need[2 * cam] < need[2 * cam + 1])
            OB_CHECK(col0 <= need[2 * cam] && need[2 * cam + 1] <= col0 + width, "window does not cover the source columns this mapper reads (octvr_mapper_source_cols)");
        m->win_col0[cam] = col0; m->win_w[cam] = width;
    });
}

octvr_status octvr_mapper_stage_ms(octvr_mapper* m, const char* stage, float* ms)
{
    return guard([&] {
        OB_CHECK(m && stage && ms && m->profiling && m->last_stream_valid, "profiling not enabled or nothing stitched");
        OB_CUDA(cudaSetDevice(m->device));
        OB_CUDA(cudaEventSynchronize(m->ev[3]));
        std::string s(stage);
        int a = 0, b = 3;
        if (s == "convert") { a = 0; b = 1; } else if (s == "gain") { a = 1; b = 2; } else if (s == "blend") { a = 2; b = 3; }
        else if (s != "total") fail(OCTVR_ERR_INVALID, "unknown stage " + s);
        OB_CUDA(cudaEventElapsedTime(ms, m->ev[a], m->ev[b]));
    });
}

octvr_status octvr_mapper_debug_gain_ns(octvr_mapper* m, unsigned long long* out6)
{
    // diagnostics only
    return guard([&] {
        OB_CHECK(m && out6 && m->d_dbg, "no gain stage");
        OB_CUDA(cudaDeviceSynchronize());
        OB_CUDA(cudaMemcpy(out6, m->d_dbg, 6 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    });
}

octvr_status octvr_mapper_debug_gain_trace(octvr_mapper* m, unsigned long long* out, int n)
{
    // diagnostics only: the first n (<= 8 + 2 * 4096) words of the gain kernel's stamp buffer; [6] is reset to "never"
    return guard([&] {
        OB_CHECK(m && out && m->d_dbg && n > 0 && n <= 8 + 2 * 4096, "no gain stage");
        OB_CUDA(cudaDeviceSynchronize());
        OB_CUDA(cudaMemcpy(out, m->d_dbg, (size_t)n * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
        const unsigned long long never = ~0ull, zero = 0ull;
        OB_CUDA(cudaMemcpy(m->d_dbg + 6, &never, 8, cudaMemcpyHostToDevice));
        OB_CUDA(cudaMemcpy(m->d_dbg + 7, &zero, 8, cudaMemcpyHostToDevice));
    });
}

octvr_status octvr_mapper_debug_ring(octvr_mapper* m, unsigned long long* out8)
{
    // diagnostics only: counters of K_blend_ring when the library is built with -DRING_DEBUG=1 (zeroed by this call)
    return guard([&] {
        OB_CHECK(m && out8 && m->d_dbg_ring, "no ring kernel");
        OB_CUDA(cudaDeviceSynchronize());
        OB_CUDA(cudaMemcpy(out8, m->d_dbg_ring, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
        OB_CUDA(cudaMemset(m->d_dbg_ring, 0, 8 * sizeof(unsigned long long)));
    });
}

// ---- device buffers shared between the processes of one node (row-band mode: every rank stores its band of the frame
//      straight into the collecting rank's buffer over NVLink peer memory; no collection step) ----
octvr_status octvr_shared_alloc(size_t bytes, int device, void** d_ptr, unsigned char handle64[64])
{
    return guard([&] {
        OB_CHECK(bytes > 0 && d_ptr && handle64, "bad argument");
        static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
        OB_CUDA(cudaSetDevice(device));
        void* p = nullptr;
        OB_CUDA(cudaMalloc(&p, bytes));
        cudaIpcMemHandle_t h;
        cudaError_t e = cudaIpcGetMemHandle(&h, p);
        if (e != cudaSuccess) { cudaFree(p); fail(OCTVR_ERR_CUDA, std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(e)); }
        OB_CUDA(cudaMemset(p, 0, bytes));
        OB_CUDA(cudaDeviceSynchronize());
        memcpy(handle64, &h, 64);
        *d_ptr = p;
    });
}

octvr_status octvr_shared_open(const unsigned char handle64[64], int device, void** d_ptr)
{
    return guard([&] {
        OB_CHECK(handle64 && d_ptr, "bad argument");
        OB_CUDA(cudaSetDevice(device));
        cudaIpcMemHandle_t h;
        memcpy(&h, handle64, 64);
        OB_CUDA(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    });
}

octvr_status octvr_shared_close(void* d_ptr, int opened)
{
    return guard([&] {
        if (!d_ptr) return;
        if (opened) OB_CUDA(cudaIpcCloseMemHandle(d_ptr)); else OB_CUDA(cudaFree(d_ptr));
    });
}

void octvr_mapper_destroy(octvr_mapper* m) { delete m; }

}  // extern "C"
