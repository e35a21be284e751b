// csrc/pack.cu -- the feather / no-blend tables of a Mapper packed ON THE DEVICE (north_star subsystem 1; the reference
// builds its tables on the CPU: blenders.cpp:531-572 feather weights, mapper.cpp:84-127 maps).
//
// Input: the template's per-camera normalised maps and masks, uploaded once.  Kernels:
//   k_pack_quantise   cv::remap's 1/32-px fixed point of fl32(map * size) (template.cpp:175-176, imgwarp.cpp:4383-4442)
//   k_chamfer         (seam.cu) cv::distanceTransform(mask, DIST_L2, 3), one CTA per camera, row-parallel min-plus scans
//   k_pack_weights    FeatherGPUBlender's W_i = N max(DT_i - border, 0) / (1e-5 + sum) in camera order (blenders.cpp:531-572),
//                     or weight 1 for the last covering camera (blend == 0, mapper.cpp:269-275)
//   k_pack_used       which (tile, camera) pairs have a contributing pixel -> the job list (prefix sums on the host: a few
//                     hundred thousand flags)
//   k_pack_boxes      per job the bounding box of its bilinear taps in the camera's source plane (the TMA box of K_blend_ring)
//   k_pack_entries    the 8-byte entries of K_blend_ring in that kernel's thread order
// The host keeps what is tiny and order-dependent: size classes of the boxes, the (camera, box size) -> tensor-map index
// and the 16-byte job records.  Byte-identical to the host packer in mapper.cpp (which stays for the layouts that are not the
// default: fused, staged, direct; OCTVR_PACK=host forces it; tests compare the two).
#include "mapper.h"
#include "prep.h"
#include <climits>
#include <cmath>
#include <cstring>
#include <map>
#include <memory>

namespace ob {
namespace {

struct PackCam {
    const float* map1; const float* map2; const uint8_t* mask;
    int2* sxy;                  // fixed-point source position per ROI pixel
    float* w;                   // chamfer distance in, blend weight out
    int rx, ry, rw, rh;         // ROI in the output frame
    int src_w, src_h;
};
struct PackParams {
    PackCam cam[MAX_CAMS];
    int n, out_w, out_h, tiles_x, tiles_y, band_y0, band_y1;
    int Rx, Ry, Rw, Rh;         // union of the ROIs
    int border; float scale;    // feather: border > 0; no blend: border == 0
    float shift;                // texel-centre convention: 0.5, else 0 (prep.h quantise_map)
};

__global__ void __launch_bounds__(256) k_pack_quantise(const __grid_constant__ PackParams p, int c)
{
    const PackCam& k = p.cam[c];
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= (size_t)k.rw * k.rh) return;
    const float fw = (float)(double)k.src_w, fh = (float)(double)k.src_h;
    float px = __fadd_rn(__fmul_rn(k.map1[i], fw), 0.f), py = __fadd_rn(__fmul_rn(k.map2[i], fh), 0.f);     // Mat * double -> f32 convertTo
    if (p.shift != 0.f) { px = __fsub_rn(px, p.shift); py = __fsub_rn(py, p.shift); }
    k.sxy[i] = make_int2(__float2int_rn(__fmul_rn(px, 32.f)), __float2int_rn(__fmul_rn(py, 32.f)));
}

__global__ void __launch_bounds__(256) k_pack_weights(const __grid_constant__ PackParams p)
{
    const int x = p.Rx + blockIdx.x * 32 + threadIdx.x, y = p.Ry + blockIdx.y * 8 + threadIdx.y;
    if (x >= p.Rx + p.Rw || y >= p.Ry + p.Rh) return;
    if (p.border > 0) {
        float sum = 1e-5f;
        for (int c = 0; c < p.n; c++) {
            const PackCam& k = p.cam[c];
            const int lx = x - k.rx, ly = y - k.ry;
            if (lx < 0 || ly < 0 || lx >= k.rw || ly >= k.rh) continue;
            const float t = __fsub_rn(k.w[(size_t)ly * k.rw + lx], (float)p.border);
            sum = __fadd_rn(t > 0.f ? t : 0.f, sum);
        }
        for (int c = 0; c < p.n; c++) {
            const PackCam& k = p.cam[c];
            const int lx = x - k.rx, ly = y - k.ry;
            if (lx < 0 || ly < 0 || lx >= k.rw || ly >= k.rh) continue;
            float t = __fsub_rn(k.w[(size_t)ly * k.rw + lx], (float)p.border);
            t = t > 0.f ? t : 0.f;
            k.w[(size_t)ly * k.rw + lx] = sum != 0.f ? __fdiv_rn(__fmul_rn(p.scale, t), sum) : 0.f;
        }
    } else {
        int owner = -1;
        for (int c = 0; c < p.n; c++) {
            const PackCam& k = p.cam[c];
            const int lx = x - k.rx, ly = y - k.ry;
            if (lx < 0 || ly < 0 || lx >= k.rw || ly >= k.rh) continue;
            if (k.mask[(size_t)ly * k.rw + lx]) owner = c;
        }
        for (int c = 0; c < p.n; c++) {
            const PackCam& k = p.cam[c];
            const int lx = x - k.rx, ly = y - k.ry;
            if (lx < 0 || ly < 0 || lx >= k.rw || ly >= k.rh) continue;
            k.w[(size_t)ly * k.rw + lx] = c == owner ? 1.f : 0.f;
        }
    }
}

// pixel p (0..511) of tile tl for camera c: ROI-local position, or false when outside the ROI / frame
__device__ __forceinline__ bool pack_pixel(const PackParams& p, const PackCam& k, int tl, int px, int& lx, int& ly)
{
    const int tx = tl % p.tiles_x, ty = tl / p.tiles_x;
    const int gx = tx * TILE_W + (px & (TILE_W - 1)), gy = ty * TILE_H + (px / TILE_W);
    lx = gx - k.rx; ly = gy - k.ry;
    return lx >= 0 && ly >= 0 && lx < k.rw && ly < k.rh && gx < p.out_w && gy < p.out_h;
}

// used[tile * n + cam] = some pixel of the tile (inside the row band) has mask != 0 and weight != 0
__global__ void __launch_bounds__(256) k_pack_used(const __grid_constant__ PackParams p, uint8_t* used)
{
    const int tl = blockIdx.x;
    for (int c = 0; c < p.n; c++) {
        const PackCam& k = p.cam[c];
        int any = 0;
        #pragma unroll
        for (int h = 0; h < 2; h++) {
            int lx, ly;
            const int px = threadIdx.x + h * 256;
            if (!pack_pixel(p, k, tl, px, lx, ly)) continue;
            const int gy = ly + k.ry;
            if (gy < p.band_y0 || gy >= p.band_y1) continue;
            const size_t o = (size_t)ly * k.rw + lx;
            any |= (k.mask[o] != 0 && k.w[o] != 0.f);
        }
        any = __syncthreads_or(any);
        if (threadIdx.x == 0) used[(size_t)tl * p.n + c] = (uint8_t)(any != 0);
    }
}

// validity and integer tap position of one table entry (make_entry in mapper.cpp)
__device__ __forceinline__ bool pack_entry(const PackCam& k, size_t o, int& ix, int& iy, int& fx, int& fy)
{
    if (!(k.mask[o] != 0 && k.w[o] != 0.f)) return false;
    const int2 s = k.sxy[o];
    ix = min(32767, max(-32768, s.x >> 5)); iy = min(32767, max(-32768, s.y >> 5));
    fx = s.x & 31; fy = s.y & 31;
    const bool x0 = ix >= 0 && ix < k.src_w, x1 = ix + 1 >= 0 && ix + 1 < k.src_w;
    const bool y0 = iy >= 0 && iy < k.src_h, y1 = iy + 1 >= 0 && iy + 1 < k.src_h;
    return (x0 || x1) && (y0 || y1);                       // at least one tap inside
}

// per job {xmin, xmax, ymin, ymax} over its valid entries (xmax, ymax include the +1 tap); empty job: xmin > xmax
__global__ void __launch_bounds__(128) k_pack_boxes(const __grid_constant__ PackParams p, const uint32_t* job_tile, const uint8_t* job_cam, int4* boxes)
{
    __shared__ int s[4];
    const int j = blockIdx.x, tl = job_tile[j];
    const PackCam& k = p.cam[job_cam[j]];
    if (threadIdx.x == 0) { s[0] = INT_MAX; s[1] = INT_MIN; s[2] = INT_MAX; s[3] = INT_MIN; }
    __syncthreads();
    int xmin = INT_MAX, xmax = INT_MIN, ymin = INT_MAX, ymax = INT_MIN;
    #pragma unroll
    for (int h = 0; h < 4; h++) {
        int lx, ly, ix, iy, fx, fy;
        if (!pack_pixel(p, k, tl, threadIdx.x + h * 128, lx, ly)) continue;
        if (!pack_entry(k, (size_t)ly * k.rw + lx, ix, iy, fx, fy)) continue;
        xmin = min(xmin, ix); xmax = max(xmax, ix + 1); ymin = min(ymin, iy); ymax = max(ymax, iy + 1);
    }
    xmin = __reduce_min_sync(0xffffffffu, xmin); xmax = __reduce_max_sync(0xffffffffu, xmax);
    ymin = __reduce_min_sync(0xffffffffu, ymin); ymax = __reduce_max_sync(0xffffffffu, ymax);
    if ((threadIdx.x & 31) == 0) { atomicMin(&s[0], xmin); atomicMax(&s[1], xmax); atomicMin(&s[2], ymin); atomicMax(&s[3], ymax); }
    __syncthreads();
    if (threadIdx.x == 0) boxes[j] = make_int4(s[0], s[1], s[2], s[3]);
}

// K_blend_ring's entries: pixel (col, row) of the tile belongs to thread (row & 7) >> 1 << 5 | col, slot q = (row & 1) | (row >> 3) << 1;
// a thread's four slots are two uint4 {code_q, code_q+1, weight_q, weight_q+1} at [(q >> 1) * 128 + thread]
__global__ void __launch_bounds__(128) k_pack_entries(const __grid_constant__ PackParams p, const uint32_t* job_tile, const uint8_t* job_cam,
                                                      const int4* job_box /* bx0, by0, bw, - */, uint4* entries)
{
    const int j = blockIdx.x, tl = job_tile[j], tid = threadIdx.x;
    const PackCam& k = p.cam[job_cam[j]];
    const int4 b = job_box[j];
    const int col = tid & 31, rp = tid >> 5;
    uint32_t code[4], wb[4];
    #pragma unroll
    for (int q = 0; q < 4; q++) {
        const int row = 2 * rp + (q & 1) + 8 * (q >> 1);
        int lx, ly, ix, iy, fx, fy;
        code[q] = 0u; wb[q] = 0u;
        if (!pack_pixel(p, k, tl, row * TILE_W + col, lx, ly)) continue;
        const size_t o = (size_t)ly * k.rw + lx;
        if (!pack_entry(k, o, ix, iy, fx, fy)) continue;
        const uint32_t off = (uint32_t)(((iy - b.y) * b.z + (ix - b.x)) * 4);
        code[q] = (off << 16) | ((uint32_t)fy << 8) | (uint32_t)fx;
        wb[q] = __float_as_uint(k.w[o]);
    }
    uint4* e = entries + (size_t)j * (TILE_PX / 2);
    e[tid] = make_uint4(code[0], code[1], wb[0], wb[1]);
    e[128 + tid] = make_uint4(code[2], code[3], wb[2], wb[3]);
}

template <class T> struct DBuf {
    T* p = nullptr;
    DBuf() {}
    explicit DBuf(size_t n) { OB_CUDA(cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T))); }
    DBuf(const T* h, size_t n) : DBuf(n) { if (n) OB_CUDA(cudaMemcpy(p, h, n * sizeof(T), cudaMemcpyHostToDevice)); }
    ~DBuf() { cudaFree(p); }
    DBuf(const DBuf&) = delete;
    T* release() { T* q = p; p = nullptr; return q; }
};

}  // namespace

bool pack_ring_gpu(octvr_mapper& m, const octvr_template& t, int blend)
{
    const int n = m.n;
    if (blend > 0 || n < 1) return false;
    if (const char* e = getenv("OCTVR_PACK")) if (std::string(e) == "host") return false;
    if (const char* e = getenv("OCTVR_BLEND")) {                   // fused / staged / direct layouts: host packer
        const std::string v(e);
        if (v == "fused" || v == "staged" || v == "direct") return false;
    }
    for (int i = 0; i < n; i++) {
        if (m.in_w[i] % 4 != 0) return false;
        if (blend < 0 && t.inputs[i].roi.w > 8192) return false;   // k_chamfer scans a row with one CTA
    }
    InitTrace tr("pack_ring_gpu");
    OB_CUDA(cudaSetDevice(m.device));
    PackParams p;
    memset(&p, 0, sizeof(p));
    p.n = n; p.out_w = t.out_w; p.out_h = t.out_h;
    p.tiles_x = (t.out_w + TILE_W - 1) / TILE_W; p.tiles_y = (t.out_h + TILE_H - 1) / TILE_H;
    p.band_y0 = m.band_y0; p.band_y1 = m.band_y1;
    p.border = blend < 0 ? -blend : 0; p.scale = (float)n; p.shift = m.texel_shift;
    const int ntiles = p.tiles_x * p.tiles_y;
    std::vector<std::unique_ptr<DBuf<float>>> d_m1(n), d_m2(n), d_w(n);
    std::vector<std::unique_ptr<DBuf<uint8_t>>> d_mask(n);
    std::vector<std::unique_ptr<DBuf<int2>>> d_sxy(n);
    Rect R = t.inputs[0].roi;
    for (int i = 0; i < n; i++) {
        const TInput& in = t.inputs[i];
        const size_t px = (size_t)in.roi.w * in.roi.h;
        d_m1[i].reset(new DBuf<float>(in.map1.d.data(), px));
        d_m2[i].reset(new DBuf<float>(in.map2.d.data(), px));
        d_mask[i].reset(new DBuf<uint8_t>(in.mask.d.data(), px));
        d_w[i].reset(new DBuf<float>(px));
        d_sxy[i].reset(new DBuf<int2>(px));
        p.cam[i] = PackCam{ d_m1[i]->p, d_m2[i]->p, d_mask[i]->p, d_sxy[i]->p, d_w[i]->p, in.roi.x, in.roi.y, in.roi.w, in.roi.h, m.in_w[i], m.in_h[i] };
        R = rect_union(R, in.roi);
    }
    p.Rx = R.x; p.Ry = R.y; p.Rw = R.w; p.Rh = R.h;
    tr.lap("upload maps + masks");
    for (int i = 0; i < n; i++) {
        const size_t px = (size_t)t.inputs[i].roi.w * t.inputs[i].roi.h;
        k_pack_quantise<<<(unsigned)((px + 255) / 256), 256>>>(p, i);
    }
    OB_CUDA(cudaGetLastError());
    if (blend < 0) {
        std::vector<const uint8_t*> masks(n); std::vector<float*> dist(n); std::vector<int> ws(n), hs(n);
        for (int i = 0; i < n; i++) { masks[i] = d_mask[i]->p; dist[i] = d_w[i]->p; ws[i] = t.inputs[i].roi.w; hs[i] = t.inputs[i].roi.h; }
        chamfer_l2_gpu_batch(masks.data(), ws.data(), hs.data(), dist.data(), n);
    }
    k_pack_weights<<<dim3((R.w + 31) / 32, (R.h + 7) / 8), dim3(32, 8)>>>(p);
    OB_CUDA(cudaGetLastError());
    DBuf<uint8_t> d_used((size_t)ntiles * n);
    k_pack_used<<<ntiles, 256>>>(p, d_used.p);
    OB_CUDA(cudaGetLastError());
    std::vector<uint8_t> used((size_t)ntiles * n);
    OB_CUDA(cudaMemcpy(used.data(), d_used.p, used.size(), cudaMemcpyDeviceToHost));
    tr.lap("quantise + chamfer + weights + used");
    std::vector<uint32_t> job_start(ntiles + 1, 0), job_tile;
    std::vector<uint8_t> job_cam;
    for (int tl = 0; tl < ntiles; tl++) {
        job_start[tl] = (uint32_t)job_cam.size();
        for (int i = 0; i < n; i++) if (used[(size_t)tl * n + i]) { job_cam.push_back((uint8_t)i); job_tile.push_back((uint32_t)tl); }
    }
    job_start[ntiles] = (uint32_t)job_cam.size();
    const size_t njobs = job_cam.size();
    if (njobs == 0 || njobs >= ((size_t)1 << 20) || njobs * TILE_PX >= ((size_t)1 << 32)) return false;
    DBuf<uint32_t> d_job_tile(job_tile.data(), njobs);
    DBuf<uint8_t> d_job_cam(job_cam.data(), njobs);
    DBuf<int4> d_boxes(njobs);
    k_pack_boxes<<<(unsigned)njobs, 128>>>(p, d_job_tile.p, d_job_cam.p, d_boxes.p);
    OB_CUDA(cudaGetLastError());
    std::vector<int4> boxes(njobs);
    OB_CUDA(cudaMemcpy(boxes.data(), d_boxes.p, njobs * sizeof(int4), cudaMemcpyDeviceToHost));
    tr.lap("job list + boxes");
    // box size classes, tensor-map slots and job records: the rules of the host packer (mapper.cpp)
    std::map<uint64_t, int> tmap_index;
    std::vector<uint4> recs((size_t)ntiles * n, make_uint4(0u, 0u, 0u, 0u));
    auto size_class = [](int v) { return v <= 64 ? (v + 7) / 8 * 8 : v <= 128 ? (v + 15) / 16 * 16 : (v + 31) / 32 * 32; };
    std::vector<int> col_lo(n, INT_MAX), col_hi(n, INT_MIN);      // source columns the taps of each camera touch
    for (int tl = 0; tl < ntiles; tl++)
        for (uint32_t j = job_start[tl]; j < job_start[tl + 1]; j++) {
            int xmin = boxes[j].x, xmax = boxes[j].y, ymin = boxes[j].z, ymax = boxes[j].w;
            if (xmin > xmax) { xmin = xmax = ymin = ymax = 0; }
            const int i = job_cam[j];
            if (boxes[j].x <= boxes[j].y) { col_lo[i] = std::min(col_lo[i], xmin); col_hi[i] = std::max(col_hi[i], xmax); }
            const int bx0 = (int)std::floor(xmin / 4.0) * 4, by0 = ymin;
            const int bw = size_class(xmax - bx0 + 1), bh = size_class(ymax - ymin + 1);
            if (bw > 256 || bh > 256 || (int64_t)bw * bh > STAGE_CAP) return false;
            if ((((size_t)bw * bh * 4 + 127) & ~(size_t)127) > (size_t)RING_BYTES) return false;
            if (!(bx0 >= -32768 && bx0 < 32768 && by0 >= -32768 && by0 < 32768)) return false;
            const uint64_t key = ((uint64_t)i << 32) | ((uint64_t)bw << 16) | (uint64_t)bh;
            auto it = tmap_index.find(key);
            if (it == tmap_index.end()) it = tmap_index.emplace(key, (int)tmap_index.size()).first;
            if (it->second >= 65536) return false;
            recs[(size_t)tl * n + (j - job_start[tl])] = make_uint4(((uint32_t)bx0 & 0xFFFFu) | ((uint32_t)by0 << 16), (uint32_t)it->second | ((uint32_t)i << 16),
                                                                    (uint32_t)(bw * bh * 4), (uint32_t)bw | (j << 12));
            boxes[j] = make_int4(bx0, by0, bw, bh);
        }
    const int ring_ctas = ring_ctas_per_sm();
    if (ring_ctas <= 0) return false;
    OB_CUDA(cudaMemcpy(d_boxes.p, boxes.data(), njobs * sizeof(int4), cudaMemcpyHostToDevice));
    DBuf<uint4> d_entries(njobs * (TILE_PX / 2));
    k_pack_entries<<<(unsigned)njobs, 128>>>(p, d_job_tile.p, d_job_cam.p, d_boxes.p, d_entries.p);
    OB_CUDA(cudaGetLastError());
    OB_CUDA(cudaDeviceSynchronize());
    tr.lap("records + entries");
    // ---- commit
    m.tiles_x = p.tiles_x; m.tiles_y = p.tiles_y; m.njobs = njobs;
    m.inv_n = blend < 0 ? (float)(1.0 / n) : 1.f;
    m.staged = true; m.ring = true; m.ring_ctas = ring_ctas;
    OB_CUDA(cudaDeviceGetAttribute(&m.sm_count, cudaDevAttrMultiProcessorCount, m.device));
    OB_CHECK((int)m.d_rgbx.size() >= n, "RGBX planes must exist before the tables are packed");
    std::vector<uint8_t> tmaps(tmap_index.size() * 128);
    for (auto& kv : tmap_index) {
        const int cam = (int)(kv.first >> 32), bw = (int)((kv.first >> 16) & 0xFFFF), bh = (int)(kv.first & 0xFFFF);
        encode_rgbx_tensor_map(tmaps.data() + (size_t)kv.second * 128, m.d_rgbx[cam], m.in_w[cam], m.in_h[cam], bw, bh);
    }
    DBuf<uint8_t> d_tm(tmaps.size() + 128);        // cudaMalloc: 256-byte aligned (a CUtensorMap needs 64)
    OB_CUDA(cudaMemcpy(d_tm.p, tmaps.data(), tmaps.size(), cudaMemcpyHostToDevice));
    m.d_tmaps = d_tm.release(); m.n_tmaps = (int)tmap_index.size();
    DBuf<uint32_t> d_js(job_start.data(), job_start.size());
    m.d_tile_job_start = d_js.release();
    DBuf<uint4> d_recs(recs.data(), recs.size());
    m.d_rjobs = d_recs.release();
    m.d_rentries = d_entries.release();
    OB_CUDA(cudaMalloc(&m.d_ring_counter, sizeof(unsigned int))); OB_CUDA(cudaMemset(m.d_ring_counter, 0, sizeof(unsigned int)));
    OB_CUDA(cudaMalloc(&m.d_dbg_ring, 8 * sizeof(unsigned long long))); OB_CUDA(cudaMemset(m.d_dbg_ring, 0, 8 * sizeof(unsigned long long)));
    m.table_bytes = (int64_t)(njobs * (TILE_PX / 2) * sizeof(uint4) + recs.size() * 16);
    for (int i = 0; i < n; i++) {
        if (col_lo[i] > col_hi[i]) { m.src_col0[i] = m.src_col1[i] = 0; continue; }
        m.src_col0[i] = std::max(0, col_lo[i]) & ~7; m.src_col1[i] = std::min(m.in_w[i], col_hi[i] + 1);
    }
    tr.lap("commit");
    return true;
}

}  // namespace ob
