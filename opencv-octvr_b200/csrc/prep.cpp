// csrc/prep.cpp -- init-time host preparation.  See prep.h for the reference lines each routine follows.
#include "prep.h"
#include <algorithm>
#include <climits>
#include <cstring>
#include <zlib.h>
#undef FAR
#include <cmath>
#include <thread>

#include <chrono>
#include <cstdlib>
#include <new>
#include <sys/mman.h>

namespace ob {

void* big_alloc(size_t bytes)
{
    constexpr size_t HUGE = (size_t)2 << 20;
    if (bytes < 2 * HUGE) { void* p = malloc(bytes ? bytes : 1); if (!p) throw std::bad_alloc(); return p; }
    void* p = aligned_alloc(HUGE, (bytes + HUGE - 1) & ~(HUGE - 1));
    if (!p) throw std::bad_alloc();
    madvise(p, bytes, MADV_HUGEPAGE);          // advisory: ignored where transparent huge pages are disabled
    return p;
}
void big_free(void* p) { free(p); }

double InitTrace::now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
InitTrace::InitTrace(const char* w) : what(w), on(getenv("OCTVR_INIT_TRACE") != nullptr), t0(on ? now() : 0.) {}
void InitTrace::lap(const char* stage)
{
    if (!on) return;
    const double t = now();
    fprintf(stderr, "[octvr init] %s: %s %.1f ms\n", what, stage, (t - t0) * 1e3);
    t0 = t;
}

namespace {
inline int round_he(float v) { return (int)lrintf(v); }        // cvRound: round-half-even
inline short to_short(int v) { return (short)std::min(32767, std::max(-32768, v)); }

template <class F> void parallel_rows(int n, F&& body)
{
    unsigned hw = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    if (n < 64 || hw == 1) { body(0, n); return; }
    std::vector<std::thread> th;
    int chunk = (n + (int)hw - 1) / (int)hw;
    for (unsigned t = 0; t < hw; t++) {
        int a = (int)t * chunk, b = std::min(n, a + chunk);
        if (a >= b) break;
        th.emplace_back([=, &body] { body(a, b); });
    }
    for (auto& t : th) t.join();
}

// resize coefficient setup shared by the u8 and f32 variants (imgwarp.cpp:3387-3447)
struct Axis { std::vector<int> ofs; std::vector<float> frac; };
Axis linear_axis(int src, int dst, bool clamp_ofs)
{
    Axis a;
    a.ofs.resize(dst); a.frac.resize(dst);
    const double scale = 1. / ((double)dst / src);
    for (int d = 0; d < dst; d++) {
        float f = (float)((d + 0.5) * scale - 0.5);
        int s = (int)std::floor(f);
        f -= s;
        if (clamp_ofs) {
            if (s < 0) { f = 0; s = 0; }
            if (s >= src - 1) { f = 0; s = src - 1; }
        }
        a.ofs[d] = s; a.frac[d] = f;
    }
    return a;
}
inline int clampi(int v, int lo, int hi) { return v < lo ? lo : v > hi ? hi : v; }
}  // namespace

Img<float> chamfer_l2(const Img<uint8_t>& mask)
{
    const int w = mask.w, h = mask.h;
    const int HV = (int)lrint(0.955f * 65536), DIAG = (int)lrint(1.3693f * 65536), FAR = INT_MAX >> 2;
    const int step = w + 2;
    std::vector<int> buf((size_t)step * (h + 2), FAR);
    auto at = [&](int y, int x) -> int& { return buf[(size_t)(y + 1) * step + (x + 1)]; };
    for (int y = 0; y < h; y++) {                 // forward raster: NW, N, NE, W
        const uint8_t* m = mask.row(y);
        for (int x = 0; x < w; x++) {
            if (!m[x]) { at(y, x) = 0; continue; }
            int best = at(y - 1, x - 1) + DIAG;
            best = std::min(best, at(y - 1, x) + HV);
            best = std::min(best, at(y - 1, x + 1) + DIAG);
            best = std::min(best, at(y, x - 1) + HV);
            at(y, x) = best;
        }
    }
    Img<float> out(w, h);
    const float scale = 1.f / 65536;
    for (int y = h - 1; y >= 0; y--) {            // backward raster: SE, S, SW, E (only where > HV)
        float* o = out.row(y);
        for (int x = w - 1; x >= 0; x--) {
            int t0 = at(y, x);
            if (t0 > HV) {
                t0 = std::min(t0, at(y + 1, x + 1) + DIAG);
                t0 = std::min(t0, at(y + 1, x) + HV);
                t0 = std::min(t0, at(y + 1, x - 1) + DIAG);
                t0 = std::min(t0, at(y, x + 1) + HV);
                at(y, x) = t0;
            }
            o[x] = (float)(t0 * scale);
        }
    }
    return out;
}

Img<uint8_t> resize_linear(const Img<uint8_t>& src, int dw, int dh)
{
    OB_CHECK(!src.empty() && dw > 0 && dh > 0, "resize: empty");
    Axis ax = linear_axis(src.w, dw, true), ay = linear_axis(src.h, dh, false);
    std::vector<short> ca(2 * (size_t)dw);
    for (int d = 0; d < dw; d++) {
        ca[2 * d] = to_short(round_he((1.f - ax.frac[d]) * 2048));
        ca[2 * d + 1] = to_short(round_he(ax.frac[d] * 2048));
    }
    Img<uint8_t> dst(dw, dh);
    parallel_rows(dh, [&](int y0, int y1) {
        std::vector<int> top(dw), bot(dw);
        for (int y = y0; y < y1; y++) {
            const uint8_t* s0 = src.row(clampi(ay.ofs[y], 0, src.h - 1));
            const uint8_t* s1 = src.row(clampi(ay.ofs[y] + 1, 0, src.h - 1));
            for (int x = 0; x < dw; x++) {
                int o = ax.ofs[x], o1 = std::min(o + 1, src.w - 1);
                top[x] = s0[o] * ca[2 * x] + s0[o1] * ca[2 * x + 1];
                bot[x] = s1[o] * ca[2 * x] + s1[o1] * ca[2 * x + 1];
            }
            const int b0 = to_short(round_he((1.f - ay.frac[y]) * 2048)), b1 = to_short(round_he(ay.frac[y] * 2048));
            uint8_t* d = dst.row(y);
            for (int x = 0; x < dw; x++)
                d[x] = (uint8_t)((((b0 * (top[x] >> 4)) >> 16) + ((b1 * (bot[x] >> 4)) >> 16) + 2) >> 2);
        }
    });
    return dst;
}

void resize_linear_tables(int src, int dst, bool clamp_ofs, std::vector<int>& ofs, std::vector<short>& coef)
{
    Axis a = linear_axis(src, dst, clamp_ofs);
    ofs = a.ofs;
    coef.resize(2 * (size_t)dst);
    for (int d = 0; d < dst; d++) {
        coef[2 * d] = to_short(round_he((1.f - a.frac[d]) * 2048));
        coef[2 * d + 1] = to_short(round_he(a.frac[d] * 2048));
    }
}

Img<float> resize_linear(const Img<float>& src, int dw, int dh)
{
    OB_CHECK(!src.empty() && dw > 0 && dh > 0, "resize: empty");
    Axis ax = linear_axis(src.w, dw, true), ay = linear_axis(src.h, dh, false);
    Img<float> dst(dw, dh);
    parallel_rows(dh, [&](int y0, int y1) {
        std::vector<float> top(dw), bot(dw);
        for (int y = y0; y < y1; y++) {
            const float* s0 = src.row(clampi(ay.ofs[y], 0, src.h - 1));
            const float* s1 = src.row(clampi(ay.ofs[y] + 1, 0, src.h - 1));
            for (int x = 0; x < dw; x++) {
                int o = ax.ofs[x], o1 = std::min(o + 1, src.w - 1);
                float a0 = 1.f - ax.frac[x], a1 = ax.frac[x];
                top[x] = s0[o] * a0 + s0[o1] * a1;
                bot[x] = s1[o] * a0 + s1[o1] * a1;
            }
            const float b0 = 1.f - ay.frac[y], b1 = ay.frac[y];
            float* d = dst.row(y);
            for (int x = 0; x < dw; x++) d[x] = top[x] * b0 + bot[x] * b1;
        }
    });
    return dst;
}

namespace {
inline int mirror101(int p, int len)
{
    if (len == 1) return 0;
    while (p < 0 || p >= len) p = p < 0 ? -p : 2 * len - 2 - p;
    return p;
}
}  // namespace

Img<float> pyrdown_f32(const Img<float>& src)
{
    const int sw = src.w, sh = src.h, dw = (sw + 1) / 2, dh = (sh + 1) / 2;
    Img<float> dst(dw, dh);
    // horizontally filtered + decimated rows, computed once per source row that is needed
    Img<float> hrow(dw, sh);
    parallel_rows(sh, [&](int y0, int y1) {
        for (int y = y0; y < y1; y++) {
            const float* s = src.row(y);
            float* r = hrow.row(y);
            for (int x = 0; x < dw; x++) {
                int c = 2 * x;
                r[x] = s[c] * 6 + (s[mirror101(c - 1, sw)] + s[mirror101(c + 1, sw)]) * 4 + s[mirror101(c - 2, sw)] + s[mirror101(c + 2, sw)];
            }
        }
    });
    const int vec_end = dw & ~7;     // the reference's SSE body covers x < (width & ~7), scalar tail after
    parallel_rows(dh, [&](int y0, int y1) {
        for (int y = y0; y < y1; y++) {
            const float* r0 = hrow.row(mirror101(2 * y - 2, sh)), *r1 = hrow.row(mirror101(2 * y - 1, sh));
            const float* r2 = hrow.row(2 * y), *r3 = hrow.row(mirror101(2 * y + 1, sh)), *r4 = hrow.row(mirror101(2 * y + 2, sh));
            float* d = dst.row(y);
            for (int x = 0; x < vec_end; x++) {
                float a = (r0[x] + r4[x]) + (r2[x] + r2[x]);
                float b = (r1[x] + r3[x]) + r2[x];
                d[x] = (a + b * 4.f) * (1.f / 256);
            }
            for (int x = vec_end; x < dw; x++)
                d[x] = (r2[x] * 6 + (r1[x] + r3[x]) * 4 + r0[x] + r4[x]) * (float)(1. / 256);
        }
    });
    return dst;
}

std::vector<Img<float>> feather_weights(const std::vector<TInput>& in, int border)
{
    const int n = (int)in.size();
    Rect R = in[0].roi;
    for (int i = 1; i < n; i++) R = rect_union(R, in[i].roi);
    Img<float> sum(R.w, R.h, 1e-5f);
    std::vector<Img<float>> w(n);
    for (int i = 0; i < n; i++) {
        w[i] = chamfer_l2(in[i].mask);
        const Rect& r = in[i].roi;
        for (int y = 0; y < r.h; y++) {
            float* wr = w[i].row(y);
            float* sr = sum.row(r.y - R.y + y) + (r.x - R.x);
            for (int x = 0; x < r.w; x++) {
                float t = wr[x] - (float)border;
                wr[x] = t > 0.f ? t : 0.f;
                sr[x] = wr[x] + sr[x];
            }
        }
    }
    const float scale = (float)n;
    for (int i = 0; i < n; i++) {
        const Rect& r = in[i].roi;
        for (int y = 0; y < r.h; y++) {
            float* wr = w[i].row(y);
            const float* sr = sum.row(r.y - R.y + y) + (r.x - R.x);
            for (int x = 0; x < r.w; x++) wr[x] = sr[x] != 0 ? scale * wr[x] / sr[x] : 0.f;
        }
    }
    return w;
}

std::vector<Img<float>> overwrite_weights(const std::vector<TInput>& in)
{
    const int n = (int)in.size();
    Rect R = in[0].roi;
    for (int i = 1; i < n; i++) R = rect_union(R, in[i].roi);
    Img<int8_t> owner(R.w, R.h, (int8_t)-1);
    for (int i = 0; i < n; i++) {
        const Rect& r = in[i].roi;
        for (int y = 0; y < r.h; y++) {
            const uint8_t* m = in[i].mask.row(y);
            int8_t* o = owner.row(r.y - R.y + y) + (r.x - R.x);
            for (int x = 0; x < r.w; x++) if (m[x]) o[x] = (int8_t)i;
        }
    }
    std::vector<Img<float>> w(n);
    for (int i = 0; i < n; i++) {
        const Rect& r = in[i].roi;
        w[i] = Img<float>(r.w, r.h, 0.f);
        for (int y = 0; y < r.h; y++) {
            const int8_t* o = owner.row(r.y - R.y + y) + (r.x - R.x);
            float* wr = w[i].row(y);
            for (int x = 0; x < r.w; x++) if (o[x] == i) wr[x] = 1.f;
        }
    }
    return w;
}

namespace { int g_seam_backend = -1; }
int seam_backend() { return g_seam_backend; }
std::vector<Img<uint8_t>> distance_seam_masks(const std::vector<TInput>& in, int out_w, int device)
{
    std::vector<Img<uint8_t>> out;
    if (distance_seam_masks_gpu(in, out_w, device, out)) { g_seam_backend = 1; return out; }
    g_seam_backend = 0;
    return distance_seam_masks_host(in, out_w);
}

std::vector<Img<uint8_t>> distance_seam_masks_host(const std::vector<TInput>& in, int out_w)
{
    const int n = (int)in.size();
    const double scale = std::min(1.0, 960.0 / out_w);
    std::vector<Rect> sr(n);
    std::vector<Img<uint8_t>> um(n);
    for (int i = 0; i < n; i++) {
        const Rect& r = in[i].roi;
        sr[i] = Rect{ (int)(r.x * scale), (int)(r.y * scale), (int)(r.w * scale), (int)(r.h * scale) };
        um[i] = resize_linear(in[i].mask, sr[i].w, sr[i].h);
    }
    Rect R = sr[0];
    for (int i = 1; i < n; i++) R = rect_union(R, sr[i]);
    std::vector<Img<float>> dist(n);
    for (int i = 0; i < n; i++) {
        if (sr[i].x == 0 && sr[i].w == R.w) {     // full-width mask: wrap-around DT on a 3x tiled copy
            Img<uint8_t> tiled(3 * sr[i].w, sr[i].h);
            for (int y = 0; y < sr[i].h; y++)
                for (int k = 0; k < 3; k++) std::copy(um[i].row(y), um[i].row(y) + sr[i].w, tiled.row(y) + k * sr[i].w);
            Img<float> td = chamfer_l2(tiled);
            dist[i] = Img<float>(sr[i].w, sr[i].h);
            for (int y = 0; y < sr[i].h; y++) std::copy(td.row(y) + sr[i].w, td.row(y) + 2 * sr[i].w, dist[i].row(y));
        } else
            dist[i] = chamfer_l2(um[i]);
    }
    for (int y = R.y; y < R.y + R.h; y++)
        for (int x = R.x; x < R.x + R.w; x++) {
            // the reference sorts candidates by distance (descending) with std::sort; for <= 16
            // cameras that is an insertion sort, so ties keep the lower index in front
            int win = -1; float wd = 0.f;
            for (int k = 0; k < n; k++) {
                int lx = x - sr[k].x, ly = y - sr[k].y;
                float d = (lx >= 0 && ly >= 0 && lx < sr[k].w && ly < sr[k].h) ? dist[k].row(ly)[lx] : -1.f;
                if (win < 0 || d > wd) { win = k; wd = d; }
            }
            for (int k = 0; k < n; k++) {
                if (k == win) continue;
                int lx = x - sr[k].x, ly = y - sr[k].y;
                if (lx >= 0 && ly >= 0 && lx < sr[k].w && ly < sr[k].h) um[k].row(ly)[lx] = 0;
            }
        }
    std::vector<Img<uint8_t>> out(n);
    for (int i = 0; i < n; i++) out[i] = resize_linear(um[i], in[i].roi.w, in[i].roi.h);
    return out;
}

void quantise_map(const Img<float>& map1, const Img<float>& map2, int src_w, int src_h,
                  Img<int32_t>& sx, Img<int32_t>& sy, float shift)
{
    sx = Img<int32_t>(map1.w, map1.h);
    sy = Img<int32_t>(map1.w, map1.h);
    const float fw = (float)(double)src_w, fh = (float)(double)src_h;
    parallel_rows(map1.h, [&](int y0, int y1) {
        for (int y = y0; y < y1; y++) {
            const float* a = map1.row(y), *b = map2.row(y);
            int32_t* ox = sx.row(y), *oy = sy.row(y);
            for (int x = 0; x < map1.w; x++) {
                float px = a[x] * fw + 0.f, py = b[x] * fh + 0.f;     // Mat * double -> f32 convertTo
                if (shift != 0.f) { px = px - shift; py = py - shift; }
                ox[x] = round_he(px * 32.f);
                oy[x] = round_he(py * 32.f);
            }
        }
    });
}


// ------------------------------------------------------------------------------------------------
// fill_poly_u8: the pixels cv::fillPoly sets (imgproc/src/drawing.cpp), written as a plain scan-line rasteriser.
// What has to match the reference, and is pinned by 310 reference-drawn polygons (tests/golden/fillpoly.npz):
//   * edges live in 16.16 fixed point; an edge from (px, py) to (qx, qy), py != qy, starts at the x of its upper end and
//     moves by trunc((qx - px) * 65536 / (qy - py)) per row; it covers the rows [min y, max y) (drawing.cpp:1195-1248);
//   * on a row, the crossings are ordered by x; an edge that starts on the row goes in front of older crossings with
//     the same x, later ties keep their order; consecutive crossings pair up into spans [ceil(xa), floor(xb)]
//     (drawing.cpp:1261-1404);
//   * rows above the image are walked (the crossings still move) but not drawn;
//   * the outline is drawn with 8-connected lines through the integer vertices, clipped like cv::clipLine
//     (drawing.cpp:80-136): an end point outside is first moved along the line onto the top / bottom image row it
//     violates, then onto the left / right column, with truncating integer division at each move.
// ------------------------------------------------------------------------------------------------
namespace {
// Segment (ax, ay) - (bx, by) against [0, w) x [0, h).  false: nothing of it is inside.
bool clip_line(int w, int h, int& ax, int& ay, int& bx, int& by)
{
    if (w <= 0 || h <= 0) return false;
    const int64_t xmax = w - 1, ymax = h - 1;
    struct P { int64_t x, y; } p{ ax, ay }, q{ bx, by };
    enum { LEFT = 1, RIGHT = 2, ABOVE = 4, BELOW = 8 };
    auto side_x = [&](const P& v) { return (v.x < 0 ? LEFT : 0) | (v.x > xmax ? RIGHT : 0); };
    auto side_y = [&](const P& v) { return (v.y < 0 ? ABOVE : 0) | (v.y > ymax ? BELOW : 0); };
    int cp = side_x(p) | side_y(p), cq = side_x(q) | side_y(q);
    if (cp & cq) return false;                              // both beyond the same border
    if ((cp | cq) == 0) return true;                        // both inside
    // rows first: each end point that is above / below slides onto that row (the other end point as it is at that moment)
    auto to_row = [&](P& v, const P& o, int code) { const int64_t row = (code & ABOVE) ? 0 : ymax; v.x += (row - v.y) * (o.x - v.x) / (o.y - v.y); v.y = row; };
    if (cp & (ABOVE | BELOW)) { to_row(p, q, cp); cp = side_x(p); }
    if (cq & (ABOVE | BELOW)) { to_row(q, p, cq); cq = side_x(q); }
    if (cp & cq) return false;
    if (cp | cq) {                                          // then columns
        auto to_col = [&](P& v, const P& o, int code) { const int64_t col = (code & LEFT) ? 0 : xmax; v.y += (col - v.x) * (o.y - v.y) / (o.x - v.x); v.x = col; };
        if (cp) to_col(p, q, cp);
        if (cq) to_col(q, p, cq);
    }
    ax = (int)p.x; ay = (int)p.y; bx = (int)q.x; by = (int)q.y;
    return true;
}
// Line(img, pt1, pt2, color, 8) = cv::LineIterator(img, pt1, pt2, 8, left_to_right = true), drawing.cpp:142-236, 238-265:
// a Bresenham walk along the major axis, drawn from the left end point
void line8(uint8_t* img, int w, int h, int ax, int ay, int bx, int by, uint8_t val)
{
    if ((unsigned)ax >= (unsigned)w || (unsigned)bx >= (unsigned)w || (unsigned)ay >= (unsigned)h || (unsigned)by >= (unsigned)h)
        if (!clip_line(w, h, ax, ay, bx, by)) return;
    int dx = bx - ax, dy = by - ay;
    if (dx < 0) { dx = -dx; dy = -dy; ax = bx; ay = by; }          // left to right
    int x = ax, y = ay;
    const int ystep = dy < 0 ? -1 : 1;
    if (dy < 0) dy = -dy;
    const bool steep = dy > dx;                                     // major axis = y: the roles of the two steps swap
    const int major = steep ? dy : dx, minor = steep ? dx : dy;
    int err = major - 2 * minor;
    for (int i = 0; i <= major; i++) {
        img[(size_t)y * w + x] = val;
        const bool diag = err < 0;                                  // step along the minor axis as well
        err += -2 * minor + (diag ? 2 * major : 0);
        if (steep) { y += ystep; if (diag) x += 1; }
        else { x += 1; if (diag) y += ystep; }
    }
}
struct Crossing { int x, step, y_end; };                  // an edge on the current row: 16.16 x, x step per row, first row it no longer covers
struct PolyEdge { int y_begin, y_end, x, step; };
}  // namespace

void fill_poly_u8(uint8_t* img, int w, int h, const int* pts, int npts, uint8_t val)
{
    if (npts <= 0) return;
    constexpr int FRAC = 16, UNIT = 1 << FRAC;
    // outline + edge table
    std::vector<PolyEdge> table;
    table.reserve((size_t)npts);
    for (int i = 0, j = npts - 1; i < npts; j = i++) {      // edge j -> i
        const int x0 = pts[2 * j], y0 = pts[2 * j + 1], x1 = pts[2 * i], y1 = pts[2 * i + 1];
        line8(img, w, h, x0, y0, x1, y1, val);
        if (y0 == y1) continue;                             // horizontal edges only contribute their outline
        const int fx0 = x0 * UNIT, fx1 = x1 * UNIT;
        PolyEdge e;
        e.step = (fx1 - fx0) / (y1 - y0);
        if (y0 < y1) { e.y_begin = y0; e.y_end = y1; e.x = fx0; } else { e.y_begin = y1; e.y_end = y0; e.x = fx1; }
        table.push_back(e);
    }
    if (table.size() < 2) return;
    // bounding box of the edges as they will be walked; nothing to do if it misses the image
    int top = INT_MAX, bottom = INT_MIN, left = INT_MAX, right = INT_MIN;
    for (const PolyEdge& e : table) {
        const int x_last = e.x + (e.y_end - e.y_begin) * e.step;
        top = std::min(top, e.y_begin); bottom = std::max(bottom, e.y_end);
        left = std::min(left, std::min(e.x, x_last)); right = std::max(right, std::max(e.x, x_last));
    }
    if (bottom < 0 || top >= h || right < 0 || left >= w * UNIT) return;
    std::sort(table.begin(), table.end(), [](const PolyEdge& a, const PolyEdge& b) {
        if (a.y_begin != b.y_begin) return a.y_begin < b.y_begin;
        if (a.x != b.x) return a.x < b.x;
        return a.step < b.step;
    });
    std::vector<Crossing> row, merged;
    size_t next = 0;                                        // first edge of the table that has not started yet
    for (int y = top; y < std::min(bottom, h); y++) {
        // crossings of this row: the ones carried over that have not ended, merged (by x) with the edges starting here;
        // a starting edge goes in front of carried crossings with the same x
        merged.clear();
        size_t c = 0;
        auto skip_ended = [&]() { while (c < row.size() && row[c].y_end == y) c++; };
        skip_ended();
        while (c < row.size() || (next < table.size() && table[next].y_begin == y)) {
            const bool starts = next < table.size() && table[next].y_begin == y;
            if (c < row.size() && (!starts || row[c].x < table[next].x)) { merged.push_back(row[c++]); skip_ended(); }
            else { merged.push_back(Crossing{ table[next].x, table[next].step, table[next].y_end }); next++; }
        }
        // spans between consecutive pairs; every paired crossing moves on to the next row
        for (size_t k = 0; k + 1 < merged.size(); k += 2) {
            Crossing& a = merged[k], &b = merged[k + 1];
            if (y >= 0) {
                const int lo = std::min(a.x, b.x), hi = std::max(a.x, b.x);
                int xa = (lo + UNIT - 1) >> FRAC, xb = hi >> FRAC;
                if (xa < w && xb >= 0) {
                    xa = std::max(xa, 0); xb = std::min(xb, w - 1);
                    if (xa <= xb) memset(img + (size_t)y * w + xa, val, (size_t)(xb - xa + 1));
                }
            }
            a.x += a.step; b.x += b.step;
        }
        std::stable_sort(merged.begin(), merged.end(), [](const Crossing& a, const Crossing& b) { return a.x < b.x; });
        row.swap(merged);
    }
}


// ------------------------------------------------------------------------------------------------
// png_decode_bgr
// ------------------------------------------------------------------------------------------------
void png_decode_bgr(const uint8_t* bytes, size_t n, int& w, int& h, std::vector<uint8_t>& bgr)
{
    static const uint8_t magic[8] = { 0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A };
    if (n < 8 || memcmp(bytes, magic, 8) != 0) fail(OCTVR_ERR_FORMAT, "png mask: not a PNG file");
    auto be32 = [&](size_t o) { return ((uint32_t)bytes[o] << 24) | ((uint32_t)bytes[o + 1] << 16) | ((uint32_t)bytes[o + 2] << 8) | bytes[o + 3]; };
    int depth = 0, ctype = -1, interlace = 0;
    std::vector<uint8_t> idat, plte;
    w = h = 0;
    for (size_t o = 8; o + 12 <= n;) {
        const uint32_t len = be32(o);
        if (len > n - o - 12) fail(OCTVR_ERR_FORMAT, "png mask: truncated chunk");
        const uint8_t* name = bytes + o + 4, *data = bytes + o + 8;
        if (!memcmp(name, "IHDR", 4) && len >= 13) {
            w = (int)be32(o + 8); h = (int)be32(o + 12); depth = data[8]; ctype = data[9]; interlace = data[12];
        } else if (!memcmp(name, "PLTE", 4)) plte.assign(data, data + len);
        else if (!memcmp(name, "IDAT", 4)) idat.insert(idat.end(), data, data + len);
        else if (!memcmp(name, "IEND", 4)) break;
        o += 12 + (size_t)len;
    }
    if (w <= 0 || h <= 0 || w > 65536 || h > 65536 || ctype < 0) fail(OCTVR_ERR_FORMAT, "png mask: bad header");
    if (interlace) fail(OCTVR_ERR_UNSUPPORTED, "png mask: interlaced PNG files are not supported");
    const int channels = ctype == 0 ? 1 : ctype == 2 ? 3 : ctype == 3 ? 1 : ctype == 4 ? 2 : ctype == 6 ? 4 : 0;
    const bool depth_ok = depth == 8 || (depth == 16 && ctype != 3) || ((depth == 1 || depth == 2 || depth == 4) && (ctype == 0 || ctype == 3));
    if (!channels || !depth_ok) fail(OCTVR_ERR_FORMAT, "png mask: unsupported colour type / bit depth");
    const size_t bpp = std::max<size_t>(1, (size_t)channels * depth / 8);            // bytes per complete pixel (filter unit)
    const size_t stride = ((size_t)w * channels * depth + 7) / 8;
    std::vector<uint8_t> raw((stride + 1) * (size_t)h);
    uLongf got = (uLongf)raw.size();
    if (uncompress(raw.data(), &got, idat.data(), (uLong)idat.size()) != Z_OK || got != raw.size()) fail(OCTVR_ERR_FORMAT, "png mask: inflate failed");
    // undo the row filters in place (PNG specification, section 9: None, Sub, Up, Average, Paeth)
    std::vector<uint8_t> zero(stride, 0);
    for (int y = 0; y < h; y++) {
        uint8_t* cur = raw.data() + (size_t)y * (stride + 1) + 1;
        const uint8_t* up = y ? cur - (stride + 1) : zero.data();
        const int f = cur[-1];
        for (size_t i = 0; i < stride; i++) {
            const int a = i >= bpp ? cur[i - bpp] : 0, b = up[i], c = i >= bpp ? up[i - bpp] : 0;
            int pred = 0;
            if (f == 1) pred = a;
            else if (f == 2) pred = b;
            else if (f == 3) pred = (a + b) >> 1;
            else if (f == 4) { const int q = a + b - c, pa = std::abs(q - a), pb = std::abs(q - b), pc = std::abs(q - c); pred = (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c); }
            else if (f != 0) fail(OCTVR_ERR_FORMAT, "png mask: bad filter type");
            cur[i] = (uint8_t)(cur[i] + pred);
        }
    }
    bgr.assign((size_t)w * h * 3, 0);
    for (int y = 0; y < h; y++) {
        const uint8_t* row = raw.data() + (size_t)y * (stride + 1) + 1;
        for (int x = 0; x < w; x++) {
            auto sample = [&](int k) -> int {                                       // k-th sample of pixel x as 8 bits (palette: the index)
                const size_t s = (size_t)x * channels + k;
                if (depth == 8) return row[s];
                if (depth == 16) return row[2 * s];
                const int per = 8 / depth, v = (row[s / per] >> (8 - depth * (int)(s % per + 1))) & ((1 << depth) - 1);
                return ctype == 3 ? v : v * 255 / ((1 << depth) - 1);
            };
            int r, g, b;
            if (ctype == 3) {
                const size_t i = (size_t)sample(0) * 3;
                if (i + 2 >= plte.size()) { r = g = b = 0; } else { r = plte[i]; g = plte[i + 1]; b = plte[i + 2]; }
            } else if (channels <= 2) r = g = b = sample(0);
            else { r = sample(0); g = sample(1); b = sample(2); }
            uint8_t* o = bgr.data() + ((size_t)y * w + x) * 3;
            o[0] = (uint8_t)b; o[1] = (uint8_t)g; o[2] = (uint8_t)r;
        }
    }
}

}  // namespace ob
