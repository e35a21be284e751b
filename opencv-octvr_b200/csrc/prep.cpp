// csrc/prep.cpp -- init-time host preparation.  See prep.h for the reference lines each routine follows.
#include "prep.h"
#include <algorithm>
#include <climits>
#include <cmath>
#include <thread>

namespace ob {

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

std::vector<Img<uint8_t>> distance_seam_masks(const std::vector<TInput>& in, int out_w)
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
                  Img<int32_t>& sx, Img<int32_t>& sy)
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
                ox[x] = round_he(px * 32.f);
                oy[x] = round_he(py * 32.f);
            }
        }
    });
}


// ------------------------------------------------------------------------------------------------
// fill_poly_u8: cv::fillPoly restated (drawing.cpp).  XY_SHIFT = 16.
// ------------------------------------------------------------------------------------------------
namespace {
// cv::clipLine(Size, pt1, pt2), drawing.cpp:80-136
bool clip_line(int w, int h, int& ax, int& ay, int& bx, int& by)
{
    if (w <= 0 || h <= 0) return false;
    const int64_t right = w - 1, bottom = h - 1;
    int64_t x1 = ax, y1 = ay, x2 = bx, y2 = by;
    int c1 = (x1 < 0) + (x1 > right) * 2 + (y1 < 0) * 4 + (y1 > bottom) * 8;
    int c2 = (x2 < 0) + (x2 > right) * 2 + (y2 < 0) * 4 + (y2 > bottom) * 8;
    if ((c1 & c2) == 0 && (c1 | c2) != 0) {
        int64_t a;
        if (c1 & 12) { a = c1 < 8 ? 0 : bottom; x1 += (a - y1) * (x2 - x1) / (y2 - y1); y1 = a; c1 = (x1 < 0) + (x1 > right) * 2; }
        if (c2 & 12) { a = c2 < 8 ? 0 : bottom; x2 += (a - y2) * (x2 - x1) / (y2 - y1); y2 = a; c2 = (x2 < 0) + (x2 > right) * 2; }
        if ((c1 & c2) == 0 && (c1 | c2) != 0) {
            if (c1) { a = c1 == 1 ? 0 : right; y1 += (a - x1) * (y2 - y1) / (x2 - x1); x1 = a; c1 = 0; }
            if (c2) { a = c2 == 1 ? 0 : right; y2 += (a - x2) * (y2 - y1) / (x2 - x1); x2 = a; c2 = 0; }
        }
        ax = (int)x1; ay = (int)y1; bx = (int)x2; by = (int)y2;
    }
    return (c1 | c2) == 0;
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
struct Edge { int y0, y1, x, dx; Edge* next; };
}  // namespace

void fill_poly_u8(uint8_t* img, int w, int h, const int* pts, int npts, uint8_t val)
{
    if (npts <= 0) return;
    const int SH = 16, ONE = 1 << SH;
    std::vector<Edge> edges;
    edges.reserve((size_t)npts + 1);
    // CollectPolyEdges, drawing.cpp:1195-1248 (shift = 0, no offset): outline + one table entry per non-horizontal edge
    int px = pts[2 * (npts - 1)] << SH, py = pts[2 * (npts - 1) + 1];
    for (int i = 0; i < npts; i++) {
        const int qx = pts[2 * i] << SH, qy = pts[2 * i + 1];
        line8(img, w, h, (px + (ONE >> 1)) >> SH, py, (qx + (ONE >> 1)) >> SH, qy, val);
        if (py != qy) {
            Edge e;
            if (py < qy) { e.y0 = py; e.y1 = qy; e.x = px; } else { e.y0 = qy; e.y1 = py; e.x = qx; }
            e.dx = (qx - px) / (qy - py);
            e.next = nullptr;
            edges.push_back(e);
        }
        px = qx; py = qy;
    }
    // FillEdgeCollection, drawing.cpp:1261-1404
    const int total = (int)edges.size();
    if (total < 2) return;
    int y_max = INT_MIN, x_max = INT_MIN, y_min = INT_MAX, x_min = INT_MAX;
    for (const Edge& e : edges) {
        const int x1 = e.x + (e.y1 - e.y0) * e.dx;
        y_min = std::min(y_min, e.y0); y_max = std::max(y_max, e.y1);
        x_min = std::min(x_min, std::min(e.x, x1)); x_max = std::max(x_max, std::max(e.x, x1));
    }
    if (y_max < 0 || y_min >= h || x_max < 0 || x_min >= (w << SH)) return;
    std::sort(edges.begin(), edges.end(), [](const Edge& a, const Edge& b) {
        return a.y0 - b.y0 ? a.y0 < b.y0 : a.x - b.x ? a.x < b.x : a.dx < b.dx;
    });
    Edge head;                                   // list head of the active edges; also the sentinel appended to the table
    head.y0 = INT_MAX; head.y1 = 0; head.x = 0; head.dx = 0; head.next = nullptr;
    edges.push_back(head);                       // no insertion after this point: pointers into the vector stay valid
    int i = 0;
    Edge* e = &edges[0];
    y_max = std::min(y_max, h);
    for (int y = e->y0; y < y_max; y++) {
        Edge *last, *prelast, *keep_prelast;
        int sort_flag = 0, draw = 0;
        const bool clipline = y < 0;
        prelast = &head;
        last = head.next;
        while (last || e->y0 == y) {
            if (last && last->y1 == y) {         // the edge ends on this row: drop it
                prelast->next = last->next;
                last = last->next;
                continue;
            }
            keep_prelast = prelast;
            if (last && (e->y0 > y || last->x < e->x)) {     // next edge of the active list
                prelast = last;
                last = last->next;
            } else if (i < total) {              // an edge starts on this row: insert it
                prelast->next = e;
                e->next = last;
                prelast = e;
                e = &edges[++i];
            } else
                break;
            if (draw) {
                if (!clipline) {
                    int x1 = keep_prelast->x, x2 = prelast->x;
                    if (x1 > x2) std::swap(x1, x2);
                    x1 = (x1 + ONE - 1) >> SH;
                    x2 = x2 >> SH;
                    if (x1 < w && x2 >= 0) {
                        if (x1 < 0) x1 = 0;
                        if (x2 >= w) x2 = w - 1;
                        for (int x = x1; x <= x2; x++) img[(size_t)y * w + x] = val;
                    }
                }
                keep_prelast->x += keep_prelast->dx;
                prelast->x += prelast->dx;
            }
            draw ^= 1;
        }
        // keep the active list ordered by x (bubble sort, as in the reference: the order decides which spans pair up)
        keep_prelast = nullptr;
        do {
            prelast = &head;
            last = head.next;
            while (last != keep_prelast && last->next != nullptr) {
                Edge* te = last->next;
                if (last->x > te->x) {
                    prelast->next = te;
                    last->next = te->next;
                    te->next = last;
                    prelast = te;
                    sort_flag = 1;
                } else {
                    prelast = last;
                    last = te;
                }
            }
            keep_prelast = prelast;
        } while (sort_flag && keep_prelast != head.next && keep_prelast != &head);
    }
}

}  // namespace ob
