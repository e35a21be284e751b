// csrc/common.h -- internal types shared by the host side and the CUDA kernels.
#pragma once
#include "../../include/octvr_b200.h"
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <vector>

namespace ob {

// ---- error plumbing -------------------------------------------------------
struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};
void set_last_error(const std::string& m);
[[noreturn]] inline void fail(int code, const std::string& m) { throw Error(code, m); }
#define OB_CUDA(expr)                                                                              \
    do {                                                                                           \
        cudaError_t e__ = (expr);                                                                  \
        if (e__ != cudaSuccess)                                                                    \
            ::ob::fail(OCTVR_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__));       \
    } while (0)
#define OB_CHECK(cond, msg)                                                                        \
    do { if (!(cond)) ::ob::fail(OCTVR_ERR_INVALID, std::string(msg) + " (" #cond ")"); } while (0)

template <class F> inline octvr_status guard(F&& f)
{
    try { f(); return OCTVR_OK; }
    catch (const Error& e) { set_last_error(e.what()); return e.code; }
    catch (const std::exception& e) { set_last_error(e.what()); return OCTVR_ERR_INVALID; }
    catch (...) { set_last_error("unknown exception"); return OCTVR_ERR_INVALID; }
}

// ---- init-time trace (OCTVR_INIT_TRACE=1: stage times of template / mapper construction on stderr) ----------
struct InitTrace {
    const char* what; bool on; double t0;
    static double now();
    explicit InitTrace(const char* w);
    void lap(const char* stage);
};

// ---- host images ------------------------------------------------------------
// Allocator of the host tables: large blocks are 2 MB aligned and advised as transparent huge pages.  What a template or a
// mapper costs on the host is mostly the first touch of fresh pages (measured: 59 MB value-initialised in 40 ms with 4 KB pages,
// 20 ms together with the copy into it with huge pages); where THP is off the advice is simply ignored.
void* big_alloc(size_t bytes);
void big_free(void* p);
template <class T> struct BigAlloc {
    using value_type = T;
    BigAlloc() = default;
    template <class U> BigAlloc(const BigAlloc<U>&) {}
    T* allocate(size_t n) { return static_cast<T*>(big_alloc(n * sizeof(T))); }
    void deallocate(T* p, size_t) { big_free(p); }
    template <class U> bool operator==(const BigAlloc<U>&) const { return true; }
    template <class U> bool operator!=(const BigAlloc<U>&) const { return false; }
};

template <class T> struct Img {
    int w = 0, h = 0;
    std::vector<T, BigAlloc<T>> d;
    Img() {}
    Img(int w_, int h_, T v = T()) : w(w_), h(h_), d((size_t)w_ * h_, v) {}
    bool empty() const { return d.empty(); }
    T* row(int y) { return d.data() + (size_t)y * w; }
    const T* row(int y) const { return d.data() + (size_t)y * w; }
};
struct Rect { int x = 0, y = 0, w = 0, h = 0; };
inline Rect rect_union(const Rect& a, const Rect& b)
{
    int x0 = a.x < b.x ? a.x : b.x, y0 = a.y < b.y ? a.y : b.y;
    int x1 = a.x + a.w > b.x + b.w ? a.x + a.w : b.x + b.w, y1 = a.y + a.h > b.y + b.h ? a.y + a.h : b.y + b.h;
    return Rect{ x0, y0, x1 - x0, y1 - y0 };
}

// vr::MapperTemplate::Input (octvr.hpp:55-62)
struct TInput {
    Rect roi;
    Img<float> map1, map2;
    Img<uint8_t> mask;
    Img<float> vignette;
};

// ---- device-side table layout ------------------------------------------------
// Output is cut into TILE_W x TILE_H tiles; a "job" is one (tile, camera) pair that has at least
// one contributing pixel.  Job j owns entries [j*TILE_PX, (j+1)*TILE_PX) of the coord and weight
// streams, so one CTA reads its tables as dense, fully coalesced 4 KB + 2 KB chunks.
constexpr int TILE_W = 32, TILE_H = 16, TILE_PX = TILE_W * TILE_H;
// coord.x : pixel offset (iy * src_pitch_px + ix) of the top-left bilinear tap in the camera's RGBX plane
// coord.y : bits 0-4 fx, 5-9 fy (1/32 px fractions, imgwarp.cpp:4383-4442), 10 VALID, 11 BORDER (some tap
//           outside the source), 12-15 per-tap inside bits (t00,t01,t10,t11)
constexpr uint32_t C_VALID = 1u << 10, C_BORDER = 1u << 11, C_TAP_SHIFT = 12;

constexpr int MAX_CAMS = 16;

}  // namespace ob
