// include/octvr.hpp -- header-only C++ shim with the reference's names over the C ABI (octvr_b200.h).
//
// Mirrors modules/octvr/include/octvr.hpp (vr::MapperTemplate :48-91, vr::AsyncMultiMapper :103-121, vr::FastMapper :123-144) and
// modules/octvr/src/mapper.hpp (vr::Mapper :29-95): same class and method names, argument meaning and error
// behaviour (std::string thrown for bad camera type / size / .dat magic -- template.cpp:30,33,53,262;
// exceptions for shape violations where the reference CV_Asserts).  OpenCV-free: planes are vr::Plane views, sizes vr::Size,
// camera options the text of their JSON object; define OCTVR_WITH_OPENCV before including to add the cv::Mat / cv::Size /
// cv::Rect_<double> overloads of push() and New() with the reference's exact signatures.  Differences that remain: the
// options are JSON text, not rapidjson::Value; Input holds pointers, not cv::Mat; morph_controlpoints is not provided.
#pragma once
#include "octvr_b200.h"
#include <array>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <unistd.h>
#include <iterator>
#include <stdexcept>
#include <string>
#include <tuple>
#include <vector>
#ifdef OCTVR_WITH_OPENCV
#include <opencv2/core.hpp>
#endif

namespace vr {

struct Error : std::runtime_error {          // stands in for cv::Exception
    int code;
    Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};
struct NotImplemented : std::exception {};    // camera.hpp:31

inline void check(octvr_status st)
{
    if (st == OCTVR_OK) return;
    const std::string msg = octvr_last_error();
    if (st == OCTVR_ERR_FORMAT) throw std::string(msg);      // the reference throws std::string here
    if (st == OCTVR_ERR_UNSUPPORTED) throw NotImplemented();
    throw Error(st, msg);
}

struct Size { int width = 0, height = 0; };
struct Rect { int x = 0, y = 0, width = 0, height = 0; };
struct RectD { double x = 0, y = 0, width = 1, height = 1; };
// one 8-bit plane (host or device memory, depending on the call)
struct Plane { uint8_t* data = nullptr; size_t step = 0; int pixel_step = 1; };
typedef std::tuple<Plane, Plane, Plane> YUV;   // std::tuple<cv::Mat, cv::Mat, cv::Mat> of the reference

inline octvr_frame to_frame(const YUV& f)
{
    octvr_frame o;
    o.y = std::get<0>(f).data; o.u = std::get<1>(f).data; o.v = std::get<2>(f).data;
    o.y_pitch = std::get<0>(f).step; o.u_pitch = std::get<1>(f).step; o.v_pitch = std::get<2>(f).step;
    o.uv_pixel_stride = std::get<1>(f).pixel_step;
    return o;
}

// Multiple input -> single output (octvr.hpp:47-91)
class MapperTemplate {
public:
    struct Input {                       // octvr.hpp:55-62
        Rect roi;
        const float* map1 = nullptr; const float* map2 = nullptr;
        const uint8_t* mask = nullptr; const uint8_t* seam_mask = nullptr;
        const float* vignette = nullptr; Size vignette_size;
    };
    Size out_size;
    // octvr.hpp:64-67: tables of the blended inputs, of the overlay inputs, and the seam masks of the blended inputs.  Views
    // into memory owned by the template, refreshed by every call that changes it.
    std::vector<Input> inputs, overlay_inputs;
    std::vector<const uint8_t*> seam_masks;

    // MapperTemplate(const std::string& to, const rapidjson::Value& to_opts, int width, int height) (octvr.hpp:72-74); the
    // options are the text of the JSON object here (the shim has no rapidjson dependency).  Throws std::string for an unknown
    // camera type or an invalid size, like the reference (template.cpp:30,33).
    MapperTemplate(const std::string& to, const std::string& to_opts_json, int width, int height, int device = 0)
    {
        check(octvr_template_create(to.c_str(), to_opts_json.c_str(), width, height, device, &h_));
        refresh();
    }
    // void add_input(const std::string& from, const rapidjson::Value& from_opts, bool overlay = false, bool use_roi = true)
    // (octvr.hpp:75-78, template.cpp:46-153); the projection of every output pixel runs as a CUDA kernel
    void add_input(const std::string& from, const std::string& from_opts_json, bool overlay = false, bool use_roi = true)
    {
        check(octvr_template_add_input(h_, from.c_str(), from_opts_json.c_str(), overlay, use_roi));
        refresh();
    }

    // MapperTemplate(to, to_opts, width, height) + add_input(...) for every entry of the config
    // (apps/octvr/dump.cpp:71-127); the projection runs as a CUDA kernel.
    static MapperTemplate from_config(const std::string& config_json, int width, int height = -1,
                                      bool use_roi = true, bool create_masks = true, int device = 0)
    {
        MapperTemplate t;
        check(octvr_template_build_json(config_json.c_str(), width, height, use_roi, create_masks, device, &t.h_));
        t.refresh();
        return t;
    }
    // explicit MapperTemplate(std::ifstream&) (template.cpp:258-314)
    explicit MapperTemplate(std::ifstream& f)
    {
        std::vector<char> buf((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
        check(octvr_template_load_dat(buf.data(), buf.size(), &h_));
        refresh();
    }
    void create_masks() { check(octvr_template_create_masks(h_)); refresh(); }            // template.cpp:155-204 (no images: DistanceSeamFinder)
    void dump(const std::string& path) { check(octvr_template_dump_file(h_, path.c_str())); }  // template.cpp:206-256
    // void dump(std::ofstream& f) (octvr.hpp:84): the same bytes appended to an open stream
    void dump(std::ofstream& f)
    {
        char name[] = "/tmp/octvr_dump_XXXXXX";
        const int fd = mkstemp(name);
        if (fd < 0) throw Error(OCTVR_ERR_INVALID, "cannot create a temporary file");
        close(fd);
        try { dump(std::string(name)); } catch (...) { std::remove(name); throw; }
        std::ifstream in(name, std::ios::binary);
        f << in.rdbuf();
        in.close();
        std::remove(name);
    }
    size_t num_inputs() const { return (size_t)octvr_template_num_inputs(h_); }
    Input input(int i) const
    {
        Input in; int roi[4], vwh[2];
        check(octvr_template_input(h_, i, roi, &in.map1, &in.map2, &in.mask, &in.seam_mask, &in.vignette, vwh));
        in.roi = Rect{ roi[0], roi[1], roi[2], roi[3] };
        in.vignette_size = Size{ vwh[0], vwh[1] };
        return in;
    }
    const octvr_template* handle() const { return h_; }

    MapperTemplate(MapperTemplate&& o) noexcept : out_size(o.out_size), inputs(std::move(o.inputs)), overlay_inputs(std::move(o.overlay_inputs)),
                                                  seam_masks(std::move(o.seam_masks)), h_(o.h_) { o.h_ = nullptr; }
    MapperTemplate& operator=(MapperTemplate&& o) noexcept
    {
        if (this != &o) {
            octvr_template_destroy(h_); h_ = o.h_; out_size = o.out_size; o.h_ = nullptr;
            inputs = std::move(o.inputs); overlay_inputs = std::move(o.overlay_inputs); seam_masks = std::move(o.seam_masks);
        }
        return *this;
    }
    MapperTemplate(const MapperTemplate&) = delete;
    MapperTemplate& operator=(const MapperTemplate&) = delete;
    ~MapperTemplate() { octvr_template_destroy(h_); }

private:
    MapperTemplate() {}
    void refresh()
    {
        check(octvr_template_out_size(h_, &out_size.width, &out_size.height));
        const int n = octvr_template_num_inputs(h_), no = octvr_template_num_overlays(h_);
        inputs.clear(); overlay_inputs.clear(); seam_masks.clear();
        for (int i = 0; i < n + no; i++) {
            const Input in = input(i);
            if (i < n) { inputs.push_back(in); seam_masks.push_back(in.seam_mask); } else overlay_inputs.push_back(in);
        }
    }
    octvr_template* h_ = nullptr;
};

// mapper.hpp:29-95.  blend: 0 do not blend, > 0 multi-band blend width, < 0 feather blend width.
class Mapper {
public:
    Mapper(const MapperTemplate& mt, std::vector<Size> in_sizes, int blend = 128, bool enable_gain_compensator = true,
           Size scale_output = Size(), int device = 0) : n_((int)in_sizes.size())
    {
        std::vector<int> wh;
        for (auto& s : in_sizes) { wh.push_back(s.width); wh.push_back(s.height); }
        check(octvr_mapper_create(mt.handle(), wh.data(), n_, blend, enable_gain_compensator, scale_output.width, scale_output.height, device, &h_));
        n_ = octvr_template_num_inputs(mt.handle());      // gains() covers the blended inputs, not the overlays
    }
    // inputs / output: DEVICE planes (the reference takes GpuMat in its packed layout; use stitch_packed for that)
    void stitch(const std::vector<YUV>& inputs, const YUV& output, std::vector<double> gains = std::vector<double>(), void* stream = nullptr)
    {
        std::vector<octvr_frame> fin;
        for (auto& f : inputs) fin.push_back(to_frame(f));
        const octvr_frame fo = to_frame(output);
        check(octvr_mapper_stitch(h_, fin.data(), (int)fin.size(), &fo, nullptr, 0, 0, 0, gains.empty() ? nullptr : gains.data(), (int)gains.size(), stream));
    }
    // the same with the preview_output argument of Mapper::stitch (mapper.cpp:308-312): the result resized (INTER_LINEAR)
    // into a DEVICE RGB888 buffer of preview_size; inputs = blended inputs followed by the template's overlay inputs
    void stitch(const std::vector<YUV>& inputs, const YUV& output, uint8_t* preview_rgb, size_t preview_step, Size preview_size,
                std::vector<double> gains = std::vector<double>(), void* stream = nullptr)
    {
        std::vector<octvr_frame> fin;
        for (auto& f : inputs) fin.push_back(to_frame(f));
        const octvr_frame fo = to_frame(output);
        check(octvr_mapper_stitch(h_, fin.data(), (int)fin.size(), &fo, preview_rgb, preview_step, preview_size.width, preview_size.height,
                                  gains.empty() ? nullptr : gains.data(), (int)gains.size(), stream));
    }
    // void stitch(std::vector<GpuMat>& inputs, GpuMat& output, GpuMat& preview, std::vector<double> gains) with the
    // W x 1.5H single-plane layout of mapper.hpp:75-83
    void stitch_packed(const std::vector<const uint8_t*>& inputs, const std::vector<size_t>& steps, uint8_t* output, size_t out_step,
                       std::vector<double> gains = std::vector<double>(), void* stream = nullptr)
    {
        check(octvr_mapper_stitch_packed(h_, inputs.data(), steps.data(), (int)inputs.size(), output, out_step,
                                         gains.empty() ? nullptr : gains.data(), (int)gains.size(), stream));
    }
    std::vector<double> gains() const            // mapper.hpp:85-87
    {
        std::vector<double> g(n_);
        check(octvr_mapper_gains(h_, g.data(), n_));
        return g;
    }
    Mapper(const Mapper&) = delete;
    Mapper& operator=(const Mapper&) = delete;
    ~Mapper() { octvr_mapper_destroy(h_); }
private:
    octvr_mapper* h_ = nullptr;
    int n_ = 0;
};

// octvr.hpp:103-121
class AsyncMultiMapper {
public:
    static AsyncMultiMapper* New(const std::vector<const MapperTemplate*>& mts, std::vector<Size> in_sizes, Size out_size,
                                 std::vector<int> blend_modes, std::vector<int> gain_modes,
                                 std::vector<RectD> output_regions, Size preview_size = Size(), int device = 0)
    {
        std::vector<const octvr_template*> th;
        for (auto* t : mts) th.push_back(t->handle());
        std::vector<int> wh;
        for (auto& s : in_sizes) { wh.push_back(s.width); wh.push_back(s.height); }
        std::vector<double> rg;
        for (auto& r : output_regions) { rg.push_back(r.x); rg.push_back(r.y); rg.push_back(r.width); rg.push_back(r.height); }
        if (blend_modes.size() != th.size() || gain_modes.size() != th.size() || output_regions.size() != th.size())
            throw Error(OCTVR_ERR_INVALID, "blend_modes / gain_modes / output_regions must have one entry per template");
        AsyncMultiMapper* a = new AsyncMultiMapper;
        octvr_status st = octvr_async_create(th.data(), (int)th.size(), wh.data(), (int)in_sizes.size(), out_size.width, out_size.height,
                                             blend_modes.data(), gain_modes.data(), rg.data(), preview_size.width, preview_size.height, device, &a->h_);
        if (st != OCTVR_OK) { delete a; check(st); }
        return a;
    }
    // static AsyncMultiMapper* New(const std::vector<MapperTemplate>& mts, ...) (octvr.hpp:105-111): the reference's container type
    static AsyncMultiMapper* New(const std::vector<MapperTemplate>& mts, std::vector<Size> in_sizes, Size out_size,
                                 std::vector<int> blend_modes, std::vector<int> gain_modes,
                                 std::vector<RectD> output_regions, Size preview_size = Size(), int device = 0)
    {
        std::vector<const MapperTemplate*> ptrs;
        for (auto& t : mts) ptrs.push_back(&t);
        return New(ptrs, in_sizes, out_size, blend_modes, gain_modes, output_regions, preview_size, device);
    }
#ifdef OCTVR_WITH_OPENCV
    // the reference's exact signature (cv::Size, cv::Rect_<double>)
    static AsyncMultiMapper* New(const std::vector<MapperTemplate>& mts, std::vector<cv::Size> in_sizes, cv::Size out_size,
                                 std::vector<int> blend_modes, std::vector<int> gain_modes,
                                 std::vector<cv::Rect_<double>> output_regions, cv::Size preview_size)
    {
        std::vector<Size> is;
        for (auto& s : in_sizes) is.push_back(Size{ s.width, s.height });
        std::vector<RectD> rs;
        for (auto& r : output_regions) rs.push_back(RectD{ r.x, r.y, r.width, r.height });
        return New(mts, is, Size{ out_size.width, out_size.height }, blend_modes, gain_modes, rs, Size{ preview_size.width, preview_size.height });
    }
#endif
    // Push one frame, in YUV420P format: HOST planes; keep them alive and untouched until the matching pop()
    virtual void push(std::vector<YUV>& inputs, YUV& output)
    {
        std::vector<octvr_frame> fin;
        for (auto& f : inputs) fin.push_back(to_frame(f));
        const octvr_frame fo = to_frame(output);
        check(octvr_async_push(h_, fin.data(), (int)fin.size(), &fo));
    }
    virtual void pop() { check(octvr_async_pop(h_)); }
    double fps() const { double v = 0; check(octvr_async_fps(h_, &v)); return v; }
    // preview frame (preview_size, RGB888, host) of the frame popped last; the reference publishes it through Qt shared
    // memory (async.cpp:119-137), here the caller reads it
    const uint8_t* preview(size_t* step = nullptr, Size* size = nullptr) const
    {
        const uint8_t* p = nullptr; size_t st = 0; int w = 0, h = 0;
        check(octvr_async_preview(h_, &p, &st, &w, &h));
        if (step) *step = st;
        if (size) { size->width = w; size->height = h; }
        return p;
    }
    virtual ~AsyncMultiMapper() { octvr_async_destroy(h_); }      // joins cleanly (the reference's destructor terminates, App. F)
#ifdef OCTVR_WITH_OPENCV
    static Plane plane(const cv::Mat& m) { return Plane{ m.data, m.step, 1 }; }
    void push(std::vector<std::tuple<cv::Mat, cv::Mat, cv::Mat>>& inputs, std::tuple<cv::Mat, cv::Mat, cv::Mat>& output)
    {
        std::vector<YUV> in;
        for (auto& t : inputs) in.emplace_back(plane(std::get<0>(t)), plane(std::get<1>(t)), plane(std::get<2>(t)));
        YUV out(plane(std::get<0>(output)), plane(std::get<1>(output)), plane(std::get<2>(output)));
        push(in, out);
    }
#endif
private:
    AsyncMultiMapper() {}
    octvr_async* h_ = nullptr;
};

// octvr.hpp:123-144, mapper_fast.cpp.  NV12 frames as DEVICE pointers to (rows + rows / 2) x step bytes (the reference's
// cv::UMat of that shape: luma rows, then interleaved chroma rows).  stitch() throws like the reference's ("not supported yet").
class FastMapper {
public:
    FastMapper(const MapperTemplate& mt, std::vector<Size> in_sizes, int device = 0)
    {
        std::vector<int> wh;
        for (auto& s : in_sizes) { wh.push_back(s.width); wh.push_back(s.height); }
        check(octvr_fast_create(mt.handle(), wh.data(), (int)in_sizes.size(), device, &h_));
    }
    void stitch(const std::vector<const uint8_t*>&, uint8_t*) { throw "not supported yet"; }      // mapper_fast.cpp:149
    void stitch_nv12(const std::vector<const uint8_t*>& inputs, const std::vector<size_t>& steps, uint8_t* output, size_t out_step, void* stream = nullptr)
    {
        if (inputs.size() != steps.size()) throw Error(OCTVR_ERR_INVALID, "one step per input");
        check(octvr_fast_stitch_nv12(h_, inputs.data(), steps.data(), (int)inputs.size(), output, out_step, stream));
    }
    FastMapper(const FastMapper&) = delete;
    FastMapper& operator=(const FastMapper&) = delete;
    ~FastMapper() { octvr_fast_destroy(h_); }
private:
    octvr_fast* h_ = nullptr;
};

}  // namespace vr
