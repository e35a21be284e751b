// csrc/async.cpp -- vr::AsyncMultiMapper (modules/octvr/include/octvr.hpp:103-121, src/async.{hpp,cpp}) behind
// the C ABI: host planes in, host planes out, BUF_SIZE = 3 frames in flight (async.cpp:261).
//
// The reference runs five host threads (copy-in, upload, map, download, copy-out; async.cpp:32-172,337-349) and its
// push() only enqueues (async.cpp:174-189).  Same contract here with two threads, because the three middle stages are
// CUDA stream work chained by events and need no thread of their own:
//   front thread  T1 + T2 + T3 + T4: takes a pushed frame, waits for a free buffer set, copies pageable planes into
//                 pinned staging (page-locked caller planes are DMA'd directly), then enqueues H2D on the upload stream,
//                 the stitch of every output region on the compute stream and D2H on the download stream;
//   back thread   T5: waits for the oldest frame's last event, copies staged output planes to the caller's (pageable)
//                 planes, marks the frame complete and frees its buffer set.
// push() validates and enqueues, whatever the planes are; pop() blocks until the oldest frame is complete.  An error
// inside a thread is reported by the pop() of that frame.
#include "mapper.h"
#include <chrono>
#include <cmath>
#include <cstring>
#include <condition_variable>
#include <deque>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>

using namespace ob;

namespace ob { void mapper_stitch_internal(octvr_mapper& m, const octvr_frame* in, int n_in, const octvr_frame* out,
                                          const double* gains, int n_gains, const double* d_gains_src, cudaStream_t s,
                                          uint8_t* d_preview, size_t preview_pitch, int preview_w, int preview_h); }

namespace {

struct Slot {
    std::vector<uint8_t*> d_in;       // per camera: contiguous I420 (w x 1.5h)
    std::vector<uint8_t*> h_in;       // pinned staging, same layout
    uint8_t* d_out = nullptr;         // whole output frame, contiguous I420
    uint8_t* h_out = nullptr;         // pinned staging
    uint8_t* d_preview = nullptr;     // preview frame, RGB888 (async.cpp:299-304), when a preview size was given
    uint8_t* h_preview = nullptr;     // pinned copy of it, valid after pop()
    cudaEvent_t uploaded = nullptr, stitched = nullptr, done = nullptr;
    bool busy = false;
    octvr_frame user_out{};           // caller's output planes (filled by the back thread when staged)
    bool out_staged = false;
};

struct Job {                          // one pushed frame: shallow copies of the caller's plane headers (async.cpp:187-188)
    std::vector<octvr_frame> in;
    octvr_frame out;
};

bool is_pinned(const void* p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}

// A few persistent helper threads for the staging copies (T1 / T5): spawning threads per plane cost more than the copies
// themselves (18 planes per C2 frame: 4.2 ms per frame, 239 frames / s with pageable caller planes).
class CopyPool {
public:
    explicit CopyPool(int n)
    {
        for (int i = 0; i < n; i++) th_.emplace_back([this] { loop(); });
    }
    int size() const { return (int)th_.size(); }
    ~CopyPool()
    {
        { std::lock_guard<std::mutex> lk(m_); stop_ = true; }
        cv_.notify_all();
        for (auto& t : th_) t.join();
    }
    // fn(i) for i in [0, count), spread over the helpers and the calling thread; returns when all are done
    void run(int count, const std::function<void(int)>& fn)
    {
        if (count <= 0) return;
        {
            std::lock_guard<std::mutex> lk(m_);
            fn_ = &fn; next_ = 0; count_ = count; done_ = 0; gen_++;
        }
        cv_.notify_all();
        work();
        std::unique_lock<std::mutex> lk(m_);
        fin_.wait(lk, [&] { return done_ == count_; });
        fn_ = nullptr;
    }
private:
    void work()
    {
        for (;;) {
            int i;
            const std::function<void(int)>* f;
            {
                std::lock_guard<std::mutex> lk(m_);
                if (!fn_ || next_ >= count_) return;
                i = next_++; f = fn_;
            }
            (*f)(i);
            std::lock_guard<std::mutex> lk(m_);
            if (++done_ == count_) fin_.notify_all();
        }
    }
    void loop()
    {
        uint64_t seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> lk(m_);
                cv_.wait(lk, [&] { return stop_ || gen_ != seen; });
                if (stop_) return;
                seen = gen_;
            }
            work();
        }
    }
    std::vector<std::thread> th_;
    std::mutex m_;
    std::condition_variable cv_, fin_;
    const std::function<void(int)>* fn_ = nullptr;
    int next_ = 0, count_ = 0, done_ = 0;
    uint64_t gen_ = 0;
    bool stop_ = false;
};

// copy a w x h byte plane (host); big planes in row slices over the pool
void host_copy_plane(CopyPool* pool, uint8_t* dst, size_t dpitch, const uint8_t* src, size_t spitch, int w, int h, int pix_step, int dst_step = 1)
{
    auto rows = [=](int y0, int y1) {
        for (int y = y0; y < y1; y++) {
            const uint8_t* s = src + (size_t)y * spitch;
            uint8_t* d = dst + (size_t)y * dpitch;
            if (pix_step == 1 && dst_step == 1) memcpy(d, s, (size_t)w);
            else for (int x = 0; x < w; x++) d[(size_t)x * dst_step] = s[(size_t)x * pix_step];
        }
    };
    const size_t bytes = (size_t)w * h;
    const int nt = !pool ? 1 : bytes > (2u << 20) ? pool->size() + 1 : bytes > (1u << 19) ? std::min(4, pool->size() + 1) : 1;
    if (nt == 1) { rows(0, h); return; }
    pool->run(nt, [&](int t) { rows(h * t / nt, h * (t + 1) / nt); });
}

}  // namespace

struct octvr_async {
    int device = 0, n_in = 0, n_out = 0, out_w = 0, out_h = 0;
    std::vector<int> in_w, in_h;
    std::vector<std::unique_ptr<octvr_mapper>> mappers;
    std::vector<int> gain_modes;
    std::vector<Rect> regions;                 // pixel rectangles of the output frame
    std::vector<Rect> preview_regions;         // the same fractions of the preview frame
    int preview_w = 0, preview_h = 0;
    const uint8_t* last_preview = nullptr;     // h_preview of the frame popped last
    static constexpr int BUF = 3;              // async.cpp:261
    Slot slots[BUF];
    uint64_t pushed = 0, issued = 0, completed = 0, popped = 0;   // frames: push() accepted / CUDA work enqueued / copy-out done / pop() returned
    std::deque<Job> queue;                     // pushed, not yet issued
    std::deque<int> status;                    // per completed, not yet popped frame: OCTVR_OK or the error of its stage
    std::string worker_error;
    cudaStream_t s_up = nullptr, s_run = nullptr, s_down = nullptr;
    std::mutex mtx;
    std::condition_variable cv;
    std::thread front, back;
    std::unique_ptr<CopyPool> pool_in, pool_out;   // helpers of the front (T1) and back (T5) threads
    bool stop = false;
    void front_loop();
    void back_loop();
    void issue(Slot& s, const Job& job);
    // fps over the last 10 frames (async.cpp:141-147)
    std::deque<std::chrono::steady_clock::time_point> stamps;
    double fps = 0;

    ~octvr_async()
    {
        { std::lock_guard<std::mutex> lk(mtx); stop = true; }
        cv.notify_all();
        if (front.joinable()) front.join();
        if (back.joinable()) back.join();
        cudaSetDevice(device);
        if (s_up) cudaStreamSynchronize(s_up);
        if (s_run) cudaStreamSynchronize(s_run);
        if (s_down) cudaStreamSynchronize(s_down);
        for (auto& s : slots) {
            cudaFree(s.d_preview); cudaFreeHost(s.h_preview);
            for (auto p : s.d_in) cudaFree(p);
            for (auto p : s.h_in) cudaFreeHost(p);
            cudaFree(s.d_out);
            if (s.h_out) cudaFreeHost(s.h_out);
            if (s.uploaded) cudaEventDestroy(s.uploaded);
            if (s.stitched) cudaEventDestroy(s.stitched);
            if (s.done) cudaEventDestroy(s.done);
        }
        mappers.clear();
        if (s_up) cudaStreamDestroy(s_up);
        if (s_run) cudaStreamDestroy(s_run);
        if (s_down) cudaStreamDestroy(s_down);
    }
};

static octvr_frame i420_frame(uint8_t* base, int w, int h)
{
    octvr_frame f;
    f.y = base; f.u = base + (size_t)w * h; f.v = f.u + (size_t)(w / 2) * (h / 2);
    f.y_pitch = (size_t)w; f.u_pitch = f.v_pitch = (size_t)(w / 2); f.uv_pixel_stride = 1;
    return f;
}

// one plane host -> device (async).  Direct DMA from page-locked caller memory, else through the staging buffer.
static void upload_plane(CopyPool* pool, uint8_t* d_dst, uint8_t* h_stage, const uint8_t* src, size_t spitch, int pix_step, int w, int h, cudaStream_t s)
{
    if (pix_step == 1 && is_pinned(src)) {
        if (spitch == (size_t)w) OB_CUDA(cudaMemcpyAsync(d_dst, src, (size_t)w * h, cudaMemcpyHostToDevice, s));   // one DMA, not h row copies
        else OB_CUDA(cudaMemcpy2DAsync(d_dst, (size_t)w, src, spitch, (size_t)w, (size_t)h, cudaMemcpyHostToDevice, s));
    } else {
        host_copy_plane(pool, h_stage, (size_t)w, src, spitch, w, h, pix_step);      // stage T1 (async.cpp:32-56)
        OB_CUDA(cudaMemcpyAsync(d_dst, h_stage, (size_t)w * h, cudaMemcpyHostToDevice, s));
    }
}

extern "C" {

octvr_status octvr_async_create(const octvr_template* const* tmpls, int n_out, const int* in_sizes_wh, int n_in, int out_w, int out_h,
                                const int* blend_modes, const int* gain_modes, const double* regions_xywh,
                                int preview_w, int preview_h, int device, octvr_async** out)
{
    return guard([&] {
        OB_CHECK(tmpls && in_sizes_wh && blend_modes && gain_modes && regions_xywh && out, "null argument");
        OB_CHECK(n_out >= 1 && n_in >= 1, "need at least one template and one input");
        OB_CHECK(out_w > 0 && out_h > 0 && out_w % 2 == 0 && out_h % 2 == 0, "output size must be even (async.cpp:266-267)");
        OB_CHECK(preview_w >= 0 && preview_h >= 0, "negative preview size");
        int count = 0;
        if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0 || device < 0 || device >= count)
            fail(OCTVR_ERR_CUDA, "no usable CUDA device (the stitch path has no CPU fallback)");
        OB_CUDA(cudaSetDevice(device));
        std::unique_ptr<octvr_async> a(new octvr_async);
        a->device = device; a->n_in = n_in; a->n_out = n_out; a->out_w = out_w; a->out_h = out_h;
        if (preview_w > 0 && preview_h > 0) { a->preview_w = preview_w; a->preview_h = preview_h; }    // preview_size.area() > 0
        // _rect_mul_size (async.cpp:20-30): rounded corner and size, clipped to the frame
        auto rect_mul = [](const double* r, int W, int H) {
            int x = (int)std::round(r[0] * W), y = (int)std::round(r[1] * H), w = (int)std::round(r[2] * W), h = (int)std::round(r[3] * H);
            if (x + w >= W) w = W - x;
            if (y + h >= H) h = H - y;
            return Rect{ x, y, w, h };
        };
        for (int i = 0; i < n_in; i++) { a->in_w.push_back(in_sizes_wh[2 * i]); a->in_h.push_back(in_sizes_wh[2 * i + 1]); }
        for (int i = 0; i < n_out; i++) {
            OB_CHECK(tmpls[i], "null template");
            // output_regions are fractions of the output frame (async.cpp:181-185, 247-259)
            const double* r = regions_xywh + 4 * i;
            const Rect px = rect_mul(r, out_w, out_h);
            if (a->preview_w) {
                const Rect pv = rect_mul(r, a->preview_w, a->preview_h);
                OB_CHECK(pv.w > 0 && pv.h > 0 && pv.x >= 0 && pv.y >= 0, "output region leaves no room in the preview frame");
                a->preview_regions.push_back(pv);
            }
            OB_CHECK(px.w > 0 && px.h > 0 && px.x >= 0 && px.y >= 0 && px.x + px.w <= out_w && px.y + px.h <= out_h, "output region outside the frame");
            OB_CHECK(px.x % 2 == 0 && px.y % 2 == 0 && px.w % 2 == 0 && px.h % 2 == 0, "output regions must be even (4:2:0)");
            a->regions.push_back(px);
            const int gm = gain_modes[i];
            OB_CHECK(gm >= -1 && gm <= i, "gain_modes[i] must be -1, i, or an earlier output (async.hpp:79)");
            a->gain_modes.push_back(gm);
            octvr_mapper* m = nullptr;
            octvr_status st = octvr_mapper_create(tmpls[i], in_sizes_wh, n_in, blend_modes[i], gm >= 0 ? 1 : 0, px.w, px.h, device, &m);
            if (st != OCTVR_OK) fail(st, octvr_last_error());
            a->mappers.emplace_back(m);
        }
        OB_CUDA(cudaStreamCreateWithFlags(&a->s_up, cudaStreamNonBlocking));
        OB_CUDA(cudaStreamCreateWithFlags(&a->s_run, cudaStreamNonBlocking));
        OB_CUDA(cudaStreamCreateWithFlags(&a->s_down, cudaStreamNonBlocking));
        for (auto& s : a->slots) {
            for (int i = 0; i < n_in; i++) {
                const size_t bytes = (size_t)a->in_w[i] * a->in_h[i] * 3 / 2;
                uint8_t* d = nullptr; uint8_t* h = nullptr;
                OB_CUDA(cudaMalloc(&d, bytes)); s.d_in.push_back(d);
                OB_CUDA(cudaMallocHost(&h, bytes)); s.h_in.push_back(h);
            }
            const size_t ob = (size_t)out_w * out_h * 3 / 2;
            OB_CUDA(cudaMalloc(&s.d_out, ob));
            OB_CUDA(cudaMallocHost(&s.h_out, ob));
            // output pre-filled with black in YUV (async.cpp:283-310): regions need not tile the frame
            OB_CUDA(cudaMemset(s.d_out, 16, (size_t)out_w * out_h));
            OB_CUDA(cudaMemset(s.d_out + (size_t)out_w * out_h, 128, (size_t)out_w * out_h / 2));
            if (a->preview_w) {
                const size_t pb = (size_t)a->preview_w * a->preview_h * 3;
                OB_CUDA(cudaMalloc(&s.d_preview, pb));
                OB_CUDA(cudaMemset(s.d_preview, 0, pb));                    // preview_mat.setTo(0)
                OB_CUDA(cudaMallocHost(&s.h_preview, pb));
                memset(s.h_preview, 0, pb);
            }
            OB_CUDA(cudaEventCreateWithFlags(&s.uploaded, cudaEventDisableTiming));
            OB_CUDA(cudaEventCreateWithFlags(&s.stitched, cudaEventDisableTiming));
            OB_CUDA(cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming));
        }
        OB_CUDA(cudaDeviceSynchronize());       // the fills above ran on the legacy stream; the pipeline's streams are non-blocking
        // helpers of the two staging stages: most of the host's cores for the inputs (37 MB per C2 frame), a few for the output
        const int hc = (int)std::max(2u, std::thread::hardware_concurrency());
        a->pool_in.reset(new CopyPool(std::max(1, std::min(11, hc - 5))));
        a->pool_out.reset(new CopyPool(std::max(1, std::min(3, hc / 4))));
        a->front = std::thread([p = a.get()] { p->front_loop(); });
        a->back = std::thread([p = a.get()] { p->back_loop(); });
        *out = a.release();
    });
}

octvr_status octvr_async_push(octvr_async* a, const octvr_frame* in, int n_inputs, const octvr_frame* out)
{
    // async.cpp:174-189: shape checks, then the frame is queued; the caller keeps the planes alive until the matching pop()
    return guard([&] {
        OB_CHECK(a && in && out, "null argument");
        OB_CHECK(n_inputs == a->n_in, "wrong number of input frames (async.cpp:176)");
        for (int i = 0; i < a->n_in; i++) {
            const octvr_frame& f = in[i];
            const size_t w = (size_t)a->in_w[i];
            OB_CHECK(f.y && f.u && f.v && f.y_pitch >= w && (f.uv_pixel_stride == 1 || f.uv_pixel_stride == 2) &&
                     f.u_pitch >= w / 2 * f.uv_pixel_stride && f.v_pitch >= w / 2 * f.uv_pixel_stride, "bad input frame");
        }
        const size_t W = (size_t)a->out_w;
        OB_CHECK(out->y && out->u && out->v && out->y_pitch >= W && (out->uv_pixel_stride == 1 || out->uv_pixel_stride == 2) &&
                 out->u_pitch >= W / 2 * out->uv_pixel_stride && out->v_pitch >= W / 2 * out->uv_pixel_stride, "bad output frame");
        Job job;
        job.in.assign(in, in + n_inputs);
        job.out = *out;
        {
            std::lock_guard<std::mutex> lk(a->mtx);
            a->queue.push_back(std::move(job));
            a->pushed++;
        }
        a->cv.notify_all();
    });
}

// T1 .. T4 of one frame (front thread): staging copies, then everything else as stream work
void octvr_async::issue(Slot& s, const Job& job)
{
    octvr_async* a = this;
    const octvr_frame* in = job.in.data();
    const octvr_frame* out = &job.out;
    // T1 + T2: caller planes -> (pinned staging ->) device, on the upload stream
    for (int i = 0; i < a->n_in; i++) {
        const int w = a->in_w[i], h = a->in_h[i];
        const octvr_frame& f = in[i];
        uint8_t* d = s.d_in[i]; uint8_t* hs = s.h_in[i];
        const size_t ysz = (size_t)w * h, csz = (size_t)(w / 2) * (h / 2);
        if (f.uv_pixel_stride == 1 && f.y_pitch == (size_t)w && f.u_pitch == (size_t)w / 2 && f.v_pitch == (size_t)w / 2 &&
            f.u == f.y + ysz && f.v == f.u + csz && is_pinned(f.y)) {
            OB_CUDA(cudaMemcpyAsync(d, f.y, ysz + 2 * csz, cudaMemcpyHostToDevice, a->s_up));       // contiguous pinned I420
            continue;
        }
        upload_plane(a->pool_in.get(), d, hs, f.y, f.y_pitch, 1, w, h, a->s_up);
        upload_plane(a->pool_in.get(), d + ysz, hs + ysz, f.u, f.u_pitch, f.uv_pixel_stride, w / 2, h / 2, a->s_up);
        upload_plane(a->pool_in.get(), d + ysz + csz, hs + ysz + csz, f.v, f.v_pitch, f.uv_pixel_stride, w / 2, h / 2, a->s_up);
    }
    OB_CUDA(cudaEventRecord(s.uploaded, a->s_up));
    // T3: every output region on the compute stream (async.cpp:70-91)
    OB_CUDA(cudaStreamWaitEvent(a->s_run, s.uploaded, 0));
    std::vector<octvr_frame> fin(a->n_in);
    for (int i = 0; i < a->n_in; i++) fin[i] = i420_frame(s.d_in[i], a->in_w[i], a->in_h[i]);
    const octvr_frame whole = i420_frame(s.d_out, a->out_w, a->out_h);
    for (int r = 0; r < a->n_out; r++) {
        const Rect& px = a->regions[r];
        octvr_frame o = whole;
        o.y += (size_t)px.y * whole.y_pitch + px.x;
        o.u += (size_t)(px.y / 2) * whole.u_pitch + px.x / 2;
        o.v += (size_t)(px.y / 2) * whole.v_pitch + px.x / 2;
        const int gm = a->gain_modes[r];
        const double* shared = (gm >= 0 && gm != r) ? a->mappers[gm]->d_gains : nullptr;   // gains of an earlier output
        uint8_t* pv = nullptr;                                          // this output's part of the preview frame (async.cpp:78-79)
        int pvw = 0, pvh = 0;
        if (a->preview_w) {
            const Rect& pr = a->preview_regions[r];
            pv = s.d_preview + ((size_t)pr.y * a->preview_w + pr.x) * 3; pvw = pr.w; pvh = pr.h;
        }
        mapper_stitch_internal(*a->mappers[r], fin.data(), a->n_in, &o, nullptr, 0, shared, a->s_run, pv, (size_t)a->preview_w * 3, pvw, pvh);
    }
    OB_CUDA(cudaEventRecord(s.stitched, a->s_run));
    // T4: device -> host on the download stream; straight into the caller's planes when they are pinned
    OB_CUDA(cudaStreamWaitEvent(a->s_down, s.stitched, 0));
    const int W = a->out_w, H = a->out_h;
    const bool direct = out->uv_pixel_stride == 1 && is_pinned(out->y) && is_pinned(out->u) && is_pinned(out->v);
    s.out_staged = !direct;
    s.user_out = *out;
    const bool tight = out->y_pitch == (size_t)W && out->u_pitch == (size_t)W / 2 && out->v_pitch == (size_t)W / 2;
    if (direct && tight && out->u == out->y + (size_t)W * H && out->v == out->u + (size_t)(W / 2) * (H / 2)) {
        OB_CUDA(cudaMemcpyAsync(out->y, s.d_out, (size_t)W * H * 3 / 2, cudaMemcpyDeviceToHost, a->s_down));      // contiguous I420
    } else if (direct && tight) {
        OB_CUDA(cudaMemcpyAsync(out->y, whole.y, (size_t)W * H, cudaMemcpyDeviceToHost, a->s_down));
        OB_CUDA(cudaMemcpyAsync(out->u, whole.u, (size_t)(W / 2) * (H / 2), cudaMemcpyDeviceToHost, a->s_down));
        OB_CUDA(cudaMemcpyAsync(out->v, whole.v, (size_t)(W / 2) * (H / 2), cudaMemcpyDeviceToHost, a->s_down));
    } else if (direct) {
        OB_CUDA(cudaMemcpy2DAsync(out->y, out->y_pitch, whole.y, whole.y_pitch, (size_t)W, (size_t)H, cudaMemcpyDeviceToHost, a->s_down));
        OB_CUDA(cudaMemcpy2DAsync(out->u, out->u_pitch, whole.u, whole.u_pitch, (size_t)W / 2, (size_t)H / 2, cudaMemcpyDeviceToHost, a->s_down));
        OB_CUDA(cudaMemcpy2DAsync(out->v, out->v_pitch, whole.v, whole.v_pitch, (size_t)W / 2, (size_t)H / 2, cudaMemcpyDeviceToHost, a->s_down));
    } else {
        OB_CUDA(cudaMemcpyAsync(s.h_out, s.d_out, (size_t)W * H * 3 / 2, cudaMemcpyDeviceToHost, a->s_down));
    }
    if (a->preview_w)
        OB_CUDA(cudaMemcpyAsync(s.h_preview, s.d_preview, (size_t)a->preview_w * a->preview_h * 3, cudaMemcpyDeviceToHost, a->s_down));
    OB_CUDA(cudaEventRecord(s.done, a->s_down));
}

void octvr_async::front_loop()
{
    cudaSetDevice(device);
    for (;;) {
        Job job;
        Slot* s;
        {
            std::unique_lock<std::mutex> lk(mtx);
            // a buffer set is free again once its frame has been copied out (T5), as in the reference's free-buffer queues
            cv.wait(lk, [&] { return stop || (!queue.empty() && !slots[issued % BUF].busy); });
            if (stop) return;
            job = std::move(queue.front());
            queue.pop_front();
            s = &slots[issued % BUF];
            s->busy = true;
        }
        int st = OCTVR_OK;
        std::string msg;
        try { issue(*s, job); }
        catch (const Error& e) { st = e.code; msg = e.what(); }
        catch (const std::exception& e) { st = OCTVR_ERR_INVALID; msg = e.what(); }
        {
            std::lock_guard<std::mutex> lk(mtx);
            s->out_staged = s->out_staged && st == OCTVR_OK;
            if (st != OCTVR_OK) { s->user_out = octvr_frame{}; worker_error = msg; }
            status.push_back(st);
            issued++;
        }
        cv.notify_all();
    }
}

void octvr_async::back_loop()
{
    cudaSetDevice(device);
    for (;;) {
        Slot* s;
        int st;
        {
            std::unique_lock<std::mutex> lk(mtx);
            cv.wait(lk, [&] { return stop || completed < issued; });
            if (stop && completed >= issued) return;
            s = &slots[completed % BUF];
            st = status[completed - popped];
        }
        if (st == OCTVR_OK) {
            if (cudaEventSynchronize(s->done) != cudaSuccess) st = OCTVR_ERR_CUDA;
            else if (s->out_staged) {              // T5 (async.cpp:113-172): staged planes -> the caller's planes
                const int W = out_w, H = out_h;
                const octvr_frame& o = s->user_out;
                const uint8_t* hy = s->h_out; const uint8_t* hu = hy + (size_t)W * H; const uint8_t* hv = hu + (size_t)(W / 2) * (H / 2);
                host_copy_plane(pool_out.get(), o.y, o.y_pitch, hy, (size_t)W, W, H, 1);
                host_copy_plane(pool_out.get(), o.u, o.u_pitch, hu, (size_t)W / 2, W / 2, H / 2, 1, o.uv_pixel_stride);
                host_copy_plane(pool_out.get(), o.v, o.v_pitch, hv, (size_t)W / 2, W / 2, H / 2, 1, o.uv_pixel_stride);
            }
        }
        {
            std::lock_guard<std::mutex> lk(mtx);
            status[completed - popped] = st;
            s->busy = false;
            last_preview = s->h_preview;
            completed++;
            auto now = std::chrono::steady_clock::now();
            stamps.push_back(now);
            if (stamps.size() > 11) stamps.pop_front();
            if (stamps.size() >= 2)
                fps = (double)(stamps.size() - 1) / std::chrono::duration<double>(stamps.back() - stamps.front()).count();
        }
        cv.notify_all();
    }
}

octvr_status octvr_async_pop(octvr_async* a)
{
    // async.cpp:191-193: blocks until the oldest pushed frame has left the last stage
    return guard([&] {
        OB_CHECK(a, "null argument");
        std::unique_lock<std::mutex> lk(a->mtx);
        OB_CHECK(a->popped < a->pushed, "pop() without a matching push()");
        a->cv.wait(lk, [&] { return a->completed > a->popped; });
        const int st = a->status.front();
        a->status.pop_front();
        a->popped++;
        if (st != OCTVR_OK) fail(st, "a pipeline stage failed: " + a->worker_error);
    });
}

octvr_status octvr_async_fps(octvr_async* a, double* fps)
{
    return guard([&] { OB_CHECK(a && fps, "null argument"); std::lock_guard<std::mutex> lk(a->mtx); *fps = a->fps; });
}

octvr_status octvr_async_preview(octvr_async* a, const uint8_t** rgb, size_t* pitch, int* w, int* h)
{
    return guard([&] {
        OB_CHECK(a && rgb, "null argument");
        std::lock_guard<std::mutex> lk(a->mtx);
        OB_CHECK(a->preview_w > 0, "no preview size was given to octvr_async_create");
        OB_CHECK(a->last_preview, "no frame popped yet");
        *rgb = a->last_preview;
        if (pitch) *pitch = (size_t)a->preview_w * 3;
        if (w) *w = a->preview_w;
        if (h) *h = a->preview_h;
    });
}

void octvr_async_destroy(octvr_async* a) { delete a; }

}  // extern "C"
