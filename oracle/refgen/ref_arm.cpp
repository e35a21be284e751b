// oracle/refgen/ref_arm.cpp -- the REFERENCE CPU arm of bench.py (test / measurement infrastructure, never shipped).
//
// A thin C entry point over the UNMODIFIED reference CPU build (static libs compiled from the
// reference's own sources) that runs the per-frame composition of vr::Mapper::stitch with the reference's own functions:
//   cv::cvtColor(YUV2RGB_I420) -> cv::remap(map * W, map * H, INTER_LINEAR) -> cv::resize(NEAREST) -> cv::detail::GainCompensator
//   -> feather (octvr recipe, blenders.cpp:531-586 + cuda/blender.cu:73-98) or cv::detail::MultiBandBlender(false, bands, CV_32F)
//   -> cv::cvtColor(RGB2YUV_I420)
// (the CPU contract of SURVEY.md 8c; the reference's own Mapper is CUDA-only and cannot run on host cores).  What the
// reference's Mapper constructor precomputes (scaled maps, feather weights, working-scale masks: mapper.cpp:84-127,
// blenders.cpp:531-572) is precomputed in refarm_create; refarm_stitch is the per-frame work.  OpenCV's own parallel_for_
// (pthreads backend in this build) spreads remap / cvtColor / resize over the host cores; the per-camera stages additionally run
// one std::thread per camera.
//
// Built by oracle/build_ref.sh into oracle/_ref/libocvref.so (git-ignored; travels to the GPU box with gpurun).
#include <opencv2/core.hpp>
#include <opencv2/imgproc.hpp>
#include <opencv2/stitching/detail/blenders.hpp>
#include <opencv2/stitching/detail/exposure_compensate.hpp>
#include "octvr.hpp"
#include <cmath>
#include <cstdint>
#include <fstream>
#include <string>
#include <thread>
#include <vector>

namespace {
struct Arm {
    vr::MapperTemplate mt;
    int in_w, in_h, blend, n;
    bool gain;
    std::vector<cv::Mat> mapx, mapy, ws, smask;
    std::vector<cv::Rect> rois, srois;
    cv::Rect R;
    std::vector<double> gains;
    explicit Arm(std::ifstream& f) : mt(f) {}
};
// the feather accumulation of one camera over a range of rows (cuda/blender.cu:73-98 on the host)
struct FeatherRows : cv::ParallelLoopBody {
    const cv::Mat& src; const cv::Mat& w; cv::Mat& dst;
    FeatherRows(const cv::Mat& s, const cv::Mat& w_, cv::Mat& d) : src(s), w(w_), dst(d) {}
    void operator()(const cv::Range& rg) const
    {
        for (int y = rg.start; y < rg.end; y++) {
            const cv::Vec3b* s = src.ptr<cv::Vec3b>(y);
            const float* wr = w.ptr<float>(y);
            cv::Vec3s* o = dst.ptr<cv::Vec3s>(y);
            for (int x = 0; x < dst.cols; x++) {
                if (wr[x] == 0) continue;
                o[x][0] += static_cast<short>(s[x][0] * wr[x]);
                o[x][1] += static_cast<short>(s[x][1] * wr[x]);
                o[x][2] += static_cast<short>(s[x][2] * wr[x]);
            }
        }
    }
};
template <class F> void per_camera(int n, F&& body)
{
    std::vector<std::thread> th;
    for (int i = 0; i < n; i++) th.emplace_back([&, i] { body(i); });
    for (auto& t : th) t.join();
}
}  // namespace

extern "C" {

void* refarm_create(const char* dat_path, int in_w, int in_h, int blend, int gain)
{
    try {
        std::ifstream f(dat_path, std::ios::binary);
        if (!f) return nullptr;
        Arm* a = new Arm(f);
        a->in_w = in_w; a->in_h = in_h; a->n = (int)a->mt.inputs.size();
        a->blend = a->n == 1 ? 0 : blend; a->gain = a->n == 1 ? false : gain != 0;          // mapper.cpp:78-82
        const int n = a->n;
        // vignette maps (mapper.cpp:108-112) are multiplied in by a CUDA kernel in the reference; this arm runs the bench rigs,
        // which have none, and refuses templates that do rather than silently skipping the stage
        for (int i = 0; i < n; i++) if (!a->mt.inputs[i].vignette.empty()) { delete a; return nullptr; }
        if (!a->mt.overlay_inputs.empty()) { delete a; return nullptr; }
        a->mapx.resize(n); a->mapy.resize(n); a->rois.resize(n); a->srois.resize(n); a->smask.resize(n); a->ws.resize(n);
        const double wsc = std::min(1.0, std::sqrt(0.1 * 1e6 / a->mt.out_size.area()));     // mapper.cpp:94
        for (int i = 0; i < n; i++) {
            a->mapx[i] = a->mt.inputs[i].map1 * in_w; a->mapy[i] = a->mt.inputs[i].map2 * in_h;   // template.cpp:174-176
            a->rois[i] = a->mt.inputs[i].roi;
            const cv::Rect& r = a->rois[i];
            a->srois[i] = cv::Rect(r.x * wsc, r.y * wsc, r.width * wsc, r.height * wsc);      // mapper.cpp:95-99
            cv::resize(a->mt.inputs[i].mask, a->smask[i], a->srois[i].size());                // mapper.cpp:113-114
        }
        a->R = a->rois[0];
        for (int i = 1; i < n; i++) a->R |= a->rois[i];
        if (a->blend < 0) {                                                                   // blenders.cpp:531-572
            cv::Mat S(a->R.size(), CV_32F, cv::Scalar(1e-5f));
            for (int i = 0; i < n; i++) {
                cv::Mat tmp;
                cv::distanceTransform(a->mt.inputs[i].mask, a->ws[i], cv::DIST_L2, 3);
                cv::subtract(a->ws[i], -a->blend, tmp);
                cv::threshold(tmp, a->ws[i], 0.f, 0.f, cv::THRESH_TOZERO);
                cv::Mat t = S(a->rois[i] - a->R.tl());
                cv::add(a->ws[i], t, t);
            }
            for (int i = 0; i < n; i++) cv::divide(a->ws[i], S(a->rois[i] - a->R.tl()), a->ws[i], (double)n);
        }
        return a;
    } catch (...) { return nullptr; }
}

int refarm_out_size(void* h, int* w, int* hgt) { Arm* a = (Arm*)h; *w = a->mt.out_size.width; *hgt = a->mt.out_size.height; return a->n; }

// frames: n standard I420 frames ((in_h * 3 / 2) x in_w bytes each); out: (H * 3 / 2) x W bytes, standard I420
int refarm_stitch(void* h, const uint8_t* const* frames, uint8_t* out_i420, double* gains_out)
{
    try {
        Arm& a = *(Arm*)h;
        const int n = a.n;
        const cv::Size out = a.mt.out_size;
        std::vector<cv::Mat> warped(n);
        per_camera(n, [&](int i) {
            cv::Mat f(a.in_h * 3 / 2, a.in_w, CV_8U, const_cast<uint8_t*>(frames[i])), rgb;
            cv::cvtColor(f, rgb, cv::COLOR_YUV2RGB_I420);
            cv::remap(rgb, warped[i], a.mapx[i], a.mapy[i], cv::INTER_LINEAR);
        });
        if (a.gain) {
            std::vector<cv::UMat> imgs(n), masks(n);
            std::vector<cv::Point> corners(n);
            for (int i = 0; i < n; i++) {
                cv::Mat si;
                cv::resize(warped[i], si, a.srois[i].size(), 0, 0, cv::INTER_NEAREST);      // mapper.cpp:235-237
                si.copyTo(imgs[i]); a.smask[i].copyTo(masks[i]);
                corners[i] = a.srois[i].tl();
            }
            cv::detail::GainCompensator gc;
            static_cast<cv::detail::ExposureCompensator&>(gc).feed(corners, imgs, masks);
            a.gains = gc.gains();
            if (gains_out) for (int i = 0; i < n; i++) gains_out[i] = a.gains[i];
            per_camera(n, [&](int i) { gc.apply(i, corners[i], warped[i], a.mt.inputs[i].mask); });
        }
        cv::Mat result(out, CV_8UC3, cv::Scalar::all(0));
        if (a.blend < 0) {
            cv::Mat acc(a.R.size(), CV_16SC3, cv::Scalar::all(0));
            for (int i = 0; i < n; i++) {                                                  // cuda/blender.cu:73-98 on the host
                cv::Mat d = acc(a.rois[i] - a.R.tl());
                cv::parallel_for_(cv::Range(0, d.rows), FeatherRows(warped[i], a.ws[i], d));
            }
            cv::Mat r8;
            acc.convertTo(r8, CV_8UC3, 1.0 / n);
            r8.copyTo(result(a.R));
        } else if (a.blend > 0) {
            const int bands = int(std::ceil(std::log((double)a.blend) / std::log(2.)) - 1.);     // mapper.cpp:161
            cv::detail::MultiBandBlender mb(false, bands, CV_32F);
            std::vector<cv::Point> corners; std::vector<cv::Size> sizes;
            for (int i = 0; i < n; i++) { corners.push_back(a.rois[i].tl()); sizes.push_back(a.rois[i].size()); }
            static_cast<cv::detail::Blender&>(mb).prepare(corners, sizes);
            for (int i = 0; i < n; i++) {
                cv::Mat im16;
                warped[i].convertTo(im16, CV_16S);
                mb.feed(im16, a.mt.seam_masks[i], a.rois[i].tl());                         // mapper.cpp:163
            }
            cv::Mat res, res_mask, r8;
            mb.blend(res, res_mask);
            res.convertTo(r8, CV_8U);
            r8.copyTo(result(a.R));
        } else {
            for (int i = 0; i < n; i++) warped[i].copyTo(result(a.rois[i]), a.mt.inputs[i].mask);   // mapper.cpp:269-275
        }
        cv::Mat yuv(out.height * 3 / 2, out.width, CV_8U, out_i420);
        cv::cvtColor(result, yuv, cv::COLOR_RGB2YUV_I420);
        return yuv.data == out_i420 ? 0 : 1;
    } catch (...) { return -1; }
}

int refarm_threads(void) { return cv::getNumThreads(); }
void refarm_destroy(void* h) { delete (Arm*)h; }

}  // extern "C"
