// oracle/refgen/ref_stitch.cpp -- GOLDEN-VECTOR GENERATOR for the whole per-frame composition
// (test infrastructure; links the UNMODIFIED reference CPU build).
//
// Loads a reference-generated "VRv11" template (octvr_dump output) through the reference's own
// vr::MapperTemplate loader, synthesises I420 frames, and runs the CPU contract of Mapper::stitch
//  with the reference's functions:
//   cv::cvtColor(YUV2RGB_I420) -> cv::remap(map*W, map*H, INTER_LINEAR) -> cv::resize(NEAREST) ->
//   cv::detail::GainCompensator -> feather (octvr recipe, blenders.cpp:531-586 + blender.cu:73-98)
//   or cv::detail::MultiBandBlender(false, bands, CV_32F) fed 16S -> cv::cvtColor(RGB2YUV_I420).
//
// After the blender, as Mapper::stitch does (mapper.cpp:279-312; evident intent for overlays):
//   overlay inputs cvtColor -> remap -> copyTo(result(roi), mask); cv::resize(result, scale_output, INTER_LINEAR) before
//   the RGB2YUV_I420 conversion; cv::resize(result, preview size, INTER_LINEAR).  Overlay frames are camera n, n+1, ...
//
// usage: ref_stitch <tmpl.dat> <in_w> <in_h> <blend> <gain 0|1> <frame kind: noise|smooth> <out.bin>
//                   [<scale_w> <scale_h> <preview_w> <preview_h>]   (0 0 = none)
#include <opencv2/core.hpp>
#include <opencv2/imgproc.hpp>
#include <opencv2/stitching/detail/blenders.hpp>
#include <opencv2/stitching/detail/exposure_compensate.hpp>
#include "octvr.hpp"
#include <cstdio>
#include <cstdint>
#include <cmath>
#include <fstream>
#include <string>
#include <vector>

static FILE* g_out;
static void put(const std::string& name, const cv::Mat& m_)
{
    cv::Mat m = m_.isContinuous() ? m_ : m_.clone();
    uint32_t nl = (uint32_t)name.size();
    fwrite(&nl, 4, 1, g_out); fwrite(name.data(), 1, nl, g_out);
    uint32_t depth = (uint32_t)m.depth(), cn = (uint32_t)m.channels();
    uint64_t rows = m.rows, cols = m.cols;
    fwrite(&depth, 4, 1, g_out); fwrite(&cn, 4, 1, g_out); fwrite(&rows, 8, 1, g_out); fwrite(&cols, 8, 1, g_out);
    fwrite(m.data, 1, m.total() * m.elemSize(), g_out);
}
static uint64_t splitmix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
// SURVEY.md 8(d) synthetic frames; standard I420 (Y, then U, then V planes) in a (1.5h, w) Mat
static cv::Mat make_frame(int cam, int w, int h, bool noise, uint64_t seed)
{
    cv::Mat f(h * 3 / 2, w, CV_8U);
    static const double expo[8] = { 0.80, 0.90, 1.00, 1.10, 1.20, 0.95, 1.05, 0.85 };
    if (noise) {
        for (size_t o = 0; o < f.total(); o++)
            f.data[o] = (uchar)(splitmix64(seed ^ ((uint64_t)cam << 32) ^ (uint64_t)o) & 0xFF);
        return f;
    }
    const double e = expo[cam % 8];
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            double v = (128 + 80 * sin(2 * M_PI * x / w * 3) * cos(2 * M_PI * y / h * 2)) * e;
            f.at<uchar>(y, x) = cv::saturate_cast<uchar>(v);
        }
    uchar* up = f.data + (size_t)w * h, *vp = up + (size_t)(w / 2) * (h / 2);
    for (int y = 0; y < h / 2; y++)
        for (int x = 0; x < w / 2; x++) {
            up[(size_t)y * (w / 2) + x] = cv::saturate_cast<uchar>(128 + 40 * sin(2 * M_PI * x / (w / 2) * 2));
            vp[(size_t)y * (w / 2) + x] = cv::saturate_cast<uchar>(128 - 40 * cos(2 * M_PI * y / (h / 2) * 3));
        }
    return f;
}

int main(int argc, char** argv)
{
    if (argc < 8) { fprintf(stderr, "usage\n"); return 2; }
    std::ifstream tf(argv[1], std::ios::binary);
    vr::MapperTemplate mt(tf);
    int in_w = atoi(argv[2]), in_h = atoi(argv[3]), blend = atoi(argv[4]);
    bool gain = atoi(argv[5]) != 0, noise = std::string(argv[6]) == "noise";
    g_out = fopen(argv[7], "wb");
    const int n = (int)mt.inputs.size();
    if (n == 1) { gain = false; blend = 0; }     // mapper.cpp:78-82
    cv::Size out = mt.out_size;

    std::vector<cv::Mat> warped(n);
    std::vector<cv::Rect> rois(n);
    for (int i = 0; i < n; i++) {
        cv::Mat f = make_frame(i, in_w, in_h, noise, 1234), rgb;
        put("frame" + std::to_string(i), f);
        cv::cvtColor(f, rgb, cv::COLOR_YUV2RGB_I420);
        cv::remap(rgb, warped[i], mt.inputs[i].map1 * in_w, mt.inputs[i].map2 * in_h, cv::INTER_LINEAR);   // template.cpp:174-176
        rois[i] = mt.inputs[i].roi;
        put("warped" + std::to_string(i), warped[i]);
    }
    if (gain) {
        double ws = std::min(1.0, sqrt(0.1 * 1e6 / out.area()));          // mapper.cpp:94
        std::vector<cv::UMat> imgs(n), masks(n);
        std::vector<cv::Point> corners(n);
        for (int i = 0; i < n; i++) {
            cv::Rect sr(rois[i].x * ws, rois[i].y * ws, rois[i].width * ws, rois[i].height * ws);   // mapper.cpp:95-99
            cv::Mat si, sm;
            cv::resize(warped[i], si, sr.size(), 0, 0, cv::INTER_NEAREST);                         // mapper.cpp:235-237
            cv::resize(mt.inputs[i].mask, sm, sr.size());                                           // mapper.cpp:113-114
            si.copyTo(imgs[i]); sm.copyTo(masks[i]);
            corners[i] = sr.tl();
        }
        cv::detail::GainCompensator gc;
        static_cast<cv::detail::ExposureCompensator&>(gc).feed(corners, imgs, masks);
        std::vector<double> g = gc.gains();
        put("gains", cv::Mat(g, true));
        for (int i = 0; i < n; i++) gc.apply(i, corners[i], warped[i], mt.inputs[i].mask);
    }
    cv::Mat result(out, CV_8UC3, cv::Scalar::all(0));
    cv::Rect R = rois[0];
    for (int i = 1; i < n; i++) R |= rois[i];
    if (blend < 0) {
        int border = -blend;
        cv::Mat S(R.size(), CV_32F, cv::Scalar(1e-5f));
        std::vector<cv::Mat> ws(n);
        for (int i = 0; i < n; i++) {
            cv::Mat tmp;
            cv::distanceTransform(mt.inputs[i].mask, ws[i], cv::DIST_L2, 3);
            cv::subtract(ws[i], border, tmp);
            cv::threshold(tmp, ws[i], 0.f, 0.f, cv::THRESH_TOZERO);
            cv::Mat t = S(rois[i] - R.tl());
            cv::add(ws[i], t, t);
        }
        cv::Mat acc(R.size(), CV_16SC3, cv::Scalar::all(0));
        for (int i = 0; i < n; i++) {
            cv::divide(ws[i], S(rois[i] - R.tl()), ws[i], (double)n);
            cv::Mat a = acc(rois[i] - R.tl());
            for (int y = 0; y < a.rows; y++) {
                const cv::Vec3b* s = warped[i].ptr<cv::Vec3b>(y);
                const float* w = ws[i].ptr<float>(y);
                cv::Vec3s* d = a.ptr<cv::Vec3s>(y);
                for (int x = 0; x < a.cols; x++) {
                    if (w[x] == 0) continue;
                    d[x][0] += static_cast<short>(s[x][0] * w[x]);
                    d[x][1] += static_cast<short>(s[x][1] * w[x]);
                    d[x][2] += static_cast<short>(s[x][2] * w[x]);
                }
            }
        }
        cv::Mat r8;
        acc.convertTo(r8, CV_8UC3, 1.0 / n);
        r8.copyTo(result(R));
    } else if (blend > 0) {
        int bands = int(ceil(log(blend) / log(2.)) - 1.);                 // mapper.cpp:161
        cv::detail::MultiBandBlender mb(false, bands, CV_32F);
        std::vector<cv::Point> corners; std::vector<cv::Size> sizes;
        for (int i = 0; i < n; i++) { corners.push_back(rois[i].tl()); sizes.push_back(rois[i].size()); }
        static_cast<cv::detail::Blender&>(mb).prepare(corners, sizes);
        for (int i = 0; i < n; i++) {
            cv::Mat im16;
            warped[i].convertTo(im16, CV_16S);
            mb.feed(im16, mt.seam_masks[i], rois[i].tl());               // seam masks as masks: mapper.cpp:163
        }
        cv::Mat res, res_mask, r8;
        mb.blend(res, res_mask);
        res.convertTo(r8, CV_8U);
        r8.copyTo(result(R));
    } else {
        for (int i = 0; i < n; i++) warped[i].copyTo(result(rois[i]), mt.inputs[i].mask);   // mapper.cpp:269-275
    }
    for (size_t k = 0; k < mt.overlay_inputs.size(); k++) {                // mapper.cpp:279-282
        const auto& ov = mt.overlay_inputs[k];
        cv::Mat f = make_frame(n + (int)k, in_w, in_h, noise, 1234), rgb, w;
        put("frame" + std::to_string(n + k), f);
        cv::cvtColor(f, rgb, cv::COLOR_YUV2RGB_I420);
        cv::remap(rgb, w, ov.map1 * in_w, ov.map2 * in_h, cv::INTER_LINEAR);
        w.copyTo(result(ov.roi), ov.mask);
    }
    put("result_rgb", result);
    const int scale_w = argc > 9 ? atoi(argv[8]) : 0, scale_h = argc > 9 ? atoi(argv[9]) : 0;
    const int prev_w = argc > 11 ? atoi(argv[10]) : 0, prev_h = argc > 11 ? atoi(argv[11]) : 0;
    cv::Mat scaled = result;
    if (scale_w > 0 && scale_h > 0 && cv::Size(scale_w, scale_h) != out) {  // mapper.cpp:290-294
        cv::resize(result, scaled, cv::Size(scale_w, scale_h), 0, 0, cv::INTER_LINEAR);
        put("result_scaled_rgb", scaled);
    }
    cv::Mat yuv;
    cv::cvtColor(scaled, yuv, cv::COLOR_RGB2YUV_I420);
    put("result_yuv", yuv);
    if (prev_w > 0 && prev_h > 0) {                                        // mapper.cpp:308-312
        cv::Mat pv;
        cv::resize(result, pv, cv::Size(prev_w, prev_h), 0, 0, cv::INTER_LINEAR);
        put("preview_rgb", pv);
    }
    fclose(g_out);
    return 0;
}
