// oracle/refgen/ref_golden.cpp -- GOLDEN-VECTOR GENERATOR (test infrastructure).
//
// Links against the UNMODIFIED reference CPU build (SURVEY.md Appendix A recipe, no-CUDA static
// libs) and runs the reference's own functions on seeded inputs, writing every input and output
// into one container file that oracle/refgen/make_golden.py turns into tests/golden/*.npz.
// Nothing here is shipped or used at run time; /root/reference is not needed once the fixtures
// are committed.
//
// usage: ref_golden <out.bin>
#include <opencv2/core.hpp>
#include <opencv2/imgproc.hpp>
#include <opencv2/stitching/detail/blenders.hpp>
#include <opencv2/stitching/detail/exposure_compensate.hpp>
#include <cstdio>
#include <cstdint>
#include <string>
#include <vector>

static FILE* g_out;

static uint64_t splitmix64(uint64_t& s)
{
    uint64_t z = (s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static void fill_u8(cv::Mat& m, uint64_t seed)
{
    uint64_t s = seed;
    for (int y = 0; y < m.rows; y++) {
        uchar* p = m.ptr(y);
        for (int x = 0; x < m.cols * (int)m.elemSize(); x++) p[x] = (uchar)(splitmix64(s) & 0xFF);
    }
}
static double urand(uint64_t& s) { return (splitmix64(s) >> 11) * (1.0 / 9007199254740992.0); }

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

// smooth-ish blob mask: union of random discs
static cv::Mat blob_mask(int w, int h, uint64_t seed, int ndisc)
{
    cv::Mat m(h, w, CV_8U, cv::Scalar(0));
    uint64_t s = seed;
    for (int k = 0; k < ndisc; k++) {
        int cx = (int)(urand(s) * w), cy = (int)(urand(s) * h), r = 4 + (int)(urand(s) * (w / 3));
        cv::circle(m, cv::Point(cx, cy), r, cv::Scalar(255), -1);
    }
    return m;
}

int main(int argc, char** argv)
{
    if (argc < 2) { fprintf(stderr, "usage: %s out.bin\n", argv[0]); return 2; }
    g_out = fopen(argv[1], "wb");
    cv::setNumThreads(1);

    // ---- remap (template.cpp:174-176 call convention) ----
    for (int cn = 1; cn <= 3; cn += 2) {
        cv::Mat src(61, 97, CV_MAKETYPE(CV_8U, cn));
        fill_u8(src, 11 + cn);
        cv::Mat mx(50, 80, CV_32F), my(50, 80, CV_32F);
        uint64_t s = 77 + cn;
        for (int y = 0; y < 50; y++)
            for (int x = 0; x < 80; x++) {
                mx.at<float>(y, x) = (float)(-3.0 + urand(s) * (97 + 6));
                my.at<float>(y, x) = (float)(-3.0 + urand(s) * (61 + 6));
            }
        // exact-integer, half-way and -1 ("masked") coordinates
        mx.at<float>(0, 0) = -1.f; my.at<float>(0, 0) = -1.f;
        mx.at<float>(0, 1) = 96.f; my.at<float>(0, 1) = 60.f;
        mx.at<float>(0, 2) = 10.5f; my.at<float>(0, 2) = 20.5f;
        mx.at<float>(0, 3) = 95.984375f; my.at<float>(0, 3) = 59.984375f;
        mx.at<float>(0, 4) = -0.015625f; my.at<float>(0, 4) = 0.f;
        mx.at<float>(0, 5) = 96.99f; my.at<float>(0, 5) = 30.f;
        cv::Mat dl, dn;
        cv::remap(src, dl, mx, my, cv::INTER_LINEAR);
        cv::remap(src, dn, mx, my, cv::INTER_NEAREST);
        std::string p = "remap_c" + std::to_string(cn) + "_";
        put(p + "src", src); put(p + "mapx", mx); put(p + "mapy", my); put(p + "linear", dl); put(p + "nearest", dn);
    }
    {   // normalised maps scaled the way template.cpp:175-176 does (Mat * int)
        cv::Mat m(7, 33, CV_32F);
        uint64_t s = 5;
        for (int i = 0; i < 7 * 33; i++) ((float*)m.data)[i] = (float)urand(s);
        m.at<float>(0, 0) = -1.f;
        cv::Mat a = m * 2704, b = m * 1520;
        put("scale_in", m); put("scale_2704", a); put("scale_1520", b);
    }

    // ---- colour (call sites apps/octvr/map.cpp:125,130) ----
    {
        int w = 64, h = 48;
        cv::Mat i420(h * 3 / 2, w, CV_8U);
        fill_u8(i420, 21);
        // force luma/chroma extremes somewhere
        i420.at<uchar>(0, 0) = 0; i420.at<uchar>(0, 1) = 255; i420.at<uchar>(0, 2) = 16; i420.at<uchar>(0, 3) = 235;
        cv::Mat rgb, rgb_nv12, bgr;
        cv::cvtColor(i420, rgb, cv::COLOR_YUV2RGB_I420);
        cv::cvtColor(i420, rgb_nv12, cv::COLOR_YUV2RGB_NV12);
        cv::cvtColor(i420, bgr, cv::COLOR_YUV2BGR_I420);
        put("color_yuv", i420); put("color_rgb_i420", rgb); put("color_rgb_nv12", rgb_nv12); put("color_bgr_i420", bgr);
        cv::Mat src(h, w, CV_8UC3), back;
        fill_u8(src, 22);
        cv::cvtColor(src, back, cv::COLOR_RGB2YUV_I420);
        put("color_rgb_in", src); put("color_yuv_out", back);
    }

    // ---- distance transform ----
    {
        cv::Mat m = blob_mask(90, 70, 31, 6), d;
        cv::distanceTransform(m, d, cv::DIST_L2, 3);
        put("dt_mask", m); put("dt_dist", d);
        cv::Mat full(40, 30, CV_8U, cv::Scalar(255)), d2;
        cv::distanceTransform(full, d2, cv::DIST_L2, 3);
        put("dt_full_dist", d2);
    }

    // ---- resize ----
    {
        cv::Mat a(120, 200, CV_8UC3), nn;
        fill_u8(a, 41);
        cv::resize(a, nn, cv::Size(22, 13), 0, 0, cv::INTER_NEAREST);
        put("resize_nn_src", a); put("resize_nn_dst", nn);
        cv::Mat m = blob_mask(223, 117, 42, 5), dn, up;
        cv::resize(m, dn, cv::Size(24, 12));
        cv::resize(dn, up, cv::Size(223, 117));
        put("resize_lin_src", m); put("resize_lin_down", dn); put("resize_lin_up", up);
        cv::Mat g(100, 160, CV_8U), gd;
        fill_u8(g, 43);
        cv::resize(g, gd, cv::Size(37, 29));
        put("resize_lin_noise_src", g); put("resize_lin_noise_dst", gd);
        cv::Mat f(32, 32, CV_32F), fu;
        uint64_t s = 44;
        for (int i = 0; i < 32 * 32; i++) ((float*)f.data)[i] = (float)(0.5 + urand(s));
        cv::resize(f, fu, cv::Size(75, 41));
        put("resize_f32_src", f); put("resize_f32_dst", fu);
    }

    // ---- pyramids ----
    {
        cv::Mat a(48, 64, CV_16SC3), dn, up;
        uint64_t s = 51;
        for (int i = 0; i < 48 * 64 * 3; i++) ((short*)a.data)[i] = (short)((int)(urand(s) * 1400) - 600);
        cv::pyrDown(a, dn);
        cv::pyrUp(dn, up, a.size());
        put("pyr_s16_src", a); put("pyr_s16_down", dn); put("pyr_s16_up", up);
        cv::Mat f(29, 37, CV_32F), fd;
        for (int i = 0; i < 29 * 37; i++) ((float*)f.data)[i] = (float)urand(s);
        cv::pyrDown(f, fd);
        put("pyr_f32_src", f); put("pyr_f32_down", fd);
        cv::Mat f2(64, 96, CV_32F), fd2;
        for (int i = 0; i < 64 * 96; i++) ((float*)f2.data)[i] = (float)urand(s);
        cv::pyrDown(f2, fd2);
        put("pyr_f32b_src", f2); put("pyr_f32b_down", fd2);
    }

    // ---- gain compensation ----
    for (int n = 2; n <= 4; n++) {
        std::vector<cv::UMat> imgs, masks;
        std::vector<cv::Point> corners;
        uint64_t s = 60 + n;
        for (int i = 0; i < n; i++) {
            int w = 40 + (int)(urand(s) * 20), h = 30 + (int)(urand(s) * 10);
            cv::Mat im(h, w, CV_8UC3);
            fill_u8(im, 600 + 10 * n + i);
            im.convertTo(im, CV_8UC3, 0.5 + 0.25 * i);
            cv::Mat m = blob_mask(w, h, 700 + 10 * n + i, 4);
            // soften some mask values so the == 255 rule matters
            for (int k = 0; k < 50; k++) m.at<uchar>((int)(urand(s) * h), (int)(urand(s) * w)) = (uchar)(urand(s) * 255);
            corners.push_back(cv::Point((int)(urand(s) * 30), (int)(urand(s) * 15)));
            put("gain" + std::to_string(n) + "_img" + std::to_string(i), im);
            put("gain" + std::to_string(n) + "_mask" + std::to_string(i), m);
            imgs.push_back(im.getUMat(cv::ACCESS_READ).clone()); masks.push_back(m.getUMat(cv::ACCESS_READ).clone());
        }
        cv::Mat cm((int)corners.size(), 2, CV_32S);
        for (int i = 0; i < n; i++) { cm.at<int>(i, 0) = corners[i].x; cm.at<int>(i, 1) = corners[i].y; }
        put("gain" + std::to_string(n) + "_corners", cm);
        cv::detail::GainCompensator gc;
        static_cast<cv::detail::ExposureCompensator&>(gc).feed(corners, imgs, masks);   // mask == 255 rule (:71-78)
        std::vector<double> g = gc.gains();
        put("gain" + std::to_string(n) + "_gains", cv::Mat(g, true));
        cv::Mat applied = imgs[0].getMat(cv::ACCESS_READ).clone();
        gc.apply(0, corners[0], applied, masks[0]);
        put("gain" + std::to_string(n) + "_applied0", applied);
    }
    {   // multiply by assorted scalars incl. ties
        cv::Mat ramp(1, 256, CV_8U);
        for (int i = 0; i < 256; i++) ramp.at<uchar>(0, i) = (uchar)i;
        double gs[] = { 0.5, 1.5, 0.873046875, 1.0000001, 1.2345678901234, 0.99999994, 2.5 };
        cv::Mat all(7, 256, CV_8U), gm(1, 7, CV_64F);
        for (int k = 0; k < 7; k++) {
            cv::Mat r = ramp.clone();
            cv::multiply(r, gs[k], r);
            r.copyTo(all.row(k));
            gm.at<double>(0, k) = gs[k];
        }
        put("mul_gains", gm); put("mul_out", all);
    }

    // ---- feather weights, octvr recipe on the CPU (blenders.cpp:531-572 / monkey_gen.cpp:44-65) ----
    {
        const int n = 3, border = 3;
        cv::Rect rois[n] = { cv::Rect(0, 4, 70, 50), cv::Rect(40, 0, 80, 60), cv::Rect(20, 30, 90, 40) };
        cv::Rect R = rois[0] | rois[1] | rois[2];
        cv::Mat S(R.size(), CV_32F, cv::Scalar(1e-5f));
        std::vector<cv::Mat> ws;
        cv::Mat rm(n, 4, CV_32S);
        for (int i = 0; i < n; i++) {
            cv::Mat m = blob_mask(rois[i].width, rois[i].height, 800 + i, 5), w, tmp;
            cv::distanceTransform(m, w, cv::DIST_L2, 3);
            cv::subtract(w, border, tmp);
            cv::threshold(tmp, w, 0.f, 0.f, cv::THRESH_TOZERO);
            cv::Mat t = S(rois[i] - R.tl());
            cv::add(w, t, t);
            ws.push_back(w);
            put("feather_mask" + std::to_string(i), m);
            rm.at<int>(i, 0) = rois[i].x; rm.at<int>(i, 1) = rois[i].y; rm.at<int>(i, 2) = rois[i].width; rm.at<int>(i, 3) = rois[i].height;
        }
        put("feather_rois", rm);
        for (int i = 0; i < n; i++) {
            cv::divide(ws[i], S(rois[i] - R.tl()), ws[i], (double)n);
            put("feather_w" + std::to_string(i), ws[i]);
        }
        // convertTo(CV_8UC3, 1/N) from 16S (blenders.cpp:581)
        cv::Mat acc(5, 400, CV_16SC3), o8;
        uint64_t s = 81;
        for (int i = 0; i < 5 * 400 * 3; i++) ((short*)acc.data)[i] = (short)(urand(s) * 1600);
        acc.convertTo(o8, CV_8UC3, 1.0 / n);
        put("feather_acc", acc); put("feather_acc_u8", o8);
    }

    // ---- CPU MultiBandBlender fed 16S (test_blenders.cpp:53-55 convention) ----
    for (int variant = 0; variant < 2; variant++) {
        const int n = 3;
        int bands = variant == 0 ? 3 : 5;
        cv::Rect rois[n] = { cv::Rect(5, 7, 120, 90), cv::Rect(70, 0, 130, 100), cv::Rect(30, 50, 150, 70) };
        if (variant == 1) { rois[0] = cv::Rect(0, 0, 200, 97); rois[1] = cv::Rect(90, 10, 167, 120); rois[2] = cv::Rect(13, 60, 230, 75); }
        cv::detail::MultiBandBlender mb(false, bands, CV_32F);
        std::vector<cv::Point> corners; std::vector<cv::Size> sizes;
        for (int i = 0; i < n; i++) { corners.push_back(rois[i].tl()); sizes.push_back(rois[i].size()); }
        static_cast<cv::detail::Blender&>(mb).prepare(corners, sizes);
        cv::Mat rm(n, 4, CV_32S);
        std::string p = "mb" + std::to_string(variant) + "_";
        for (int i = 0; i < n; i++) {
            cv::Mat im(rois[i].size(), CV_8UC3);
            fill_u8(im, 900 + 10 * variant + i);
            cv::GaussianBlur(im, im, cv::Size(9, 9), 3);
            cv::Mat m = blob_mask(rois[i].width, rois[i].height, 950 + 10 * variant + i, 6);
            cv::Mat msoft;
            cv::GaussianBlur(m, msoft, cv::Size(5, 5), 1.5);   // soft-edged like octvr's resized seam masks
            cv::Mat im16;
            im.convertTo(im16, CV_16S);
            mb.feed(im16, msoft, rois[i].tl());
            put(p + "img" + std::to_string(i), im); put(p + "mask" + std::to_string(i), msoft);
            rm.at<int>(i, 0) = rois[i].x; rm.at<int>(i, 1) = rois[i].y; rm.at<int>(i, 2) = rois[i].width; rm.at<int>(i, 3) = rois[i].height;
        }
        cv::Mat res, res_mask, res8;
        mb.blend(res, res_mask);
        res.convertTo(res8, CV_8U);
        cv::Mat bm(1, 1, CV_32S); bm.at<int>(0, 0) = bands;
        put(p + "rois", rm); put(p + "bands", bm); put(p + "result16", res); put(p + "result8", res8); put(p + "result_mask", res_mask);
    }

    fclose(g_out);
    return 0;
}
