// oracle/refgen/ref_fast.cpp -- GOLDEN-VECTOR GENERATOR for vr::FastMapper (test infrastructure; links the UNMODIFIED
// reference CPU build).
//
// 1. Runs the reference's own FastMapper constructor (modules/octvr/src/mapper_fast.cpp:27-109, compiled into
//    libopencv_octvr.a) on a reference-generated "VRv11" template (octvr_dump -n: FastMapper needs full-frame inputs) and
//    writes out the tables it built: map1s / map2s (cv::convertMaps CV_16SC2 + CV_16UC1), half_map1s / half_map2s,
//    feather_masks, half_feather_masks.  The members are private; this tool reads them through `#define private public`,
//    which does not change the class layout.
// 2. FastMapper::stitch_nv12 (mapper_fast.cpp:153-195) needs an OpenCL device (cv::remap_weighted asserts without one,
//    imgproc/src/imgwarp.cpp:4635-4690) and this build has none, so the per-frame step is evaluated here with the arithmetic
//    of the OpenCL kernel imgproc/src/opencl/remap_weighted.cl:20-77 (WT = float, DST_T = ushort, convertToDstT =
//    convert_ushort_sat_rte, `dst += ...` wrapping in 16 bits) on the reference-built tables, followed by the reference's
//    own cv::Mat::convertTo(CV_8U, 1.0 / 255.0) and cv::merge.  Every float operation of the kernel is exact except the
//    multiplication by the weight (one IEEE rounding) and the conversion (round to nearest even), so the result does not
//    depend on the OpenCL compiler.
//
// usage: ref_fast <tmpl.dat> <in_w> <in_h> <out.bin>
#include <opencv2/core.hpp>
#include <opencv2/imgproc.hpp>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <fstream>
#include <sstream>
#include <iostream>
#include <memory>
#include <queue>
#include <tuple>
#include <string>
#include <vector>
#include "rapidjson/document.h"
#define private public
#include "octvr.hpp"
#undef private

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

// remap_weighted.cl:20-77 for SRC_T = uchar, DST_T = ushort
static void remap_weighted_cl(const cv::Mat& src, cv::Mat& dst, const cv::Mat& map1, const cv::Mat& map2, const cv::Mat& wmap)
{
    const int src_cols = src.cols, src_rows = src.rows;
    for (int y = 0; y < dst.rows; y++)
        for (int x = 0; x < dst.cols; x++) {
            const cv::Vec2s m = map1.at<cv::Vec2s>(y, x);
            const int ax = m[0], ay = m[1];
            const unsigned short m2 = (unsigned short)(map2.at<unsigned short>(y, x) & 1023);
            const float ux = (float)(m2 & 31) / 32.f, uy = (float)(m2 >> 5) / 32.f;
            auto pix = [&](int gx, int gy) -> float {
                if (gx >= src_cols || gy >= src_rows || gx < 0 || gy < 0) return 0.f;
                return (float)src.at<uchar>(gy, gx);
            };
            const float a = pix(ax, ay), b = pix(ax + 1, ay), c = pix(ax, ay + 1), d = pix(ax + 1, ay + 1);
            volatile float t0 = a * (1 - ux) * (1 - uy), t1 = b * ux * (1 - uy), t2 = c * (1 - ux) * uy, t3 = d * ux * uy;
            volatile float v = t0 + t1 + t2 + t3;
            volatile float p = v * (float)wmap.at<uchar>(y, x);
            float r = nearbyintf(p);                                   // convert_ushort_sat_rte
            unsigned short q = r <= 0.f ? 0 : r >= 65535.f ? 65535 : (unsigned short)r;
            dst.at<unsigned short>(y, x) = (unsigned short)(dst.at<unsigned short>(y, x) + q);
        }
}

int main(int argc, char** argv)
{
    if (argc < 5) { fprintf(stderr, "usage\n"); return 2; }
    std::ifstream f(argv[1], std::ios::binary);
    vr::MapperTemplate mt(f);
    const int iw = atoi(argv[2]), ih = atoi(argv[3]);
    const int n = (int)mt.inputs.size();
    std::vector<cv::Size> in_sizes(n, cv::Size(iw, ih));
    vr::FastMapper fm(mt, in_sizes);                                   // the reference's constructor, unmodified
    g_out = fopen(argv[4], "wb");
    const cv::Size out = mt.out_size;
    for (int i = 0; i < n; i++) {
        put("map1_" + std::to_string(i), fm.map1s[i].getMat(cv::ACCESS_READ));
        put("map2_" + std::to_string(i), fm.map2s[i].getMat(cv::ACCESS_READ));
        put("hmap1_" + std::to_string(i), fm.half_map1s[i].getMat(cv::ACCESS_READ));
        put("hmap2_" + std::to_string(i), fm.half_map2s[i].getMat(cv::ACCESS_READ));
        put("feather" + std::to_string(i), fm.feather_masks[i].getMat(cv::ACCESS_READ));
        put("hfeather" + std::to_string(i), fm.half_feather_masks[i].getMat(cv::ACCESS_READ));
    }
    // NV12-shaped noise frames: (1.5 h, w) bytes, luma rows then interleaved chroma rows
    std::vector<cv::Mat> frames(n);
    for (int i = 0; i < n; i++) {
        frames[i].create(ih + ih / 2, iw, CV_8U);
        for (size_t o = 0; o < frames[i].total(); o++)
            frames[i].data[o] = (uchar)(splitmix64(0xFA57ull ^ ((uint64_t)i << 32) ^ (uint64_t)o) & 0xFF);
    }
    // stitch_nv12, mapper_fast.cpp:153-195
    cv::Mat s_c0(out, CV_16U, cv::Scalar(0));
    std::vector<cv::Mat> s_c1c2{ cv::Mat(out.height / 2, out.width / 2, CV_16U, cv::Scalar(0)), cv::Mat(out.height / 2, out.width / 2, CV_16U, cv::Scalar(0)) };
    for (int i = 0; i < n; i++) {
        cv::Mat c0 = frames[i].rowRange(0, ih);
        cv::Mat c1c2 = frames[i].rowRange(ih, ih + ih / 2);
        std::vector<cv::Mat> ch;
        cv::split(c1c2.reshape(2), ch);
        remap_weighted_cl(c0, s_c0, fm.map1s[i].getMat(cv::ACCESS_READ), fm.map2s[i].getMat(cv::ACCESS_READ), fm.feather_masks[i].getMat(cv::ACCESS_READ));
        remap_weighted_cl(ch[1], s_c1c2[0], fm.half_map1s[i].getMat(cv::ACCESS_READ), fm.half_map2s[i].getMat(cv::ACCESS_READ), fm.half_feather_masks[i].getMat(cv::ACCESS_READ));
        remap_weighted_cl(ch[0], s_c1c2[1], fm.half_map1s[i].getMat(cv::ACCESS_READ), fm.half_map2s[i].getMat(cv::ACCESS_READ), fm.half_feather_masks[i].getMat(cv::ACCESS_READ));
    }
    cv::Mat output(out.height + out.height / 2, out.width, CV_8U);
    cv::Mat o_c0 = output.rowRange(0, out.height);
    s_c0.convertTo(o_c0, CV_8U, 1.0 / 255.0);
    cv::Mat merged;
    cv::merge(s_c1c2, merged);
    cv::Mat o_c1c2 = output.rowRange(out.height, out.height + out.height / 2).reshape(2);
    merged.convertTo(o_c1c2, CV_8U, 1.0 / 255.0);
    put("acc_c0", s_c0);
    put("result", output);
    fclose(g_out);
    return 0;
}
