// oracle/refgen/ref_fillpoly.cpp -- GOLDEN-VECTOR GENERATOR for cv::fillPoly (test infrastructure; links the UNMODIFIED
// reference CPU build).  Camera::Camera draws its selection rectangle and its polygonal exclude /
// include masks with cv::fillPoly(mask, {points}, value) (modules/octvr/src/camera.cpp:96-167).
//
// usage: ref_fillpoly <cases.txt> <out.bin>      cases.txt: one polygon per line "w h n x0 y0 x1 y1 ..."
#include <opencv2/core.hpp>
#include <opencv2/imgproc.hpp>
#include <cstdio>
#include <cstdint>
#include <fstream>
#include <sstream>
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

int main(int argc, char** argv)
{
    if (argc < 3) return 2;
    std::ifstream in(argv[1]);
    g_out = fopen(argv[2], "wb");
    std::string line;
    int k = 0;
    while (std::getline(in, line)) {
        std::istringstream ss(line);
        int w, h, n;
        if (!(ss >> w >> h >> n)) continue;
        std::vector<cv::Point2i> pts(n);
        for (auto& p : pts) ss >> p.x >> p.y;
        cv::Mat img(h, w, CV_8U, cv::Scalar(7));
        cv::fillPoly(img, std::vector<std::vector<cv::Point2i>>({ pts }), 200);
        put("m" + std::to_string(k++), img);
    }
    fclose(g_out);
    return 0;
}
