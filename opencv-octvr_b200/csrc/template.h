// csrc/template.h -- vr::MapperTemplate's data (octvr.hpp:48-91) behind the C ABI.
#pragma once
#include "common.h"
#include <memory>

struct octvr_template {
    int out_w = 0, out_h = 0;
    std::vector<ob::TInput> inputs, overlays;
    std::vector<ob::Img<uint8_t>> seam_masks;   // only for inputs (octvr.hpp:67)
    std::vector<uint8_t> visible;               // visible_mask (octvr.hpp:68), W*H, used while adding inputs
    // set by MapperTemplate(to, to_opts, width, height): the output camera model, kept for add_input (mapgen.cu)
    std::shared_ptr<void> out_cam;
    int device = -1;
};

namespace ob {
octvr_template* template_from_dat(const uint8_t* bytes, size_t n);
void template_to_dat(octvr_template& t, const std::string& path);
void template_ensure_seams(octvr_template& t);
// mapgen.cu / camera.cpp : JSON -> template with the projection evaluated on the GPU
octvr_template* template_from_json(const std::string& json, int width, int height, bool use_roi,
                                   bool with_seams, int device);
// MapperTemplate(to, to_opts, width, height) and add_input(from, from_opts, overlay, use_roi) (octvr.hpp:72-79,
// template.cpp:23-153); options are JSON object texts
octvr_template* template_create(const std::string& to, const std::string& to_opts_json, int width, int height, int device);
void template_add_input(octvr_template& t, const std::string& from, const std::string& from_opts_json, bool overlay, bool use_roi);
}  // namespace ob
