// csrc/camera.h -- camera model parameters (vr::Camera and its 11 subclasses, modules/octvr/src/camera.{hpp,cpp},
// src/cameras/ *) flattened into a POD that the map-generation kernel reads from constant memory.
#pragma once
#include "common.h"
#include "json.h"

namespace ob {

enum CamType { CAM_NORMAL = 0, CAM_PERSPECTIVE, CAM_PINHOLE, CAM_FISHEYE, CAM_EQUIRECT, CAM_FULLFRAME_FISHEYE,
               CAM_OCAM, CAM_STUPIDOVAL, CAM_CUBIC, CAM_EQAREA_NORTH, CAM_EQAREA_SOUTH, CAM_INVALID = -1 };

struct CamModel {
    int type;
    double rot[9];        // rotate_matrix, row-major (camera.cpp:49-73)
    double rot_inv[9];    // rotate_matrix.inv() (used by image_to_obj, camera.cpp:202-210,296-315)
    double min_lon, max_lon;
    // model parameters (meaning per type, see camera.cpp in this directory)
    double p[16];
    int ip[8];
    double pol[64], invpol[64];
    int n_pol, n_invpol;
    double dist[14];
    int n_dist;
    // optional masks in the input image's pixel grid (device pointers once uploaded)
    const uint8_t* exclude_mask; int ex_w, ex_h;
    const uint8_t* include_mask; int in_w, in_h;
};

struct CamHost {
    CamModel m;
    std::vector<uint8_t> exclude, include;      // host copies of the masks
    bool has_vignette = false;
    float vig[4] = { 0, 0, 0, 0 };              // a,b,c,d after the exposure division (vignette.cpp:18-37)
};

CamHost camera_from_json(const std::string& type, const Json& opts);
double camera_aspect_ratio(const CamModel& m);
// vignette.cpp:39-54
Img<float> vignette_map(const float abcd[4], int width, int height);

}  // namespace ob
