// csrc/multiband.cu -- CPU-MultiBandBlender-compatible Laplacian pyramid blend on the GPU.  (placeholder)
#include "mapper.h"
namespace ob {
struct Multiband { int dummy; };
Multiband* multiband_create(octvr_mapper&, const octvr_template&, const std::vector<Img<int32_t>>&, const std::vector<Img<int32_t>>&)
{
    fail(OCTVR_ERR_UNSUPPORTED, "multiband blending is not implemented yet");
}
void multiband_stitch(octvr_mapper&, const octvr_frame*, cudaStream_t) {}
int multiband_launches(const octvr_mapper&) { return 0; }
void multiband_destroy(Multiband* mb) { delete mb; }
}  // namespace ob
