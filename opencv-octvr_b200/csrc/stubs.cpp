// csrc/stubs.cpp -- entry points still to be implemented this round.
#include "template.h"
extern "C" {
octvr_status octvr_async_create(const octvr_template* const*, int, const int*, int, int, int, const int*, const int*, const double*, int, int, int, octvr_async**)
{ return ob::guard([] { ob::fail(OCTVR_ERR_UNSUPPORTED, "async pipeline is not implemented yet"); }); }
octvr_status octvr_async_push(octvr_async*, const octvr_frame*, int, const octvr_frame*) { return OCTVR_ERR_UNSUPPORTED; }
octvr_status octvr_async_pop(octvr_async*) { return OCTVR_ERR_UNSUPPORTED; }
octvr_status octvr_async_fps(octvr_async*, double*) { return OCTVR_ERR_UNSUPPORTED; }
void octvr_async_destroy(octvr_async*) {}
}
