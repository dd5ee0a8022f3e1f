// orbfe_match.cu -- OrbMatcher / Frame-grid entry points of include/orbfe.h (placeholder TU,
// replaced by the real kernels in the next milestone).
#include "../../include/orbfe.h"
#include "orbfe_host.h"
extern "C" {
int orbfe_descriptor_distance(int, const uint8_t*, const uint8_t*, int, int32_t*) { return orbfe_fail(ORBFE_ERR_INVALID, "not implemented"); }
int orbfe_frame_create(int, int, const orbfe_keypoint*, const uint8_t*, const float*, float, float, float, float, int, const float*, orbfe_frame**) { return orbfe_fail(ORBFE_ERR_INVALID, "not implemented"); }
int orbfe_frame_destroy(orbfe_frame*) { return ORBFE_OK; }
int orbfe_features_in_area(orbfe_frame*, float, float, float, int, int, int32_t*, int, int*) { return orbfe_fail(ORBFE_ERR_INVALID, "not implemented"); }
int orbfe_search_for_initialization(orbfe_frame*, orbfe_frame*, float*, int32_t*, int, float, int, int*) { return orbfe_fail(ORBFE_ERR_INVALID, "not implemented"); }
int orbfe_search_by_projection_mappoints(orbfe_frame*, int, const uint8_t*, const float*, const float*, const float*, const int32_t*, const float*, const uint8_t*, const uint8_t*, const uint8_t*, int, float, int32_t*, int*) { return orbfe_fail(ORBFE_ERR_INVALID, "not implemented"); }
int orbfe_search_by_projection_lastframe(orbfe_frame*, int, const uint8_t*, const float*, const float*, const float*, const int32_t*, const float*, const uint8_t*, const uint8_t*, float, int, int, const uint8_t*, float, int, int32_t*, int*) { return orbfe_fail(ORBFE_ERR_INVALID, "not implemented"); }
}
