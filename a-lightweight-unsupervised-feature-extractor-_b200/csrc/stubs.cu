// TEMPORARY: entry points not implemented yet (removed as each lands).
#include "common.cuh"
#define NI(name) return b200::fail(B200_EINVAL, name ": not implemented yet")
extern "C" {
int b200_app_cost_topk_f32(const float*, const int32_t*, const float*, const float*, int, int, int, int, int, float*, int, void*) { NI("app_cost"); }
int b200_pair_cost_f32(const float*, const float*, const float*, const float*, const float*, int, int, float, float, float, float, float, float, float*, float*, float*, float*, float*, int, void*) { NI("pair_cost"); }
int b200_tracker_create(b200_tracker**, int, int, int, const b200_tracker_conf*) { NI("tracker"); }
void b200_tracker_destroy(b200_tracker*) {}
int b200_tracker_reset(b200_tracker*, void*) { NI("tracker"); }
int b200_tracker_result_stride(const b200_tracker*) { return 0; }
int b200_tracker_step(b200_tracker*, const int32_t*, const double*, const double*, const float*, const int32_t*, int32_t*, void*) { NI("tracker"); }
int b200_tracker_step_host(b200_tracker*, const int32_t*, const double*, const double*, const float*, const int32_t*, int32_t*, void*) { NI("tracker"); }
int b200_tracker_export(b200_tracker*, int, int32_t*, double*, double*, uint8_t*, float*, float*, int32_t*, int32_t*, int32_t*, double*, double*, double*, int32_t*, void*) { NI("tracker"); }
}
