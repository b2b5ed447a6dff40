// Entry points whose kernels are not built yet return B2_ERR_STATE with a message (never a silent fallback).
#include "common.cuh"
void sg_destroy(b2_context*) {}
#define NOT_BUILT(ctx) b2_fail(ctx, B2_ERR_STATE, std::string(__func__) + ": not built yet")
extern "C" {
int b2_superglue_set_weights(b2_context* c, const float*, size_t) { return NOT_BUILT(c); }
int b2_superglue_match_dev(b2_context* c, const float*, const float*, const float*, int, int, int, const float*,
                           const float*, const float*, int, int, int, int, float, uint32_t*, float*, int*, void*) { return NOT_BUILT(c); }
int b2_superglue_match_host(b2_context* c, const float*, const float*, const float*, int, int, int, const float*,
                            const float*, const float*, int, int, int, int, float, uint32_t*, float*, int*) { return NOT_BUILT(c); }
}
