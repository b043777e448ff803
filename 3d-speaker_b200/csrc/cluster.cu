// Spectral-clustering back end (placeholder entry points; implemented in cluster kernels).
#include "common.cuh"
using namespace spk;
extern "C" int spk_affinity_laplacian(const float *, int64_t, int64_t, int64_t, float *, void *, int64_t, void *) {
    set_error("not implemented"); return SPK_ERR_UNSUPPORTED; }
extern "C" int64_t spk_affinity_workspace_bytes(int64_t, int64_t) { return 0; }
extern "C" int spk_eig_smallest(const float *, int64_t, int32_t, float *, float *, void *, int64_t, void *) {
    set_error("not implemented"); return SPK_ERR_UNSUPPORTED; }
extern "C" int64_t spk_eig_workspace_bytes(int64_t, int32_t) { return 0; }
extern "C" int spk_kmeans(const float *, int64_t, int32_t, int32_t, const float *, int32_t, float, int32_t *, float *,
                          void *, int64_t, void *) { set_error("not implemented"); return SPK_ERR_UNSUPPORTED; }
extern "C" int64_t spk_kmeans_workspace_bytes(int64_t, int32_t, int32_t) { return 0; }
extern "C" int spk_cosine_pairs(const float *, int64_t, int64_t, const int32_t *, const int32_t *, int64_t, float *, void *) {
    set_error("not implemented"); return SPK_ERR_UNSUPPORTED; }
