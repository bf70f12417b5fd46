// placeholder until the tcgen05 path lands
#include "common.cuh"
namespace snerf {
size_t tc_workspace_bytes(const MlpDims&, const snerf_mlp_desc&, int, int, uint32_t) { return 0; }
size_t tc_packed_bytes(const snerf_mlp_desc&) { return 256; }
int tc_pack(const snerf_mlp_desc&, const float* const*, void*, cudaStream_t) { return fail(SNERF_ERR_UNSUPPORTED, "tensor path not built"); }
int tc_forward(const snerf_mlp_desc&, const float* const*, const void*, const float*, const float*, const float*, const float*, const float*, float*, float*, void*, size_t, int, int, uint32_t, cudaStream_t) { return fail(SNERF_ERR_UNSUPPORTED, "tensor path not built"); }
int tc_backward(const snerf_mlp_desc&, const float* const*, const void*, const float*, const float*, const float*, const float*, const float*, const float*, const float*, const float*, float* const*, void*, size_t, int, int, uint32_t, cudaStream_t) { return fail(SNERF_ERR_UNSUPPORTED, "tensor path not built"); }
int tc_selftest(float*, cudaStream_t) { return fail(SNERF_ERR_UNSUPPORTED, "tensor path not built"); }
}
extern "C" int snerf_has_tensor_path(void) { return 0; }
