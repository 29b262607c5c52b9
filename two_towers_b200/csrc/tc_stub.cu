// tc_stub.cu -- temporary: TT_PREC_BF16 entry points until the tcgen05 kernels land.
#include "tensor_core.cuh"
namespace tt {
size_t tc_mlp_workspace(int64_t, int, int) { return 256; }
int tc_mlp_fwd(const float*, const float*, const float*, const float*, const float*, int64_t, int, int, float*, float*, float*, __nv_bfloat16*, void*, size_t, cudaStream_t) { set_error("TT_PREC_BF16 mlp not built"); return TT_ERR_UNSUPPORTED; }
int tc_mlp_bwd(const float*, const float*, const float*, const float*, const float*, const float*, int64_t, int, int, float*, float*, float*, float*, float*, void*, size_t, cudaStream_t) { set_error("TT_PREC_BF16 mlp not built"); return TT_ERR_UNSUPPORTED; }
size_t tc_inbatch_workspace(int64_t, int64_t, int) { return 256; }
int tc_inbatch_fwd(const float*, const float*, const __nv_bfloat16*, const __nv_bfloat16*, int64_t, int64_t, int, float, int64_t, float, float*, float*, float*, void*, size_t, cudaStream_t) { set_error("TT_PREC_BF16 inbatch not built"); return TT_ERR_UNSUPPORTED; }
int tc_inbatch_bwd(const float*, const float*, const __nv_bfloat16*, const __nv_bfloat16*, const float*, int64_t, int64_t, int, float, int64_t, float, const float*, float*, float*, void*, size_t, cudaStream_t) { set_error("TT_PREC_BF16 inbatch not built"); return TT_ERR_UNSUPPORTED; }
}
