// Tensor-core (tcgen05) implementations of the MLP tiles.  Placeholder dispatch until the
// kernels land: reports CGNN_ERR_UNSUPPORTED so callers fail loudly instead of falling back.
#include "common.cuh"
#include "mlp_common.cuh"

namespace cgnn {

int tc_mlp_fwd(MlpTask& a, int precision, void* ws, int64_t wsb, cudaStream_t s) {
    (void)a; (void)precision; (void)ws; (void)wsb; (void)s;
    set_error("tensor-core precision modes are not built yet");
    return CGNN_ERR_UNSUPPORTED;
}
int tc_mlp_bwd(MlpTask& a, const cgnn_mlp_grad* g, void* ws, int64_t wsb, int precision, cudaStream_t s) {
    (void)a; (void)g; (void)ws; (void)wsb; (void)precision; (void)s;
    set_error("tensor-core precision modes are not built yet");
    return CGNN_ERR_UNSUPPORTED;
}
int64_t tc_mlp_bwd_workspace(const cgnn_mlp* mlp) { (void)mlp; return 0; }
int64_t tc_edge_fwd_workspace(const cgnn_mlp* mlp, int64_t n, int precision) { (void)mlp; (void)n; (void)precision; return 0; }

}  // namespace cgnn
