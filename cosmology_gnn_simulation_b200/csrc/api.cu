// extern "C" surface of libcgnn.so: argument validation, error string, precision dispatch.
#include <stdarg.h>

#include <atomic>
#include <mutex>

#include "common.cuh"
#include "mlp_common.cuh"

namespace cgnn {

static thread_local char g_error[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int num_sms() {
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
            sms = 148;
    }
    return sms;
}

// tensor-core implementations (mp_tc.cu); return CGNN_ERR_UNSUPPORTED for shapes they do not cover
int tc_mlp_fwd(MlpTask& a, int precision, void* ws, int64_t wsb, cudaStream_t s);
int64_t tc_edge_fwd_workspace(const cgnn_mlp* mlp, int64_t n, int precision);
int64_t tc_node_fwd_workspace(const cgnn_mlp* mlp, int64_t n, int precision);
int tc_mlp_bwd(MlpTask& a, const cgnn_mlp_grad* g, void* ws, int64_t wsb, int precision, cudaStream_t s);
int64_t tc_bwd_workspace(const cgnn_mlp* mlp, int64_t n, int64_t n_nodes, int k, int precision);
int64_t tc_rows_workspace(const cgnn_mlp* mlp, int64_t rows, int precision, int backward);

void set_debug_stamps(unsigned long long* buf, int tiles, int launches);

static bool is_tc(int precision) { return precision == CGNN_PREC_BF16X3 || precision == CGNN_PREC_BF16 || precision == CGNN_PREC_BF16X3_G16; }

static int run_fwd(MlpTask& a, int precision, cudaStream_t s, void* ws = nullptr, int64_t wsb = 0) {
    if (precision == CGNN_PREC_FP32) return simt_mlp_fwd(a, s);
    if (is_tc(precision)) {
        int rc = tc_mlp_fwd(a, precision, ws, wsb, s);
        // row-wise MLPs of shapes the tensor-core chain does not cover run the FP32 kernels (see cgnn.h)
        return (rc == CGNN_ERR_UNSUPPORTED && a.mode == MODE_ROWS) ? simt_mlp_fwd(a, s) : rc;
    }
    set_error("unknown precision %d", precision);
    return CGNN_ERR_INVALID;
}
static int run_bwd(MlpTask& a, const cgnn_mlp_grad* g, void* ws, int64_t wsb, int precision, cudaStream_t s) {
    if (precision == CGNN_PREC_FP32) return simt_mlp_bwd(a, g, ws, wsb, s);
    if (is_tc(precision)) {
        // the tensor-core backward covers what tc_mlp_bwd implements; everything else recomputes and
        // differentiates with the FP32 kernels (same math, higher precision)
        int rc = tc_mlp_bwd(a, g, ws, wsb, precision, s);
        return rc == CGNN_ERR_UNSUPPORTED ? simt_mlp_bwd(a, g, ws, wsb, s) : rc;
    }
    set_error("unknown precision %d", precision);
    return CGNN_ERR_INVALID;
}

}  // namespace cgnn

using namespace cgnn;

extern "C" const char* cgnn_last_error(void) { return g_error; }
extern "C" const char* cgnn_version(void) { return "cgnn 0.1 (sm_100a)"; }
extern "C" int64_t cgnn_launch_count(void) { return g_launches.load(); }
extern "C" void cgnn_debug_stamps(unsigned long long* stamps, int32_t tiles, int32_t launches) {
    set_debug_stamps(stamps, stamps ? tiles : 0, stamps ? launches : 0);
}

extern "C" int64_t cgnn_mlp_rows_workspace_bytes(const cgnn_mlp* mlp, int64_t rows, int32_t precision, int32_t backward) {
    if (mlp_validate(mlp, "cgnn_mlp_rows_workspace_bytes")) return -1;
    int64_t a = backward ? simt_mlp_bwd_workspace(mlp) : 0;
    if (precision == CGNN_PREC_FP32) return a;
    int64_t b = tc_rows_workspace(mlp, rows, precision, backward);
    return a > b ? a : b;
}

extern "C" int cgnn_mlp_rows_fwd(const cgnn_mlp* mlp, const float* x, int64_t rows, float* out, void* workspace,
                                 int64_t workspace_bytes, int32_t precision, cgnn_stream stream) {
    int rc = mlp_validate(mlp, "cgnn_mlp_rows_fwd");
    if (rc) return rc;
    CGNN_CHECK_ARG(x && out && rows >= 0, "cgnn_mlp_rows_fwd: bad arguments");
    if (rows == 0) return CGNN_OK;
    MlpTask a{};
    a.mlp = mlp_to_dev(mlp); a.mode = MODE_ROWS; a.n = rows; a.x = x; a.out = out;
    return run_fwd(a, precision, (cudaStream_t)stream, workspace, workspace_bytes);
}

extern "C" int64_t cgnn_mlp_bwd_workspace_bytes(const cgnn_mlp* mlp) {
    if (mlp_validate(mlp, "cgnn_mlp_bwd_workspace_bytes")) return -1;
    return simt_mlp_bwd_workspace(mlp);
}

extern "C" int cgnn_mlp_rows_bwd(const cgnn_mlp* mlp, const cgnn_mlp_grad* grad, const float* x, int64_t rows,
                                 const float* dout, float* dx, void* workspace, int64_t workspace_bytes,
                                 int32_t precision, cgnn_stream stream) {
    int rc = mlp_validate(mlp, "cgnn_mlp_rows_bwd");
    if (rc) return rc;
    CGNN_CHECK_ARG(grad && x && dout && rows >= 1, "cgnn_mlp_rows_bwd: bad arguments");
    MlpTask a{};
    a.mlp = mlp_to_dev(mlp); a.mode = MODE_ROWS; a.n = rows; a.x = x; a.dout = dout; a.dx = dx;
    a.need_input_grad = dx != nullptr;
    return run_bwd(a, grad, workspace, workspace_bytes, precision, (cudaStream_t)stream);
}

static int check_latent(const cgnn_mlp* mlp, int mult, const char* who) {
    CGNN_CHECK_ARG(mlp->ln_gamma != nullptr, "%s: the processor MLPs end in a LayerNorm", who);
    CGNN_CHECK_ARG(mlp->in_dim == mult * mlp->out_dim, "%s: in_dim must be %d * latent (got %d vs %d)", who, mult, mlp->in_dim, mlp->out_dim);
    CGNN_CHECK_ARG(mlp->out_dim % 4 == 0, "%s: latent width must be a multiple of 4", who);
    return CGNN_OK;
}

extern "C" int64_t cgnn_mp_edge_fwd_workspace_bytes(const cgnn_mlp* mlp, int64_t n_nodes, int32_t precision) {
    if (mlp_validate(mlp, "cgnn_mp_edge_fwd_workspace_bytes")) return -1;
    if (precision == CGNN_PREC_FP32) return 0;
    return tc_edge_fwd_workspace(mlp, n_nodes, precision);
}

extern "C" int cgnn_mp_edge_fwd(const cgnn_mlp* mlp, const float* h, const float* e_in, const int32_t* senders,
                                int64_t n, int64_t n_nodes, int32_t k, int32_t k_valid, float* e_out, float* agg_edge, void* workspace,
                                int64_t workspace_bytes, int32_t precision, cgnn_stream stream) {
    int rc = mlp_validate(mlp, "cgnn_mp_edge_fwd");
    if (rc) return rc;
    if ((rc = check_latent(mlp, 3, "cgnn_mp_edge_fwd"))) return rc;
    CGNN_CHECK_ARG(h && e_in && senders && (e_out || agg_edge) && n >= 1 && n_nodes >= n, "cgnn_mp_edge_fwd: bad arguments");
    CGNN_CHECK_ARG(k >= 1 && k <= 64, "cgnn_mp_edge_fwd: need 1 <= k <= 64");
    CGNN_CHECK_ARG(k_valid >= 0 && k_valid <= k && (k_valid == 0 || k_valid == k || is_tc(precision)),
                   "cgnn_mp_edge_fwd: k_valid must be 0 or in 1..k (padded in-degrees are a tensor-core feature; the FP32 kernels take any k)");
    MlpTask a{};
    a.mlp = mlp_to_dev(mlp); a.mode = MODE_EDGE; a.n = n; a.n_nodes = n_nodes; a.k = k; a.k_valid = k_valid; a.L = mlp->out_dim;
    a.h = h; a.e_in = e_in; a.senders = senders; a.out = e_out; a.agg_out = agg_edge;
    return run_fwd(a, precision, (cudaStream_t)stream, workspace, workspace_bytes);
}

extern "C" int cgnn_aggregate_senders(const float* h, const int32_t* senders, int64_t n, int32_t k, int32_t latent,
                                      float* agg, cgnn_stream stream) {
    CGNN_CHECK_ARG(h && senders && agg && n >= 1 && k >= 1, "cgnn_aggregate_senders: bad arguments");
    CGNN_CHECK_ARG(latent >= 4 && latent % 4 == 0, "cgnn_aggregate_senders: latent must be a multiple of 4");
    return simt_aggregate_senders(h, senders, n, k, latent, agg, (cudaStream_t)stream);
}

extern "C" int64_t cgnn_mp_node_fwd_workspace_bytes(const cgnn_mlp* mlp, int64_t n, int32_t precision) {
    if (mlp_validate(mlp, "cgnn_mp_node_fwd_workspace_bytes")) return -1;
    if (precision == CGNN_PREC_FP32) return 0;
    return tc_node_fwd_workspace(mlp, n, precision);
}

extern "C" int cgnn_mp_node_fwd(const cgnn_mlp* mlp, const float* h, const float* agg, int64_t n, float* h_out,
                                void* workspace, int64_t workspace_bytes, int32_t precision, cgnn_stream stream) {
    int rc = mlp_validate(mlp, "cgnn_mp_node_fwd");
    if (rc) return rc;
    if ((rc = check_latent(mlp, 2, "cgnn_mp_node_fwd"))) return rc;
    CGNN_CHECK_ARG(h && agg && h_out && n >= 1, "cgnn_mp_node_fwd: bad arguments");
    MlpTask a{};
    a.mlp = mlp_to_dev(mlp); a.mode = MODE_NODE; a.n = n; a.L = mlp->out_dim;
    a.h = h; a.agg = agg; a.out = h_out;
    return run_fwd(a, precision, (cudaStream_t)stream, workspace, workspace_bytes);
}

extern "C" int64_t cgnn_mp_bwd_workspace_bytes(const cgnn_mlp* mlp, int64_t n, int64_t n_nodes, int32_t k, int32_t precision) {
    if (mlp_validate(mlp, "cgnn_mp_bwd_workspace_bytes")) return -1;
    int64_t a = simt_mlp_bwd_workspace(mlp);
    if (precision == CGNN_PREC_FP32) return a;
    int64_t b = tc_bwd_workspace(mlp, n, n_nodes, k, precision);
    return a > b ? a : b;
}

extern "C" int cgnn_mp_node_bwd(const cgnn_mlp* mlp, const cgnn_mlp_grad* grad, const float* h, const float* agg,
                                const float* dh_next, int64_t n, float* dh, float* dagg, void* workspace,
                                int64_t workspace_bytes, int32_t precision, cgnn_stream stream) {
    int rc = mlp_validate(mlp, "cgnn_mp_node_bwd");
    if (rc) return rc;
    if ((rc = check_latent(mlp, 2, "cgnn_mp_node_bwd"))) return rc;
    CGNN_CHECK_ARG(grad && h && agg && dh_next && dh && dagg && n >= 1, "cgnn_mp_node_bwd: bad arguments");
    MlpTask a{};
    a.mlp = mlp_to_dev(mlp); a.mode = MODE_NODE; a.n = n; a.L = mlp->out_dim;
    a.h = h; a.agg = agg; a.dout = dh_next; a.dh = dh; a.dagg_out = dagg; a.need_input_grad = 1;
    return run_bwd(a, grad, workspace, workspace_bytes, precision, (cudaStream_t)stream);
}

extern "C" int cgnn_mp_edge_bwd(const cgnn_mlp* mlp, const cgnn_mlp_grad* grad, const float* h, const float* e_in,
                                const int32_t* senders, const int32_t* t_rowptr, const int32_t* t_perm, int64_t n,
                                int64_t n_nodes, int32_t k, int32_t k_valid,
                                const float* de_next, const float* dagg, float* de, float* dh, float* gs,
                                void* workspace, int64_t workspace_bytes, int32_t precision, cgnn_stream stream) {
    int rc = mlp_validate(mlp, "cgnn_mp_edge_bwd");
    if (rc) return rc;
    if ((rc = check_latent(mlp, 3, "cgnn_mp_edge_bwd"))) return rc;
    CGNN_CHECK_ARG(grad && h && e_in && senders && t_rowptr && t_perm && dagg && de && dh && n >= 1 && n_nodes >= n,
                   "cgnn_mp_edge_bwd: bad arguments");
    CGNN_CHECK_ARG(k >= 1 && k <= 32, "cgnn_mp_edge_bwd: need 1 <= k <= 32 (got %d)", k);      // the FP32 backward tile holds 32 rows
    CGNN_CHECK_ARG(k_valid >= 0 && k_valid <= k && (k_valid == 0 || k_valid == k || is_tc(precision)), "cgnn_mp_edge_bwd: bad k_valid");
    MlpTask a{};
    a.mlp = mlp_to_dev(mlp); a.mode = MODE_EDGE; a.n = n; a.n_nodes = n_nodes; a.k = k; a.k_valid = k_valid; a.L = mlp->out_dim;
    a.h = h; a.e_in = e_in; a.senders = senders; a.de_next = de_next; a.dagg = dagg;
    a.de = de; a.dh = dh; a.gs = gs; a.need_input_grad = 1;
    a.t_rowptr = t_rowptr; a.t_perm = t_perm;
    if (is_tc(precision)) {
        rc = tc_mlp_bwd(a, grad, workspace, workspace_bytes, precision, (cudaStream_t)stream);
        if (rc != CGNN_ERR_UNSUPPORTED) return rc;          // the tensor-core path also did the sender scatter
    }
    // FP32 kernels: gs[e] = dIn[:, :L] per edge, then the deterministic scatter over the transpose
    CGNN_CHECK_ARG(gs != nullptr, "cgnn_mp_edge_bwd: the FP32 kernels need the per-edge scratch `gs` [E][latent]");
    if ((rc = simt_mlp_bwd(a, grad, workspace, workspace_bytes, (cudaStream_t)stream))) return rc;
    return simt_scatter_to_senders(gs, 0, t_rowptr, t_perm, n_nodes, k, mlp->out_dim, dh, (cudaStream_t)stream);
}

extern "C" int cgnn_scatter_to_senders(const float* src, int32_t src_is_per_receiver, const int32_t* rowptr,
                                       const int32_t* perm, int64_t n, int32_t k, int32_t latent, float* dh,
                                       cgnn_stream stream) {
    CGNN_CHECK_ARG(src && rowptr && perm && dh && n >= 1 && k >= 1, "cgnn_scatter_to_senders: bad arguments");
    CGNN_CHECK_ARG(latent >= 4 && latent % 4 == 0, "cgnn_scatter_to_senders: latent must be a multiple of 4");
    return simt_scatter_to_senders(src, src_is_per_receiver, rowptr, perm, n, k, latent, dh, (cudaStream_t)stream);
}
