// hop_cabi.cu -- extern "C" surface of libhop_b200.so (see include/hop_b200.h for the contract).
#include <cstring>
#include <cstdlib>
#include <mutex>
#include <string>
#include <vector>

#include "hop_common.cuh"
#include "hop_select_body.cuh"
#include "../../include/hop_b200.h"

namespace hop {
int dispatch_select_generic(int d, int m, int mode, const SelectArgs& p, cudaStream_t st);
long tpp_min_batch(long set_to);
int dispatch_select_fused(int n, int m, const FusedArgs& p, cudaStream_t st);
int dispatch_rollout(int B, int sys, const double* params_host, int N, const double* x0, const double* U, long ustride,
                     double max_norm, double* X, cudaStream_t st);
int dispatch_linearize(int B, int sys, const double* params_host, int N, const double* X, const double* U, long ustride,
                       int central, double epsx, double epsu, double relx, double relu, int f0_from_x, const int* skip,
                       double* A, double* Bm, cudaStream_t st);
extern int g_linearize_variant;
extern int g_backward_variant;
extern int g_linesearch_variant;
extern int g_fused_small_variant;
extern int g_generic_pre;
extern int g_generic_nodiag;
struct DdpConst {
    const double *xg, *w, *u_ref, *Q, *R, *Qf;
    unsigned wrap_mask;
};
int dispatch_backward_linesearch(int sys, int B, const double* params_host, int N, const double* A, const double* Bm,
                                 const double* X, const double* U, const DdpConst& c, const int* T, const double* lm,
                                 const int* done, double* kl, double* Kl, int* ok, int* bw_err, double* Xn, double* Un,
                                 double* Jn, int* acc, bool ordered, cudaEvent_t mid, cudaStream_t st);
int dispatch_bruteforce(int n, int m, int B, int N, int T_max, const double* A, const double* Bm, const double* X, const double* U,
                        long ustride, const DdpConst& c, double lm, double* J_out, int* status, cudaStream_t st);
int dispatch_cost(int n, int m, int B, int N, const double* X, const double* U, const DdpConst& c, const int* T, double* J,
                  cudaStream_t st);
int dispatch_linesearch(int sys, int B, const double* params_host, int N, const double* X, const double* U, const DdpConst& c,
                        const int* T, const double* kl, const double* Kl, const int* ok, double* Xn, double* Un, double* Jn,
                        int* acc, cudaStream_t st);
int launch_init_state(int B, double lm_init, double* lm, int* done, int* n_hist, int* status_out, cudaStream_t st);
int launch_tile_u(int B, int N, int m, const double* u_ref, double* U, cudaStream_t st);
int launch_after_select(int B, const int* sel_status, int* done, int* status_out, cudaStream_t st);
int launch_warm_update(int B, int cap, const int* T_sel, const int* ok, const int* acc, const double* Jn, const int* bw_err,
                       int* done, int* T_bar, double* J_hist, int* T_hist, int* n_hist, int* copy, int* status_out,
                       cudaStream_t st);
int launch_ddp_update(int B, int cap, const int* T_sel, const int* ok, const int* acc, const double* Jn, const int* bw_err,
                      int* done, int* T_bar, double* lm, double* J_hist, int* T_hist, int* n_hist, int* copy,
                      int* status_out, int* n_active, cudaStream_t st);
int launch_copy_accepted(int B, size_t per_x, size_t per_u, const int* copy, const double* Xn, const double* Un, double* X,
                         double* U, cudaStream_t st);
int launch_finalize(int B, int cap, const int* n_hist, const int* T_hist, const int* T_bar, int* T_star, cudaStream_t st);
int launch_chol(int B, int d, int c, const double* A, const double* Bm, double* X, double jitter, int max_tries, int* status,
                cudaStream_t st);
int launch_affine_residuals(int B, int sys, const double* params_host, int N, const double* X, const double* U, long ustride,
                            double* a, cudaStream_t st);
int launch_build_augmented(int B, int N, int n, int m, const double* A, const double* Bm, const double* a, const double* X,
                           const double* U, long ustride, const double* xg, const double* w, const double* u_ref,
                           const double* Q, unsigned wrap_mask, double q_reg, double rho_reg, double* A_aug, double* B_aug,
                           double* Q_aug, cudaStream_t st);
int launch_build_terminal(int B, int N, int n, const double* X, const double* xg, const double* Qf, unsigned wrap_mask,
                          double rho_reg, double* QT, cudaStream_t st);
int sys_dims(int sys, int* n, int* m);
int stream_alloc(void** ptr, size_t bytes, cudaStream_t st, const char* what);
void stream_free(void* ptr, cudaStream_t st);
int launch_scan_consts(int B, int d, int m, const double* R_inv1, double* R_inv, double* z0, cudaStream_t st);

static thread_local std::string g_err;
void set_last_error(const char* msg) { g_err = msg ? msg : ""; }
int report_cuda(cudaError_t e, const char* where) {
    if (e == cudaSuccess) return 0;
    g_err = std::string(where) + ": " + cudaGetErrorString(e);
    return (int)e;
}
static int need_device() {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        (void)cudaGetLastError();
        set_last_error("no CUDA device visible: libhop_b200 has no CPU fallback");
        return HOP_E_NO_DEVICE;
    }
    return 0;
}
static const double kJitter = 1e-9;   // utils.py:69 defaults
static const int kMaxTries = 8;
}  // namespace hop

using namespace hop;

extern "C" {

int hop_abi_version(void) { return HOP_ABI_VERSION; }
const char* hop_version(void) { return "hop_b200 0.1 (sm_100a, fp64)"; }
const char* hop_last_error_string(void) { return g_err.c_str(); }
int hop_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { (void)cudaGetLastError(); return 0; }
    return n;
}
int hop_select_supported(int d, int m) {
    return (d == 3 && m == 1) || (d == 4 && m == 2) || (d == 5 && m == 1) || (d == 12 && m == 4) || (d == 13 && m == 4);
}
int hop_select_supported_mode(int d, int m, int mode) {
    if (mode == HOP_MODE_EXACT || mode == HOP_MODE_FP32) return d >= 1 && d <= 16 && m >= 1 && m <= 16;
    if (mode == HOP_MODE_SCAN) return (d == 12 || d == 13) && m == 4;
    if (mode == HOP_MODE_FAST || mode == HOP_MODE_GJ) return hop_select_supported(d, m);
    return 0;
}

int hop_select_f64(int B, int N, int d, int m, int T_min, int T_max, const double* A_aug, const double* B_aug,
                   const double* Q_aug, const double* R_inv, long rinv_step_stride, const double* z0, const double* QT,
                   const double* w_explicit, int mode, double* J_out, int* Tstar_out, double* Jstar_out, int* status,
                   void* stream) {
    if (B < 0 || N < 1 || T_min < 1 || T_max < T_min || T_max > N ||
        mode < HOP_MODE_EXACT || mode > HOP_MODE_FP32 ||
        (rinv_step_stride != 0 && rinv_step_stride != (long)m * m)) {
        set_last_error("hop_select_f64: bad argument (need 1 <= T_min <= T_max <= N, rinv_step_stride in {0, m*m})");
        return HOP_E_BADARG;
    }
    if (int rc = need_device()) return rc;
    if (B == 0) return 0;
    SelectArgs p{B, N, T_min, T_max, kJitter, kMaxTries, A_aug, B_aug, Q_aug, R_inv, z0, QT, rinv_step_stride, w_explicit,
                 J_out, Tstar_out, Jstar_out, status, nullptr, nullptr, nullptr, g_generic_nodiag};
    return dispatch_select_generic(d, m, mode, p, (cudaStream_t)stream);
}

int hop_select_fused_f64(int B, int N, int n, int m, int T_min, int T_max, const double* A, const double* Bm,
                         const double* a_resid, const double* X, const double* U, long u_batch_stride,
                         const double* xg, const double* w,
                         const double* u_ref, const double* Q, const double* R, const double* Qf, unsigned wrap_mask,
                         double q_reg, double rho_reg, int mode, double* J_out, int* Tstar_out, double* Jstar_out,
                         int* status, void* stream) {
    if (B < 0 || N < 1 || T_min < 1 || T_max < T_min || T_max > N || (mode != HOP_MODE_EXACT && mode != HOP_MODE_FAST && mode != HOP_MODE_GJ && mode != HOP_MODE_SCAN)) {
        set_last_error("hop_select_fused_f64: bad argument (need 1 <= T_min <= T_max <= N, mode in {EXACT, FAST, GJ, SCAN})");
        return HOP_E_BADARG;
    }
    if (int rc = need_device()) return rc;
    if (B == 0) return 0;
    if (mode == HOP_MODE_SCAN) {
        // the scan kernel works at the LQR boundary: materialise the augmented blocks (augmented.py:10-87) in a stream-ordered
        // scratch allocation, then run the chunked parallel scan over the horizon.  Small batches only (that is where it pays).
        if (!hop_select_supported_mode(n + 1, m, HOP_MODE_SCAN)) {
            set_last_error("hop_select_fused_f64: HOP_MODE_SCAN is instantiated for (n, m) = (11,4) and (12,4) only");
            return HOP_E_UNSUPPORTED_DIMS;
        }
        cudaStream_t st = (cudaStream_t)stream;
        const int d = n + 1;
        const size_t dd = (size_t)B * N * d * d, dm = (size_t)B * N * d * m;
        const size_t total = sizeof(double) * (3 * dd + dm + (size_t)B * m * m + (size_t)B * d + (size_t)m * m) + sizeof(int);
        void* ws = nullptr;
        if (int rc = stream_alloc(&ws, total, st, "cudaMallocAsync(scan workspace)")) return rc;
        double* A_aug = (double*)ws; double* Q_aug = A_aug + dd; double* QT = Q_aug + dd; double* B_aug = QT + dd;
        double* R_inv = B_aug + dm; double* z0 = R_inv + (size_t)B * m * m; double* R1 = z0 + (size_t)B * d;
        int* st1 = (int*)(R1 + m * m);
        int rc = launch_build_augmented(B, N, n, m, A, Bm, a_resid, X, U, u_batch_stride, xg, w, u_ref, Q, wrap_mask, q_reg,
                                        rho_reg, A_aug, B_aug, Q_aug, st);
        if (!rc) rc = launch_build_terminal(B, N, n, X, xg, Qf, wrap_mask, rho_reg, QT, st);
        if (!rc) rc = launch_chol(1, m, 0, R, nullptr, R1, kJitter, kMaxTries, st1, st);              // augmented.py:23
        if (!rc) rc = launch_scan_consts(B, d, m, R1, R_inv, z0, st);                                 // augmented.py:59
        if (!rc) {
            SelectArgs q{B, N, T_min, T_max, kJitter, kMaxTries, A_aug, B_aug, Q_aug, R_inv, z0, QT, 0, nullptr,
                         J_out, Tstar_out, Jstar_out, status, nullptr, nullptr, nullptr, g_generic_nodiag};
            rc = dispatch_select_generic(d, m, HOP_MODE_SCAN, q, st);
        }
        stream_free(ws, st);
        return rc;
    }
    FusedArgs p{B, N, T_min, T_max, kJitter, kMaxTries, A, Bm, a_resid, X, U, u_batch_stride, xg, w, u_ref, Q, R, Qf,
                wrap_mask, q_reg, rho_reg, mode, nullptr, J_out, Tstar_out, Jstar_out, status};
    return dispatch_select_fused(n, m, p, (cudaStream_t)stream);
}

int hop_rollout_f64(int B, int sys, const double* params_host, int N, const double* x0, const double* U,
                    long u_batch_stride, double max_state_norm, double* X, void* stream) {
    if (B < 0 || N < 1 || !params_host) { set_last_error("hop_rollout_f64: bad argument"); return HOP_E_BADARG; }
    if (int rc = need_device()) return rc;
    if (B == 0) return 0;
    return dispatch_rollout(B, sys, params_host, N, x0, U, u_batch_stride, max_state_norm, X, (cudaStream_t)stream);
}

int hop_linearize_f64(int B, int sys, const double* params_host, int N, const double* X, const double* U,
                      long u_batch_stride, int central, double epsx, double epsu, double relx, double relu, double* A,
                      double* Bm, void* stream) {
    if (B < 0 || N < 1 || !params_host) { set_last_error("hop_linearize_f64: bad argument"); return HOP_E_BADARG; }
    if (int rc = need_device()) return rc;
    if (B == 0) return 0;
    return dispatch_linearize(B, sys, params_host, N, X, U, u_batch_stride, central, epsx, epsu, relx, relu, 0, nullptr, A, Bm,
                              (cudaStream_t)stream);
}

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

unsigned long long hop_select_from_x0_workspace_bytes(int B, int N, int n, int m) {
    const size_t sX = align256(sizeof(double) * (size_t)B * (N + 1) * n);
    const size_t sA = align256(sizeof(double) * (size_t)B * N * n * n);
    const size_t sB = align256(sizeof(double) * (size_t)B * N * n * m);
    return (unsigned long long)(sX + sA + sB);
}

int hop_select_from_x0_f64(int B, int sys, const double* params_host, int N, int T_min, int T_max, const double* x0,
                           const double* U, long u_batch_stride, const double* xg, const double* w,
                           const double* u_ref, const double* Q, const double* R, const double* Qf, unsigned wrap_mask,
                           int central, int mode, void* workspace, unsigned long long workspace_bytes, double* J_out,
                           int* Tstar_out, double* Jstar_out, int* status, void* stream) {
    int n = 0, m = 0;
    if (sys_dims(sys, &n, &m)) { set_last_error("hop_select_from_x0_f64: unknown system id"); return HOP_E_BADARG; }
    if (workspace_bytes < hop_select_from_x0_workspace_bytes(B, N, n, m) || (!workspace && B > 0)) {
        set_last_error("hop_select_from_x0_f64: workspace too small");
        return HOP_E_WORKSPACE;
    }
    if (B == 0) return 0;
    char* ws = (char*)workspace;
    double* X = (double*)ws;
    double* A = (double*)(ws + align256(sizeof(double) * (size_t)B * (N + 1) * n));
    double* Bm = (double*)((char*)A + align256(sizeof(double) * (size_t)B * N * n * n));
    int rc = hop_rollout_f64(B, sys, params_host, N, x0, U, u_batch_stride, 1e6, X, stream);            // solver.py:42
    if (rc) return rc;
    // X comes from the rollout above, so F(X_k, U_k) = X[k+1] bit for bit: the linearisation may read f0 from it
    rc = dispatch_linearize(B, sys, params_host, N, X, U, u_batch_stride, central, 1e-5, 1e-5, 1e-6, 1e-6, 1, nullptr, A, Bm,
                            (cudaStream_t)stream);                                                       // linearization.py:177,216
    if (rc) return rc;
    // a_resid = NULL: on a trajectory produced by the rollout above F(X_k,U_k) - X_{k+1} is exactly 0
    return hop_select_fused_f64(B, N, n, m, T_min, T_max, A, Bm, nullptr, X, U, u_batch_stride, xg, w, u_ref, Q, R, Qf,
                                wrap_mask, 1e-9, 1e-12, mode, J_out, Tstar_out, Jstar_out, status, stream);
}

// ---- utilities of the reference API (utils.py, linearization.py:269, augmented.py) -----------------------
int hop_chol_inv_f64(int B, int d, const double* A, double* X, double jitter, int max_tries, int* status, void* stream) {
    if (d < 1 || d > 16 || B < 0) { set_last_error("hop_chol_inv_f64: need 1 <= d <= 16"); return HOP_E_BADARG; }
    if (int rc = need_device()) return rc;
    if (B == 0) return 0;
    return launch_chol(B, d, 0, A, nullptr, X, jitter, max_tries, status, (cudaStream_t)stream);
}
int hop_chol_solve_f64(int B, int d, int c, const double* A, const double* Bm, double* X, double jitter, int max_tries,
                       int* status, void* stream) {
    if (d < 1 || d > 16 || c < 1 || B < 0) { set_last_error("hop_chol_solve_f64: need 1 <= d <= 16, c >= 1"); return HOP_E_BADARG; }
    if (int rc = need_device()) return rc;
    if (B == 0) return 0;
    return launch_chol(B, d, c, A, Bm, X, jitter, max_tries, status, (cudaStream_t)stream);
}
int hop_affine_residuals_f64(int B, int sys, const double* params_host, int N, const double* X, const double* U,
                             long u_batch_stride, double* a_out, void* stream) {
    if (int rc = need_device()) return rc;
    if (B <= 0) return B == 0 ? 0 : HOP_E_BADARG;
    return launch_affine_residuals(B, sys, params_host, N, X, U, u_batch_stride, a_out, (cudaStream_t)stream);
}
int hop_build_augmented_f64(int B, int N, int n, int m, const double* A, const double* Bm, const double* a_resid,
                            const double* X, const double* U, long u_batch_stride, const double* xg, const double* w,
                            const double* u_ref, const double* Q, unsigned wrap_mask, double q_reg, double rho_reg,
                            double* A_aug, double* B_aug, double* Q_aug, void* stream) {
    if (n < 1 || n > 15 || m < 1 || m > 8) { set_last_error("hop_build_augmented_f64: need n <= 15, m <= 8"); return HOP_E_BADARG; }
    if (int rc = need_device()) return rc;
    if (B <= 0) return B == 0 ? 0 : HOP_E_BADARG;
    return launch_build_augmented(B, N, n, m, A, Bm, a_resid, X, U, u_batch_stride, xg, w, u_ref, Q, wrap_mask, q_reg, rho_reg,
                                  A_aug, B_aug, Q_aug, (cudaStream_t)stream);
}
int hop_build_terminal_f64(int B, int N, int n, const double* X, const double* xg, const double* Qf, unsigned wrap_mask,
                           double rho_reg, double* QT, void* stream) {
    if (n < 1 || n > 15) { set_last_error("hop_build_terminal_f64: need n <= 15"); return HOP_E_BADARG; }
    if (int rc = need_device()) return rc;
    if (B <= 0) return B == 0 ? 0 : HOP_E_BADARG;
    return launch_build_terminal(B, N, n, X, xg, Qf, wrap_mask, rho_reg, QT, (cudaStream_t)stream);
}

// ---- HOP-DDP pieces and the batched solver loop --------------------------------------------------------
int hop_cost_f64(int B, int N, int n, int m, const double* X, const double* U, const double* xg, const double* w,
                 const double* u_ref, const double* Q, const double* R, const double* Qf, unsigned wrap_mask,
                 const int* T_star, double* J_out, void* stream) {
    if (int rc = need_device()) return rc;
    if (B <= 0) return B == 0 ? 0 : HOP_E_BADARG;
    if (N < 1 || !X || !U || !xg || !w || !u_ref || !Q || !R || !Qf || !T_star || !J_out) {
        set_last_error("hop_cost_f64: null pointer or N < 1");
        return HOP_E_BADARG;
    }
    DdpConst c{xg, w, u_ref, Q, R, Qf, wrap_mask};
    return dispatch_cost(n, m, B, N, X, U, c, T_star, J_out, (cudaStream_t)stream);
}

int hop_bruteforce_jt_f64(int B, int N, int n, int m, int T_max, const double* A, const double* Bm, const double* X,
                          const double* U, long u_batch_stride, const double* xg, const double* w, const double* u_ref,
                          const double* Q, const double* R, const double* Qf, unsigned wrap_mask, double lm_lambda,
                          double* J_out, int* status, void* stream) {
    if (B < 0 || N < 1 || T_max < 1 || T_max > N) { set_last_error("hop_bruteforce_jt_f64: need 1 <= T_max <= N"); return HOP_E_BADARG; }
    if (int rc = need_device()) return rc;
    if (B == 0) return 0;
    DdpConst c{xg, w, u_ref, Q, R, Qf, wrap_mask};
    return dispatch_bruteforce(n, m, B, N, T_max, A, Bm, X, U, u_batch_stride, c, lm_lambda, J_out, status, (cudaStream_t)stream);
}

int hop_backward_linesearch_f64(int B, int sys, const double* params_host, int N, const double* A, const double* Bm,
                                const double* X, const double* U, const double* xg, const double* w, const double* u_ref,
                                const double* Q, const double* R, const double* Qf, unsigned wrap_mask, const int* T_star,
                                const double* lm, double* k_out, double* K_out, int* ok_out, int* err_out, double* X_new,
                                double* U_new, double* J_new, int* accepted, void* stream) {
    if (int rc = need_device()) return rc;
    if (B <= 0) return B == 0 ? 0 : HOP_E_BADARG;
    if (N < 1 || !params_host || !A || !Bm || !X || !U || !xg || !w || !u_ref || !Q || !R || !Qf || !T_star || !lm || !k_out ||
        !K_out || !ok_out || !err_out || !X_new || !U_new || !J_new || !accepted) {
        set_last_error("hop_backward_linesearch_f64: null pointer or N < 1");
        return HOP_E_BADARG;
    }
    DdpConst c{xg, w, u_ref, Q, R, Qf, wrap_mask};
    // the stand-alone entry point is the API-parity path: the reference's summation order
    return dispatch_backward_linesearch(sys, B, params_host, N, A, Bm, X, U, c, T_star, lm, nullptr, k_out, K_out, ok_out,
                                        err_out, X_new, U_new, J_new, accepted, true, nullptr, (cudaStream_t)stream);
}

int hop_linesearch_f64(int B, int sys, const double* params_host, int N, const double* X, const double* U, const double* xg,
                       const double* w, const double* u_ref, const double* Q, const double* R, const double* Qf,
                       unsigned wrap_mask, const int* T_star, const double* k_list, const double* K_list, const int* ok,
                       double* X_new, double* U_new, double* J_new, int* accepted, void* stream) {
    if (int rc = need_device()) return rc;
    if (B <= 0) return B == 0 ? 0 : HOP_E_BADARG;
    if (N < 1 || !params_host || !X || !U || !xg || !w || !u_ref || !Q || !R || !Qf || !T_star || !k_list || !K_list || !X_new ||
        !U_new || !J_new || !accepted) {
        set_last_error("hop_linesearch_f64: null pointer or N < 1");
        return HOP_E_BADARG;
    }
    DdpConst c{xg, w, u_ref, Q, R, Qf, wrap_mask};
    return dispatch_linesearch(sys, B, params_host, N, X, U, c, T_star, k_list, K_list, ok, X_new, U_new, J_new, accepted,
                               (cudaStream_t)stream);
}

namespace {
struct IlqrWs {
    double *A, *Bm, *Xn, *Un, *kl, *Kl, *lm, *Jn, *Jstar;
    int *T_bar, *T_sel, *ok, *acc, *bw_err, *done, *copy, *sel_status, *n_active;
    size_t total;
};
IlqrWs ilqr_layout(char* base, int B, int N, int n, int m) {
    IlqrWs w;
    size_t off = 0;
    auto take = [&](size_t bytes) { char* q = base ? base + off : nullptr; off += align256(bytes); return q; };
    w.A = (double*)take(sizeof(double) * (size_t)B * N * n * n);
    w.Bm = (double*)take(sizeof(double) * (size_t)B * N * n * m);
    w.Xn = (double*)take(sizeof(double) * (size_t)B * (N + 1) * n);
    w.Un = (double*)take(sizeof(double) * (size_t)B * N * m);
    w.kl = (double*)take(sizeof(double) * (size_t)B * N * m);
    w.Kl = (double*)take(sizeof(double) * (size_t)B * N * m * n);
    w.lm = (double*)take(sizeof(double) * (size_t)B);
    w.Jn = (double*)take(sizeof(double) * (size_t)B);
    w.Jstar = (double*)take(sizeof(double) * (size_t)B);
    w.T_bar = (int*)take(sizeof(int) * (size_t)B);
    w.T_sel = (int*)take(sizeof(int) * (size_t)B);
    w.ok = (int*)take(sizeof(int) * (size_t)B);
    w.acc = (int*)take(sizeof(int) * (size_t)B);
    w.bw_err = (int*)take(sizeof(int) * (size_t)B);
    w.done = (int*)take(sizeof(int) * (size_t)B);
    w.copy = (int*)take(sizeof(int) * (size_t)B);
    w.sel_status = (int*)take(sizeof(int) * (size_t)B);
    w.n_active = (int*)take(sizeof(int) * 4);
    w.total = off;
    return w;
}
}  // namespace

unsigned long long hop_ilqr_workspace_bytes(int B, int N, int n, int m) {
    return (unsigned long long)ilqr_layout(nullptr, B, N, n, m).total;
}

int hop_ilqr_timeopt_f64(int B, int sys, const double* params_host, int N, int T_min, int T_max, const double* x0,
                         const double* U_init, const double* xg, const double* w, const double* u_ref, const double* Q,
                         const double* R, const double* Qf, unsigned wrap_mask, int max_iter, double lm_init, int central,
                         int mode, void* workspace, unsigned long long workspace_bytes, double* X, double* U,
                         double* J_hist, int* T_hist, int* n_hist, double* J_curve, int* T_star, int* status,
                         int* iters_run_host, double* timers_host, void* stream) {
    int n = 0, m = 0;
    if (sys_dims(sys, &n, &m)) { set_last_error("hop_ilqr_timeopt_f64: unknown system id"); return HOP_E_BADARG; }
    if (B < 0 || N < 1 || T_min < 1 || T_max < T_min || T_max > N || max_iter < 0 ||
        (mode != HOP_MODE_EXACT && mode != HOP_MODE_FAST && mode != HOP_MODE_GJ)) {
        set_last_error("hop_ilqr_timeopt_f64: bad argument (need 1 <= T_min <= T_max <= N, max_iter >= 0, mode in {EXACT, FAST, GJ})");
        return HOP_E_BADARG;
    }
    if (int rc = need_device()) return rc;
    if (B == 0) return 0;
    if (!workspace || workspace_bytes < hop_ilqr_workspace_bytes(B, N, n, m)) {
        set_last_error("hop_ilqr_timeopt_f64: workspace too small");
        return HOP_E_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    IlqrWs ws = ilqr_layout((char*)workspace, B, N, n, m);
    const int cap = max_iter + 1;
    const long ustride = (long)N * m;
    DdpConst c{xg, w, u_ref, Q, R, Qf, wrap_mask};
    int rc;
#define HOP_TRY(x) do { rc = (x); if (rc) return rc; } while (0)
    auto select = [&](const int* skip) -> int {
        FusedArgs p{B, N, T_min, T_max, kJitter, kMaxTries, ws.A, ws.Bm, nullptr, X, U, ustride, xg, w, u_ref, Q, R, Qf,
                    wrap_mask, 1e-9, 1e-12, mode, skip, J_curve, ws.T_sel, ws.Jstar, ws.sel_status};
        return dispatch_select_fused(n, m, p, st);
    };
    // optional per-phase device timing (the reference's timers dict: linearize / select / backward / forward): five events per
    // outer iteration, all read after the loop -- nothing here makes the host wait for the device
    const int n_sets = max_iter + 1;
    struct Events {   // destroyed on every exit path
        std::vector<cudaEvent_t> e;
        cudaEvent_t done[2] = {nullptr, nullptr};
        ~Events() {
            for (auto& x : e) if (x) cudaEventDestroy(x);
            for (auto& x : done) if (x) cudaEventDestroy(x);
        }
    } evs;
    double tsum[4] = {0.0, 0.0, 0.0, 0.0};
    if (timers_host) {
        evs.e.assign((size_t)5 * n_sets, nullptr);
        for (auto& e : evs.e) HOP_TRY(report_cuda(cudaEventCreate(&e), "cudaEventCreate"));
    }
    for (auto& e : evs.done) HOP_TRY(report_cuda(cudaEventCreateWithFlags(&e, cudaEventDisableTiming), "cudaEventCreate"));
    int set = 0;                                                                     // event set of the current iteration
    auto ev = [&](int i) { return evs.e[(size_t)5 * set + i]; };
    auto mark = [&](int i) { if (timers_host) cudaEventRecord(ev(i), st); };
    // pinned landing zone of the per-iteration "instances still active" counter (one per host thread, kept for the process)
    static thread_local int* h_active = nullptr;
    if (!h_active) HOP_TRY(report_cuda(cudaHostAlloc((void**)&h_active, 2 * sizeof(int), cudaHostAllocDefault), "cudaHostAlloc"));
    HOP_TRY(launch_init_state(B, lm_init, ws.lm, ws.done, n_hist, status, st));
    if (U_init) HOP_TRY(report_cuda(cudaMemcpyAsync(U, U_init, sizeof(double) * (size_t)B * N * m, cudaMemcpyDeviceToDevice, st), "copy U_init"));
    else HOP_TRY(launch_tile_u(B, N, m, u_ref, U, st));                                                  // solver.py:480-481
    HOP_TRY(dispatch_rollout(B, sys, params_host, N, x0, U, ustride, 1e6, X, st));                       // :492
    mark(0);
    HOP_TRY(dispatch_linearize(B, sys, params_host, N, X, U, ustride, central, 1e-5, 1e-5, 1e-6, 1e-6, 1, nullptr, ws.A, ws.Bm, st));
    mark(1);
    HOP_TRY(select(nullptr));                                                                            // :516-522
    HOP_TRY(launch_after_select(B, ws.sel_status, ws.done, status, st));
    mark(2);
    HOP_TRY(dispatch_backward_linesearch(sys, B, params_host, N, ws.A, ws.Bm, X, U, c, ws.T_sel, ws.lm, ws.done, ws.kl, ws.Kl,
                                         ws.ok, ws.bw_err, ws.Xn, ws.Un, ws.Jn, ws.acc, mode == HOP_MODE_EXACT, timers_host ? ev(3) : nullptr, st));   // :541-551
    HOP_TRY(launch_warm_update(B, cap, ws.T_sel, ws.ok, ws.acc, ws.Jn, ws.bw_err, ws.done, ws.T_bar, J_hist, T_hist, n_hist,
                               ws.copy, status, st));
    HOP_TRY(launch_copy_accepted(B, (size_t)(N + 1) * n, (size_t)N * m, ws.copy, ws.Xn, ws.Un, X, U, st));
    mark(4);
    // Outer loop (:564).  The early exit ("every instance has stopped") is decided ONE ITERATION LATE: iteration `it` is
    // enqueued before the host looks at the counter of iteration it - 1, so the stream never drains while the host waits.
    // When that counter is zero the already enqueued iteration finds every instance done and changes nothing.
    int iters = 0, enqueued = 0;
    for (int it = 0; it < max_iter; ++it) {                                                              // :564
        const int slot = it & 1;
        set = it + 1;
        ++enqueued;
        mark(0);
        HOP_TRY(dispatch_linearize(B, sys, params_host, N, X, U, ustride, central, 1e-5, 1e-5, 1e-6, 1e-6, 1, ws.done, ws.A, ws.Bm, st));
        mark(1);
        HOP_TRY(select(ws.done));                                                                        // :581-590
        HOP_TRY(launch_after_select(B, ws.sel_status, ws.done, status, st));
        mark(2);
        HOP_TRY(dispatch_backward_linesearch(sys, B, params_host, N, ws.A, ws.Bm, X, U, c, ws.T_sel, ws.lm, ws.done, ws.kl,
                                             ws.Kl, ws.ok, ws.bw_err, ws.Xn, ws.Un, ws.Jn, ws.acc, mode == HOP_MODE_EXACT, timers_host ? ev(3) : nullptr, st)); // :594-604
        HOP_TRY(report_cuda(cudaMemsetAsync(ws.n_active + slot, 0, sizeof(int), st), "memset n_active"));
        HOP_TRY(launch_ddp_update(B, cap, ws.T_sel, ws.ok, ws.acc, ws.Jn, ws.bw_err, ws.done, ws.T_bar, ws.lm, J_hist, T_hist,
                                  n_hist, ws.copy, status, ws.n_active + slot, st));                      // :735-748
        HOP_TRY(launch_copy_accepted(B, (size_t)(N + 1) * n, (size_t)N * m, ws.copy, ws.Xn, ws.Un, X, U, st));
        mark(4);
        HOP_TRY(report_cuda(cudaMemcpyAsync(h_active + slot, ws.n_active + slot, sizeof(int), cudaMemcpyDeviceToHost, st), "read n_active"));
        HOP_TRY(report_cuda(cudaEventRecord(evs.done[slot], st), "cudaEventRecord"));
        iters = it + 1;
        if (it >= 1) {                                                                                   // look at iteration it - 1
            HOP_TRY(report_cuda(cudaEventSynchronize(evs.done[slot ^ 1]), "hop_ilqr_timeopt_f64"));
            if (h_active[slot ^ 1] == 0) { iters = it; break; }
        }
    }
    HOP_TRY(report_cuda(cudaStreamSynchronize(st), "hop_ilqr_timeopt_f64"));
    if (timers_host)
        for (int q = 0; q <= enqueued; ++q)
            for (int i = 0; i < 4; ++i) {
                float ms = 0.f;
                cudaEventElapsedTime(&ms, evs.e[(size_t)5 * q + i], evs.e[(size_t)5 * q + i + 1]);
                tsum[i] += 1e-3 * ms;
            }
    HOP_TRY(launch_finalize(B, cap, n_hist, T_hist, ws.T_bar, T_star, st));
#undef HOP_TRY
    if (iters_run_host) *iters_run_host = iters;
    if (timers_host)
        for (int i = 0; i < 4; ++i) timers_host[i] = tsum[i];
    return 0;
}

// ---- FP64 pipe probe (roofline denominator; MEASURED_PEAKS.json has no FP64 figure) ---------------
__global__ void k_probe_dfma(int iters, double seed, double* sink) {
    double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 1.0 - 1e-9, c = 1e-9;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
            a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
        }
    }
    const double r = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
    if (r == 12345.678) sink[0] = r;   // never true; keeps the chain alive
}

int hop_test_set_backward_variant(int variant) {
    const int old = g_backward_variant;
    if (variant >= 0 && variant <= 3) g_backward_variant = variant;
    return old;
}

int hop_test_set_linesearch_variant(int variant) {
    const int old = g_linesearch_variant;
    if (variant == 0 || variant == 1) g_linesearch_variant = variant;
    return old;
}

int hop_test_set_fused_small_variant(int variant) {
    const int old = g_fused_small_variant;
    if (variant >= 0 && variant <= 3) g_fused_small_variant = variant;
    return old;
}

int hop_test_set_generic_diag(int on) {
    const int old = g_generic_nodiag ? 0 : 1;
    if (on == 0 || on == 1) g_generic_nodiag = on ? 0 : 1;
    return old;
}

int hop_test_set_generic_pre(int on) {
    const int old = g_generic_pre;
    if (on >= -1 && on <= 1) g_generic_pre = on;
    return old;
}

long hop_test_set_tpp_min_batch(long min_batch) { return hop::tpp_min_batch(min_batch); }

int hop_test_set_linearize_variant(int variant) {
    const int old = g_linearize_variant;
    if (variant >= 0 && variant <= 2) g_linearize_variant = variant;
    return old;
}

int hop_probe_fp64_tflops(int iters, double* tflops_out, double* ms_out) {
    if (int rc = need_device()) return rc;
    if (iters < 1 || !tflops_out) { set_last_error("hop_probe_fp64_tflops: bad argument"); return HOP_E_BADARG; }
    cudaDeviceProp prop;
    int dev = 0;
    cudaGetDevice(&dev);
    cudaGetDeviceProperties(&prop, dev);
    const int threads = 512, blocks = prop.multiProcessorCount * 4;
    double* sink = nullptr;
    if (int rc = report_cuda(cudaMalloc(&sink, sizeof(double)), "cudaMalloc(probe)")) return rc;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k_probe_dfma<<<blocks, threads>>>(iters / 8 + 1, 0.5, sink);   // warm-up
    cudaEventRecord(e0);
    k_probe_dfma<<<blocks, threads>>>(iters, 0.5, sink);
    cudaEventRecord(e1);
    cudaError_t e = cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(sink);
    if (e != cudaSuccess) return report_cuda(e, "hop_probe_fp64_tflops");
    const double flops = 2.0 * 8.0 * 16.0 * (double)iters * (double)threads * (double)blocks;
    *tflops_out = flops / (ms * 1e-3) / 1e12;
    if (ms_out) *ms_out = ms;
    return 0;
}

// ---- host-buffer variant ------------------------------------------------------------------------
// The batch is cut into up to four chunks that alternate between two streams: while chunk c runs its three kernels,
// the J(T) block of chunk c-1 (the bulk of the device -> host traffic, 8 T_max bytes per instance) drains over PCIe
// and the inputs of chunk c+1 arrive, so only the first upload and the last download are exposed.  It also bounds
// the linearisation workspace (A, B: 1.7 KB per instance and step) to two chunks instead of the whole batch.
// Measured (profiles/r2_experiments.txt): one step = 44.3 ms against 43.7 ms of kernels (selection 39.4 + linearisation 3.6 +
// rollout 0.74).  HOP_HOST_PRIO_PREP=1 moves the HBM-bound rollout / linearisation of every chunk to a HIGH-PRIORITY
// stream so that they could run under the FP64-bound selection of the previous chunk; on B200 that bought nothing
// (44.7 ms with 8 chunks, 44.5 ms with 4: the selection kernel holds every SM's register file, a freed slot goes to a
// linearisation block and the displaced selection work returns later), so it stays an opt-in A/B switch.
namespace {
constexpr int kHostChunksCap = 16;
struct HostCtx {
    std::mutex mu;
    cudaStream_t stream[2] = {nullptr, nullptr};
    cudaStream_t pre = nullptr;                       // high priority: rollout + linearisation
    cudaEvent_t consts_ready = nullptr;
    cudaEvent_t up[kHostChunksCap] = {}, prep[kHostChunksCap] = {}, sel[kHostChunksCap] = {};
    void* buf = nullptr;
    size_t cap = 0;
    int device = -1;
    void reset() {
        if (buf) cudaFree(buf);
        for (auto& st : stream) { if (st) cudaStreamDestroy(st); st = nullptr; }
        if (pre) cudaStreamDestroy(pre);
        if (consts_ready) cudaEventDestroy(consts_ready);
        for (int i = 0; i < kHostChunksCap; ++i) {
            if (up[i]) cudaEventDestroy(up[i]);
            if (prep[i]) cudaEventDestroy(prep[i]);
            if (sel[i]) cudaEventDestroy(sel[i]);
            up[i] = prep[i] = sel[i] = nullptr;
        }
        buf = nullptr; cap = 0; pre = nullptr; consts_ready = nullptr;
    }
};
HostCtx g_host;
constexpr int kHostChunkMin = 16384;   // instances: below this a chunk no longer fills the machine for several waves
constexpr int kHostChunksMax = 4;
}  // namespace

int hop_select_from_x0_host_f64(int B, int sys, const double* params_host, int N, int T_min, int T_max,
                                const double* x0, const double* U, long u_batch_stride, const double* xg,
                                const double* w, const double* u_ref, const double* Q, const double* R,
                                const double* Qf, unsigned wrap_mask, int central, int mode, double* J_out,
                                int* Tstar_out, double* Jstar_out, int* status) {
    int n = 0, m = 0;
    if (sys_dims(sys, &n, &m)) { set_last_error("hop_select_from_x0_host_f64: unknown system id"); return HOP_E_BADARG; }
    if (int rc = need_device()) return rc;
    if (B <= 0) return B == 0 ? 0 : HOP_E_BADARG;
    std::lock_guard<std::mutex> lock(g_host.mu);
    int dev = 0;
    cudaGetDevice(&dev);
    if (g_host.device != dev) {   // one cached context per process; re-created when the current device changes
        g_host.reset();
        g_host.device = dev;
    }
    for (auto& st : g_host.stream)
        if (!st) { if (int rc = report_cuda(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking), "cudaStreamCreate")) return rc; }
    if (!g_host.pre) {
        int lo = 0, hi = 0;                                                      // (numerically lower = higher priority)
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        if (int rc = report_cuda(cudaStreamCreateWithPriority(&g_host.pre, cudaStreamNonBlocking, hi), "cudaStreamCreateWithPriority")) return rc;
    }
    if (!g_host.consts_ready) {
        if (int rc = report_cuda(cudaEventCreateWithFlags(&g_host.consts_ready, cudaEventDisableTiming), "cudaEventCreate")) return rc;
        for (int i = 0; i < kHostChunksCap; ++i)
            for (cudaEvent_t* e : {&g_host.up[i], &g_host.prep[i], &g_host.sel[i]})
                if (int rc = report_cuda(cudaEventCreateWithFlags(e, cudaEventDisableTiming), "cudaEventCreate")) return rc;
    }
    static const int chunks_max = getenv("HOP_HOST_CHUNKS") ? atoi(getenv("HOP_HOST_CHUNKS")) : kHostChunksMax;   // A/B switch
    int chunks = (B + kHostChunkMin - 1) / kHostChunkMin;
    chunks = chunks < 1 ? 1 : (chunks > chunks_max ? chunks_max : chunks);
    if (chunks < 1) chunks = 1;
    if (chunks > kHostChunksCap) chunks = kHostChunksCap;
    const int per = (((B + chunks - 1) / chunks) + 3) & ~3;       // whole CTAs of the selection kernel
    const int lanes = chunks > 1 ? 2 : 1;                         // workspaces (= streams) in use
    const bool shared_U = (u_batch_stride == 0);
    const size_t nU = shared_U ? (size_t)N * m : (size_t)B * N * m;
    const size_t s_x0 = align256(sizeof(double) * (size_t)B * n), s_xg = s_x0, s_w = align256(sizeof(double) * (size_t)B);
    const size_t s_U = align256(sizeof(double) * nU), s_c = align256(sizeof(double) * (size_t)(m + 2 * n * n + m * m));
    const size_t s_J = align256(sizeof(double) * (size_t)B * T_max), s_T = align256(sizeof(int) * (size_t)B);
    const size_t s_Js = align256(sizeof(double) * (size_t)B);
    const size_t s_ws = align256((size_t)hop_select_from_x0_workspace_bytes(per, N, n, m));
    const size_t total = s_x0 + s_xg + s_w + s_U + s_c + s_J + 2 * s_T + s_Js + lanes * s_ws;
    if (total > g_host.cap) {
        if (g_host.buf) cudaFree(g_host.buf);
        g_host.buf = nullptr; g_host.cap = 0;
        if (int rc = report_cuda(cudaMalloc(&g_host.buf, total), "cudaMalloc(host-variant arena)")) return rc;
        g_host.cap = total;
    }
    char* q = (char*)g_host.buf;
    double* d_x0 = (double*)q; q += s_x0;
    double* d_xg = (double*)q; q += s_xg;
    double* d_w = (double*)q; q += s_w;
    double* d_U = (double*)q; q += s_U;
    double* d_c = (double*)q; q += s_c;
    double* d_J = (double*)q; q += s_J;
    int* d_T = (int*)q; q += s_T;
    int* d_st = (int*)q; q += s_T;
    double* d_Js = (double*)q; q += s_Js;
    char* d_ws = q;
    double* d_uref = d_c; double* d_Q = d_uref + m; double* d_R = d_Q + n * n; double* d_Qf = d_R + m * m;
    cudaStream_t s0 = g_host.stream[0];
    int rc = 0;
    auto copy = [&](void* dst, const void* src, size_t bytes, cudaMemcpyKind kind, cudaStream_t st, const char* what) {
        if (rc == 0) rc = report_cuda(cudaMemcpyAsync(dst, src, bytes, kind, st), what);
    };
    copy(d_uref, u_ref, sizeof(double) * m, cudaMemcpyHostToDevice, s0, "H2D u_ref");
    copy(d_Q, Q, sizeof(double) * n * n, cudaMemcpyHostToDevice, s0, "H2D Q");
    copy(d_R, R, sizeof(double) * m * m, cudaMemcpyHostToDevice, s0, "H2D R");
    copy(d_Qf, Qf, sizeof(double) * n * n, cudaMemcpyHostToDevice, s0, "H2D Qf");
    if (shared_U) copy(d_U, U, sizeof(double) * nU, cudaMemcpyHostToDevice, s0, "H2D U");
    if (rc == 0) rc = report_cuda(cudaEventRecord(g_host.consts_ready, s0), "cudaEventRecord");
    if (rc == 0 && lanes > 1) rc = report_cuda(cudaStreamWaitEvent(g_host.stream[1], g_host.consts_ready, 0), "cudaStreamWaitEvent");
    if (rc == 0) rc = report_cuda(cudaStreamWaitEvent(g_host.pre, g_host.consts_ready, 0), "cudaStreamWaitEvent");
    static const bool serial_prep = !(getenv("HOP_HOST_PRIO_PREP") && atoi(getenv("HOP_HOST_PRIO_PREP")) != 0);   // A/B switch (default: serial)
    for (int c = 0; c < chunks && rc == 0; ++c) {
        const int b0 = c * per, cb = (B - b0) < per ? (B - b0) : per;
        if (cb <= 0) break;
        const int lane = c % lanes;
        cudaStream_t st = g_host.stream[lane];
        cudaStream_t sp = serial_prep ? st : g_host.pre;
        const size_t o = (size_t)b0;
        copy(d_x0 + o * n, x0 + o * n, sizeof(double) * (size_t)cb * n, cudaMemcpyHostToDevice, st, "H2D x0");
        copy(d_xg + o * n, xg + o * n, sizeof(double) * (size_t)cb * n, cudaMemcpyHostToDevice, st, "H2D xg");
        copy(d_w + o, w + o, sizeof(double) * (size_t)cb, cudaMemcpyHostToDevice, st, "H2D w");
        const double* dU = d_U;
        if (!shared_U) {
            copy(d_U + o * N * m, U + o * N * m, sizeof(double) * (size_t)cb * N * m, cudaMemcpyHostToDevice, st, "H2D U");
            dU = d_U + o * N * m;
        }
        if (rc) break;
        // workspace of this lane: X | A | Bm (hop_select_from_x0_workspace_bytes layout)
        char* wsl = d_ws + lane * s_ws;
        double* Xc = (double*)wsl;
        double* Ac = (double*)(wsl + align256(sizeof(double) * (size_t)per * (N + 1) * n));
        double* Bc = (double*)((char*)Ac + align256(sizeof(double) * (size_t)per * N * n * n));
        if (!serial_prep) {
            if ((rc = report_cuda(cudaEventRecord(g_host.up[c], st), "cudaEventRecord"))) break;
            if ((rc = report_cuda(cudaStreamWaitEvent(sp, g_host.up[c], 0), "cudaStreamWaitEvent"))) break;
            if (c >= lanes && (rc = report_cuda(cudaStreamWaitEvent(sp, g_host.sel[c - lanes], 0), "cudaStreamWaitEvent"))) break;   // workspace re-use
        }
        if ((rc = dispatch_rollout(cb, sys, params_host, N, d_x0 + o * n, dU, u_batch_stride, 1e6, Xc, sp))) break;         // solver.py:42
        // X comes from the rollout above, so F(X_k, U_k) = X[k+1] bit for bit: the linearisation may read f0 from it
        if ((rc = dispatch_linearize(cb, sys, params_host, N, Xc, dU, u_batch_stride, central, 1e-5, 1e-5, 1e-6, 1e-6, 1, nullptr,
                                     Ac, Bc, sp))) break;                                                                  // linearization.py:177,216
        if (!serial_prep) {
            if ((rc = report_cuda(cudaEventRecord(g_host.prep[c], sp), "cudaEventRecord"))) break;
            if ((rc = report_cuda(cudaStreamWaitEvent(st, g_host.prep[c], 0), "cudaStreamWaitEvent"))) break;
        }
        // a_resid = NULL: on a trajectory produced by the rollout above F(X_k,U_k) - X_{k+1} is exactly 0
        rc = hop_select_fused_f64(cb, N, n, m, T_min, T_max, Ac, Bc, nullptr, Xc, dU, u_batch_stride, d_xg + o * n, d_w + o, d_uref,
                                  d_Q, d_R, d_Qf, wrap_mask, 1e-9, 1e-12, mode, d_J + o * T_max, d_T + o, d_Js + o, d_st + o, st);
        if (rc) break;
        if ((rc = report_cuda(cudaEventRecord(g_host.sel[c], st), "cudaEventRecord"))) break;
        if (J_out) copy(J_out + o * T_max, d_J + o * T_max, sizeof(double) * (size_t)cb * T_max, cudaMemcpyDeviceToHost, st, "D2H J");
        copy(Tstar_out + o, d_T + o, sizeof(int) * (size_t)cb, cudaMemcpyDeviceToHost, st, "D2H T*");
        if (Jstar_out) copy(Jstar_out + o, d_Js + o, sizeof(double) * (size_t)cb, cudaMemcpyDeviceToHost, st, "D2H J*");
        if (status) copy(status + o, d_st + o, sizeof(int) * (size_t)cb, cudaMemcpyDeviceToHost, st, "D2H status");
    }
    cudaError_t ep = cudaStreamSynchronize(g_host.pre);
    cudaError_t e0 = cudaStreamSynchronize(g_host.stream[0]);
    cudaError_t e1 = lanes > 1 ? cudaStreamSynchronize(g_host.stream[1]) : cudaSuccess;
    if (rc) return rc;
    return report_cuda(e0 != cudaSuccess ? e0 : (e1 != cudaSuccess ? e1 : ep), "hop_select_from_x0_host_f64");
}

}  // extern "C"
