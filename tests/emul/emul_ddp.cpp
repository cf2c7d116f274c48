// emul_ddp.cpp -- host entry points that run the backward-pass kernel bodies (hop_ddp_core.cuh: one warp per problem with the
// reference's summation order; hop_ddp_mma.cuh: matrix products as DMMA fragments) under the SIMT emulator.  Test infrastructure.
#include <vector>

#include "hop_ddp_mma.cuh"

namespace {
struct BwJob {
    const double *A, *Bm, *X, *U;
    hop::ddp::CostConst c;
    int T;
    double lm;
    double *k_out, *K_out;
    int ok, rc, variant;
    double* smem;
};
template <int n, int m>
void bw_lane(void* a) {
    auto* j = (BwJob*)a;
    int ok = 0;
    const int lane = hop::simt::lane_id();
    const int rc = (j->variant == 3)
        ? hop::ddp::backward_pass_mma<n, m>(j->A, j->Bm, j->X, j->U, j->c, j->T, j->lm, j->k_out, j->K_out, &ok, j->smem, lane)
        : hop::ddp::backward_pass_warp<n, m>(j->A, j->Bm, j->X, j->U, j->c, j->T, j->lm, j->k_out, j->K_out, &ok, j->smem, lane);
    if (lane == 0) { j->ok = ok; j->rc = rc; }
}
}  // namespace

// variant 2: ordered warp kernel, 3: tensor-pipe kernel.  Returns the DDP_* code (or -1 / -2), *ok as the kernels set it.
extern "C" int emul_backward(int n, int m, int variant, int T, double lm, const double* A, const double* Bm, const double* X,
                             const double* U, const double* xg, const double* u_ref, const double* Q, const double* R,
                             const double* Qf, double w, unsigned wrap_mask, double* k_out, double* K_out, int* ok) {
    if (!(n == 12 && m == 4)) return -2;
    std::vector<double> smem(hop::ddp::BwSmem<12, 4>::SIZE, -7.0);
    const unsigned diag = (hop::ddp::is_diagonal<12>(Q) ? 1u : 0u) | (hop::ddp::is_diagonal<4>(R) ? 2u : 0u) | (hop::ddp::is_diagonal<12>(Qf) ? 4u : 0u);
    BwJob j{A, Bm, X, U, {xg, u_ref, Q, R, Qf, w, wrap_mask, diag}, T, lm, k_out, K_out, 0, 0, variant, smem.data()};
    if (hop::simt::run_warp(bw_lane<12, 4>, &j)) return -1;
    *ok = j.ok;
    return j.rc;
}
