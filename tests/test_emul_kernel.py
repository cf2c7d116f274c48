"""Runs the CUDA kernel *source* (hop_select_body.cuh) on the host through the fiber-based SIMT
emulator of tests/emul and checks it against the oracle / reference goldens.  CPU only; this is how
the kernel logic is exercised in a container without a GPU.  The GPU run of the same checks is in
tests/test_gpu_parity.py."""
import numpy as np
import pytest

import emul
import oracle as O
from _common import CASE_NAMES, J_TOL, golden, rel, s2_batch
from hop import cases


@pytest.mark.parametrize("d,m,N", [(3, 1, 32), (4, 2, 64), (5, 1, 64), (12, 4, 32), (13, 4, 48)])
def test_emulated_generic_kernel_matches_oracle_on_s2(d, m, N):
    A, B, Q, R, z0, w, QT = s2_batch(range(5), d, m, N)      # 5 instances: ragged last warp on every G
    Rinv = np.stack([O.chol_inv(r) for r in R])
    J, T, Js, st = emul.select_generic(A, B, Q, Rinv, z0, QT, 1, N, w_explicit=w)
    Jo, sto = O.propagator_batch(A, B, Q, Rinv, z0, QT)
    assert not st.any() and not sto.any()
    assert rel(J, Jo) <= 1e-12
    tot = Jo + w[:, None] * np.arange(1, N + 1)
    assert np.array_equal(T, np.argmin(tot, axis=1) + 1)
    assert np.allclose(Js, tot.min(axis=1), rtol=1e-12)


@pytest.mark.parametrize("name", ["DoubleIntegrator", "Segway_Balance", "Quadrotor"])
def test_emulated_fused_kernel_matches_reference_goldens(name):
    g = golden("case_" + name)
    F, x0, xg, u_ref, Q, R, alpha, w, N, T_min, T_max, wrap_idx, _ = cases.make_case(name, N=int(g["N"]))
    if name == "Quadrotor":
        T_max = 64                    # keep the emulation fast; the curve prefix does not depend on T_max
    n = x0.size
    J, T, Js, st = emul.select_fused(g["A_fwd"][None], g["B_fwd"][None], g["a_resid"][None], g["X"][None], g["U"][None],
                                     xg[None], np.array([w]), u_ref, Q, R, O.as_terminal_weight(alpha, n),
                                     O.wrap_mask(wrap_idx), T_min, T_max)
    Jr, Tr = g["J_curve0"], int(g["T0"])
    tol_win, tol_star, dT = J_TOL[name]
    assert (st[0] & 0xFF) == 0
    assert abs(int(T[0]) - Tr) <= dT
    assert abs(J[0, Tr - 1] - Jr[Tr - 1]) <= tol_star * abs(Jr[Tr - 1])
    if tol_win is not None:
        assert rel(J[0, T_min - 1:T_max], Jr[T_min - 1:T_max]) <= tol_win


def test_emulated_kernel_ladder_fallback_and_nonfinite_status():
    """Edge cases the reference handles by exceptions / the jitter ladder (utils.py:75,81-93)."""
    d, m, N = 4, 2, 8
    A, B, Q, R, z0, w, QT = s2_batch(range(3), d, m, N)
    Rinv = np.stack([O.chol_inv(r) for r in R])
    Q[1, 3] = np.diag([1.0, 1.0, 1.0, -1e-4])        # needs the ladder (PD again once eps >= 1e-3)
    QT[2, 5] = -np.eye(d)                             # never PD -> LU fallback on (A + 0.1 I)
    J, T, Js, st = emul.select_generic(A, B, Q, Rinv, z0, QT, 1, N)
    Jo, sto = O.propagator_batch(A, B, Q, Rinv, z0, QT)
    assert st[0] == 0 and st[1] == 0x100 and st[2] == 0x300 and not sto.any()
    assert rel(J, Jo) <= 1e-9
    A[0, 2, 1, 1] = np.nan
    J, T, Js, st = emul.select_generic(A, B, Q, Rinv, z0, QT, 1, N)
    _, sto = O.propagator_batch(A, B, Q, Rinv, z0, QT)
    assert (st[0] & 0xFF) == 1 and sto[0] == 1        # FloatingPointError in the reference
    assert (st[1] & 0xFF) == 0 and (st[2] & 0xFF) == 0


@pytest.mark.parametrize("d,m,N", [(12, 4, 12), (13, 4, 16)])
def test_emulated_dmma_generic_kernel_matches_oracle(d, m, N):
    """One-problem-per-warp DMMA mapping (hop_select_mma_body.cuh): register-fragment layout, shuffle
    transposes and the Gauss-Jordan pivot exchange, checked on the host emulator."""
    A, B, Q, R, z0, w, QT = s2_batch(range(2), d, m, N)
    Rinv = np.stack([O.chol_inv(r) for r in R])
    Q[1, 3] = np.diag(np.r_[np.ones(d - 1), -1e-4])       # ladder
    QT[1, 5] = -np.eye(d)                                  # LU fallback
    J, T, Js, st = emul.select_generic(A, B, Q, Rinv, z0, QT, 1, N, w_explicit=w, mma=True)
    Jo, sto = O.propagator_batch(A, B, Q, Rinv, z0, QT)
    assert st[0] == 0 and st[1] == 0x300 and not sto.any()
    assert rel(J, Jo) <= 1e-9
    assert np.array_equal(T, np.argmin(Jo + w[:, None] * np.arange(1, N + 1), axis=1) + 1)


@pytest.mark.parametrize("mode", [0, 1, 2, 3, 4, 5])   # 5: 4 with two pivots per loop trip + high-word zeroing; 2 / 3 / 4 = pipelined FAST schedule: unrolled / looped / looped + shared-memory pivot exchange
def test_emulated_dmma_fused_kernel_matches_reference_golden(mode):
    g = golden("case_Quadrotor")
    F, x0, xg, u_ref, Q, R, alpha, w, N, T_min, T_max, wrap_idx, _ = cases.make_case("Quadrotor", N=128)
    T_max = 60
    J, T, Js, st = emul.select_fused(g["A_fwd"][None], g["B_fwd"][None], g["a_resid"][None], g["X"][None], g["U"][None],
                                     xg[None], np.array([w]), u_ref, Q, R, O.as_terminal_weight(alpha, 12),
                                     O.wrap_mask(wrap_idx), T_min, T_max, mma=True, mode=mode)
    Jr, Tr = g["J_curve0"], int(g["T0"])
    assert (st[0] & 0xFF) == 0 and int(T[0]) == Tr
    assert abs(J[0, Tr - 1] - Jr[Tr - 1]) <= 1e-8 * abs(Jr[Tr - 1])
    assert rel(J[0, T_min - 1:T_max], Jr[T_min - 1:T_max]) <= 1e-6


def test_emulated_pipelined_kernel_falls_back_to_sequential_body():
    """A non-finite trajectory entry aborts the pipelined sweep mid-horizon; the sequential body then owns the
    status word (reference: FloatingPointError, utils.py:75).  T_max = N exercises the X-only last stage."""
    g = golden("case_Quadrotor")
    F, x0, xg, u_ref, Q, R, alpha, w, N, T_min, T_max, wrap_idx, _ = cases.make_case("Quadrotor", N=128)
    Nn = 12
    A, B, a, X, U = (g[k][:Nn].copy() for k in ("A_fwd", "B_fwd", "a_resid", "X", "U"))
    X = g["X"][:Nn + 1].copy()
    args = (xg[None], np.array([w]), u_ref, Q, R, O.as_terminal_weight(alpha, 12), O.wrap_mask(wrap_idx), 1, Nn)
    J1, T1, _, st1 = emul.select_fused(A[None], B[None], a[None], X[None], U[None], *args, mma=True, mode=1)
    J2, T2, _, st2 = emul.select_fused(A[None], B[None], a[None], X[None], U[None], *args, mma=True, mode=2)
    assert st1[0] == 0 and st2[0] == 0 and T1[0] == T2[0] and rel(J2, J1) <= 1e-5
    A[7, 3, 3] = np.nan
    J1, T1, _, st1 = emul.select_fused(A[None], B[None], a[None], X[None], U[None], *args, mma=True, mode=1)
    J2, T2, _, st2 = emul.select_fused(A[None], B[None], a[None], X[None], U[None], *args, mma=True, mode=2)
    assert (st2[0] & 0xFF) == 1 and st1[0] == st2[0]
    assert np.array_equal(J1, J2, equal_nan=True) and T1[0] == T2[0]


@pytest.mark.parametrize("d,m,N,T_max", [(12, 4, 24, 24), (13, 4, 21, 19), (13, 4, 6, 5)])
def test_emulated_scan_mode_matches_sequential_sweep(d, m, N, T_max):
    """HOP_MODE_SCAN (chunked parallel scan over the horizon, 8 warps per problem): chunk 0 is bit-identical to the
    sequential sweep, later chunks differ by re-association only (well-conditioned S2: <= 1e-9), T* identical.
    Ragged chunking: T_max not a multiple of the warp count, and T_max smaller than it."""
    A, B, Q, R, z0, w, QT = s2_batch(range(2), d, m, N)
    Rinv = np.stack([O.chol_inv(r) for r in R])
    Js, Ts, Jss, sts = emul.select_generic(A, B, Q, Rinv, z0, QT, 1, T_max, w_explicit=w, mma=True)
    Jp, Tp, Jsp, stp = emul.select_generic(A, B, Q, Rinv, z0, QT, 1, T_max, w_explicit=w, scan=True)
    Lc = -(-T_max // 8)
    assert not sts.any() and not stp.any()
    assert np.array_equal(Jp[:, :Lc], Js[:, :Lc])
    assert rel(Jp, Js) <= 1e-9
    assert np.array_equal(Tp, Ts) and np.allclose(Jsp, Jss, rtol=1e-9)
    Jo, _ = O.propagator_batch(A, B, Q, Rinv, z0, QT, T_use=T_max)
    assert rel(Jp, Jo) <= 1e-9


@pytest.mark.parametrize("d,m,N", [(12, 4, 12), (13, 4, 16)])
def test_emulated_pipelined_generic_kernel_matches_oracle_and_falls_back(d, m, N):
    """Software-pipelined LQR-boundary body (hop_select_gpipe_body.cuh): three interleaved sweeps + a dual sweep per
    step, elimination with right-hand side z0, cp.async staging.  Instance 1 needs the jitter ladder and the LU
    fallback, instance 2 has a non-finite input: both must come out of the sequential cold path with its status word."""
    A, B, Q, R, z0, w, QT = s2_batch(range(3), d, m, N)
    Rinv = np.stack([O.chol_inv(r) for r in R])
    Q[1, 3] = np.diag(np.r_[np.ones(d - 1), -1e-4])       # ladder
    QT[1, 5] = -np.eye(d)                                  # LU fallback
    A[2, 4, 1, 1] = np.nan
    J, T, Js, st = emul.select_generic(A, B, Q, Rinv, z0, QT, 1, N, w_explicit=w, pipe=True)
    Jq, Tq, Jsq, stq = emul.select_generic(A, B, Q, Rinv, z0, QT, 1, N, w_explicit=w, mma=True)
    Jo, sto = O.propagator_batch(A, B, Q, Rinv, z0, QT)
    assert st[0] == 0 and st[1] == 0x300 and (st[2] & 0xFF) == 1 and sto[2] == 1
    assert np.array_equal(st, stq)
    assert rel(J[:2], Jo[:2]) <= 1e-9 and np.array_equal(J[1:], Jq[1:], equal_nan=True)
    assert np.array_equal(T[:2], np.argmin(Jo[:2] + w[:2, None] * np.arange(1, N + 1), axis=1) + 1)


@pytest.mark.parametrize("d,m,N", [(3, 1, 32), (4, 2, 64), (5, 1, 48)])
def test_emulated_thread_per_problem_kernel_is_bit_identical_to_the_lane_group_kernel(d, m, N):
    """hop_select_tpp_body.cuh (one problem per thread, blocks in registers) performs, element for element, the IEEE
    operations of the lane-group kernel: J, T*, J* and status must be identical bit for bit -- also through the jitter
    ladder, the LU fallback and a non-finite input (utils.py:75,81-93)."""
    A, B, Q, R, z0, w, QT = s2_batch(range(5), d, m, N)
    Rinv = np.stack([O.chol_inv(r) for r in R])
    Q[1, 3] = np.diag(np.r_[np.ones(d - 1), -1e-4])       # ladder
    QT[2, 5] = -np.eye(d)                                  # LU fallback
    A[3, 7, 1, 1] = np.nan                                 # FloatingPointError in the reference
    Jg, Tg, Jsg, stg = emul.select_generic(A, B, Q, Rinv, z0, QT, 2, N - 1, w_explicit=w)
    Jt, Tt, Jst, stt = emul.select_generic(A, B, Q, Rinv, z0, QT, 2, N - 1, w_explicit=w, tpp=True)
    assert stg[1] == 0x100 and stg[2] == 0x300 and (stg[3] & 0xFF) == 1 and stg[0] == 0 and stg[4] == 0
    assert np.array_equal(stt, stg) and np.array_equal(Tt, Tg)
    assert np.array_equal(Jt, Jg, equal_nan=True) and np.array_equal(Jst, Jsg, equal_nan=True)
    Jo, _ = O.propagator_batch(A[:1], B[:1], Q[:1], Rinv[:1], z0[:1], QT[:1], T_use=N - 1)
    assert rel(Jt[:1], Jo) <= 1e-12


@pytest.mark.parametrize("name", ["DoubleIntegrator", "Segway_Balance", "Cartpole_SwingUp"])
def test_emulated_element_per_lane_fused_kernel_is_bit_identical_to_the_lane_group_kernel(name):
    """hop_select_epl_body.cuh (a warp per problem, lane 5r + c owns element (r, c), everything by shuffles) performs the
    lane-group kernel's IEEE operations element for element: J, T*, J*, status identical bit for bit on the nominal
    trajectories of the reference cases, with per-instance goals / weights, a residual a_k, and a NaN in one instance."""
    g = golden("case_" + name)
    F, x0, xg, u_ref, Q, R, alpha, w, N, T_min, T_max, wrap_idx, _ = cases.make_case(name, N=int(g["N"]))
    n = x0.size
    T_max = min(T_max, 60); T_min = min(T_min, 20)
    rng = np.random.default_rng(1)
    B = 3
    A = np.repeat(g["A_fwd"][None], B, 0); Bm = np.repeat(g["B_fwd"][None], B, 0)
    X = np.repeat(g["X"][None], B, 0) + 1e-3 * rng.standard_normal((B,) + g["X"].shape)
    U = np.repeat(g["U"][None], B, 0); ar = np.repeat(g["a_resid"][None], B, 0) + 1e-4 * rng.standard_normal((B,) + g["a_resid"].shape)
    xgs = xg[None] + 0.01 * rng.standard_normal((B, n)); ws = w * np.array([1.0, 0.5, 2.0])
    A[2, 5, 0, 0] = np.nan
    args = (A, Bm, ar, X, U, xgs, ws, u_ref, Q, R, O.as_terminal_weight(alpha, n), O.wrap_mask(wrap_idx), T_min, T_max)
    Jg, Tg, Jsg, stg = emul.select_fused(*args)
    Je, Te, Jse, ste = emul.select_fused(*args, epl=True)
    assert (stg[2] & 0xFF) == 1 and (stg[0] & 0xFF) == 0
    assert np.array_equal(ste, stg) and np.array_equal(Te, Tg)
    assert np.array_equal(Je, Jg, equal_nan=True) and np.array_equal(Jse, Jsg, equal_nan=True)


# ------------------------------------------------------------------ HOP_MODE_EXACT: reference operation order (hop_select_ref_body.cuh)
@pytest.mark.parametrize("d,m,N", [(1, 1, 8), (3, 1, 32), (4, 2, 48), (5, 1, 40), (7, 3, 16), (12, 4, 24), (13, 4, 32), (16, 5, 10)])
def test_emulated_exact_mode_is_bit_identical_to_the_oracle_on_s2(d, m, N):
    """HOP_MODE_EXACT issues the reference's operations one IEEE rounding at a time (Cholesky in dpotf2 order, two
    substitutions, left-to-right products, no FMA): on identical inputs J(T) equals the plain-C oracle bit for bit -- for
    every d, m <= 16 (run-time dimensions), through the jitter ladder, the LU fallback, T_min > 1 and T_max < N."""
    A, B, Q, R, z0, w, QT = s2_batch(range(3), d, m, N)
    Rinv = np.stack([O.chol_inv(r) for r in R])
    if N > 8:
        Q[1, 3] = np.diag(np.r_[np.ones(d - 1), -1e-4])       # ladder
        QT[2, 5] = -np.eye(d)                                  # LU fallback
    J, T, Js, st = emul.select_generic(A, B, Q, Rinv, z0, QT, 2, N - 1, w_explicit=w, ref=True)
    Jo, sto = O.propagator_batch(A, B, Q, Rinv, z0, QT, T_use=N - 1)
    assert not sto.any() and not (st & 0xFF).any()
    if N > 8:
        assert st[0] == 0 and st[1] & 0x100 and st[2] == 0x300
    assert np.array_equal(J, Jo)
    tot = Jo + w[:, None] * np.arange(1, N)
    assert np.array_equal(T, np.argmin(tot[:, 1:], axis=1) + 2)
    assert np.array_equal(Js, tot[np.arange(3), T - 1])


def test_emulated_exact_mode_status_on_non_finite_input():
    """A non-finite block raises FloatingPointError in the reference (utils.py:75): status 1, the curve from that step on is NaN,
    NaN wins the argmin as in np.argmin; the other instances are unaffected."""
    d, m, N = 4, 2, 8
    A, B, Q, R, z0, w, QT = s2_batch(range(2), d, m, N)
    Rinv = np.stack([O.chol_inv(r) for r in R])
    A[0, 2, 1, 1] = np.nan
    J, T, Js, st = emul.select_generic(A, B, Q, Rinv, z0, QT, 1, N, ref=True)
    Jo, sto = O.propagator_batch(A, B, Q, Rinv, z0, QT)
    assert (st[0] & 0xFF) == 1 and sto[0] == 1 and st[1] == 0
    assert np.isfinite(J[0, :2]).all() and np.isnan(J[0, 2:]).all() and T[0] == 3 and np.isnan(Js[0])
    assert np.array_equal(J[1], Jo[1])
    Q[1, 0] = np.diag([np.inf, 1.0, 1.0, 1.0])                  # an infinite diagonal entry is non-finite input, not a PD matrix
    J, T, Js, st = emul.select_generic(A, B, Q, Rinv, z0, QT, 1, N, ref=True)
    _, sto = O.propagator_batch(A, B, Q, Rinv, z0, QT)
    assert (st[1] & 0xFF) == 1 and sto[1] == 1


@pytest.mark.parametrize("name", CASE_NAMES)
@pytest.mark.parametrize("traj", ["nominal", "converged", "residuals"])
def test_emulated_exact_fused_mode_is_bit_identical_to_the_oracle(name, traj):
    """Fused form (augmented.py:10-87 built in the slab): the four reference cases on the nominal and the converged
    trajectory of the reference's own solve, and with a perturbed trajectory whose affine residuals a_k are NOT zero."""
    g = golden("case_" + name)
    F, x0, xg, u_ref, Q, R, alpha, w, N, T_min, T_max, wrap_idx, _ = cases.make_case(name, N=int(g["N"]))
    n = x0.size
    T_max = min(T_max, 72 if name == "Quadrotor" else T_max)   # keeps the emulation fast; the curve prefix does not depend on T_max
    ar = None
    if traj == "nominal":
        X, U, A, Bm = g["X"], g["U"], g["A_fwd"], g["B_fwd"]
    elif traj == "converged":
        X, U = g["sol_X"], g["sol_U"]
        A, Bm = O.linearize(F.hop_sys, F.hop_params, X, U)
    else:
        rng = np.random.default_rng(5)
        U = g["U"] + 0.01 * rng.standard_normal(g["U"].shape)
        X = O.rollout(F.hop_sys, F.hop_params, x0, U) + 1e-4 * rng.standard_normal(g["X"].shape)
        A, Bm = O.linearize(F.hop_sys, F.hop_params, X, U)
        ar = O.affine_residuals(F.hop_sys, F.hop_params, X, U)
        assert np.abs(ar).max() > 1e-5
    J, T, Js, st = emul.select_fused(A[None], Bm[None], None if ar is None else ar[None], X[None], U[None], xg[None],
                                     np.array([w]), u_ref, Q, R, O.as_terminal_weight(alpha, n), O.wrap_mask(wrap_idx),
                                     T_min, T_max, ref=True)
    Jo, To = O.select_fused(A, Bm, X, U, xg, u_ref, Q, R, alpha, w, T_min, T_max, wrap_idx, a_resid=ar)
    assert (st[0] & 0xFF) == 0
    assert np.array_equal(J[0], Jo) and int(T[0]) == To and Js[0] == Jo[To - 1]
    if traj == "nominal":                                      # and the reference itself
        tol_win, tol_star, dT = J_TOL[name]
        Tr = int(g["T0"])
        if Tr <= T_max:
            assert abs(int(T[0]) - Tr) <= dT
            assert abs(J[0, Tr - 1] - g["J_curve0"][Tr - 1]) <= tol_star * abs(g["J_curve0"][Tr - 1])


@pytest.mark.parametrize("d,m,N", [(4, 2, 64), (13, 4, 64)])
def test_emulated_fp32_mode_stays_within_its_stated_tolerance(d, m, N):
    """HOP_MODE_FP32: the same sweep in IEEE single precision; on the well-conditioned family J within 1e-4, T* reproduced."""
    A, B, Q, R, z0, w, QT = s2_batch(range(4), d, m, N)
    Rinv = np.stack([O.chol_inv(r) for r in R])
    J, T, Js, st = emul.select_generic(A, B, Q, Rinv, z0, QT, 1, N, w_explicit=w, fp32=True)
    Jo, _ = O.propagator_batch(A, B, Q, Rinv, z0, QT)
    assert not st.any() and rel(J, Jo) <= 1e-4
    assert np.array_equal(T, np.argmin(Jo + w[:, None] * np.arange(1, N + 1), axis=1) + 1)


@pytest.mark.parametrize("d,m,N", [(12, 4, 12), (13, 4, 16)])
def test_emulated_diagonal_block_fast_path_is_bit_identical_to_the_sweep(d, m, N):
    """hop_select_gpipe_body.cuh inverts DIAGONAL input blocks (Q_k, QT_t: the whole synthetic family S2) element-wise instead
    of by two Gauss-Jordan sweeps.  The sweep of a diagonal matrix computes exactly those reciprocals, so J, T*, J*, status
    must not change by a bit -- also when only some blocks are diagonal (instance 1: a dense Q_3, a dense QT_5) and when a
    diagonal entry is not positive (instance 2: ladder -> sequential cold path)."""
    A, B, Q, R, z0, w, QT = s2_batch(range(3), d, m, N)
    Rinv = np.stack([O.chol_inv(r) for r in R])
    rng = np.random.default_rng(8)
    Mx = rng.standard_normal((d, d)); Q[1, 3] = Mx @ Mx.T / d + np.eye(d)          # dense SPD blocks
    Mx = rng.standard_normal((d, d)); QT[1, 5] = Mx @ Mx.T + 50.0 * np.eye(d)
    Q[2, 4] = np.diag(np.r_[np.ones(d - 1), -1e-4])                                 # diagonal, not PD: jitter ladder
    a = emul.select_generic(A, B, Q, Rinv, z0, QT, 1, N, w_explicit=w, pipe=True)
    b = emul.select_generic(A, B, Q, Rinv, z0, QT, 1, N, w_explicit=w, pipe=True, no_diag=True)
    for x, y in zip(a, b):
        assert np.array_equal(x, y, equal_nan=True)
    Jo, sto = O.propagator_batch(A, B, Q, Rinv, z0, QT)
    assert a[3][0] == 0 and a[3][1] == 0 and a[3][2] & 0x100 and not sto.any()
    assert rel(a[0], Jo) <= 1e-9


def test_emulated_pipelined_fast_kernel_with_a_dense_running_weight():
    """The pipelined FAST body forms F_k^T = A_k E_k in closed form (column scaling + rank-1 term) when K = (sym(Q) + q_reg I +
    eps I)^-1 is diagonal -- every reference case -- and through the tensor pipe otherwise.  Both branches against the oracle on
    the quadrotor nominal trajectory: the case's diagonal Q, and a dense SPD Q (same problem otherwise)."""
    g = golden("case_Quadrotor")
    F, x0, xg, u_ref, Q, R, alpha, w, N, T_min, T_max, wrap_idx, _ = cases.make_case("Quadrotor", N=128)
    T_max = 56
    rng = np.random.default_rng(12)
    Mx = rng.standard_normal((12, 12))
    Qd = Q + 0.05 * (Mx @ Mx.T) / 12.0
    for Qc in (Q, Qd):
        J, T, Js, st = emul.select_fused(g["A_fwd"][None], g["B_fwd"][None], g["a_resid"][None], g["X"][None], g["U"][None],
                                         xg[None], np.array([w]), u_ref, Qc, R, O.as_terminal_weight(alpha, 12),
                                         O.wrap_mask(wrap_idx), T_min, T_max, mma=True, mode=5)
        Jo, To = O.select_fused(g["A_fwd"], g["B_fwd"], g["X"], g["U"], xg, u_ref, Qc, R, alpha, w, T_min, T_max, wrap_idx)
        assert (st[0] & 0xFF) == 0
        assert rel(J[0, T_min - 1:T_max], Jo[T_min - 1:T_max]) <= 1e-6
        assert abs(J[0, To - 1] - Jo[To - 1]) <= 5e-8 * abs(Jo[To - 1])
        assert int(T[0]) == To or abs(Jo[int(T[0]) - 1] - Jo[To - 1]) <= 1e-7 * abs(Jo[To - 1])


@pytest.mark.parametrize("variant", ["lanes", "tpp", "mma", "pipe"])
def test_emulated_gauss_jordan_kernels_treat_an_infinite_diagonal_as_non_finite_input(variant):
    """utils.py:75 raises FloatingPointError on a non-finite chol_inv input before any attempt.  A +Inf diagonal entry used to
    pass the pivot test of the Gauss-Jordan kernels (p = +Inf > 0, 1/p = 0) and come back as an 'inverse' with a zeroed
    row; the pivot test now rejects +Inf, the finiteness check runs, and the status is the reference's (1)."""
    d, m, N = (4, 2, 8) if variant in ("lanes", "tpp") else (13, 4, 8)
    A, B, Q, R, z0, w, QT = s2_batch(range(2), d, m, N)
    Rinv = np.stack([O.chol_inv(r) for r in R])
    Q[1, 2] = np.diag(np.r_[np.inf, np.ones(d - 1)])
    kw = {"lanes": {}, "tpp": {"tpp": True}, "mma": {"mma": True}, "pipe": {"pipe": True}}[variant]
    J, T, Js, st = emul.select_generic(A, B, Q, Rinv, z0, QT, 1, N, **kw)
    _, sto = O.propagator_batch(A, B, Q, Rinv, z0, QT)
    assert sto[1] == 1 and (st[1] & 0xFF) == 1 and st[0] == 0


@pytest.mark.parametrize("traj", ["nominal", "converged"])
def test_emulated_tensor_pipe_backward_pass(traj):
    """hop_ddp_mma.cuh (12 x 12 products of solver.backward_pass_truncated as DMMA fragments) on the host emulator: the ordered
    warp kernel reproduces the oracle bit for bit, the tensor-pipe kernel stays within 1e-12 relative of it (FMA accumulation
    in k-blocks instead of unfused left-to-right sums) and takes the same ok decision, also for a short horizon."""
    g = golden("case_Quadrotor")
    F, x0, xg, u_ref, Q, R, alpha, w, N, T_min, T_max, wrap_idx, _ = cases.make_case("Quadrotor", N=128)
    Qf = O.as_terminal_weight(alpha, 12)
    if traj == "nominal":
        X, U, A, Bm, T = g["X"], g["U"], g["A_fwd"], g["B_fwd"], int(g["T0"])
    else:
        X, U = g["sol_X"], g["sol_U"]
        A, Bm = O.linearize(F.hop_sys, F.hop_params, X, U)
        T = int(g["sol_T_star"])
    for Tq in (T, 3):
        ko, Ko, oko = O.backward_pass(A, Bm, X, U, xg, u_ref, Q, R, alpha, Tq, 1e-3, wrap_idx)
        k2, K2, ok2, rc2 = emul.backward_pass(A, Bm, X, U, xg, u_ref, Q, R, Qf, w, O.wrap_mask(wrap_idx), Tq, 1e-3, 2)
        k3, K3, ok3, rc3 = emul.backward_pass(A, Bm, X, U, xg, u_ref, Q, R, Qf, w, O.wrap_mask(wrap_idx), Tq, 1e-3, 3)
        assert oko and ok2 and ok3 and rc2 == 0 and rc3 == 0
        assert np.array_equal(K2, Ko) and np.array_equal(k2, ko)
        assert np.abs(K3 - K2).max() <= 1e-12 * np.abs(K2).max() and np.abs(k3 - k2).max() <= 1e-12 * np.abs(k2).max()
    Xn = X.copy(); Xn[5, 3] = np.nan                                        # non-finite trajectory: ok = False in both
    assert not emul.backward_pass(A, Bm, Xn, U, xg, u_ref, Q, R, Qf, w, O.wrap_mask(wrap_idx), T, 1e-3, 3)[2]
    assert not emul.backward_pass(A, Bm, Xn, U, xg, u_ref, Q, R, Qf, w, O.wrap_mask(wrap_idx), T, 1e-3, 2)[2]
