"""GPU: the drop-in modules reproduce the reference's single-instance call flow
(systems -> rollout -> linearize -> augmented -> propagator_all_Jt_aug -> ilqr_timeopt, run_suite CLI)."""
import importlib
import os
import subprocess
import sys

import numpy as np
import pytest

from _common import J_TOL, golden, rel

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DROPIN = os.path.join(ROOT, "time-opt-ilqr_b200", "dropin")


@pytest.fixture(scope="module")
def m():
    sys.path.insert(0, DROPIN)
    mods = {k: importlib.import_module(k) for k in ("utils", "linearization", "augmented", "horizon_selection", "solver",
                                                    "systems", "run_suite", "ilqr_propagator")}
    yield mods
    sys.path.remove(DROPIN)


def test_utils_match_reference(m):
    u = golden("utils")
    for A, X, kind in zip(u["mats"], u["invs"], u["kinds"]):
        d = int(str(kind).split(":")[1])
        Xo = m["utils"].chol_inv(A[:d, :d])
        assert np.linalg.norm(Xo - X[:d, :d]) <= 1e-9 * np.linalg.norm(X[:d, :d]), kind
    for A, B, X, (d, c) in zip(u["sA"], u["sB"], u["sX"], u["sdims"]):
        assert np.abs(m["utils"].chol_solve(A[:d, :d], B[:d, :c]) - X[:d, :c]).max() < 1e-11
    with pytest.raises(FloatingPointError):
        m["utils"].chol_inv(np.array([[1.0, np.nan], [0.0, 1.0]]))
    with pytest.raises(np.linalg.LinAlgError):
        m["utils"].chol_solve(-np.eye(3), np.ones(3))


@pytest.mark.parametrize("name,maker", [("DoubleIntegrator", "make_double_integrator"), ("Segway_Balance", "make_segway_balance"),
                                        ("Quadrotor", "make_quadrotor")])
def test_reference_call_flow_single_instance(m, name, maker):
    g = golden("case_" + name)
    N = int(g["N"])
    F, x0, xg, u_ref, Q, R, alpha, w, N, T_min, T_max, wrap_idx, extra = getattr(m["systems"], maker)(N=N)
    T_max = min(T_max, N)
    U = np.tile(u_ref.reshape(1, -1), (N, 1))
    X = m["solver"].rollout(F, x0, U)
    assert np.abs(X - g["X"]).max() <= 1e-9 * max(1.0, np.abs(g["X"]).max())
    A_list, B_list = m["linearization"].linearize_forward_diff_traj(F, X, U)
    assert isinstance(A_list, list) and A_list[0].shape == (x0.size, x0.size)
    assert np.abs(np.stack(A_list) - g["A_fwd"]).max() <= 2e-9 * max(1.0, np.abs(g["A_fwd"]).max())
    A_aug, B_aug, Q_aug, R_list, z0, R_inv = m["augmented"].build_augmented_sequence_QR(F, A_list, B_list, X, U, xg, u_ref, Q, R, w,
                                                                                       wrap_idx=wrap_idx)
    QT = m["augmented"].build_terminal_aug_list(X, xg, alpha, wrap_idx=wrap_idx)
    ks = g["ks"]
    assert np.abs(np.stack([A_aug[k] for k in ks]) - g["A_aug_ks"]).max() <= 2e-9 * max(1.0, np.abs(g["A_aug_ks"]).max())
    assert np.abs(np.stack([B_aug[k] for k in ks]) - g["B_aug_ks"]).max() <= 2e-9 * max(1.0, np.abs(g["B_aug_ks"]).max())
    assert all(a.shape == (x0.size + 1, x0.size + 1) for a in A_aug) and all(b.shape == (x0.size + 1, u_ref.size) for b in B_aug)
    assert np.array_equal(z0, g["z0"])
    assert np.abs(np.stack([Q_aug[k] for k in ks]) - g["Q_aug_ks"]).max() <= 1e-9 * np.abs(g["Q_aug_ks"]).max()
    assert np.abs(np.stack([QT[k] for k in ks]) - g["QT_ks"]).max() <= 1e-9 * np.abs(g["QT_ks"]).max()
    assert np.abs(R_inv - g["R_inv"]).max() <= 1e-12 * np.abs(g["R_inv"]).max()
    J = m["horizon_selection"].propagator_all_Jt_aug(A_aug, B_aug, Q_aug, R_list, z0, QT, T_use=T_max, R_inv_cached=R_inv)
    T = int(np.argmin(J[T_min - 1:T_max]) + T_min)
    tol_win, tol_star, dT = J_TOL[name]
    Tr = int(g["T0"])
    assert T == Tr and abs(J[Tr - 1] - g["J_curve0"][Tr - 1]) <= tol_star * abs(g["J_curve0"][Tr - 1])
    # R_list path (no cached inverse) gives the same curve
    J2 = m["horizon_selection"].propagator_all_Jt_aug(A_aug, B_aug, Q_aug, [np.array(R_list[0]) for _ in R_list], z0, QT, T_use=T_max)
    assert rel(J2[T_min - 1:], J[T_min - 1:]) <= 1e-9
    k_list, K_list, ok = m["solver"].backward_pass_truncated(A_list, B_list, X, U, xg, u_ref, Q, R, alpha, Tr, lm_lambda=1e-3,
                                                             wrap_idx=wrap_idx)
    assert ok and np.abs(np.stack(K_list) - g["K_list"]).max() <= 1e-9 * np.abs(g["K_list"]).max()
    X1, U1, J1, acc = m["solver"].forward_linesearch_fixedT(F, X, U, xg, u_ref, Q, R, alpha, w, Tr, k_list, K_list, wrap_idx=wrap_idx)
    assert acc == bool(g["acc1"]) and abs(J1 - float(g["J1"])) <= 1e-9 * abs(float(g["J1"]))
    res = m["solver"].ilqr_timeopt_ourmethod(F, x0, xg, u_ref, Q, R, alpha, w, N, T_min, T_max, max_iter=12, S_window=20,
                                             use_central_diff=False, wrap_idx=wrap_idx, extra_stage_cost=None)
    assert sorted(res) == ["J_curve", "J_hist", "T_hist", "T_star", "U", "X", "onepass_error", "timers"]
    assert res["T_hist"] == list(g["sol_T_hist"]) and rel(res["J_hist"], g["sol_J_hist"]) <= 1e-9
    assert sorted(res["timers"]) == ["backward", "forward", "linearize", "select"] and res["timers"]["select"] > 0


@pytest.mark.parametrize("maker,T_ref,J_ref,n_ref", [
    ("make_double_integrator", 25, 6.544382184867515, 3),            # /root/reference/plots/summary.csv:2  (DoubleIntegrator, propagator)
    ("make_quadrotor", 51, 449.1438881199965, 9),                    # /root/reference/plots/summary.csv:11 (Quadrotor_Hover, propagator)
])
def test_legacy_monolith_reproduces_the_shipped_results(m, maker, T_ref, J_ref, n_ref):
    """The only goldens the reference SHIPS: plots/summary.csv, written by ilqr_propagator.main() with central differences,
    max_iter=20, lm_init=1e-3, S_window=10 (ilqr_propagator.py:776-780).  Through the legacy-name adapter: T*, n_iterations
    = len(J_hist) identical, J* within 1e-9.  Deviation from the monolith, irrelevant here: in-kernel ladder of 8 tries instead
    of 4 (neither case ever fails a Cholesky)."""
    lp = m["ilqr_propagator"]
    case = getattr(lp, maker)()
    assert len(case) == 12                                                     # the monolith's 12-tuple (ilqr_propagator.py:668)
    F, x0, xg, u_ref, Q, R, alpha, w, N, T_min, T_max, wrap_idx = case
    res = lp.ilqr_timeopt(F, x0, xg, u_ref, Q, R, alpha, w, N, T_min, T_max, method="propagator", max_iter=20, lm_init=1e-3,
                          S_window=10, use_central_diff=True, wrap_idx=wrap_idx)
    assert sorted(res) == ["J_curve", "J_hist", "T_hist", "T_star", "U", "X", "consistency_check", "timers"]
    assert res["T_star"] == T_ref and len(res["J_hist"]) == n_ref
    assert abs(res["J_hist"][-1] - J_ref) <= 1e-9 * J_ref
    cc = res["consistency_check"]                                              # summary.csv: 4.3e-4 (DI), 7.9e-2 (Quadrotor)
    assert 0.0 < cc["max_abs_diff"] < (1e-2 if T_ref == 25 else 1.0) and cc["rmse"] <= cc["max_abs_diff"]
    assert res["J_curve"].shape == (T_max,)
    # stand-alone linear algebra with the monolith's defaults: 4 tries, then the plain inverse of A + 1e-5 I (the modular
    # 8-try ladder would end at A + 0.1 I): checked against the monolith's own output for diag(1, -1, 1)
    Ainv = lp.chol_inv(np.diag([1.0, -1.0, 1.0]))
    assert np.allclose(np.diag(Ainv), [1.0 / (1.0 + 1e-5), 1.0 / (-1.0 + 1e-5), 1.0 / (1.0 + 1e-5)], rtol=1e-12)
    with pytest.raises(np.linalg.LinAlgError):
        lp.chol_solve(-np.eye(3), np.ones(3))


def test_run_suite_cli_writes_the_reference_csv_schema(m, tmp_path):
    env = dict(os.environ, PYTHONHASHSEED="0")
    out = tmp_path / "res"
    subprocess.check_call([sys.executable, os.path.join(DROPIN, "run_suite.py"), "--cases", "DoubleIntegrator,Segway_Balance",
                           "--trials", "5", "--max-iter", "12", "--outdir", str(out)], env=env, cwd=DROPIN)
    import pandas as pd
    df = pd.read_csv(out / "summary_all.csv")
    assert list(df.columns) == ["case", "trial", "solver", "status", "T_star", "J_star", "total_time", "final_err", "success",
                                "n_iter", "solver_error", "best_J", "cost_ratio_best", "time_base", "time_ratio_base"]
    di = df[(df.case == "DoubleIntegrator") & (df.trial == 0)].iloc[0]
    assert di.T_star == 25 and abs(di.J_star - 6.54438218486751) <= 1e-9 * 6.54438218486751 and di.n_iter == 3
    sg = df[(df.case == "Segway_Balance") & (df.trial == 0)].iloc[0]
    assert sg.T_star == 40 and abs(sg.J_star - 4.642932072056589) <= 1e-9 * 4.642932072056589
    agg = pd.read_csv(out / "summary_agg.csv")
    assert list(agg.columns) == ["case", "solver", "n", "success_rate", "T_median", "J_median", "time_median", "ratio_cost_median",
                                 "ratio_time_median"]
    assert (out / "Segway_Balance" / "summary_all.csv").exists()
