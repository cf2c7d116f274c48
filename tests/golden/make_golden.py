#!/usr/bin/env python
"""Generate golden vectors from the REAL reference (dmmsjtu-umich/time-opt-ilqr).

Run in the build container only (the reference is mounted read-only at /root/reference and does
not exist on the GPU box):

    OPENBLAS_NUM_THREADS=1 python tests/golden/make_golden.py

It imports the unmodified reference modules (utils, linearization, augmented, horizon_selection,
solver, systems, run_suite) and records inputs + outputs of every function on the HOP hot path
 into compressed .npz files next to this script.  tests/test_oracle.py pins the C
oracle against these files; tests/test_gpu_parity.py compares the CUDA path against them as well.
Nothing here is imported by the product.
"""
import os
import sys

os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
REF = os.environ.get("HOP_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
sys.dont_write_bytecode = True

import numpy as np  # noqa: E402

import augmented  # noqa: E402
import horizon_selection  # noqa: E402
import linearization  # noqa: E402
import solver  # noqa: E402
import systems  # noqa: E402
import utils  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def stack(lst):
    return np.stack([np.asarray(a, dtype=float) for a in lst])


def save(name, **kw):
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **kw)
    print(f"{name}.npz  {os.path.getsize(path) / 1024:.1f} KiB")


# ------------------------------------------------------------------ utils: chol_inv / chol_solve / wrap
def gold_utils():
    rng = np.random.default_rng(1234)
    mats, invs, kinds = [], [], []
    for d in (1, 2, 3, 4, 5, 12, 13):
        for kind in ("spd", "illcond", "semidef", "indef", "indef_big"):
            M = rng.standard_normal((d, d))
            if kind == "spd":
                A = M @ M.T + d * np.eye(d)
            elif kind == "illcond":
                s = np.logspace(0, -10, d)
                Qm, _ = np.linalg.qr(M)
                A = (Qm * s) @ Qm.T
            elif kind == "semidef":
                v = rng.standard_normal((d, max(d - 1, 1)))
                A = v @ v.T if d > 1 else np.zeros((1, 1))
            elif kind == "indef":
                A = M @ M.T - 1e-5 * np.eye(d) * (np.arange(d) == d - 1) * 50.0
                A[-1, -1] -= 2e-4 + (M @ M.T)[-1, -1] * 0  # small negative direction for some d
                A = A - 3e-4 * np.outer(np.ones(d), np.ones(d)) / d
            else:
                A = M + M.T  # strongly indefinite: exhausts the ladder -> LU fallback
            A = A + 1e-3 * rng.standard_normal((d, d))  # slightly non-symmetric input (sym() is part of the function)
            mats.append(np.pad(A, ((0, 13 - d), (0, 13 - d))))
            invs.append(np.pad(utils.chol_inv(A), ((0, 13 - d), (0, 13 - d))))
            kinds.append(f"{kind}:{d}")
    # chol_solve
    sA, sB, sX = [], [], []
    for d, c in ((1, 1), (1, 4), (4, 1), (4, 12), (2, 4)):
        M = rng.standard_normal((d, d))
        A = M @ M.T + 0.1 * np.eye(d)
        Bm = rng.standard_normal((d, c))
        sA.append(np.pad(A, ((0, 4 - d), (0, 4 - d))))
        sB.append(np.pad(Bm, ((0, 4 - d), (0, 12 - c))))
        sX.append(np.pad(utils.chol_solve(A, Bm), ((0, 4 - d), (0, 12 - c))))
    angles = np.concatenate([np.linspace(-25, 25, 101), [np.pi, -np.pi, 3 * np.pi, -3 * np.pi, 0.0, 1e-300, 2 * np.pi]])
    wrapped = np.array([utils.angle_normalize(float(a)) for a in angles])
    save("utils", mats=stack(mats), invs=stack(invs), kinds=np.array(kinds),
         sA=stack(sA), sB=stack(sB), sX=stack(sX), sdims=np.array([(1, 1), (1, 4), (4, 1), (4, 12), (2, 4)]),
         angles=angles, wrapped=wrapped)


# ------------------------------------------------------------------ cases
def quadrotor_128():
    t = list(systems.make_quadrotor(N=128))
    t[10] = min(t[10], t[8])  # T_max clipped to N
    return tuple(t)


CASES = {
    "DoubleIntegrator": systems.make_double_integrator,
    "Cartpole_SwingUp": systems.make_cartpole_swingup,
    "Quadrotor": quadrotor_128,
    "Segway_Balance": systems.make_segway_balance,
}


def gold_dynamics():
    rng = np.random.default_rng(99)
    out = {}
    for name, maker in CASES.items():
        F, x0, xg, u_ref = maker()[:4]
        n, m = x0.size, u_ref.size
        xs = x0[None, :] + rng.standard_normal((40, n)) * 0.7
        us = u_ref[None, :] + rng.standard_normal((40, m)) * 0.5
        if name == "Quadrotor":
            xs[35, 7] = np.pi / 2            # Euler singularity guard
            xs[36, 10] = 2e3                 # omega guard
            xs[37, 0] = 2e6                  # norm guard
            xs[38, 3] = np.nan               # non-finite guard
        if name in ("Cartpole_SwingUp", "Segway_Balance"):
            xs[30:35, 2] = [3.1, -3.1, 6.0, -6.0, np.pi]   # exercise the floored-mod wrap
            xs[30:35, 3] = [5.0, -5.0, 20.0, -20.0, 0.0]
        fs = np.stack([np.asarray(F(x, u), dtype=float) for x, u in zip(xs, us)])
        out[name + "_x"] = xs
        out[name + "_u"] = us
        out[name + "_f"] = fs
    save("dynamics", **out)


def selection(F, A_list, B_list, X, U, xg, u_ref, Q, R, alpha, w, T_min, T_max, wrap_idx):
    A_aug, B_aug, Q_aug, R_list, z0, R_inv = augmented.build_augmented_sequence_QR(
        F, A_list, B_list, X, U, xg, u_ref, Q, R, w, wrap_idx=wrap_idx)
    QT = augmented.build_terminal_aug_list(X, xg, alpha, wrap_idx=wrap_idx)
    J = horizon_selection.propagator_all_Jt_aug(A_aug, B_aug, Q_aug, R_list, z0, QT, T_use=T_max, R_inv_cached=R_inv)
    T = int(np.argmin(J[T_min - 1:T_max]) + T_min)
    return A_aug, B_aug, Q_aug, z0, R_inv, QT, J, T


def gold_cases():
    for name, maker in CASES.items():
        F, x0, xg, u_ref, Q, R, alpha, w, N, T_min, T_max, wrap_idx, _ = maker()
        n, m = x0.size, u_ref.size
        U = np.tile(u_ref.reshape(1, -1), (N, 1))
        X = solver.rollout(F, x0, U)
        A_f, B_f = linearization.linearize_forward_diff_traj(F, X, U)
        A_c, B_c = linearization.linearize_central_diff_traj(F, X, U)
        a = stack(linearization.compute_affine_residuals(F, X, U)).reshape(N, n)
        A_aug, B_aug, Q_aug, z0, R_inv, QT, J0, T0 = selection(F, A_f, B_f, X, U, xg, u_ref, Q, R, alpha, w, T_min, T_max, wrap_idx)
        # warm-start step exactly as solver.py:541-551
        k_list, K_list, ok = solver.backward_pass_truncated(A_f, B_f, X, U, xg, u_ref, Q, R, alpha, T0,
                                                            lm_lambda=1e-3, wrap_idx=wrap_idx)
        assert ok
        X1, U1, J1, acc1 = solver.forward_linesearch_fixedT(F, X, U, xg, u_ref, Q, R, alpha, w, T0, k_list, K_list,
                                                            wrap_idx=wrap_idx)
        cost0 = solver.cost_timeopt_true(X, U, xg, u_ref, Q, R, alpha, w, T0, wrap_idx)
        # full solve, run_suite defaults (run_suite.py:233-237): max_iter=12, forward differences
        res = solver.ilqr_timeopt_ourmethod(F, x0, xg, u_ref, Q, R, alpha, w, N, T_min, T_max, max_iter=12,
                                            S_window=20, use_central_diff=False, wrap_idx=wrap_idx)
        # selection on the converged trajectory (re-linearised)
        Xc, Uc = res["X"], res["U"]
        A_cv, B_cv = linearization.linearize_forward_diff_traj(F, Xc, Uc)
        *_, Jc, Tc = selection(F, A_cv, B_cv, Xc, Uc, xg, u_ref, Q, R, alpha, w, T_min, T_max, wrap_idx)
        # brute-force comparator on the nominal (loose sanity only)
        J_bf = solver.bruteforce_all_Jt_backward_expansion(A_f, B_f, X, U, xg, u_ref, Q, R, alpha, w, min(T_max, 48),
                                                           wrap_idx=wrap_idx)
        ks = sorted(set([0, 1, N // 2, T_max - 1]))
        save("case_" + name,
             x0=x0, xg=xg, u_ref=u_ref, Q=Q, R=R, alpha=np.asarray(alpha, dtype=float), w=w, N=N, T_min=T_min,
             T_max=T_max, wrap_idx=np.asarray(wrap_idx, dtype=int), dt=F.dt,
             U=U, X=X, A_fwd=stack(A_f), B_fwd=stack(B_f), A_cen=stack(A_c), B_cen=stack(B_c), a_resid=a,
             ks=np.asarray(ks), A_aug_ks=stack([A_aug[k] for k in ks]), B_aug_ks=stack([B_aug[k] for k in ks]),
             Q_aug_ks=stack([Q_aug[k] for k in ks]), QT_ks=stack([QT[k] for k in ks]), z0=z0, R_inv=R_inv,
             J_curve0=J0, T0=T0, cost0=cost0,
             k_list=stack(k_list), K_list=stack(K_list), X1=X1, U1=U1, J1=J1, acc1=acc1,
             sol_X=Xc, sol_U=Uc, sol_J_hist=np.asarray(res["J_hist"]), sol_T_hist=np.asarray(res["T_hist"]),
             sol_J_curve=np.asarray(res["J_curve"]), sol_T_star=res["T_star"],
             conv_J_curve=Jc, conv_T=Tc, J_bruteforce48=J_bf)


def gold_s1_batch():
    """S1 workload: quadrotor N=128, x0_b = x0 + sigma*xi_b, iteration-0 selection."""
    F, x0, xg, u_ref, Q, R, alpha, w, N, T_min, T_max, wrap_idx, _ = quadrotor_128()
    sigma = np.array([0.4, 0.4, 0.4] + [0.0] * 9)
    xi = np.random.default_rng(0).standard_normal((16, 12))
    U = np.tile(u_ref.reshape(1, -1), (N, 1))
    Js, Ts, x0s = [], [], []
    for b in range(16):
        xb = x0 + sigma * xi[b]
        X = solver.rollout(F, xb, U)
        A_f, B_f = linearization.linearize_forward_diff_traj(F, X, U)
        *_, J, T = selection(F, A_f, B_f, X, U, xg, u_ref, Q, R, alpha, w, T_min, T_max, wrap_idx)
        Js.append(J); Ts.append(T); x0s.append(xb)
    save("s1_quadrotor_batch", x0=stack(x0s), J=stack(Js), T=np.asarray(Ts))


def _s1_ref_one(xb):
    F, x0, xg, u_ref, Q, R, alpha, w, N, T_min, T_max, wrap_idx, _ = quadrotor_128()
    U = np.tile(u_ref.reshape(1, -1), (N, 1))
    X = solver.rollout(F, xb, U)
    A_f, B_f = linearization.linearize_forward_diff_traj(F, X, U)
    *_, J, T = selection(F, A_f, B_f, X, U, xg, u_ref, Q, R, alpha, w, T_min, T_max, wrap_idx)
    cols = T + np.arange(-2, 3)
    J5 = np.array([J[c - 1] if 1 <= c <= T_max else np.nan for c in cols])
    gap = np.sort(J[T_min - 1:T_max])[:2]
    return T, J5, (gap[1] - gap[0]) / abs(gap[0])


def gold_s1_ref4096(B=4096, seed=0):
    """S1 at scale through the REAL reference: 4096 instances (the first 4096 of bench.py's rank-0 batch: same generator,
    same seed), one core-minute each 17; stored: the seed, T*, J at T*-2..T*+2 and the reference's own argmin gap."""
    import multiprocessing as mp
    F, x0 = quadrotor_128()[:2]
    sigma = np.array([0.4, 0.4, 0.4] + [0.0] * 9)
    xi = np.random.default_rng(seed).standard_normal((B, 12))
    x0s = x0[None] + sigma[None] * xi
    with mp.Pool(os.cpu_count()) as pool:
        res = pool.map(_s1_ref_one, list(x0s), chunksize=16)
    save("s1_quadrotor_ref4096", seed=seed, T=np.asarray([r[0] for r in res], dtype=np.int32),
         J_pm2=np.stack([r[1] for r in res]), gap=np.asarray([r[2] for r in res]))


def gold_resid(B=4):
    """Non-zero affine residuals: X is perturbed AFTER the rollout and U around u_ref, so a_k = F(X_k, U_k) - X_{k+1} != 0
    (linearization.py:269-270; ~1e-4) and du != 0 enter A_aug[:, n] = a_k - B_k du (augmented.py:50).  The reference computes a_k
    itself inside build_augmented_sequence_QR; it is stored so that the CUDA entry points get the same input."""
    F, x0, xg, u_ref, Q, R, alpha, w, N, T_min, T_max, wrap_idx, _ = quadrotor_128()
    rng = np.random.default_rng(77)
    sigma = np.array([0.4, 0.4, 0.4] + [0.0] * 9)
    out = {k: [] for k in ("X", "U", "A", "B", "a_resid", "J", "T")}
    for b in range(B):
        U = np.tile(u_ref.reshape(1, -1), (N, 1)) + 0.02 * rng.standard_normal((N, u_ref.size))
        X = solver.rollout(F, x0 + sigma * rng.standard_normal(12), U)
        X = X + 1e-4 * rng.standard_normal(X.shape)
        A_f, B_f = linearization.linearize_forward_diff_traj(F, X, U)
        a = stack(linearization.compute_affine_residuals(F, X, U)).reshape(N, x0.size)
        *_, J, T = selection(F, A_f, B_f, X, U, xg, u_ref, Q, R, alpha, w, T_min, T_max, wrap_idx)
        for k, v in zip(out, (X, U, stack(A_f), stack(B_f), a, J, T)):
            out[k].append(v)
    save("resid_Quadrotor", **{k: (np.asarray(v, dtype=np.int64) if k == "T" else stack(v)) for k, v in out.items()})


def _ddp_ref_one(arg):
    name, x0 = arg
    maker = quadrotor_128 if name == "Quadrotor" else CASES[name]
    F, _x0, xg, u_ref, Q, R, alpha, w, N, T_min, T_max, wrap_idx, _ = maker()
    res = solver.ilqr_timeopt_ourmethod(F, x0, xg, u_ref, Q, R, alpha, w, N, T_min, T_max, max_iter=12, S_window=20,
                                        use_central_diff=False, wrap_idx=wrap_idx)
    return np.asarray(res["T_hist"], dtype=np.int32), np.asarray(res["J_hist"], dtype=float), int(res["T_star"])


def gold_ddp_batch():
    """Full HOP-DDP solves of the REAL reference (solver.ilqr_timeopt_ourmethod, run_suite defaults: max_iter = 12, forward
    differences) on sampled initial states: the configurations 2-4 of BASELINE.json at a size the reference finishes in
    minutes -- Segway 25 trials (sigma .02, run_suite.py:73), Cartpole 48 instances (x0 ~ N(0, diag(.1,.1,.2,.2)^2), the
    distribution tests/run_configs.py uses: the reference's own sigma is zero), Quadrotor N = 128 16 instances (sigma .4 on the
    position).  Stored: x0, T_hist / J_hist (padded to 13), n_hist, T_star."""
    import multiprocessing as mp
    out = {}
    jobs = []
    rng = np.random.default_rng(2024)
    x0s = {}
    base = CASES["Segway_Balance"]()[1]
    x0s["Segway_Balance"] = base[None] + 0.02 * rng.standard_normal((25, 4)); x0s["Segway_Balance"][0] = base
    x0s["Cartpole_SwingUp"] = np.array([rng.normal(0, .1, 48), rng.normal(0, .1, 48), rng.normal(0, .2, 48), rng.normal(0, .2, 48)]).T
    x0s["Cartpole_SwingUp"][0] = CASES["Cartpole_SwingUp"]()[1]
    base = quadrotor_128()[1]
    x0s["Quadrotor"] = base[None] + np.array([0.4, 0.4, 0.4] + [0.0] * 9)[None] * rng.standard_normal((16, 12))
    for name, xs in x0s.items():
        jobs += [(name, x) for x in xs]
    with mp.Pool(os.cpu_count()) as pool:
        res = pool.map(_ddp_ref_one, jobs, chunksize=1)
    i = 0
    for name, xs in x0s.items():
        B = len(xs)
        Th = np.zeros((B, 13), dtype=np.int32); Jh = np.full((B, 13), np.nan); nh = np.zeros(B, dtype=np.int32); Ts = np.zeros(B, dtype=np.int32)
        for b in range(B):
            T_hist, J_hist, T_star = res[i]; i += 1
            nh[b] = len(T_hist); Th[b, :nh[b]] = T_hist; Jh[b, :nh[b]] = J_hist; Ts[b] = T_star
        out.update({name + "_x0": xs, name + "_T_hist": Th, name + "_J_hist": Jh, name + "_n_hist": nh, name + "_T_star": Ts})
    save("ddp_batch", **out)


def s2_instance(seed, d, m, N):
    """Synthetic HOP-LQR generator S2 -- draw order: all A, all B, all Q, R, z0, w."""
    r = np.random.default_rng(seed)
    A = np.eye(d)[None] + 0.05 * r.standard_normal((N, d, d)) / np.sqrt(d)
    B = 0.05 * r.standard_normal((N, d, m))
    Q = np.stack([np.diag(r.uniform(0.5, 2.0, d)) for _ in range(N)])
    R = np.diag(r.uniform(0.05, 0.5, m))
    z0 = r.standard_normal(d)
    w = r.uniform(0.01, 0.1)
    QT = np.tile(50.0 * np.eye(d), (N, 1, 1))
    return A, B, Q, R, z0, w, QT


def gold_s2():
    out = {}
    for (d, m, N) in ((4, 2, 64), (12, 4, 128), (13, 4, 128), (13, 4, 256), (5, 1, 64), (3, 1, 32)):
        for seed in (0, 1):
            A, B, Q, R, z0, w, QT = s2_instance(seed, d, m, N)
            R_inv = utils.chol_inv(R)
            J = horizon_selection.propagator_all_Jt_aug(list(A), list(B), list(Q), [R] * N, z0, list(QT), T_use=N,
                                                        R_inv_cached=R_inv)
            out[f"J_d{d}_m{m}_N{N}_s{seed}"] = J
            out[f"w_d{d}_m{m}_N{N}_s{seed}"] = w
    save("s2_synthetic", **out)


def gold_signatures():
    """Public signatures of the reference's hot-path API (parameter names, kinds, defaults)."""
    import inspect
    import json
    import run_suite
    mods = {"utils": utils, "linearization": linearization, "augmented": augmented, "horizon_selection": horizon_selection,
            "solver": solver, "systems": systems, "run_suite": run_suite}
    names = {
        "utils": ["_sym", "as_terminal_weight", "chol_inv", "chol_solve", "angle_normalize", "wrap_error"],
        "linearization": ["linearize_central_diff_traj", "linearize_forward_diff_traj", "compute_affine_residuals"],
        "augmented": ["build_augmented_sequence_QR", "build_terminal_aug_list"],
        "horizon_selection": ["propagator_all_Jt_aug"],
        "solver": ["rollout", "cost_timeopt_true", "backward_pass_truncated", "forward_linesearch_fixedT", "ilqr_timeopt",
                   "ilqr_timeopt_ourmethod", "ilqr_timeopt_baseline1", "ilqr_timeopt_baseline2"],
        "systems": ["make_double_integrator", "make_cartpole_swingup", "make_quadrotor", "make_segway_balance"],
        "run_suite": ["sample_x", "run_case", "main"],
    }
    out = {}
    for mod, fns in names.items():
        for fn in fns:
            sig = inspect.signature(getattr(mods[mod], fn))
            out[f"{mod}.{fn}"] = [[p.name, p.kind.name, None if p.default is inspect.Parameter.empty else repr(p.default)]
                                  for p in sig.parameters.values()]
    out["run_suite.CASES"] = [c[0] for c in run_suite.CASES]
    out["run_suite.SOLVERS"] = sorted(run_suite.SOLVERS)
    with open(os.path.join(OUT, "signatures.json"), "w") as fh:
        json.dump(out, fh, indent=1, sort_keys=True)
    print("signatures.json", len(out), "entries")


if __name__ == "__main__":
    ALL = {"signatures": gold_signatures, "utils": gold_utils, "dynamics": gold_dynamics, "cases": gold_cases,
           "s1_batch": gold_s1_batch, "s2": gold_s2, "s1_ref4096": gold_s1_ref4096, "resid": gold_resid, "ddp_batch": gold_ddp_batch}
    for name in (sys.argv[1:] or list(ALL)):      # no arguments: everything (s1_ref4096 takes ~5 core-minutes)
        ALL[name]()
