"""The REAL reference (dmmsjtu-umich/time-opt-ilqr, pure Python + numpy) as a checker / timed baseline -- TEST
INFRASTRUCTURE ONLY (tests/, __graft_entry__.smoke() and bench.py's CPU legs).

Recipe: `stage()` (called by __graft_entry__.build() in the build container, where /root/reference is mounted) copies the
reference's modules, unmodified, into oracle/_ref/.  That directory is git-ignored -- reference sources never enter the
repository's history -- but it is not gpurun-ignored, so it travels to the GPU box with the built libraries, and the
benchmark can run the reference itself there (`cpu_baseline.reference_python`, `--impl reference`).  When oracle/_ref is
absent everything here reports "unavailable" and the callers fall back to the C port (oracle/hop_oracle.c).
"""
from __future__ import annotations

import os
import shutil
import sys
import time
import types

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(_HERE, "_ref")
MODULES = ("utils", "linearization", "augmented", "horizon_selection", "solver", "systems", "run_suite", "ilqr_propagator")


def stage(src: str = "/root/reference") -> bool:
    """Copy the reference's modules into oracle/_ref (no-op when `src` does not exist).  Returns availability."""
    if os.path.isdir(src):
        os.makedirs(REF_DIR, exist_ok=True)
        for mod in MODULES:
            f = os.path.join(src, mod + ".py")
            if os.path.exists(f):
                shutil.copyfile(f, os.path.join(REF_DIR, mod + ".py"))
        plots = os.path.join(src, "plots", "summary.csv")          # the only goldens the reference ships
        if os.path.exists(plots):
            shutil.copyfile(plots, os.path.join(REF_DIR, "plots_summary.csv"))
    return available()


def available() -> bool:
    return all(os.path.exists(os.path.join(REF_DIR, m + ".py")) for m in MODULES[:6])


def _import():
    """Import the staged reference modules under their own names (they import each other by bare name)."""
    if not available():
        raise RuntimeError("oracle/_ref is not staged (run __graft_entry__.build() where /root/reference exists)")
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    for name in ("matplotlib", "matplotlib.pyplot"):              # ilqr_propagator imports pyplot at module level; not installed
        if name not in sys.modules:
            try:
                __import__(name)
            except ImportError:
                sys.modules[name] = types.ModuleType(name)
    if not hasattr(sys.modules["matplotlib"], "pyplot"):
        sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    import importlib
    return {m: importlib.import_module(m) for m in ("utils", "linearization", "augmented", "horizon_selection", "solver", "systems")}


_CASE = None


def _quadrotor_128():
    global _CASE
    if _CASE is None:
        m = _import()
        t = list(m["systems"].make_quadrotor(N=128))
        t[10] = min(t[10], t[8])                                   # T_max clipped to N
        _CASE = (m, tuple(t))
    return _CASE


def s1_select_one(x0):
    """The reference's own code path for one S1 instance: solver.rollout -> linearization.linearize_forward_diff_traj ->
    augmented.build_augmented_sequence_QR + build_terminal_aug_list -> horizon_selection.propagator_all_Jt_aug -> argmin
    (solver.py:492,504-522).  Returns (T*, J[T_max])."""
    m, (F, _x0, xg, u_ref, Q, R, alpha, w, N, T_min, T_max, wrap_idx, _) = _quadrotor_128()
    U = np.tile(u_ref.reshape(1, -1), (N, 1))
    X = m["solver"].rollout(F, np.asarray(x0, dtype=float), U)
    A, B = m["linearization"].linearize_forward_diff_traj(F, X, U)
    A_aug, B_aug, Q_aug, R_list, z0, R_inv = m["augmented"].build_augmented_sequence_QR(
        F, A, B, X, U, xg, u_ref, Q, R, w, wrap_idx=wrap_idx)
    QT = m["augmented"].build_terminal_aug_list(X, xg, alpha, wrap_idx=wrap_idx)
    J = m["horizon_selection"].propagator_all_Jt_aug(A_aug, B_aug, Q_aug, R_list, z0, QT, T_use=T_max, R_inv_cached=R_inv)
    return int(np.argmin(J[T_min - 1:T_max]) + T_min), np.asarray(J, dtype=float)


def _worker(x0):
    return s1_select_one(x0)


class Pool:
    """Process pool over the host cores (the reference is single-threaded Python; OPENBLAS_NUM_THREADS=1)."""

    def __init__(self, procs=None):
        import multiprocessing as mp
        os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
        self.procs = int(procs or os.cpu_count() or 1)
        self.pool = mp.get_context("fork").Pool(self.procs)
        self.pool.map(_worker_warm, range(self.procs))

    def s1_select(self, x0s):
        """(T [B], J [B, T_max], seconds) for a batch of initial states."""
        t = time.perf_counter()
        res = self.pool.map(_worker, list(np.asarray(x0s, dtype=float)), chunksize=max(1, len(x0s) // (4 * self.procs)))
        dt = time.perf_counter() - t
        return np.asarray([r[0] for r in res], dtype=np.int32), np.stack([r[1] for r in res]), dt

    def close(self):
        self.pool.close()
        self.pool.join()


def _worker_warm(_):
    try:                                                           # OPENBLAS_NUM_THREADS is read at numpy import time: too late in a forked worker
        from threadpoolctl import threadpool_limits
        threadpool_limits(1)
    except Exception:
        pass
    _quadrotor_128()
    return 0


# ---- the reference's brute-force curve (solver.py:293-358), config 5's CPU comparator -------------------------------
def _bf_one(arg):
    (A, B, X, U, xg, u_ref, Q, R, alpha, w), N = arg
    solver = _import()["solver"]
    return solver.bruteforce_all_Jt_backward_expansion(list(A), list(B), X, U, xg, u_ref, Q, R, alpha, w, N, lm_lambda=0.0)


def bruteforce_rate(insts, N, procs=None):
    """solves/s of solver.bruteforce_all_Jt_backward_expansion over `insts` (tuples A, B, X, U, xg, u_ref, Q, R, alpha, w)
    with one worker process per core."""
    import multiprocessing as mp
    procs = int(procs or os.cpu_count() or 1)
    with mp.get_context("fork").Pool(procs) as pool:
        pool.map(_worker_warm_import, range(procs))
        t = time.perf_counter()
        pool.map(_bf_one, [(a, N) for a in insts])
        return len(insts) / (time.perf_counter() - t)


def _worker_warm_import(_):
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(1)
    except Exception:
        pass
    _import()
    return 0
