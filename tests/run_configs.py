#!/usr/bin/env python
"""Runs the five BASELINE.json configurations on the B200 path at their stated sizes and prints one JSON line each
(device time, throughput, parity against the CPU oracle on a bounded sample).  These are the parity / scale cases
around the headline workload that bench.py times; nothing here feeds bench.py.

  python tests/run_configs.py [--configs 1,2,3,4,5] [--small]
  torchrun --nproc-per-node G ... tests/run_configs.py --configs 4,5     # batch sharded over G GPUs (hop.dist)

The oracle (oracle/, plain-C restatement of the reference) is the checker only; it never computes a reported result.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))   # repo root (this file lives in tests/: it uses the oracle as a checker)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "time-opt-ilqr_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def f_alg(N, d, m):
    return N * (5 * d**3 + 2 * d * m * m + 2 * d * d * m) + (N - 1) * (11 * d**3 + 3 * d * d) + N * (7 * d**3 + 4 * d * d)


def b_alg(N, d, m):
    return 8 * (N * (3 * d * d + d * m) + m * m + d) + 8 * N + 4


def timed(fn, reps=1, best=False):
    """Device time of fn() with CUDA events: the mean of `reps` back-to-back calls, or with best=True the fastest of `reps`
    individually timed calls (robust against a one-off allocator / clock-ramp hiccup on one rank of a multi-GPU run)."""
    torch.cuda.synchronize()
    if best:
        ts = []
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e-3)
        return out, min(ts)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return out, e0.elapsed_time(e1) * 1e-3 / reps


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def emit(rec):
    rank, G = world()
    rec["n_gpus"] = G
    if rank == 0:
        print(json.dumps(rec), flush=True)


def max_over_ranks(t, dev):
    rank, G = world()
    if G == 1:
        return t
    x = torch.tensor([t], dtype=torch.float64, device=dev)
    dist.all_reduce(x, op=dist.ReduceOp.MAX)
    return float(x.item())


def cfg1(args, dev):
    """DoubleIntegrator HOP-LQR horizon selection, single trial, through the drop-in run_suite CLI."""
    import subprocess
    import tempfile
    import pandas as pd
    from _common import golden
    rank, _ = world()
    if rank != 0:
        return
    out = tempfile.mkdtemp()
    env = dict(os.environ, PYTHONPATH=os.path.join(ROOT, "time-opt-ilqr_b200", "dropin"), PYTHONHASHSEED="0")
    t0 = time.perf_counter()
    subprocess.check_call([sys.executable, os.path.join(ROOT, "time-opt-ilqr_b200", "dropin", "run_suite.py"), "--cases",
                           "DoubleIntegrator", "--trials", "1", "--outdir", out], env=env, stdout=subprocess.DEVNULL)
    wall = time.perf_counter() - t0
    df = pd.read_csv(os.path.join(out, "summary_all.csv"))
    g = golden("case_DoubleIntegrator")
    row = df[df["solver"] == "ourmethod"].iloc[0] if "solver" in df.columns else df.iloc[0]
    emit({"config": 1, "what": "DoubleIntegrator, run_suite.py --cases DoubleIntegrator --trials 1 (drop-in CLI)",
          "wall_s_incl_process_start": wall, "T_star": int(row["T_star"]), "J_star": float(row["J_star"]),
          "reference_T_star": int(g["sol_T_hist"][-1]), "reference_J_star": float(g["sol_J_hist"][-1]),
          "T_star_equal": int(row["T_star"]) == int(g["sol_T_hist"][-1]),
          "rel_J_star": abs(float(row["J_star"]) - float(g["sol_J_hist"][-1])) / abs(float(g["sol_J_hist"][-1]))})


def _ddp(name, x0s, dev, max_iter, check, label, cfg, N=None, mode=None):
    import oracle as O
    from _common import rel
    from hop import api, cases, dist as hdist
    case = cases.make_case(name, N=N) if N else cases.make_case(name)
    F, x0, xg, u_ref, Q, R, alpha, w, Nn, T_min, T_max, wrap_idx, _ = case
    rank, G = world()
    lo, hi = hdist.shard_bounds(len(x0s), rank, G)
    mine = torch.as_tensor(x0s[lo:hi], device=dev)
    mode = api.MODE_FAST if mode is None else mode
    run = lambda: api.ilqr_timeopt_batched(case, mine, max_iter=max_iter, use_central_diff=False, mode=mode)  # noqa: E731
    run()                                                                       # warm-up (allocator, module load)
    r, dt = timed(run)
    dt = max_over_ranks(dt, dev)
    nh = r["n_hist"].cpu().numpy(); Th = r["T_hist"].cpu().numpy(); Jh = r["J_hist"].cpu().numpy()
    st = r["status"].cpu().numpy()
    rec = {"config": cfg, "what": label, "instances": len(x0s), "max_iter": max_iter, "device_s": dt,
           "solves_per_s": len(x0s) / dt, "outer_iterations_run": r["iters"], "crashed": int(((st & 0xFF) != 0).sum()),
           "phase_seconds_rank0": r["timers"]}
    if rank == 0 and check > 0 and not os.environ.get("HOP_CFG_NO_PARITY"):           # (timing-only A/B runs skip the CPU census)
        k = min(check, hi - lo)
        rec["parity_vs_oracle"] = ddp_census(case, Nn, x0s[:k], max_iter, {"fast": (nh, Th, Jh, r["T_star"].cpu().numpy())}, dev)
    emit(rec)


def ddp_census(case, Nn, x0s, max_iter, runs, dev):
    """HOP-DDP parity census against the oracle on identical initial states.

    The iterated solve amplifies rounding on ill-conditioned embeddings (cartpole: Q has a zero weight, |E_k| ~ 5e8): the
    REFERENCE COMPUTATION ITSELF does not reproduce its T_hist when it is perturbed at rounding level (its
    curve moves by 4e-2 when only the OpenBLAS kernel set changes).  The census measures that directly.  Six perturbed
    oracle runs -- the selection sweep in x87 extended precision, and the fp64 oracle started from x0 +- 1e-15,
    x0 (1 +- 4e-16) and x0 + 3e-15 -- give (i) the oracle's own flip rate per perturbation and (ii) the set of WELL-POSED
    instances (T_hist identical in all seven runs).  A device run is then one more rounding-level perturbation (CUDA's
    sin/cos differ from glibc's by <= 2 ulp; FAST also re-orders the selection arithmetic): its mismatch rate is reported
    next to the oracle's own flip rates, and its T_hist on the well-posed instances must be the oracle's.
    `runs` maps a label to a device result (n_hist, T_hist, J_hist, T_star); HOP_MODE_EXACT is always run here in addition."""
    from oracle import census
    from _common import rel
    from hop import api
    k = len(x0s)
    th = os.cpu_count() or 1
    o, well, self_flip, t_cpu = census.ddp_oracle_census(case, Nn, x0s, max_iter, th)
    ex = api.ilqr_timeopt_batched(case, torch.as_tensor(x0s, device=dev), max_iter=max_iter, use_central_diff=False,
                                  mode=api.MODE_EXACT)
    runs = dict(runs, exact=(ex["n_hist"].cpu().numpy(), ex["T_hist"].cpu().numpy(), ex["J_hist"].cpu().numpy(),
                             ex["T_star"].cpu().numpy()))
    out = {"checked": k, "well_posed": int(well.sum()),
           "well_posed_rule": "T_hist identical among the fp64 oracle and six rounding-level perturbations of it (fp80 selection "
                              "sweep; x0 +- 1e-15; x0 (1 +- 4e-16); x0 + 3e-15)",
           "oracle_self_flip_rate": self_flip,
           "oracle_solves_per_s_all_host_threads": k / t_cpu, "host_threads": th}
    for label, (nh, Th, Jh, Ts) in runs.items():
        same_T = np.array([nh[b] == o["n_hist"][b] and np.array_equal(Th[b, :nh[b]], o["T_hist"][b, :o["n_hist"][b]]) for b in range(k)])
        relJ = [rel(Jh[b, :nh[b]], o["J_hist"][b, :nh[b]]) for b in range(k) if same_T[b]]
        out[label] = {"T_hist_identical": int(same_T.sum()), "mismatch_rate": float(1.0 - same_T.mean()),
                      "T_hist_identical_among_well_posed": int((same_T & well).sum()),
                      "well_posed_but_different": [int(b) for b in np.nonzero(well & ~same_T)[0][:16]],
                      "max_rel_J_hist_where_T_identical": float(max(relJ)) if relJ else None,
                      "T_star_identical": int((Ts[:k] == o["T_star"][:k]).sum())}
    return out


def cfg2(args, dev):
    from hop import cases
    case = cases.make_case("Segway_Balance")
    rng = np.random.default_rng(0)
    x0s = case[1][None] + 0.02 * rng.standard_normal((25, 4))                   # run_suite.py:73 sigma_x0 = .02 x 4
    x0s[0] = case[1]                                                            # trial 0 is the nominal case (run_suite.py:109-111)
    _ddp("Segway_Balance", x0s, dev, 12, 25, "Segway_Balance HOP-DDP, 25 trials, max-iter 12, one batched device call", 2)


def cfg3(args, dev):
    B = int(os.environ.get("HOP_CFG3_B", "0")) or (256 if args.small else 4096)    # HOP_CFG3_B: batch-size A/B runs
    rng = np.random.default_rng(0)                                              # the reference's cartpole sigma is 0
    x0s = np.array([rng.normal(0, .1, B), rng.normal(0, .1, B), rng.normal(0, .2, B), rng.normal(0, .2, B)]).T
    _ddp("Cartpole_SwingUp", x0s, dev, 12, B, f"Cartpole swing-up HOP-DDP (augmented homogeneous state), {B} random initial "
         "states x0 ~ N(0, diag(.1,.1,.2,.2)^2), max-iter 12; T* is ill-posed at the reference's own noise level "
         "(argmin gap 1.3e-5 < 3e-5, SURVEY.md s.9)", 3)


def cfg4(args, dev):
    from _common import s1_x0
    B = int(os.environ.get("HOP_CFG4_B", "0")) or (1024 if args.small else 16384)    # HOP_CFG4_B: per-rank load of a strong-scaled run on one GPU
    _ddp("Quadrotor", s1_x0(B, seed=4), dev, 12, 1024, f"Quadrotor 12-DOF HOP-DDP, N=128, {B} initial states, max-iter 12, "
         "batch sharded over the ranks", 4, N=128)


def _bruteforce_cpu_timing(d, m, N, n_inst=None):
    """Config 5's CPU comparator: the reference's brute-force curve (solver.py:293-358 semantics: one Riccati sweep per
    candidate horizon, O(N^2 d^3)) on the sub-variant S2c (time-invariant Q_k = Q_0, X_{k+1} = A_k X_k, X_0 = z0, U = 0,
    alpha = 50, lm_lambda = 0), timed on all host cores: the C port through a thread pool (the ctypes call
    releases the GIL) and, when oracle/_ref is staged, the reference's own Python function on one process per core."""
    import concurrent.futures as cf
    import oracle as O
    from oracle import ref_py
    from _common import s2_instance
    th = os.cpu_count() or 1
    n_inst = n_inst or 2 * th

    def inst(seed):
        A, B, Q, R, z0, w, QT = s2_instance(seed, d, m, N)
        X = np.zeros((N + 1, d)); X[0] = z0
        for k in range(N):
            X[k + 1] = A[k] @ X[k]
        return A, B, X, np.zeros((N, m)), np.zeros(d), np.zeros(m), Q[0], R, 50.0, w
    insts = [inst(s) for s in range(n_inst)]
    one = lambda a: O.bruteforce_all_Jt(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8], a[9], N, lm_lambda=0.0)  # noqa: E731
    one(insts[0])
    t0 = time.perf_counter()
    with cf.ThreadPoolExecutor(th) as ex:
        Jb = list(ex.map(one, insts))
    t_port = time.perf_counter() - t0
    out = {"instances": n_inst, "host_threads": th, "port_solves_per_s": n_inst / t_port,
           "what": "bruteforce_all_Jt_backward_expansion semantics (solver.py:293-358) on S2c, all host cores"}
    # loose sanity: the propagator's 1e-9 jitter has no counterpart in the Riccati sweep (~1e-8)
    A, B, X, U0, xg0, ur0, Q0, R, alpha, w = insts[0]
    Jp = O.propagator_all_Jt(A, B, np.tile(Q0, (N, 1, 1)), O.chol_inv(R), X[0], np.tile(50.0 * np.eye(d), (N, 1, 1)))
    out["rel_propagator_vs_bruteforce_curve"] = float(np.max(np.abs(Jp + w * np.arange(1, N + 1) - Jb[0]) / np.abs(Jb[0])))
    if ref_py.available():
        k = min(n_inst, th)
        out["reference_python_solves_per_s"] = ref_py.bruteforce_rate(insts[:k], N, th)
        out["reference_python_instances"] = k
    return out


def cfg5(args, dev):
    """Synthetic batched HOP-LQR sweep (LQR-boundary entry point hop_select_f64): 2^20 instances for EVERY (d, N).  At d = 12 / 13
    the LQR-boundary inputs of 2^20 instances (0.25 - 1.2 TB) exceed the HBM of 8 GPUs, so the batch is processed as tiles of
    at most ~40 GB per rank: a tile is generated on the device from its seed, selected, checked and dropped; the reported
    time is the sum of the selection calls (generation is not timed)."""
    import oracle as O
    from _common import rel, s2_batch
    from hop import api, dist as hdist
    rank, G = world()
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    for d, m in ((4, 2), (12, 4), (13, 4)):
        for N in (64, 128, 256):
            per = b_alg(N, d, m)
            Btot = (1 << 20) if not args.small else (1 << 14)
            lo, hi = hdist.shard_bounds(Btot, rank, G)
            Bmine = hi - lo
            tile = max(1, min(Bmine, int(40e9 // per)))
            ntiles = -(-Bmine // tile)
            dt_sum, nonzero, first = 0.0, 0, None
            for ti in range(ntiles):
                B = min(tile, Bmine - ti * tile)
                gen = torch.Generator(device=dev); gen.manual_seed(1234 + 1000 * rank + ti)
                A = torch.eye(d, dtype=torch.float64, device=dev).expand(B, N, d, d) + \
                    0.05 / np.sqrt(d) * torch.randn((B, N, d, d), dtype=torch.float64, device=dev, generator=gen)
                Bm = 0.05 * torch.randn((B, N, d, m), dtype=torch.float64, device=dev, generator=gen)
                Q = torch.diag_embed(0.5 + 1.5 * torch.rand((B, N, d), dtype=torch.float64, device=dev, generator=gen))
                QT = (50.0 * torch.eye(d, dtype=torch.float64, device=dev)).expand(B, N, d, d).contiguous()
                Rd = 0.05 + 0.45 * torch.rand((B, m), dtype=torch.float64, device=dev, generator=gen)
                Rinv = torch.diag_embed(1.0 / (Rd + 1e-9))                      # chol_inv of a diagonal R (utils.py:79-85)
                z0 = torch.randn((B, d), dtype=torch.float64, device=dev, generator=gen)
                w = 0.01 + 0.09 * torch.rand((B,), dtype=torch.float64, device=dev, generator=gen)
                run = lambda: api.propagator_all_Jt_aug_batched(A, Bm, Q, Rinv, z0, QT, 1, N, w_explicit=w, mode=api.MODE_FAST)  # noqa: E731
                if ti == 0:
                    run()                                                       # warm-up
                sel, dt = timed(run, reps=3, best=True) if ntiles == 1 else timed(run)
                dt_sum += dt
                nonzero += int((sel.status != 0).sum())
                if ti == 0 and rank == 0:
                    k = 64
                    Jo, sto = O.propagator_batch(A[:k].cpu().numpy(), Bm[:k].cpu().numpy(), Q[:k].cpu().numpy(), Rinv[:k].cpu().numpy(),
                                                 z0[:k].cpu().numpy(), QT[:k].cpu().numpy(), nthreads=os.cpu_count() or 1)
                    tot = Jo + w[:k].cpu().numpy()[:, None] * np.arange(1, N + 1)
                    first = {"checked": k, "max_rel_J": rel(sel.J[:k].cpu().numpy(), Jo),
                             "T_star_identical": int((sel.T_star[:k].cpu().numpy() == np.argmin(tot, 1) + 1).sum())}
                    ex = api.propagator_all_Jt_aug_batched(A[:k], Bm[:k], Q[:k], Rinv[:k], z0[:k], QT[:k], 1, N, w_explicit=w[:k],
                                                           mode=api.MODE_EXACT)
                    first["exact_mode_bit_identical_to_oracle"] = bool(np.array_equal(ex.J.cpu().numpy(), Jo))
                del A, Bm, Q, QT, Rinv, z0, w, sel
                torch.cuda.empty_cache()
            dt = max_over_ranks(dt_sum, dev)
            rec = {"config": 5, "what": f"synthetic HOP-LQR d={d} m={m} N={N}", "batch": Btot, "tiles_per_rank": ntiles,
                   "tile_instances": tile, "device_s": dt,
                   "solves_per_s": Btot / dt, "algorithmic_TFLOPs": Btot * f_alg(N, d, m) / dt / 1e12,
                   "algorithmic_GBs_per_gpu": Bmine * per / dt / 1e9, "hbm_frac_per_gpu": Bmine * per / dt / 1e9 / hbm,
                   "status_nonzero": nonzero}
            if rank == 0:
                # and the reference's own generator (numpy default_rng per instance) on a few instances
                Ar, Br, Qr, Rr, zr, wr, QTr = s2_batch(range(8), d, m, N)
                Rir = np.stack([O.chol_inv(r) for r in Rr])
                s8 = api.propagator_all_Jt_aug_batched(*(torch.as_tensor(x, device=dev) for x in (Ar, Br, Qr, Rir, zr, QTr)), 1, N,
                                                       w_explicit=torch.as_tensor(wr, device=dev), mode=api.MODE_FAST)
                Jo8, _ = O.propagator_batch(Ar, Br, Qr, Rir, zr, QTr)
                first["max_rel_J_reference_generator"] = rel(s8.J.cpu().numpy(), Jo8)
                rec["parity_vs_oracle"] = first
                rec["cpu_bruteforce"] = _bruteforce_cpu_timing(d, m, N)
                rec["speedup_vs_cpu_bruteforce_port_all_cores"] = rec["solves_per_s"] / rec["cpu_bruteforce"]["port_solves_per_s"]
            emit(rec)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="1,2,3,4,5")
    ap.add_argument("--small", action="store_true", help="reduced sizes (smoke run)")
    args = ap.parse_args()
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        dist.init_process_group("nccl", device_id=dev)
    import oracle as O
    O.build()
    fns = {1: cfg1, 2: cfg2, 3: cfg3, 4: cfg4, 5: cfg5}
    for c in (int(x) for x in args.configs.split(",")):
        fns[c](args, dev)
    if dist.is_initialized():
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
