"""CPU, world_size = 2, gloo: the batch-sharding / final-gather logic of hop.dist (the N > 1 path of bench.py).
The per-slice compute is stood in for by the CPU oracle (tests may use it); on the GPUs it is the CUDA path."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle as O
from _common import s1_x0
from hop import cases
from hop import dist as hdist


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _oracle_select(case, x0):
    F, _x0, xg, u_ref, Q, R, alpha, w, N, T_min, T_max, wrap_idx, _ = case
    J, T, st = O.select_from_x0_batch(F.hop_sys, F.hop_params, N, T_min, T_max, x0.numpy(), np.tile(u_ref, (N, 1)), xg, u_ref, Q, R,
                                      alpha, w, wrap_idx, nthreads=2)
    Js = J[np.arange(len(T)), T - 1]
    return torch.from_numpy(T), torch.from_numpy(Js), torch.from_numpy(st)


def _worker(rank, world, port, B, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    case = cases.make_case("Quadrotor", N=48)
    x0 = torch.from_numpy(s1_x0(B, seed=3))
    T, J, st = hdist.sharded_select(lambda x: _oracle_select(case, x), x0)
    lo, hi = hdist.shard_bounds(B, rank, world)
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), T=T.numpy(), J=J.numpy(), st=st.numpy(), lo=lo, hi=hi)
    dist.destroy_process_group()


def test_shard_bounds_cover_the_batch_exactly():
    for B in (0, 1, 7, 64, 65537):
        for world in (1, 2, 4, 8):
            spans = [hdist.shard_bounds(B, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1 and sizes == hdist.shard_sizes(B, world)


def test_two_rank_sharded_selection_equals_unsharded(tmp_path):
    B, world = 7, 2                      # ragged: 4 + 3
    port = _free_port()
    mp.spawn(_worker, args=(world, port, B, str(tmp_path)), nprocs=world, join=True)
    case = cases.make_case("Quadrotor", N=48)
    T0, J0, st0 = _oracle_select(case, torch.from_numpy(s1_x0(B, seed=3)))
    r0, r1 = np.load(tmp_path / "r0.npz"), np.load(tmp_path / "r1.npz")
    assert (int(r0["lo"]), int(r0["hi"]), int(r1["lo"]), int(r1["hi"])) == (0, 4, 4, 7)
    for r in (r0, r1):                   # every rank holds the full gathered result
        assert np.array_equal(r["T"], T0.numpy()) and np.array_equal(r["J"], J0.numpy()) and np.array_equal(r["st"], st0.numpy())
