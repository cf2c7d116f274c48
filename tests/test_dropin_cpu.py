"""CPU-only: the drop-in modules import without a GPU and expose the reference's public signatures
(parameter names, kinds and defaults recorded from the real reference in tests/golden/signatures.json)."""
import importlib
import inspect
import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DROPIN = os.path.join(ROOT, "time-opt-ilqr_b200", "dropin")
SIGS = json.load(open(os.path.join(ROOT, "tests", "golden", "signatures.json")))


@pytest.fixture(scope="module")
def dropin():
    sys.path.insert(0, DROPIN)
    mods = {m: importlib.import_module(m) for m in ("utils", "linearization", "augmented", "horizon_selection", "solver",
                                                    "systems", "run_suite", "ilqr_propagator")}
    yield mods
    sys.path.remove(DROPIN)


def test_signatures_match_the_reference(dropin):
    checked = 0
    for key, ref in SIGS.items():
        mod, fn = key.split(".")
        if fn in ("CASES", "SOLVERS"):
            continue
        sig = inspect.signature(getattr(dropin[mod], fn))
        mine = [[p.name, p.kind.name, None if p.default is inspect.Parameter.empty else repr(p.default)]
                for p in sig.parameters.values()]
        assert mine == ref, f"{key}: {mine} != {ref}"
        checked += 1
    assert checked >= 25


def test_case_and_solver_registries_match(dropin):
    rs = dropin["run_suite"]
    assert [c[0] for c in rs.CASES] == SIGS["run_suite.CASES"]
    assert sorted(rs.SOLVERS) == SIGS["run_suite.SOLVERS"]


def test_baselines_and_python_closures_are_refused_loudly(dropin):
    import numpy as np
    s = dropin["solver"]
    case = dropin["systems"].make_double_integrator()
    with pytest.raises(NotImplementedError):
        s.ilqr_timeopt(*case[:11], method="onepass")
    with pytest.raises(TypeError):
        dropin["linearization"].linearize_forward_diff_traj(lambda x, u: x, np.zeros((3, 2)), np.zeros((2, 1)))


def test_host_dynamics_are_callable_like_the_reference(dropin):
    import numpy as np
    F, x0, xg, u_ref = dropin["systems"].make_quadrotor()[:4]
    xn = F(x0, u_ref)
    assert xn.shape == (12,) and F.dt == 0.05 and np.allclose(xn[:3], x0[:3])
