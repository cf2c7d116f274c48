"""Benchmark systems of the reference (systems.py) as *device-registered* dynamics.

The reference hands the solver a Python closure ``F(x, u)`` (systems.py:30-33,72-95,170-210,
321-333).  A closure cannot run inside a CUDA kernel, so every system here is a ``Dynamics``
object: still callable as ``F(x, u)`` with ``F.dt`` (drop-in for user code), but it also carries
``F.hop_sys`` (device function id, see include/hop_b200.h ``hop_sys_t``) and ``F.hop_params`` (the
parameter vector the device function reads).  The ``make_*`` factories return the reference's
13-tuple ``(F, x0, xg, u_ref, Q, R, alpha, w, N, T_min, T_max, wrap_idx, extra)``.

Out of scope: ``make_pointmass_navigation`` (needs the Python ``extra_stage_cost``
callback, commented out of the reference's CASES) and the legacy Ballbot.
"""
from __future__ import annotations

import math

import numpy as np

SYS_DOUBLE_INTEGRATOR, SYS_CARTPOLE, SYS_QUADROTOR, SYS_SEGWAY = 0, 1, 2, 3
SYS_DIMS = {SYS_DOUBLE_INTEGRATOR: (2, 1), SYS_CARTPOLE: (4, 1), SYS_QUADROTOR: (12, 4), SYS_SEGWAY: (4, 1)}
NPARAMS = 16
_TWO_PI = 2.0 * math.pi


def _wrap_pi(a: float) -> float:
    """Floored-modulo wrap to [-pi, pi) (utils.py:127-128)."""
    return (a + math.pi) % _TWO_PI - math.pi


class Dynamics:
    """Callable discrete dynamics with a device twin (``hop_sys``/``hop_params``)."""

    def __init__(self, sys_id: int, params, dt: float, name: str):
        self.hop_sys = int(sys_id)
        p = np.zeros(NPARAMS, dtype=np.float64)
        p[: len(params)] = np.asarray(params, dtype=np.float64)
        self.hop_params = p
        self.dt = float(dt)
        self.name = name
        self.n, self.m = SYS_DIMS[self.hop_sys]

    def __call__(self, x, u):
        x = np.asarray(x, dtype=float).reshape(-1)
        u = np.asarray(u, dtype=float).reshape(-1)
        return _HOST_STEP[self.hop_sys](self.hop_params, x, u)

    def __repr__(self):
        return f"Dynamics({self.name}, dt={self.dt})"


# --- host-side evaluation (user convenience; the solver path runs the device twins) ------------
def _step_double_integrator(p, x, u):
    dt = p[0]
    return np.array([x[0] + dt * x[1], x[1] + dt * u[0]], dtype=float)


def _step_cartpole(p, x, u):
    dt, g, m_pole, length, total_mass, pml = (float(v) for v in p[:6])
    pos, vel, th, om = (float(v) for v in x)
    force = float(u[0])
    a = th - math.pi                      # stored angle 0 = down; internal model 0 = up
    ca, sa = math.cos(a), math.sin(a)
    tmp = (force + pml * om * om * sa) / total_mass
    den = length * (4.0 / 3.0 - m_pole * ca * ca / total_mass)
    th_acc = (g * sa - ca * tmp) / den
    x_acc = tmp - pml * th_acc * ca / total_mass
    return np.array([pos + dt * vel, vel + dt * x_acc, _wrap_pi(th + dt * om), om + dt * th_acc], dtype=float)


def _step_segway(p, x, u):
    dt, a_tau, a_th, b_tau, b_th = (float(v) for v in p[:5])
    pos, vel, th, om = (float(v) for v in x)
    tau = float(u[0])
    acc = a_tau * tau + a_th * th
    alp = b_tau * tau + b_th * th
    return np.array([pos + dt * vel, vel + dt * acc, _wrap_pi(th + dt * om), om + dt * alp], dtype=float)


def _step_quadrotor(p, x, u):
    if x.size != 12 or u.size != 4:
        raise ValueError("Quadrotor expects x in R^12 and u in R^4")
    dt, mass, g, Ix, Iy, Iz, iIx, iIy, iIz, kv, kw, cmin, wmax, nmax = (float(v) for v in p[:14])
    nan = np.full(12, np.nan)
    if not (np.all(np.isfinite(x)) and np.all(np.isfinite(u))):
        return nan
    if math.sqrt(float(x @ x)) > nmax:
        return nan
    phi, th, psi = (float(v) for v in x[6:9])
    wp, wq, wr = (float(v) for v in x[9:12])
    cth = math.cos(th)
    if abs(cth) < cmin or max(abs(wp), abs(wq), abs(wr)) > wmax:
        return nan
    sph, cph, sth = math.sin(phi), math.cos(phi), math.sin(th)
    sps, cps, tth = math.sin(psi), math.cos(psi), math.tan(th)
    sec = 1.0 / cth
    thrust = float(u[0])
    # body z-axis in the world frame (third column of Rz Ry Rx)
    bz = ((-sps) * (-sph) + (cps * sth) * cph, cps * (-sph) + (sps * sth) * cph, cth * cph)
    xd = np.empty(12)
    xd[0:3] = x[3:6]
    xd[3] = bz[0] * thrust / mass - 0.0 - kv * x[3]
    xd[4] = bz[1] * thrust / mass - 0.0 - kv * x[4]
    xd[5] = bz[2] * thrust / mass - g - kv * x[5]
    xd[6] = wp + (sph * tth) * wq + (cph * tth) * wr
    xd[7] = cph * wq + (-sph) * wr
    xd[8] = (sph * sec) * wq + (cph * sec) * wr
    h0, h1, h2 = Ix * wp, Iy * wq, Iz * wr
    xd[9] = iIx * (u[1] - (wq * h2 - wr * h1)) - kw * wp
    xd[10] = iIy * (u[2] - (wr * h0 - wp * h2)) - kw * wq
    xd[11] = iIz * (u[3] - (wp * h1 - wq * h0)) - kw * wr
    return x + dt * xd


_HOST_STEP = {
    SYS_DOUBLE_INTEGRATOR: _step_double_integrator,
    SYS_CARTPOLE: _step_cartpole,
    SYS_QUADROTOR: _step_quadrotor,
    SYS_SEGWAY: _step_segway,
}


# --- factories (same 13-tuples as systems.py) --------------------------------------------------
def make_double_integrator(dt: float = 0.05, N: int = 120):
    """systems.py:28-50: x=[pos, vel], u=[acc]."""
    F = Dynamics(SYS_DOUBLE_INTEGRATOR, [dt], dt, "DoubleIntegrator")
    return (F, np.array([1.0, 0.0]), np.array([2.0, 0.0]), np.array([0.0]),
            np.diag([1.0, 0.1]), np.array([[1e-2]]), 50.0, 0.02, N, 10, 80, [], None)


def make_cartpole_swingup(dt: float = 0.02, N: int = 360):
    """systems.py:57-112: x=[cart_pos, cart_vel, theta (0=down), theta_dot], u=[force]."""
    g, m_cart, m_pole, length = 9.81, 1.0, 0.1, 0.5
    F = Dynamics(SYS_CARTPOLE, [dt, g, m_pole, length, m_cart + m_pole, m_pole * length], dt, "Cartpole_SwingUp")
    return (F, np.zeros(4), np.array([0.0, 0.0, math.pi, 0.0]), np.array([0.0]),
            np.diag([0.01, 0.2, 0.0, 0.2]), np.array([[0.02]]), np.diag([5.0, 5.0, 800.0, 40.0]), 0.03,
            N, 40, 320, [2], None)


def make_quadrotor(dt: float = 0.05, N: int = 160):
    """systems.py:119-230: 12-D Euler-angle quadrotor, u=[thrust, tau_x, tau_y, tau_z]."""
    mass, g = 1.0, 9.81
    Ix, Iy, Iz = 0.02, 0.02, 0.04
    kv, kw = 0.05, 0.01
    F = Dynamics(SYS_QUADROTOR, [dt, mass, g, Ix, Iy, Iz, 1.0 / Ix, 1.0 / Iy, 1.0 / Iz, kv, kw, 1e-3, 1e3, 1e6],
                 dt, "Quadrotor")
    x0 = np.zeros(12)
    x0[0:3] = 2.0
    return (F, x0, np.zeros(12), np.array([mass * g, 0.0, 0.0, 0.0]),
            np.diag([5.0, 5, 5, 1, 1, 1, 20, 20, 10, 1, 1, 1]), np.diag([1e-3, 1e-2, 1e-2, 1e-2]), 300.0, 0.005,
            N, 40, 160, [6, 7, 8], None)


def make_segway_balance(dt: float = 0.02, N: int = 240):
    """systems.py:303-349: wheeled inverted pendulum, linearised plant, theta wrapped."""
    g, r, M, m, l = 9.81, 0.15, 1.0, 2.0, 0.5
    I = (1.0 / 3.0) * m * l * l
    a1, a2, a3 = M + m, m * l, I + m * l * l
    den = a1 * a3 - a2 * a2
    a_tau = a3 / (r * den) - a2 / den
    a_th = -(a2 * m * g * l) / den
    b_tau = -a2 / (r * den) + a1 / den
    b_th = (a1 * m * g * l) / den
    F = Dynamics(SYS_SEGWAY, [dt, a_tau, a_th, b_tau, b_th], dt, "Segway_Balance")
    return (F, np.array([0.05, 0.0, 0.08, 0.0]), np.zeros(4), np.array([0.0]),
            np.diag([1.0, 0.1, 25.0, 1.0]), np.array([[0.25]]), np.diag([20.0, 2.0, 250.0, 10.0]), 1e-4,
            N, 40, 200, [2], None)


def make_pointmass_navigation(*_a, **_k):
    raise NotImplementedError("pointmass navigation needs a Python extra_stage_cost callback; not on the B200 path "
                              "(it is commented out of the reference CASES, run_suite.py:43)")


MAKERS = {
    "DoubleIntegrator": make_double_integrator,
    "Cartpole_SwingUp": make_cartpole_swingup,
    "Quadrotor": make_quadrotor,
    "Quadrotor_Hover": make_quadrotor,   # legacy name (ilqr_propagator.py:766)
    "Segway_Balance": make_segway_balance,
}


def make_case(name: str, N: int | None = None, **kw):
    """Case by name; ``N`` overrides the nominal length and clips ``T_max`` to it."""
    maker = MAKERS[name]
    tup = list(maker(N=N, **kw) if N is not None else maker(**kw))
    tup[10] = min(int(tup[10]), int(tup[8]))
    return tuple(tup)
