# -*- coding: utf-8 -*-
"""Drop-in for the reference's systems.py: the same make_* factories and 13-tuples, with `F` a
device-registered dynamics object (still callable as F(x, u), still exposing F.dt)."""
from _bridge import cases

make_double_integrator = cases.make_double_integrator
make_cartpole_swingup = cases.make_cartpole_swingup
make_quadrotor = cases.make_quadrotor
make_segway_balance = cases.make_segway_balance
make_pointmass_navigation = cases.make_pointmass_navigation
