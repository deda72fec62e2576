"""Exact time bookkeeping: the library in slab mode (dgrhs_set_slab)
forms step and substep times the way the reference's Slab / Time / TimeStepId do
(Slab.hpp advance, Time.cpp:114-117, TimeStepId.cpp:72-82), bit for bit; the C++ value
types of host/SpectreTime.hpp drive it through DgTimeLoop."""
import os
import subprocess

import numpy as np
import pytest

from oracle import oracle as orc
from spectre_b200 import analytic, domain, lib

pytestmark = pytest.mark.gpu

# a slab whose ends trigger rounding errors (the values of the reference's Test_Time.cpp)
SLAB = (0.68138945475734402635, 0.68138945475734402635 + 3e-3)


@pytest.mark.parametrize("name,stepper,order,per_slab", [
    ("AB3", lib.STEPPER_ADAMS_BASHFORTH, 3, 3), ("AB4", lib.STEPPER_ADAMS_BASHFORTH, 4, 1),
    ("RK3", lib.STEPPER_RK3_HESTHAVEN, 3, 2), ("Rk3Owren", lib.STEPPER_RK3_OWREN, 0, 3),
    ("DP5", lib.STEPPER_DORMAND_PRINCE5, 0, 2)])
def test_slab_mode_times_bit_exact(name, stepper, order, per_slab):
    N, steps = 4, 7
    brick = domain.Brick([0, 0, 0], [2 * np.pi] * 3, [1, 1, 1], N)
    x, J, nb = brick.coords(), brick.inverse_jacobian(), brick.neighbors()
    u0 = analytic.plane_wave(x, 0.0)
    stat = np.zeros((brick.n_elements, 1, brick.n))
    ctx = lib.Context(lib.SYSTEM_SCALAR_WAVE, N, brick.n_elements)
    ctx.set_geometry(J, x, nb)
    ctx.set_static_fields(stat)
    ctx.set_state(u0)
    ctx.set_stepper(stepper, order, 123.0, 4.5e-3)   # t0 and dt are replaced by the slab's
    ctx.set_slab(SLAB[0], SLAB[1], per_slab)
    gpu_times, done_steps = [], 0
    while done_steps < steps:
        t = ctx.begin_substep()
        gpu_times.append(t)
        ctx.compute_time_derivative(t)
        done_steps += ctx.end_substep()
    cpu_times = []

    def rhs(u, t):
        cpu_times.append(t)
        return orc.dg_rhs(0, N, u, J, stat, nb)

    ev = orc.Evolution(rhs, u0, None, None, name, slab=(SLAB[0], SLAB[1], per_slab))
    for _ in range(steps):
        ev.step()
    assert gpu_times == cpu_times          # every RHS time, self-start included, bit for bit
    assert ctx.time == ev.time
    if per_slab == 3:   # these times are not those of t0 + k dt
        assert ctx.time != SLAB[0] + steps * ((SLAB[1] - SLAB[0]) / per_slab)
    got = ctx.get_state()
    assert np.max(np.abs(got - ev.u)) / np.max(np.abs(ev.u)) < 1e-12
    ctx.close()


def test_slab_misuse():
    ctx = lib.Context(lib.SYSTEM_SCALAR_WAVE, 3, 1)
    with pytest.raises(lib.DgrhsError, match="set_stepper has not been called"):
        ctx.set_slab(0.0, 1.0, 1)
    ctx.set_stepper(lib.STEPPER_RK3_HESTHAVEN, 3, 0.0, 1e-3)
    with pytest.raises(lib.DgrhsError, match="bad slab"):
        ctx.set_slab(1.0, 0.0, 1)
    with pytest.raises(lib.DgrhsError, match="bad slab"):
        ctx.set_slab(0.0, 1.0, 0)
    ctx.close()


def test_cpp_time_loop_with_time_step_ids():
    """DgTimeLoop (TimeStepId / next_time_id / Slab in C++) next to dgrhs_take_steps:
    identical substep times (checked inside the loop) and bit-identical states."""
    from tests.test_abi_cpu import _build_time_types_test
    out = subprocess.run([_build_time_types_test(), "gpu"], capture_output=True, text=True)
    assert out.returncode == 0 and "all checks passed" in out.stdout, out.stdout + out.stderr
