"""BASELINE.json configs[4] on the BinaryCompactObject domain (44 blocks: wedges around the
two excised objects, cube wedges, ten bulged frustums, ten outer (half-)wedges; non-aligned
neighbours everywhere, three external spheres): GPU through the C-ABI vs the oracle with
orient_variables_on_slice, DirichletAnalytic ghost states and, with per-group refinement,
2:1 mortars between the block groups."""
import numpy as np
import pytest

from oracle import oracle as orc
from spectre_b200 import bco, evolution, lib

pytestmark = pytest.mark.gpu

TOL = 1e-12
GH_BLOCKS = [slice(0, 10), slice(10, 20), slice(20, 50)]


def _relerr(a, b, blocks):
    return max(np.max(np.abs(a[:, s] - b[:, s])) / np.max(np.abs(b[:, s])) for s in blocks)


def _oracle_inputs(problem, ev, u0, x, J, N):
    H = np.zeros((len(x), 4, N ** 3))
    dH = np.zeros((len(x), 16, N ** 3))
    for e in range(len(x)):
        H[e], dH[e] = orc.analytic_christoffel_gauge(N, u0[e], J[e])
    return H, dH


@pytest.mark.parametrize("N,refinement", [(4, 0), (6, 0), (5, "groups")])
def test_gh_rhs_and_steps_on_binary_domain_match_oracle(N, refinement):
    if refinement == "groups":
        # the cubes one level finer in the angular directions (Inspiral.yaml:95-101 pattern)
        refinement = {g: (0, 0, 0) for g in bco.BinaryCompactObject.GROUPS}
        refinement["ObjectACube"] = refinement["ObjectBCube"] = (1, 1, 0)
    problem = evolution.gh_binary_problem(refinement, N)
    ev = evolution.Evolution(problem, lib.STEPPER_ADAMS_BASHFORTH, 3, 1e-3)
    ctx, part = ev.ctx, ev.part
    assert part.oriented and len(part.external_faces) == 22
    ids = part.global_ids
    x, J, stat = problem.coords(ids), problem.inverse_jacobian(ids), problem.static(ids)
    u0 = problem.u0(ids, 0.0)
    rng = np.random.default_rng(N)
    u = u0 + 1e-3 * rng.uniform(-1, 1, u0.shape)
    ctx.set_state(u)
    ctx.compute_time_derivative(0.0)
    got = ctx.get_time_derivative()
    H, dH = _oracle_inputs(problem, ev, u0, x, J, N)
    ext = ev.boundary_ghost_data(problem, 0.0)[:, :50]
    full = np.concatenate([stat, H, dH], axis=1)
    kw = dict(gauge_params=orc.GAUGE_GIVEN, ext_u=ext, nbr_dir=part.local_neighbor_direction,
              face_perm=part.local_face_permutation)
    if len(problem.mortars):
        kw["mortars"] = part.local_mortars
    ref = orc.dg_rhs(1, N, u, J, full, part.local_neighbors, **kw)
    assert _relerr(got, ref, GH_BLOCKS) < TOL
    # two AB3 steps (with the self-start) from the perturbed state
    ctx.set_stepper(lib.STEPPER_ADAMS_BASHFORTH, 3, 0.0, 1e-3)
    ev.take_steps(2)
    o = orc.Evolution(lambda v, t: orc.dg_rhs(1, N, v, J, full, part.local_neighbors, **kw), u,
                      0.0, 1e-3, "AB3")
    o.step()
    o.step()
    assert _relerr(ctx.get_state(), o.u, GH_BLOCKS) < TOL
    ctx.close()
