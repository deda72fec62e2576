"""Local time stepping, CPU side: the oracle's Adams LTS coefficients against the known
answers of the reference's tests/Unit/Time/TimeSteppers/Test_AdamsLts.cpp:416-760 (explicit
schemes), and properties of the oracle's LTS evolution (equal steps = the GTS oracle,
conservation of the coupling, convergence with mixed steps)."""
from fractions import Fraction as Fr

import numpy as np
import pytest

from oracle import lts
from oracle import oracle as orc


def _check(got, expected):
    assert set(got) == set((Fr(a), Fr(b)) for a, b in expected), (got, expected)
    for (a, b), v in expected.items():
        assert got[(Fr(a), Fr(b))] == pytest.approx(v, rel=1e-13, abs=1e-14)


AB3 = [5.0 / 12.0, -4.0 / 3.0, 23.0 / 12.0]


def test_gts_orders_1_and_3():
    # Test_AdamsLts.cpp:431-467
    _check(lts.lts_coefficients([0], [0], 0, 1, 1), {(0, 0): 1.0})
    _check(lts.lts_coefficients([0, 1, 2], [0, 1, 2], 2, 3, 3),
           {(0, 0): AB3[0], (1, 1): AB3[1], (2, 2): AB3[2]})


def test_single_side_order_3():
    # Test_AdamsLts.cpp:505-527: the remote side is first order (one value)
    _check(lts.lts_coefficients([0, 1, 2], [0], 2, 3, 3, 1, 3),
           {(0, 0): AB3[0], (1, 0): AB3[1], (2, 0): AB3[2]})


def test_two_to_one_order_3():
    # Test_AdamsLts.cpp:556-596
    large, small = [-8, -4, 0], [-4, -2, 0, 2]
    _check(lts.lts_coefficients(large, small, 0, 4, 3), {
        (0, 2): 115.0 / 16.0, (0, 0): 7.0 / 6.0, (0, -2): -11.0 / 16.0, (-4, 2): -115.0 / 24.0,
        (-4, -2): -11.0 / 8.0, (-4, -4): 5.0 / 6.0, (-8, 2): 23.0 / 16.0, (-8, -2): 11.0 / 48.0})
    _check(lts.lts_coefficients(small, large, 0, 2, 3), {
        (0, 0): 23.0 / 6.0, (-2, 0): -1.0, (-2, -4): -2.0, (-4, -4): 5.0 / 6.0,
        (-2, -8): 1.0 / 3.0})
    _check(lts.lts_coefficients(small, large, 2, 4, 3), {
        (2, 0): 115.0 / 16.0, (0, 0): -8.0 / 3.0, (-2, 0): 5.0 / 16.0, (2, -4): -115.0 / 24.0,
        (-2, -4): 5.0 / 8.0, (2, -8): 23.0 / 16.0, (-2, -8): -5.0 / 48.0})


def test_lts_to_gts_order_2():
    # Test_AdamsLts.cpp:598-613
    _check(lts.lts_coefficients([-2, 0], [-1, 0], 0, 1, 2),
           {(0, 0): 3.0 / 2.0, (0, -1): -1.0 / 4.0, (-2, -1): -1.0 / 4.0})


def test_three_to_one_order_2():
    # Test_AdamsLts.cpp:615-657
    large, small = [-3, 0], [-1, 0, 1, 2]
    _check(lts.lts_coefficients(large, small, 0, 3, 2), {
        (-3, -1): -1.0 / 6.0, (-3, 1): -1.0 / 3.0, (-3, 2): -1.0, (0, -1): -1.0 / 3.0,
        (0, 0): 1.0, (0, 1): 4.0 / 3.0, (0, 2): 5.0 / 2.0})
    _check(lts.lts_coefficients(small, large, 0, 1, 2),
           {(-1, -3): -1.0 / 6.0, (-1, 0): -1.0 / 3.0, (0, 0): 3.0 / 2.0})
    _check(lts.lts_coefficients(small, large, 1, 2, 2),
           {(0, 0): -1.0 / 2.0, (1, -3): -1.0 / 2.0, (1, 0): 2.0})
    _check(lts.lts_coefficients(small, large, 2, 3, 2),
           {(1, -3): 1.0 / 6.0, (1, 0): -2.0 / 3.0, (2, -3): -1.0, (2, 0): 5.0 / 2.0})


def test_unaligned_order_2():
    # Test_AdamsLts.cpp:685-725
    a, b = [1, 3, 4], [2, 3, 5]
    _check(lts.lts_coefficients(a, b, 3, 4, 2),
           {(1, 2): -1.0 / 4.0, (3, 2): -1.0 / 4.0, (3, 3): 3.0 / 2.0})
    _check(lts.lts_coefficients(a, b, 4, 6, 2),
           {(3, 3): -1.0 / 2.0, (3, 5): -3.0 / 2.0, (4, 2): -3.0 / 2.0, (4, 3): 11.0 / 4.0,
            (4, 5): 11.0 / 4.0})
    _check(lts.lts_coefficients(b, a, 3, 5, 2),
           {(2, 1): -1.0 / 4.0, (2, 3): -1.0 / 4.0, (2, 4): -3.0 / 2.0, (3, 3): 1.0, (3, 4): 3.0})
    _check(lts.lts_coefficients(b, a, 5, 6, 2),
           {(3, 4): -1.0 / 4.0, (5, 3): -3.0 / 2.0, (5, 4): 11.0 / 4.0})


def test_self_start_history_order_4():
    # Test_AdamsLts.cpp:778-795: the history after self-start is not sorted in time (the
    # values at 2 and 3 come from the self-start slab and precede those at 0 and 1)
    steps = [2, 3, 0, 1]
    _check(lts.lts_coefficients(steps, steps, 1, 2, 4),
           {(2, 2): 13.0 / 24.0, (3, 3): -1.0 / 24.0, (0, 0): -1.0 / 24.0, (1, 1): 13.0 / 24.0})


@pytest.mark.parametrize("order", [2, 3, 4])
def test_coefficients_are_conservative_and_consistent(order):
    """What the reference's LTS design rests on (AdamsLts.hpp:90-125): over a common
    interval the two sides integrate the same coupling terms, and for a coupling that does
    not depend on its arguments the coefficients sum to the step."""
    k = order
    coarse = [4 * i for i in range(-(k - 1), 1)]
    fine_all = [2 * i for i in range(-2 * (k - 1), 2)]
    big = lts.lts_coefficients(coarse, fine_all, 0, 4, k, exact=True)
    s1 = lts.lts_coefficients([t for t in fine_all if t <= 0], coarse, 0, 2, k, exact=True)
    s2 = lts.lts_coefficients(fine_all, coarse, 2, 4, k, exact=True)
    assert sum(big.values()) == 4
    assert sum(s1.values()) == 2 and sum(s2.values()) == 2
    both = {}
    for part in (s1, s2):
        for (a, b), v in part.items():
            both[(b, a)] = both.get((b, a), 0) + v
    both = {key: v for key, v in both.items() if v != 0}
    assert both == {key: v for key, v in big.items() if v != 0}


def _sw_problem(N, levels_of):
    from spectre_b200 import analytic, domain
    brick = domain.Brick([0, 0, 0], [2 * np.pi] * 3, [1, 1, 1], N)
    x, J, nb = brick.coords(), brick.inverse_jacobian(), brick.neighbors()
    stat = np.zeros((brick.n_elements, 1, brick.n))
    levels = np.array([levels_of(x[e].mean(axis=1)) for e in range(brick.n_elements)])
    return analytic, x, J, nb, stat, levels


def test_equal_levels_reproduce_the_gts_oracle():
    """all elements on one level: every boundary is `lts_coefficients_for_gts`
    (AdamsLts.cpp:307-327) and the evolution equals the GTS Adams-Bashforth one up to the
    order of the additions (volume and boundary parts are added separately)"""
    N, dt, order = 4, 2e-3, 3
    analytic, x, J, nb, stat, levels = _sw_problem(N, lambda c: 0)
    u0 = analytic.plane_wave(x, 0.0)
    ev = lts.LtsEvolution(0, N, J, stat, nb, levels, order, 0.0, dt, u0,
                          lambda j: analytic.plane_wave(x, -j * dt))
    ev.take_coarse_steps(4)
    # the same history start for the GTS evolution
    hist = [(orc.dg_rhs(0, N, analytic.plane_wave(x, -j * dt), J, stat, nb)) for j in (2, 1)]
    u = u0.copy()
    for _ in range(4):
        hist.append(orc.dg_rhs(0, N, u, J, stat, nb))
        c = orc._AB_CONST[3]
        u = u + dt * (c[0] * hist[-3] + c[1] * hist[-2] + c[2] * hist[-1])
    assert np.max(np.abs(ev.u - u)) < 1e-13 * np.max(np.abs(u))


def test_mixed_levels_converge_to_the_analytic_solution():
    """half of the brick takes two (four) steps per coarse step; the error of the LTS
    evolution is that of the spatial discretisation, as for the GTS evolution with the fine
    step (TimeStepperTestUtils check_convergence_order's role for the coupled system)"""
    N, dt, order = 6, 8e-3, 3
    analytic, x, J, nb, stat, levels = _sw_problem(N, lambda c: int(c[0] > np.pi) + int(c[1] > np.pi))
    assert sorted(set(levels)) == [0, 1, 2]
    u0 = analytic.plane_wave(x, 0.0)
    stride = 2 ** (levels.max() - levels)
    tick = dt / 4

    def past(j):
        return np.stack([analytic.plane_wave(x[e], -j * stride[e] * tick)
                         for e in range(len(levels))])
    ev = lts.LtsEvolution(0, N, J, stat, nb, levels, order, 0.0, dt, u0, past)
    ev.take_coarse_steps(5)
    exact = analytic.plane_wave(x, 5 * dt)
    err = np.max(np.abs(ev.u - exact))
    # GTS with the fine step everywhere
    ev_f = lts.LtsEvolution(0, N, J, stat, nb, 0 * levels, order, 0.0, tick, u0,
                            lambda j: analytic.plane_wave(x, -j * tick))
    ev_f.take_coarse_steps(20)
    err_f = np.max(np.abs(ev_f.u - exact))
    assert err < 3 * err_f + 1e-6 and err < 2e-3
    # the coupling is evaluated more than once per face and step only at LTS boundaries
    assert ev.corrections_evaluated > 0


def test_library_lts_coefficients_match_the_oracle_and_the_reference():
    """dgrhs_adams_lts_coefficients (host code of libdgrhs.so, doubles) against the oracle's
    exact rationals on the reference's cases and on 2:1 / 4:1 / 8:1 steady-state patterns of
    orders 1..8."""
    from spectre_b200 import lib
    cases = [([0, 1, 2], [0, 1, 2], 2, 3, 3, 3, 3), ([0, 1, 2], [0], 2, 3, 3, 1, 3),
             ([-8, -4, 0], [-4, -2, 0, 2], 0, 4, 3, 3, 3),
             ([-4, -2, 0, 2], [-8, -4, 0], 2, 4, 3, 3, 3), ([-2, 0], [-1, 0], 0, 1, 2, 2, 2),
             ([-3, 0], [-1, 0, 1, 2], 0, 3, 2, 2, 2), ([1, 3, 4], [2, 3, 5], 3, 4, 2, 2, 2),
             ([1, 3, 4], [2, 3, 5], 4, 6, 2, 2, 2), ([2, 3, 5], [1, 3, 4], 3, 5, 2, 2, 2),
             ([2, 3, 5], [1, 3, 4], 5, 6, 2, 2, 2), ([2, 3, 0, 1], [2, 3, 0, 1], 1, 2, 4, 4, 4)]
    for k in range(1, 9):
        for r in (2, 4, 8):
            coarse = [r * i for i in range(-(k - 1), 1)]
            fine = list(range(-(k - 1), r))
            cases.append((coarse, fine, 0, r, k, k, k))
            cases.append((fine[:k + 1], coarse, 1, 2, k, k, k))
    for local, remote, start, end, lo, ro, so in cases:
        ref = lts.lts_coefficients(local, remote, start, end, lo, ro, so)
        got = lib.adams_lts_coefficients(local, remote, start, end, lo, ro, so, origin=0.3,
                                         tick_size=1.0)
        assert set(got) == set((int(a), int(b)) for a, b in ref)
        scale = max(abs(v) for v in ref.values())
        for (a, b), v in ref.items():
            assert abs(got[(int(a), int(b))] - v) < 1e-9 * scale, (local, remote, a, b)
    # the terms come out in the reference's sorted order and scale with the tick size
    got = lib.adams_lts_coefficients([-8, -4, 0], [-4, -2, 0, 2], 0, 4, 3, tick_size=0.25)
    assert list(got) == sorted(got)
    assert got[(0, 2)] == pytest.approx(0.25 * 115.0 / 16.0, rel=1e-13)


def test_element_size_cfl_chooser_and_levels():
    """StepChoosers::ElementSizeCfl evaluated at the start (ElementSizeCfl.hpp:76-92,
    SizeOfElement.cpp:44-59, Characteristics.cpp:188-196) and the levels it gives on a thick
    Kerr-Schild shell: the expected value of Test_ElementSizeCfl.cpp:73-95 is the formula
    itself, safety * stable_step * min size / (speed * Dim)."""
    from spectre_b200 import analytic, domain, lib
    from spectre_b200 import lts as hlts
    # size_of_element: affine brick and the exact wedge map
    brick = domain.Brick([0, 0, 0], [2.0, 4.0, 8.0], [1, 1, 2], 4)
    np.testing.assert_allclose(hlts.size_of_element(brick), [[1.0, 2.0, 2.0]] * brick.n_elements,
                               rtol=1e-14)
    shell = domain.SphericalShell(2.0, 32.0, (0, 2), 5, radial_distribution="Logarithmic",
                                  order="radial")
    size = hlts.size_of_element(shell)
    # radial face centres lie on the wedge axis: four layers 2-4-8-16-32
    np.testing.assert_allclose(sorted(set(np.round(size[:, 2], 10))), [2.0, 4.0, 8.0, 16.0])
    # the formula
    goal = hlts.element_size_cfl(np.array([[1.0, 2.0, 0.5]]), [2.0], 3.0 / 11.0, 0.8)
    assert goal[0] == pytest.approx(0.8 * (3.0 / 11.0) * 0.5 / (2.0 * 3.0), rel=1e-15)
    # largest characteristic speed of Kerr-Schild (M = 1): lapse^2 = 1/(1+2/r), |beta| =
    # (2/r) lapse -> |beta| + lapse = sqrt(1 + 2/r) at the innermost point, gamma1 = -1
    x = shell.coords()
    u = analytic.kerr_schild(x, 1.0)
    speed = hlts.gh_largest_characteristic_speed(u, -np.ones((shell.n_elements, shell.n)))
    r = np.sqrt((x ** 2).sum(axis=1))
    lapse = 1.0 / np.sqrt(1.0 + 2.0 / r)
    np.testing.assert_allclose(speed, ((2.0 / r) * lapse + lapse).max(axis=1), rtol=1e-12)
    stable = lib.stepper_properties(lib.STEPPER_ADAMS_BASHFORTH, 3)[3]
    goal = hlts.element_size_cfl(size, speed, stable, 0.5)
    levels = hlts.levels_from_step_limit(goal, dt_coarse=goal.max())
    assert levels.min() == 0 and len(set(levels.tolist())) == 4   # one level per radial layer
    assert np.all(np.diff(levels) <= 0)    # inside-out element order: finer steps inside
    assert np.all(goal.max() / 2.0 ** levels <= goal * (1 + 1e-12))


def _refined_sw_problem(N):
    from spectre_b200 import analytic, domain
    rb = domain.RefinedBrick([0, 0, 0], [2 * np.pi] * 3, [1, 1, 1], N, [(0, 0, 0)])
    x, J, nb, mt = rb.coords(), rb.inverse_jacobian(), rb.neighbors(), rb.mortars()
    stat = np.zeros((rb.n_elements, 1, N ** 3))
    size = x[:, 0].max(axis=1) - x[:, 0].min(axis=1)
    return analytic, rb, x, J, nb, mt, stat, size


def test_lts_with_mortars_equal_levels_reproduce_the_gts_oracle():
    """h-refined brick (one cell split into eight), every element on one level: the mortar
    couplings through the boundary histories sum to the GTS right-hand side with mortars"""
    N, dt, order = 4, 1e-3, 3
    analytic, rb, x, J, nb, mt, stat, _ = _refined_sw_problem(N)
    assert len(mt) > 0
    u0 = analytic.plane_wave(x, 0.0)
    ev = lts.LtsEvolution(0, N, J, stat, nb, np.zeros(rb.n_elements, int), order, 0.0, dt, u0,
                          lambda j: analytic.plane_wave(x, -j * dt), mortars=mt)
    ev.take_coarse_steps(3)
    hist = [orc.dg_rhs(0, N, analytic.plane_wave(x, -j * dt), J, stat, nb, mortars=mt)
            for j in (2, 1)]
    u = u0.copy()
    c = orc._AB_CONST[3]
    for _ in range(3):
        hist.append(orc.dg_rhs(0, N, u, J, stat, nb, mortars=mt))
        u = u + dt * (c[0] * hist[-3] + c[1] * hist[-2] + c[2] * hist[-1])
    assert np.max(np.abs(ev.u - u)) < 1e-13 * np.max(np.abs(u))


def test_lts_with_mortars_fine_elements_take_half_steps():
    """the canonical LTS set-up: the eight children of the refined cell take two steps per
    step of the unrefined elements; the solution stays as close to the analytic one as the
    GTS evolution with the fine step everywhere"""
    N, dt, order = 6, 8e-3, 3
    analytic, rb, x, J, nb, mt, stat, size = _refined_sw_problem(N)
    levels = (size < 0.75 * size.max()).astype(int)
    assert levels.sum() == 8
    from spectre_b200 import lts as hlts
    perm, nbp = hlts.order_by_level(levels, nb)
    inv = np.empty_like(perm)
    inv[perm] = np.arange(len(perm))
    mtp = np.array(mt).copy()
    mtp[:, 0], mtp[:, 2] = inv[mtp[:, 0]], inv[mtp[:, 2]]
    x, J, levels = x[perm], J[perm], levels[perm]
    u0 = analytic.plane_wave(x, 0.0)

    def past(j):
        return np.stack([analytic.plane_wave(x[e], -j * dt / 2 ** levels[e])
                         for e in range(len(levels))])
    ev = lts.LtsEvolution(0, N, J, stat, nbp, levels, order, 0.0, dt, u0, past, mortars=mtp)
    ev.take_coarse_steps(4)
    exact = analytic.plane_wave(x, 4 * dt)
    err = np.max(np.abs(ev.u - exact))
    ev_f = lts.LtsEvolution(0, N, J, stat, nbp, 0 * levels, order, 0.0, dt / 2, u0,
                            lambda j: analytic.plane_wave(x, -j * dt / 2), mortars=mtp)
    ev_f.take_coarse_steps(8)
    err_f = np.max(np.abs(ev_f.u - exact))
    assert err < 3 * err_f + 1e-6 and err < 5e-3


def test_partition_reorder_keeps_the_connectivity():
    """Partition.reorder (elements sorted by step-size level): neighbours, orientations,
    mortars and boundary slots name the same faces as before"""
    from spectre_b200 import domain
    rb = domain.RefinedBrick([0, 0, 0], [1, 1, 1], [1, 1, 1], 3, [(0, 0, 0)],
                             periodic=(True, True, False))
    part0 = domain.Partition(rb.neighbors(), 1, 0, boundary_slots=True, mortars=rb.mortars())
    part = domain.Partition(rb.neighbors(), 1, 0, boundary_slots=True, mortars=rb.mortars())
    rng = np.random.default_rng(0)
    perm = rng.permutation(part.n_local)
    part.reorder(perm)
    old_of = perm                                  # new -> old
    np.testing.assert_array_equal(part.global_ids, part0.global_ids[perm])
    for new in range(part.n_local):
        for d in range(6):
            a, b = part.local_neighbors[new, d], part0.local_neighbors[old_of[new], d]
            assert (a < 0 and a == b) or (a >= 0 and old_of[a] == b)
    for row, row0 in zip(part.local_mortars, part0.local_mortars):
        assert (old_of[row[0]], row[1], old_of[row[2]], row[3], row[4], row[5]) == tuple(row0)
    assert sorted((int(old_of[le]), d, s) for le, d, s in part.external_faces) == \
        sorted(part0.external_faces)
    with pytest.raises(ValueError):
        domain.Partition(rb.neighbors(), 2, 0, mortars=rb.mortars()).reorder(perm[:4])


def _check_ids(got, expected):
    norm = lambda k: (Fr(k[0]), Fr(k[1])) if isinstance(k, tuple) else Fr(k)
    want = {(norm(a), norm(b)): v for (a, b), v in expected.items()}
    assert set(got) == set(want), (got, want)
    for k, v in want.items():
        assert got[k] == pytest.approx(v, rel=1e-13, abs=1e-14)


def test_implicit_schemes_reference_known_answers():
    """Adams-Moulton (implicit) schemes of lts_coefficients, the predictor value of a step
    written as the pair (step time, step size): Test_AdamsLts.cpp:469-503 (AM GTS order 3),
    :529-554 (AM single-side order 3), :727-776 (AM 2:1 order 2)"""
    am = dict(local_implicit=True, remote_implicit=True, small_implicit=True)
    steps = [0, 1, (1, 1)]
    am3 = [-1.0 / 12.0, 2.0 / 3.0, 5.0 / 12.0]
    _check_ids(lts.lts_coefficients(steps, steps, 1, 2, 3, **am),
               {(0, 0): am3[0], (1, 1): am3[1], ((1, 1), (1, 1)): am3[2]})
    # the predictor of the same step is an explicit step of one order less
    _check_ids(lts.lts_coefficients(steps, steps, 1, 2, 2), {(0, 0): -0.5, (1, 1): 1.5})
    _check_ids(lts.lts_coefficients(steps, [0], 1, 2, 3, 1, 3, local_implicit=True,
                                    small_implicit=True),
               {(0, 0): am3[0], (1, 0): am3[1], ((1, 1), 0): am3[2]})
    _check_ids(lts.lts_coefficients(steps, [0], 1, 2, 2, 1, 2), {(0, 0): -0.5, (1, 0): 1.5})
    large, small = [0, (0, 2)], [0, (0, 1), 1, (1, 1)]
    _check_ids(lts.lts_coefficients(large, [0], 0, 2, 1), {(0, 0): 2.0})
    _check_ids(lts.lts_coefficients(large, small, 0, 2, 2, **am),
               {(0, 0): 0.5, (0, (0, 1)): 0.25, ((0, 2), (0, 1)): 0.25, (0, 1): 0.25,
                ((0, 2), 1): 0.25, ((0, 2), (1, 1)): 0.5})
    _check_ids(lts.lts_coefficients(small, large, 0, 1, 1), {(0, 0): 1.0})
    _check_ids(lts.lts_coefficients(small, large, 0, 1, 2, **am),
               {(0, 0): 0.5, ((0, 1), 0): 0.25, ((0, 1), (0, 2)): 0.25})
    _check_ids(lts.lts_coefficients(small, large, 1, 2, 1), {(1, 0): 1.0})
    _check_ids(lts.lts_coefficients(small, large, 1, 2, 2, **am),
               {(1, 0): 0.25, (1, (0, 2)): 0.25, ((1, 1), (0, 2)): 0.5})


def test_library_implicit_schemes_match_the_oracle_and_the_reference():
    """dgrhs_adams_lts_coefficients_general with Adams-Moulton schemes: the reference's cases
    (Test_AdamsLts.cpp:469-503, :529-554, :727-776) and 2:1 / 4:1 predictor-corrector
    patterns of orders 2..8 against the oracle's exact rationals"""
    from spectre_b200 import lib
    am = dict(local_implicit=True, remote_implicit=True, small_implicit=True)
    got = lib.adams_lts_coefficients([0, 1, (1, 1)], [0, 1, (1, 1)], 1, 2, 3, **am)
    assert got == pytest.approx({(0, 0): -1.0 / 12.0, (1, 1): 2.0 / 3.0,
                                 ((1, 1), (1, 1)): 5.0 / 12.0}, rel=1e-14)
    large, small = [0, (0, 2)], [0, (0, 1), 1, (1, 1)]
    got = lib.adams_lts_coefficients(large, small, 0, 2, 2, **am)
    assert got == pytest.approx({(0, 0): 0.5, (0, (0, 1)): 0.25, ((0, 2), (0, 1)): 0.25,
                                 (0, 1): 0.25, ((0, 2), 1): 0.25, ((0, 2), (1, 1)): 0.5},
                                rel=1e-14)
    assert list(got) == sorted(got, key=lambda k: tuple(
        (i, 0, 0) if not isinstance(i, tuple) else (i[0], 1, i[1]) for i in k))
    cases = []
    for k in range(2, 9):
        for r in (2, 4):
            # steady state: coarse steps of r ticks with the predictor value of the current
            # step, fine steps of one tick with theirs
            coarse = [r * i for i in range(-(k - 2), 1)] + [(0, r)]
            fine = []
            for t in range(-(k - 2), r):
                fine += [t, (t, 1)]
            cases.append((coarse, fine, 0, r, k))                       # the coarse corrector
            cases.append(([i for i in fine if (i[0] if isinstance(i, tuple) else i) <= 0],
                          coarse, 0, 1, k))                             # first fine corrector
    for local, remote, start, end, k in cases:
        ref = lts.lts_coefficients(local, remote, start, end, k, **am)
        got = lib.adams_lts_coefficients(local, remote, start, end, k, origin=-0.7,
                                         tick_size=0.5, **am)
        norm = lambda i: (Fr(i[0]), Fr(i[1])) if isinstance(i, tuple) else Fr(i)
        assert {(norm(a), norm(b)) for a, b in got} == set(ref)
        scale = max(abs(v) for v in ref.values())
        for (a, b), v in got.items():
            assert abs(v - 0.5 * ref[(norm(a), norm(b))]) < 1e-9 * scale, (local, remote, a, b)
    with pytest.raises(lib.DgrhsError, match="substep data"):
        lib.adams_lts_coefficients([0, 1], [0, 1], 1, 2, 3, **am)


def test_choose_lts_step_size_known_answers():
    """Test_ChooseLtsStepSize.cpp:12-24 (slab [1, 4], duration 3): the step is slab / 2^n"""
    from spectre_b200 import lts as hlts
    dur = 3.0
    assert hlts.choose_lts_step_power(4.0, dur) == 0
    assert hlts.choose_lts_step_power(10.0, dur) == 0
    assert hlts.choose_lts_step_power(2.0, dur) == 1
    assert hlts.choose_lts_step_power(1.4, dur) == 2
    assert hlts.choose_lts_step_power(2.0, dur, fraction_denominator=4) == 2   # at start + 1/4
    assert hlts.choose_lts_step_power(np.inf, dur) == 0
    # arrays, exact powers of two (log2(2^n + eps) must not round up twice)
    np.testing.assert_array_equal(
        hlts.choose_lts_step_power(np.array([1.0, 0.5, 0.25, 0.125, 0.1249]), 1.0),
        [0, 1, 2, 3, 4])
    np.testing.assert_array_equal(hlts.levels_from_step_limit([0.3, 0.06, 1.0], 0.25), [0, 3, 0])


def test_cfl_step_chooser_and_minimum_grid_spacing():
    """StepChoosers::Cfl (Cfl.hpp:69-79) on domain::minimum_grid_spacing
    (MinimumGridSpacing.cpp:31-71): on an affine brick the smallest spacing is the first LGL
    interval of the shortest side; on a wedge no pair of index neighbours is closer than the
    value found by brute force over all pairs of the inner face"""
    from spectre_b200 import domain, lib
    from spectre_b200 import lts as hlts
    N = 5
    brick = domain.Brick([0, 0, 0], [2.0, 4.0, 8.0], [1, 1, 1], N)
    xi, _ = lib.collocation_points_and_weights(N)
    first = 0.5 * (xi[1] - xi[0])        # first LGL interval of a unit-length side
    np.testing.assert_allclose(hlts.minimum_grid_spacing(brick.coords(), N), first * 1.0,
                               rtol=1e-13)
    shell = domain.SphericalShell(2.0, 6.0, (0, 0), N)
    x = shell.coords()
    got = hlts.minimum_grid_spacing(x, N)
    for e in range(shell.n_elements):
        pts = x[e].T
        d = np.sqrt(((pts[:, None, :] - pts[None, :, :]) ** 2).sum(axis=2))
        np.fill_diagonal(d, np.inf)
        assert got[e] == pytest.approx(d.min(), rel=1e-12)
    step = hlts.cfl_step(np.array([0.1, 0.2]), np.array([1.0, 2.0]), 3.0 / 11.0, 0.8)
    np.testing.assert_allclose(step, [0.8 * (3.0 / 11.0) * 0.1 / 3.0, 0.8 * (3.0 / 11.0) * 0.2 / 6.0],
                               rtol=1e-15)
    assert hlts.limit_increase(-0.25, 2.0) == 0.5
