"""YAML front end (spectre_b200/input_file.py): the option subset of the
reference's input files that the accelerated path understands.  The CPU tests
parse this repository's own fixtures (tests/inputs/, written in the reference's
option schema) and, when the reference checkout is present, the reference's
PlaneWave3D.yaml / GaugeWave3D.yaml / KerrSchild.yaml unchanged; the GPU tests
run them."""
import os

import numpy as np
import pytest

from spectre_b200 import domain, input_file, lib

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/tests/InputFiles"
# byte-for-byte copies of the reference's input files (tests/golden/gen_input_files.py): the
# GPU box has no /root/reference
GOLDEN = os.path.join(HERE, "golden", "inputs")


def _reference_input(rel):
    """The reference's own file when the checkout is present, else the committed copy."""
    path = os.path.join(REF, rel)
    return path if os.path.exists(path) else os.path.join(GOLDEN, os.path.basename(rel))


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present")
def test_committed_input_files_are_the_reference_files():
    for rel in ("ScalarWave/PlaneWave3D.yaml", "GeneralizedHarmonic/GaugeWave3D.yaml",
                "GeneralizedHarmonic/KerrSchild.yaml"):
        with open(os.path.join(REF, rel), "rb") as a, \
                open(os.path.join(GOLDEN, os.path.basename(rel)), "rb") as b:
            assert a.read() == b.read(), rel


def test_parse_own_fixtures():
    r = input_file.load(os.path.join(HERE, "inputs", "GaugeWaveBjorhus.yaml"))
    assert r.system == lib.SYSTEM_GH and r.stepper == lib.STEPPER_RK3_HESTHAVEN
    assert (r.t0, r.dt, r.n_steps, r.observe_interval) == (0.0, 5e-4, 4, 2)
    assert r.filter == (36.0, 64) and r.gauge == lib.GAUGE_HARMONIC
    assert r.face_bc == {0: "DirichletAnalytic", 1: "ConstraintPreservingPhysical"}
    assert isinstance(r.domain, domain.Brick) and r.domain.periodic == (False, True, True)
    assert r.static[0] == 1.0 and r.static[1] == -1.0 and callable(r.static[2])
    p = r.problem()
    assert p.boundary_time_dependent and p.bjorhus(0, 1) == "ConstraintPreservingPhysical"
    assert p.bjorhus(0, 0) is None and p.dirichlet_analytic(0, 0) and not p.dirichlet_analytic(0, 1)
    s = input_file.load(os.path.join(HERE, "inputs", "ScalarWaveRk4.yaml"))
    assert s.system == lib.SYSTEM_SCALAR_WAVE and s.stepper == lib.STEPPER_RK4
    assert s.n_steps == 10 and s.t0 == 0.1 and s.domain.N == 7 and s.static == (0.0,)


def test_parse_reference_input_files():
    sw = input_file.load(_reference_input("ScalarWave/PlaneWave3D.yaml"))
    assert (sw.system, sw.stepper, sw.order, sw.dt, sw.n_steps) == \
        (lib.SYSTEM_SCALAR_WAVE, lib.STEPPER_ADAMS_BASHFORTH, 3, 1e-3, 50)
    gw = input_file.load(_reference_input("GeneralizedHarmonic/GaugeWave3D.yaml"))
    assert gw.gauge == lib.GAUGE_ANALYTIC_GAUGE_WAVE and gw.gauge_params == (0.1, 1.0)
    assert gw.static == (1.0, -1.0, 1.0) and gw.filter is None and gw.n_steps == 2
    assert not gw.lts_executable and not gw.step_choosers_ignored
    # EvolveGhSingleBlackHole is an LTS executable and the file lists step choosers: a run
    # with fixed global steps is not the reference's run and must be asked for explicitly
    with pytest.raises(input_file.InputFileError, match="local time stepping"):
        input_file.load(_reference_input("GeneralizedHarmonic/KerrSchild.yaml"))
    ks = input_file.load(_reference_input("GeneralizedHarmonic/KerrSchild.yaml"),
                         allow_gts_fixed_step=True)
    assert ks.lts_executable
    assert isinstance(ks.domain, domain.SphericalShell) and ks.domain.n_elements == 6
    assert ks.domain.radii == [1.9, 2.3] and ks.domain.distributions == ["Logarithmic"]
    assert (ks.stepper, ks.order, ks.dt) == (lib.STEPPER_ADAMS_BASHFORTH, 4, 2e-4)
    assert ks.steps_per_slab == 5 and ks.n_steps == 15 and ks.filter == (36.0, 64)
    assert ks.analytic_christoffel and ks.face_bc == {4: "DirichletAnalytic", 5: "DirichletAnalytic"}
    assert set(ks.step_choosers_ignored) == {"LimitIncrease", "ElementSizeCfl", "ErrorControl"}


def test_unsupported_options_are_errors(tmp_path):
    import yaml
    with open(os.path.join(HERE, "inputs", "ScalarWaveRk4.yaml")) as f:
        meta, opts = list(yaml.safe_load_all(f))
    for mutate, msg in (
            (lambda o: o["SpatialDiscretization"]["BoundaryCorrection"].__setitem__("Rusanov", None)
             or o["SpatialDiscretization"]["BoundaryCorrection"].pop("UpwindPenalty"),
             "BoundaryCorrection Rusanov"),
            (lambda o: o["DomainCreator"]["Brick"].__setitem__("InitialGridPoints", [4, 5, 5]),
             "anisotropic"),
            (lambda o: o["Evolution"].__setitem__("TimeStepper", {"AdamsMoultonPc": {"Order": 3}}),
             "TimeStepper AdamsMoultonPc"),
            (lambda o: o["SpatialDiscretization"]["DiscontinuousGalerkin"].__setitem__(
                "Quadrature", "Gauss"), "GaussLobatto")):
        import copy
        o = copy.deepcopy(opts)
        mutate(o)
        with pytest.raises(input_file.InputFileError, match=msg):
            input_file.Run(meta, o)


@pytest.mark.gpu
def test_run_own_fixtures():
    from oracle import oracle as orc
    r = input_file.load(os.path.join(HERE, "inputs", "ScalarWaveRk4.yaml"))
    obs = r.run()
    assert [o[0] for o in obs] == [0, 10] and obs[-1][1] == pytest.approx(0.12)
    assert obs[0][2]["Error(Psi)"] == 0.0
    # the same run with the oracle: identical error norms
    dom = r.domain
    x, J, nb = dom.coords(), dom.inverse_jacobian(), dom.neighbors()
    stat = np.zeros((dom.n_elements, 1, dom.N ** 3))
    oev = orc.Evolution(lambda v, t: orc.dg_rhs(0, dom.N, v, J, stat, nb), r.u0(x, r.t0), r.t0,
                        r.dt, "RK4")
    for _ in range(r.n_steps):
        oev.step()
    exact = r.u0(x, oev.time)
    npts = exact.shape[0] * exact.shape[2]
    for name, (a, b) in zip(("Psi", "Pi", "Phi"), ((0, 1), (1, 2), (2, 5))):
        want = np.sqrt(np.sum((oev.u[:, a:b] - exact[:, a:b]) ** 2) / npts)
        assert obs[-1][2][f"Error({name})"] == pytest.approx(want, rel=1e-8)
    assert obs[-1][2]["Error(Psi)"] < 1e-4 and obs[-1][2]["Error(Phi)"] < 1e-2
    g = input_file.load(os.path.join(HERE, "inputs", "GaugeWaveBjorhus.yaml"))
    obs = g.run()
    assert [o[0] for o in obs] == [0, 2, 4]
    # 6 points per element over a 2-wavelength box: errors at the 1e-4 level, growing slowly
    assert all(v < 2e-2 for v in obs[-1][2].values()) and obs[-1][2]["Error(Pi)"] > 0.0


def _oracle_error_norms(run, n_steps):
    """The same input file evolved by the CPU oracle: Error(...) L2 norms as ObserveNorms
    reports them (NormType L2Norm, Components Sum, ObserveNorms.hpp:60-80)."""
    from oracle import oracle as orc
    from spectre_b200 import evolution
    problem = run.problem()
    part = domain.Partition(problem.neighbors, 1, 0, boundary_slots=problem.dirichlet_analytic,
                            neighbor_direction=problem.orientations[0],
                            face_permutation=problem.orientations[1], mortars=problem.mortars)
    ids, N = part.global_ids, problem.N
    x, J, stat = problem.coords(ids), problem.inverse_jacobian(ids), problem.static(ids)
    u0 = problem.u0(ids, run.t0)
    kw = {}
    if part.oriented:
        kw = dict(nbr_dir=part.local_neighbor_direction, face_perm=part.local_face_permutation)

    def gauge_fields(t):
        H = np.zeros((len(ids), 4, N ** 3))
        dH = np.zeros((len(ids), 16, N ** 3))
        for e in range(len(ids)):
            H[e], dH[e] = orc.analytic_christoffel_gauge(N, run.u0(x[e], t), J[e])
        return np.concatenate([stat, H, dH], axis=1)
    static_sf = gauge_fields(run.t0) if run.analytic_christoffel else None

    def rhs(v, t):
        ext = (evolution.boundary_ghost_data(problem, part, t, 55)[:, :50]
               if part.external_faces else None)
        if run.gauge == lib.GAUGE_ANALYTIC_GAUGE_WAVE:
            sf = gauge_fields(t)      # AnalyticChristoffel of the time-dependent solution
        elif run.analytic_christoffel:
            sf = static_sf
        else:
            return orc.dg_rhs(1, N, v, J, stat, part.local_neighbors, ext_u=ext, **kw)
        return orc.dg_rhs(1, N, v, J, sf, part.local_neighbors, gauge_params=orc.GAUGE_GIVEN,
                          ext_u=ext, **kw)
    post = None
    if run.filter:
        F = orc.exponential_filter_matrix(N, *run.filter)
        post = lambda v: orc.apply_filter(N, v, F)
    oev = orc.Evolution(rhs, u0, run.t0, run.dt, f"AB{run.order}", post_update=post)
    for _ in range(n_steps):
        oev.step()
    exact = problem.u0(ids, oev.time)
    npts = exact.shape[0] * exact.shape[2]
    return {f"Error({nm})": float(np.sqrt(np.sum((oev.u[:, a:b] - exact[:, a:b]) ** 2) / npts))
            for nm, (a, b) in zip(("SpacetimeMetric", "Pi", "Phi"), ((0, 10), (10, 20), (20, 50)))}


@pytest.mark.gpu
def test_run_reference_input_files():
    """GaugeWave3D.yaml and KerrSchild.yaml of the reference, unchanged (the committed
    byte-for-byte copies when /root/reference is absent): the printed Error(...) norms equal
    the CPU oracle's for the same file to 1e-12 relative (plus 1e-15 absolute: the norms
    themselves are differences of nearly equal numbers)."""
    gw = input_file.load(_reference_input("GeneralizedHarmonic/GaugeWave3D.yaml"))
    obs = gw.run()
    assert [o[0] for o in obs] == [0, 2]
    assert all(v < 1e-3 for v in obs[-1][2].values())
    want = _oracle_error_norms(gw, gw.n_steps)
    for k, v in want.items():
        assert obs[-1][2][k] == pytest.approx(v, rel=1e-12, abs=1e-15), k
    # EvolveGhSingleBlackHole is LTS with step choosers: only on request, fixed global steps
    ks = input_file.load(_reference_input("GeneralizedHarmonic/KerrSchild.yaml"),
                         allow_gts_fixed_step=True)
    obs = ks.run()
    assert obs[-1][0] == 15 and all(np.isfinite(v) and v < 0.2 for v in obs[-1][2].values())
    want = _oracle_error_norms(ks, ks.n_steps)
    for k, v in want.items():
        assert obs[-1][2][k] == pytest.approx(v, rel=1e-12, abs=1e-15), k


def test_sphere_per_block_refinement_options(tmp_path):
    """The four forms of Sphere.InitialRefinement (Sphere.hpp:222-233, ExpandOverBlocks):
    number, [phi, theta, r], one triple per block, map over block / group names.  Per-block
    values give oriented 2:1 mortars between the wedges."""
    import yaml
    with open(_reference_input("GeneralizedHarmonic/KerrSchild.yaml")) as f:
        meta, opts = list(yaml.safe_load_all(f))

    def load_with(**sphere_options):
        o = yaml.safe_load(yaml.safe_dump(opts))
        o["DomainCreator"]["Sphere"].update(sphere_options)
        path = tmp_path / "input.yaml"
        with open(path, "w") as f:
            yaml.safe_dump_all([meta, o], f)
        return input_file.load(str(path), allow_gts_fixed_step=True)
    r = load_with(InitialRefinement=[1, 1, 2])
    assert r.domain.n_elements == 6 * 4 * 4 and len(r.domain.mortars()) == 0
    per_block = [[1, 1, 1]] + [[0, 0, 0]] * 5
    r = load_with(InitialRefinement=per_block)
    assert r.domain.n_elements == 8 + 5 and len(r.domain.mortars()) == 16
    by_name = load_with(InitialRefinement={"Shell0UpperZ": [1, 1, 1], "Shell0LowerZ": [0, 0, 0],
                                           "Shell0UpperY": 0, "Shell0LowerY": 0,
                                           "Shell0UpperX": 0, "Shell0LowerX": 0})
    np.testing.assert_array_equal(by_name.domain.mortars(), r.domain.mortars())
    assert ((r.domain.mortars()[:, 3] >> 3) != 0).any()
    two = load_with(RadialPartitioning=[2.1], RadialDistribution=["Logarithmic", "Linear"],
                    InitialRefinement={"Shell0": [1, 1, 0], "Shell1": [0, 0, 1]})
    assert two.domain.n_layers == 2 and two.domain.n_elements == 6 * 4 + 6 * 2
    assert len(two.domain.mortars()) == 6 * 4
    for bad, msg in (({"Shell0": 0, "Shell0UpperZ": 1}, "duplicate block name"),
                     ({"Shell0UpperZ": 1}, "is missing"),
                     ({"Shell7": 1}, "unknown block or group"),
                     ([[0, 0, 0]] * 5, "you supplied 5 values"),
                     ([1, 0, 0], "different angular refinement levels")):
        with pytest.raises(input_file.InputFileError, match=msg):
            load_with(InitialRefinement=bad)
    with pytest.raises(input_file.InputFileError, match="p-refinement"):
        load_with(InitialGridPoints={"Shell0UpperZ": 6, "Shell0LowerZ": 5, "Shell0UpperY": 5,
                                     "Shell0LowerY": 5, "Shell0UpperX": 5, "Shell0LowerX": 5})
    # finer in the angle on one side, finer in radius on the other: no such mortar here
    with pytest.raises(ValueError, match="unsupported non-conforming interface"):
        load_with(InitialRefinement=[[1, 1, 0]] + [[0, 0, 1]] * 5).domain.mortars()


def test_one_sided_periodic_and_multi_dimension_bjorhus_are_errors():
    import copy
    import yaml
    with open(os.path.join(HERE, "inputs", "GaugeWaveBjorhus.yaml")) as f:
        meta, opts = list(yaml.safe_load_all(f))
    bcs = opts["DomainCreator"]["Brick"]["BoundaryConditions"]
    o = copy.deepcopy(opts)
    o["DomainCreator"]["Brick"]["BoundaryConditions"][1] = {
        "Lower": "Periodic", "Upper": {"DirichletAnalytic": None}}
    with pytest.raises(input_file.InputFileError, match="only one side"):
        input_file.Run(meta, o)
    o = copy.deepcopy(opts)
    bj = {"ConstraintPreservingBjorhus": {"Type": "ConstraintPreserving"}}
    o["DomainCreator"]["Brick"]["BoundaryConditions"] = [{"Lower": bj, "Upper": bj},
                                                         {"Lower": bj, "Upper": bj}, bcs[2]]
    with pytest.raises(input_file.InputFileError, match="more than one dimension"):
        input_file.Run(meta, o)


def _binary_domain_options():
    """DomainCreator block in the schema of support/Pipelines/Bbh/Inspiral.yaml:54-112
    (static maps, CubeScale 1)."""
    excise = {"ExciseWithBoundaryCondition": {"DemandOutgoingCharSpeeds": None}}
    groups = ("ObjectAShell", "ObjectACube", "ObjectBShell", "ObjectBCube", "Envelope",
              "OuterShell")
    return {"BinaryCompactObject": {
        "ObjectA": {"InnerRadius": 0.8, "OuterRadius": 4.0, "XCoord": 8.0, "Interior": excise,
                    "UseLogarithmicMap": True},
        "ObjectB": {"InnerRadius": 0.8, "OuterRadius": 4.0, "XCoord": -8.0, "Interior": excise,
                    "UseLogarithmicMap": True},
        "CenterOfMassOffset": [0.0, 0.0],
        "Envelope": {"Radius": 60.0, "RadialDistribution": "Logarithmic"},
        "OuterShell": {"Radius": 300.0, "RadialDistribution": "Linear", "OpeningAngle": 120.0,
                       "BoundaryCondition": {"ConstraintPreservingBjorhus":
                                             {"Type": "ConstraintPreservingPhysical"}}},
        "UseEquiangularMap": True, "CubeScale": 1.0,
        "InitialRefinement": {g: ([1, 1, 0] if g.endswith("Cube") else [0, 0, 0])
                              for g in groups},
        "InitialGridPoints": 4, "TimeDependentMaps": None}}


def test_binary_compact_object_creator_options(tmp_path):
    """DomainCreator: BinaryCompactObject in the option schema of Inspiral.yaml:54-112 (with
    the KerrSchild.yaml evolution options around it): the 44-block domain with per-group
    refinement, the three boundary conditions on their spheres, and explicit errors for what
    the path cannot represent."""
    import yaml
    with open(_reference_input("GeneralizedHarmonic/KerrSchild.yaml")) as f:
        meta, opts = list(yaml.safe_load_all(f))

    def load_with(**changes):
        o = yaml.safe_load(yaml.safe_dump(opts))
        o["DomainCreator"] = _binary_domain_options()
        o["DomainCreator"]["BinaryCompactObject"].update(changes)
        path = tmp_path / "bbh.yaml"
        with open(path, "w") as f:
            yaml.safe_dump_all([meta, o], f)
        return input_file.load(str(path), allow_gts_fixed_step=True)
    r = load_with()
    assert r.domain.n_blocks == 44 and r.domain.n_elements == 32 + 12 * 4
    assert len(r.domain.mortars()) == 88
    assert r.outgoing and not r.ghost
    nbr = r.domain.neighbors()
    ext = [(e, d) for e in range(r.domain.n_elements) for d in range(6) if nbr[e, d] == -1]
    kinds = [r.bjorhus(e, d) for e, d in ext]
    assert kinds.count("ConstraintPreservingPhysical") == 10 and kinds.count(None) == 12
    p = r.problem()
    assert p.neighbors.shape == (80, 6)
    for changes, msg in (({"CubeScale": 1.2}, "CubeScale"),
                         ({"CenterOfMassOffset": [0.1, 0.0]}, "CenterOfMassOffset"),
                         ({"InitialGridPoints": {g: ([7, 7, 5] if g == "Envelope" else 5) for g in
                                                 ("ObjectAShell", "ObjectACube", "ObjectBShell",
                                                  "ObjectBCube", "Envelope", "OuterShell")}},
                          "p-refinement"),
                         ({"TimeDependentMaps": {"InitialTime": 0.0}}, "time-dependent maps"),
                         ({"Envelope": {"Radius": 20.0, "RadialDistribution": "Linear"}},
                          "envelope radius is too small")):
        with pytest.raises(input_file.InputFileError, match=msg):
            load_with(**changes)


@pytest.mark.gpu
def test_run_kerr_schild_with_local_time_stepping(tmp_path):
    """KerrSchild.yaml belongs to an LTS executable: --lts-fixed-levels runs it through the LTS
    entry points with the steps the reference starts with (largest slab / 2^n below
    InitialTimeStep and the ElementSizeCfl goal; the later step changes of LimitIncrease /
    ErrorControl are not reproduced and are reported as ignored).  The original file has one
    step-size level; a thick-shell variant of it (outer radius 30.4 M, four radial layers) has
    several.  The oracle's LtsEvolution of the same set-up gives the same error norms."""
    import yaml
    from oracle import lts as olts
    from oracle import oracle as orc
    from spectre_b200 import evolution
    from spectre_b200 import lts as hlts
    path = _reference_input("GeneralizedHarmonic/KerrSchild.yaml")
    ks = input_file.load(path, lts_fixed_levels=True)
    assert ks.step_choosers_ignored == ["LimitIncrease", "ErrorControl"]
    obs = ks.run_lts()
    assert [o[0] for o in obs] == [0, 1, 2, 3] and obs[-1][1] == pytest.approx(0.003)
    assert set(ks.lts_levels.tolist()) == {0} and ks.lts_dt_coarse == pytest.approx(1.25e-4)
    assert all(np.isfinite(v) and v < 0.2 for v in obs[-1][2].values())
    # thick shell: several levels
    with open(path) as f:
        meta, opts = list(yaml.safe_load_all(f))
    sph = opts["DomainCreator"]["Sphere"]
    sph["OuterRadius"] = 30.4
    sph["InitialRefinement"] = [0, 0, 2]
    opts["Evolution"]["InitialTimeStep"] = 0.1     # the ElementSizeCfl goals lie below it
    opts["Evolution"]["InitialSlabSize"] = 0.1
    thick = tmp_path / "KerrSchildThick.yaml"
    with open(thick, "w") as f:
        yaml.safe_dump_all([meta, opts], f)
    run = input_file.load(str(thick), lts_fixed_levels=True)
    obs = run.run_lts(n_slabs=2)
    levels = run.lts_levels
    assert len(set(levels.tolist())) >= 2 and np.all(np.diff(levels) >= 0)
    assert all(np.isfinite(v) and v < 0.2 for v in obs[-1][2].values())
    # the oracle on the same elements, levels and past states
    problem = run.problem()
    lev = hlts.LtsEvolution.__new__(hlts.LtsEvolution)   # only for its element order
    part = domain.Partition(problem.neighbors, 1, 0, boundary_slots=problem.dirichlet_analytic,
                            neighbor_direction=problem.orientations[0],
                            face_permutation=problem.orientations[1], mortars=problem.mortars)
    ids0 = part.global_ids
    speed = hlts.gh_largest_characteristic_speed(problem.u0(ids0, 0.0), problem.static(ids0)[:, 1])
    stable = lib.stepper_properties(lib.STEPPER_ADAMS_BASHFORTH, run.order)[3]
    goal = np.minimum(hlts.element_size_cfl(hlts.size_of_element(problem.brick, ids0), speed,
                                            stable, run.element_size_cfl), run.dt)
    n = hlts.levels_from_step_limit(goal, run.slab_size, max_level=60)
    part.reorder(np.argsort(n - n.min(), kind="stable"))
    ids, N = part.global_ids, problem.N
    np.testing.assert_array_equal(np.sort(n - n.min()), levels)
    x, J, stat = problem.coords(ids), problem.inverse_jacobian(ids), problem.static(ids)
    u0 = problem.u0(ids, 0.0)
    H = np.zeros((len(ids), 4, N ** 3))
    dH = np.zeros((len(ids), 16, N ** 3))
    for e in range(len(ids)):
        H[e], dH[e] = orc.analytic_christoffel_gauge(N, u0[e], J[e])
    ext = evolution.boundary_ghost_data(problem, part, 0.0, 55)[:, :50]
    F = orc.exponential_filter_matrix(N, *run.filter)
    ev = olts.LtsEvolution(1, N, J, np.concatenate([stat, H, dH], axis=1), part.local_neighbors,
                           levels, run.order, 0.0, run.lts_dt_coarse, u0, lambda j: u0,
                           gauge_params=orc.GAUGE_GIVEN, ext_u=ext,
                           nbr_dir=part.local_neighbor_direction,
                           face_perm=part.local_face_permutation,
                           post_update=lambda v: orc.apply_filter(N, v, F))
    ev.take_coarse_steps(2 * int(round(run.slab_size / run.lts_dt_coarse)))
    assert ev.time() == pytest.approx(obs[-1][1], rel=1e-13)
    npts = u0.shape[0] * u0.shape[2]
    for nm, (a, b) in zip(("SpacetimeMetric", "Pi", "Phi"), ((0, 10), (10, 20), (20, 50))):
        want = float(np.sqrt(np.sum((ev.u[:, a:b] - u0[:, a:b]) ** 2) / npts))
        assert obs[-1][2][f"Error({nm})"] == pytest.approx(want, rel=1e-8, abs=1e-15), nm
