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


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present")
def test_parse_reference_input_files():
    sw = input_file.load(f"{REF}/ScalarWave/PlaneWave3D.yaml")
    assert (sw.system, sw.stepper, sw.order, sw.dt, sw.n_steps) == \
        (lib.SYSTEM_SCALAR_WAVE, lib.STEPPER_ADAMS_BASHFORTH, 3, 1e-3, 50)
    gw = input_file.load(f"{REF}/GeneralizedHarmonic/GaugeWave3D.yaml")
    assert gw.gauge == lib.GAUGE_ANALYTIC_GAUGE_WAVE and gw.gauge_params == (0.1, 1.0)
    assert gw.static == (1.0, -1.0, 1.0) and gw.filter is None and gw.n_steps == 2
    ks = input_file.load(f"{REF}/GeneralizedHarmonic/KerrSchild.yaml")
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


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present")
def test_run_reference_input_files():
    gw = input_file.load(f"{REF}/GeneralizedHarmonic/GaugeWave3D.yaml")
    obs = gw.run()
    assert [o[0] for o in obs] == [0, 2]
    assert all(v < 1e-5 for v in obs[-1][2].values())
    ks = input_file.load(f"{REF}/GeneralizedHarmonic/KerrSchild.yaml")
    obs = ks.run()
    assert obs[-1][0] == 15 and all(np.isfinite(v) and v < 0.2 for v in obs[-1][2].values())


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present")
def test_sphere_per_block_refinement_options(tmp_path):
    """The four forms of Sphere.InitialRefinement (Sphere.hpp:222-233, ExpandOverBlocks):
    number, [phi, theta, r], one triple per block, map over block / group names.  Per-block
    values give oriented 2:1 mortars between the wedges."""
    import yaml
    with open(f"{REF}/GeneralizedHarmonic/KerrSchild.yaml") as f:
        meta, opts = list(yaml.safe_load_all(f))

    def load_with(**sphere_options):
        o = yaml.safe_load(yaml.safe_dump(opts))
        o["DomainCreator"]["Sphere"].update(sphere_options)
        path = tmp_path / "input.yaml"
        with open(path, "w") as f:
            yaml.safe_dump_all([meta, o], f)
        return input_file.load(str(path))
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
