"""Golden vectors for ConstraintPreservingBjorhus (Type: ConstraintPreserving)
from the REFERENCE's numpy twins (authoring container only; needs /root/reference):

    python tests/golden/gen_bjorhus_golden.py

Imports tests/Unit/Evolution/Systems/GeneralizedHarmonic/BoundaryConditions/
Bjorhus.py and .../GeneralizedHarmonic/TestFunctions.py from the reference's
test tree and writes tests/golden/bjorhus.npz: random per-point inputs (the
argument list of dt_*_ConstraintPreserving_static_mesh, as the reference's
Test_Bjorhus.cpp feeds them: every argument an independent random tensor) and
the outputs: two_index_constraint, f_constraint, and the corrections to
dt spacetime_metric / Pi / Phi for both types (ConstraintPreserving and
ConstraintPreservingPhysical).
"""
import os
import sys

import numpy as np

sys.path.insert(0, "/root/reference/tests/Unit")
HERE = os.path.dirname(os.path.abspath(__file__))


def sym(t):
    return 0.5 * (t + np.swapaxes(t, -1, -2))


def main():
    import Evolution.Systems.GeneralizedHarmonic.BoundaryConditions.Bjorhus as bj
    import Evolution.Systems.GeneralizedHarmonic.TestFunctions as ght
    rng = np.random.default_rng(20241018)
    npts = 40
    keys = ["normal_covector", "normal_vector", "spacetime_metric", "pi", "phi", "coords",
            "gamma1", "gamma2", "lapse", "shift", "inverse_spacetime_metric",
            "spacetime_unit_normal_vector", "three_index_constraint", "gauge_source",
            "spacetime_deriv_gauge_source", "dt_spacetime_metric", "dt_pi", "dt_phi",
            "d_spacetime_metric", "d_pi", "d_phi"]
    data = {k: [] for k in keys}
    out = {k: [] for k in ("two_index_constraint", "f_constraint", "corr_g", "corr_pi", "corr_phi",
                           "char_speeds", "phys_corr_pi", "phys_corr_phi")}
    u = lambda *shape: rng.uniform(-1.0, 1.0, shape)
    for p in range(npts):
        d = {
            "normal_covector": u(3), "normal_vector": u(3),
            "spacetime_metric": sym(u(4, 4)), "pi": sym(u(4, 4)), "phi": sym(u(3, 4, 4)),
            "coords": u(3) * 10.0, "gamma1": float(u()), "gamma2": float(u()),
            "lapse": float(rng.uniform(0.5, 1.5)), "shift": u(3),
            "inverse_spacetime_metric": sym(u(4, 4)), "spacetime_unit_normal_vector": u(4),
            "three_index_constraint": sym(u(3, 4, 4)), "gauge_source": u(4),
            "spacetime_deriv_gauge_source": u(4, 4), "dt_spacetime_metric": sym(u(4, 4)),
            "dt_pi": sym(u(4, 4)), "dt_phi": sym(u(3, 4, 4)),
            "d_spacetime_metric": sym(u(3, 4, 4)), "d_pi": sym(u(3, 4, 4)),
            "d_phi": sym(u(3, 3, 4, 4)),
        }
        args = [d[k].copy() if isinstance(d[k], np.ndarray) else d[k] for k in keys]
        out["corr_g"].append(bj.dt_spacetime_metric_static_mesh(*[
            a.copy() if isinstance(a, np.ndarray) else a for a in args]))
        out["corr_pi"].append(bj.dt_pi_ConstraintPreserving_static_mesh(*[
            a.copy() if isinstance(a, np.ndarray) else a for a in args]))
        out["corr_phi"].append(bj.dt_phi_ConstraintPreserving_static_mesh(*[
            a.copy() if isinstance(a, np.ndarray) else a for a in args]))
        # Type ConstraintPreservingPhysical (dt g is the same for both types)
        out["phys_corr_pi"].append(bj.dt_pi_ConstraintPreservingPhysical_static_mesh(*[
            a.copy() if isinstance(a, np.ndarray) else a for a in args]))
        out["phys_corr_phi"].append(bj.dt_phi_ConstraintPreservingPhysical_static_mesh(*[
            a.copy() if isinstance(a, np.ndarray) else a for a in args]))
        t_lo = np.zeros(4)
        t_lo[0] = -d["lapse"]
        ig = d["inverse_spacetime_metric"][1:, 1:] + np.outer(d["shift"], d["shift"]) / d["lapse"] ** 2
        out["two_index_constraint"].append(ght.two_index_constraint(
            d["spacetime_deriv_gauge_source"], t_lo, d["spacetime_unit_normal_vector"], ig,
            d["inverse_spacetime_metric"], d["pi"], d["phi"], d["d_pi"], d["d_phi"], d["gamma2"],
            d["three_index_constraint"]))
        out["f_constraint"].append(ght.f_constraint(
            d["gauge_source"], d["spacetime_deriv_gauge_source"], t_lo,
            d["spacetime_unit_normal_vector"], ig, d["inverse_spacetime_metric"], d["pi"],
            d["phi"], d["d_pi"], d["d_phi"], d["gamma2"], d["three_index_constraint"]))
        out["char_speeds"].append([
            ght.char_speed_upsi(d["gamma1"], d["lapse"], d["shift"], d["normal_covector"]),
            ght.char_speed_uzero(d["gamma1"], d["lapse"], d["shift"], d["normal_covector"]),
            ght.char_speed_uplus(d["gamma1"], d["lapse"], d["shift"], d["normal_covector"]),
            ght.char_speed_uminus(d["gamma1"], d["lapse"], d["shift"], d["normal_covector"])])
        for k in keys:
            data[k].append(d[k])
    save = {"in_" + k: np.array(v) for k, v in data.items()}
    save.update({"out_" + k: np.array(v) for k, v in out.items()})
    np.savez(os.path.join(HERE, "bjorhus.npz"), **save)
    cs = save["out_char_speeds"]
    print("points:", npts, "with an incoming speed:", int((cs.min(axis=1) < 0).sum()),
          "max |corr_pi|:", np.abs(save["out_corr_pi"]).max())


if __name__ == "__main__":
    main()
