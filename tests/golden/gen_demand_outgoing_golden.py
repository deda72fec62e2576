"""Golden vectors for DemandOutgoingCharSpeeds from the REFERENCE's numpy twin
(authoring container only; needs /root/reference):

    python tests/golden/gen_demand_outgoing_golden.py

Writes tests/golden/demand_outgoing.npz: random (gamma_1, lapse, shift, outward
unit normal covector) per point, the four characteristic speeds and the verdict
of tests/Unit/Evolution/Systems/GeneralizedHarmonic/BoundaryConditions/
DemandOutgoingCharSpeeds.py (`characteristic_speeds`, `error`).
"""
import importlib.util
import os

import numpy as np

REF = ("/root/reference/tests/Unit/Evolution/Systems/GeneralizedHarmonic/BoundaryConditions/"
       "DemandOutgoingCharSpeeds.py")
HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    spec = importlib.util.spec_from_file_location("ref_doc", REF)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    rng = np.random.default_rng(20241017)
    npts = 200
    gamma1 = rng.uniform(-1.5, 0.5, npts)
    gamma1[::7] = -1.0
    lapse = rng.uniform(0.2, 1.5, npts)
    shift = rng.uniform(-2.0, 2.0, (npts, 3))
    normal = rng.normal(size=(npts, 3))
    normal /= np.linalg.norm(normal, axis=1)[:, None]
    speeds = np.zeros((npts, 4))
    violated = np.zeros(npts, dtype=bool)
    for k in range(npts):
        speeds[k] = ref.characteristic_speeds(gamma1[k], lapse[k], shift[k], normal[k])
        violated[k] = ref.error(None, normal[k], None, gamma1[k], lapse[k], shift[k]) is not None
    assert violated.any() and not violated.all()
    np.savez(os.path.join(HERE, "demand_outgoing.npz"), gamma1=gamma1, lapse=lapse, shift=shift,
             normal=normal, speeds=speeds, violated=violated)
    print("violated:", int(violated.sum()), "of", npts)


if __name__ == "__main__":
    main()
