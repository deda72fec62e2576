"""Generate tests/golden/gh_dudt_spec.json from the reference's own test.

Run in the authoring container only (needs /root/reference and g++):
    python tests/golden/gen_dudt_golden.py

The reference test tests/Unit/Evolution/Systems/GeneralizedHarmonic/
Test_DuDt.cpp:306-464 fills its input tensors from std::mt19937 gen(1.) with
std::uniform_real_distribution<>(-10, 10), component by component in Tensor
storage order, two grid points per component, and compares the GH right-hand
side with 100 numbers computed by SpEC.  This script (a) regenerates that
random stream with libstdc++ and (b) parses the expected numbers out of the
reference test source, so the oracle can be pinned on the GPU box where the
reference is absent.
"""
import json
import os
import re
import subprocess
import tempfile

REF = "/root/reference/tests/Unit/Evolution/Systems/GeneralizedHarmonic/Test_DuDt.cpp"
HERE = os.path.dirname(os.path.abspath(__file__))

# (name, number of independent components) in the order of Test_DuDt.cpp:314-352
TENSORS = [
    ("psi", 10), ("pi", 10), ("phi", 30), ("d_psi", 30), ("d_pi", 30), ("d_phi", 90),
    ("gauge_function", 4), ("spacetime_deriv_gauge_function", 16), ("gamma0", 1),
    ("gamma1", 1), ("gamma2", 1), ("lapse", 1), ("shift", 3), ("inverse_spatial_metric", 6),
    ("inverse_psi", 10), ("christoffel_first_kind", 40), ("christoffel_second_kind", 40),
    ("trace_christoffel_first_kind", 4), ("normal_one_form", 4), ("normal_vector", 4),
]

CPP = r"""
#include <cstdio>
#include <random>
int main(int argc, char** argv) {
  std::mt19937 gen(1.);
  const int n = %d;
  for (int i = 0; i < n; ++i)
    std::printf("%%.17g\n", std::uniform_real_distribution<>(-10, 10)(gen));
}
"""


def main():
    total = 2 * sum(c for _, c in TENSORS)
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "g.cpp")
        open(src, "w").write(CPP % total)
        exe = os.path.join(d, "g")
        subprocess.check_call(["g++", "-O1", "-o", exe, src])
        stream = [float(x) for x in subprocess.check_output([exe]).split()]
    inputs = {}
    k = 0
    for name, ncomp in TENSORS:
        vals = stream[k:k + 2 * ncomp]
        k += 2 * ncomp
        # [comp][point]
        inputs[name] = [[vals[2 * c], vals[2 * c + 1]] for c in range(ncomp)]
    text = open(REF).read()
    pat = re.compile(
        r"CHECK\((dt_psi|dt_pi|dt_phi)\.get\(([0-9, ]+)\)\[(\d)\] ==\s*\w+\((-?[0-9.]+)\)\);")
    expected = []
    for m in pat.finditer(text):
        idx = [int(t) for t in m.group(2).split(",")]
        expected.append({"tensor": m.group(1), "index": idx, "point": int(m.group(3)),
                         "value": float(m.group(4))})
    assert len(expected) == 100, len(expected)
    out = {"source": "Test_DuDt.cpp:306-464 (SpEC values), mt19937(1) U(-10,10) stream",
           "n_pts": 2, "inputs": inputs, "expected": expected}
    with open(os.path.join(HERE, "gh_dudt_spec.json"), "w") as f:
        json.dump(out, f, indent=0)
    print("wrote", len(expected), "expected values,", total, "inputs")


if __name__ == "__main__":
    main()
