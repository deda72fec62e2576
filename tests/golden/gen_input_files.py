#!/usr/bin/env python
"""Copies the three input files of the reference that the accelerated path runs unchanged
into tests/golden/inputs/ (byte for byte), so that the YAML front-end tests also run on
the GPU box, where /root/reference does not exist:

    tests/InputFiles/ScalarWave/PlaneWave3D.yaml            (EvolveScalarWave3D)
    tests/InputFiles/GeneralizedHarmonic/GaugeWave3D.yaml   (EvolveGhNoBlackHole3D)
    tests/InputFiles/GeneralizedHarmonic/KerrSchild.yaml    (EvolveGhSingleBlackHole)

They are test fixtures in the reference's option schema (data, not source code).
usage: python tests/golden/gen_input_files.py [/root/reference]
"""
import hashlib
import os
import shutil
import sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
FILES = ("ScalarWave/PlaneWave3D.yaml", "GeneralizedHarmonic/GaugeWave3D.yaml",
         "GeneralizedHarmonic/KerrSchild.yaml")

out_dir = os.path.join(HERE, "inputs")
os.makedirs(out_dir, exist_ok=True)
with open(os.path.join(out_dir, "SHA256SUMS"), "w") as sums:
    for rel in FILES:
        src = os.path.join(REF, "tests", "InputFiles", rel)
        dst = os.path.join(out_dir, os.path.basename(rel))
        shutil.copyfile(src, dst)
        digest = hashlib.sha256(open(dst, "rb").read()).hexdigest()
        sums.write(f"{digest}  {os.path.basename(rel)}  (tests/InputFiles/{rel})\n")
        print(dst, digest[:16])
