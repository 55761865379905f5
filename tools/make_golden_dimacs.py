"""Copies a few small DIMACS fixtures of the reference (data files: src/MinCostFlow.Problems/Resources/**/*.min, *.sol) into
tests/golden/dimacs/ so that the native reader (csrc/mcf_io.cpp) is pinned on the reference's own files on the GPU box too,
where /root/reference does not exist.  Run in the build container:  python tools/make_golden_dimacs.py"""
import os
import shutil

REF = "/root/reference/src/MinCostFlow.Problems/Resources"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "dimacs")
PICK = ["grid/grid_5x5", "transport/transport_2x3", "circulation/cycle_shortcut", "netgen/netgen_8_08a", "path/path_5node", "assignment/assignment_3x3"]

os.makedirs(OUT, exist_ok=True)
for stem in PICK:
    for ext in (".min", ".sol"):
        src = os.path.join(REF, stem + ext)
        if os.path.exists(src):
            shutil.copyfile(src, os.path.join(OUT, os.path.basename(stem) + ext))
            os.chmod(os.path.join(OUT, os.path.basename(stem) + ext), 0o644)
print(sorted(os.listdir(OUT)))
