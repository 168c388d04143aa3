"""Generates the committed fixtures under tests/golden/ from the reference tree.

Run in the build container only (reads /root/reference, which does not exist on the GPU box):
    python tests/golden/make_golden.py

dumpR3_frame0.json : frame 0 of the reference's own trajectory dump
    (CUDA-Parallel-MC/CUDA-Parallel-MC/dumpR3.txt:1-73) = the output of init_r
    (kernel.cu:78-89, identical formula to start.cu:47-58) for N=64, L=10.  Pins the lattice
    formula and the index order ix fastest, then iy, then iz.
"""
import json
import os

REF = "/root/reference/CUDA-Parallel-MC/CUDA-Parallel-MC/dumpR3.txt"
HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    with open(REF) as fh:
        lines = [next(fh).strip() for _ in range(9 + 64)]
    assert lines[0].startswith("ITEM: TIMESTEP") and lines[1] == "0" and lines[3] == "64"
    lo, hi = (float(v) for v in lines[5].split())
    atoms = [[float(v) for v in ln.split()[2:5]] for ln in lines[9:]]
    out = {"source": "dumpR3.txt frame 0 (reference init_r, N=64, L=10)", "N": 64, "L": hi - lo,
           "xyz": atoms}
    with open(os.path.join(HERE, "dumpR3_frame0.json"), "w") as fh:
        json.dump(out, fh)
    print("wrote dumpR3_frame0.json", len(atoms), "atoms, L =", hi - lo)


if __name__ == "__main__":
    main()
