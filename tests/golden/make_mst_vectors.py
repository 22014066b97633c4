#!/usr/bin/env python
"""Golden vectors of the reference's minimum-spanning-tree test (test/tstMinimumSpanningTreeGoldenTest.cpp:86-150):
the 1000 points and the 999 expected edges of test/mst_golden_test_{points,edges}.csv, and the total weights the
test expects for k = 5, 10, 15 (computed there with the hdbscan Python package).  Run in the build container (the
reference tree is not on the GPU box); writes tests/golden/mst_golden.npz."""
import os

import numpy as np

REF = "/root/reference/test"
HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    pts = np.loadtxt(os.path.join(REF, "mst_golden_test_points.csv"), delimiter=",", dtype=np.float64, comments="#")
    e = np.loadtxt(os.path.join(REF, "mst_golden_test_edges.csv"), delimiter=",", dtype=np.float64, comments="#")
    assert pts.shape == (1000, 3) and e.shape == (999, 3)
    np.savez_compressed(os.path.join(HERE, "mst_golden.npz"),
                        points=pts, edges=e[:, :2].astype(np.int32), weights=e[:, 2],
                        # tstMinimumSpanningTreeGoldenTest.cpp:126-129
                        total_weight_k=np.array([5, 10, 15], np.int32),
                        total_weight=np.array([102.68084503576422, 138.0244333174116, 162.51948793942978]))


if __name__ == "__main__":
    main()
