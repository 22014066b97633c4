"""DBSCAN known answers from test/tstDBSCAN.cpp and examples/dbscan (data in
tests/golden/vectors.py): verifier accept/reject vectors pin the oracle's
verifier; the end-to-end cases run on both engines."""
import numpy as np
import pytest

import oracle
from tests import dbscan_checks
from tests.engines import SearchException
from tests.golden import vectors as V

ALGO = {"dbscan": 0, "dbscan*": 1}


@pytest.mark.parametrize("case", range(len(V.VERIFIER_CASES)))
def test_verifier_vectors(case):
    pts, eps, minpts, labels, algo, ok = V.VERIFIER_CASES[case]
    verdict = oracle.dbscan_verify(pts, eps, minpts, np.array(labels, np.int32), ALGO[algo]) == 0
    assert verdict == ok


@pytest.mark.parametrize("impl", [0, 1], ids=["fdbscan", "densebox"])
@pytest.mark.parametrize("algo", [0, 1], ids=["dbscan", "dbscan_star"])
def test_dbscan_runs(engine, impl, algo):
    for pts, eps, minpts in V.DBSCAN_RUN_CASES:
        labels = engine.dbscan(pts, eps, minpts, impl, algo)
        assert oracle.dbscan_verify(pts, eps, minpts, labels, algo) == 0
        dbscan_checks.check_small_bruteforce(pts, eps, minpts, labels, algo)


@pytest.mark.parametrize("impl", [0, 1], ids=["fdbscan", "densebox"])
def test_dbscan_example(engine, impl):
    for eps, minpts, accepted in V.EXAMPLE_DBSCAN:
        labels = engine.dbscan(V.EXAMPLE_DBSCAN_POINTS, eps, minpts, impl, 0)
        # compare partitions up to relabelling against the documented outputs
        def canon(l):
            m = {}
            return [(-1 if x < 0 else m.setdefault(x, len(m))) for x in l]
        assert canon(labels.tolist()) in [canon(a) for a in accepted]


@pytest.mark.parametrize("impl", [0, 1], ids=["fdbscan", "densebox"])
def test_dbscan_benchmark_input(engine, impl):
    pts = V.BENCH_INPUT_POINTS
    labels = engine.dbscan(pts, 1.4, 2, impl, 0)
    assert oracle.dbscan_verify(pts, 1.4, 2, labels, 0) == 0


def test_dbscan_preconditions(engine):
    with pytest.raises(SearchException):
        engine.dbscan(V.DB_P2, 0.0, 2)
    with pytest.raises(SearchException):
        engine.dbscan(V.DB_P2, 1.0, 1)


@pytest.mark.parametrize("impl", [0, 1], ids=["fdbscan", "densebox"])
@pytest.mark.parametrize("minpts", [2, 3, 5, 10])
def test_dbscan_random_clustered(engine, impl, minpts):
    rng = np.random.default_rng(7 + minpts)
    centers = rng.uniform(1, 9, (6, 3))
    pts = np.concatenate([c + 0.25 * rng.standard_normal((150, 3)) for c in centers]
                         + [rng.uniform(1, 9, (100, 3))]).astype(np.float32)
    for algo in (0, 1):
        labels = engine.dbscan(pts, 0.2, minpts, impl, algo)
        dbscan_checks.check_small_bruteforce(pts, 0.2, minpts, labels, algo)
        assert oracle.dbscan_verify(pts, 0.2, minpts, labels, algo) == 0
