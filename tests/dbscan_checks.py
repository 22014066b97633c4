"""DBSCAN parity checks: exact core partition and noise
set against the oracle, border points validated against their core neighbours,
plus the reference's five-property verifier (oracle.dbscan_verify)."""
import numpy as np

import oracle
from tests import brute


def check_against_oracle(xyz, eps, minpts, labels, impl=0, algo=0, verify=True):
    xyz = np.ascontiguousarray(xyz, np.float32)
    n = len(xyz)
    ref, core = oracle.dbscan(xyz, eps, minpts, impl=0, algo=algo, return_core=True)
    labels = np.asarray(labels)
    assert labels.shape == (n,)
    if minpts == 2:
        # every non-noise point is core: partition is unique, labels = min index
        assert np.array_equal(labels, ref)
    else:
        # core points: label = smallest core index of the component (exact)
        assert np.array_equal(labels[core], ref[core])
        noncore = ~core
        # noise set identical
        if algo == 1:
            assert np.all(labels[noncore] == -1)
        else:
            assert np.array_equal(labels[noncore] == -1, ref[noncore] == -1)
    if verify:
        assert oracle.dbscan_verify(xyz, eps, minpts, labels, algo) == 0


def check_small_bruteforce(xyz, eps, minpts, labels, algo=0):
    """Independent of the oracle's tree: O(n^2) float32 neighbour matrix."""
    xyz = np.ascontiguousarray(xyz, np.float32)
    n = len(xyz)
    adj = brute.dist_point_point(xyz, xyz) <= np.float32(eps)
    core = adj.sum(1) >= minpts
    # components of the core graph
    comp = -np.ones(n, np.int64)
    for s in range(n):
        if core[s] and comp[s] < 0:
            stack = [s]
            comp[s] = s
            while stack:
                u = stack.pop()
                for v in np.nonzero(adj[u] & core)[0]:
                    if comp[v] < 0:
                        comp[v] = s
                        stack.append(v)
    labels = np.asarray(labels)
    for i in range(n):
        if core[i]:
            assert labels[i] == comp[i], (i, labels[i], comp[i])
        else:
            cn = np.nonzero(adj[i] & core)[0]
            if len(cn) == 0 or algo == 1:
                assert labels[i] == -1, i
            else:
                assert labels[i] in set(comp[cn].tolist()), i
