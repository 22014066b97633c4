"""Unit tests of the label-merge helpers of the distributed DBSCAN against plain Python restatements of
sortAndFilterMergePairs / relabel (cluster/detail/ArborX_DistributedDBSCANHelpers.hpp:501-561, 619-660)."""
import numpy as np
import torch

from arborx_b200.distributed_dbscan import _ghost_distance, _relabel, _sort_and_filter


def ref_sort_and_filter(pairs):
    pairs = sorted(set(map(tuple, pairs)))
    out = []
    i = 0
    while i < len(pairs):
        frm, to = pairs[i]
        out.append((frm, to))
        j = i + 1
        while j < len(pairs) and pairs[j][0] == frm:
            if pairs[j][1] != pairs[j - 1][1]:
                out.append((pairs[j][1], to))
            j += 1
        i = j
    return sorted(set(out))


def ref_relabel(pairs, labels):
    first = {}
    for frm, to in pairs:  # pairs sorted: the first pair of a `from` has its lowest `to`
        first.setdefault(frm, to)
    out = []
    for lab in labels:
        while lab in first:
            lab = first[lab]
        out.append(lab)
    return out


def test_sort_and_filter_and_relabel_random():
    rng = np.random.default_rng(0)
    for trial in range(200):
        m = int(rng.integers(0, 40))
        hi = int(rng.integers(2, 30))
        a = rng.integers(0, hi, m)
        b = rng.integers(0, hi, m)
        keep = a > b  # merge pairs always point to a smaller label
        pairs = np.stack([a[keep], b[keep]], 1).astype(np.int64).reshape(-1, 2)
        got = _sort_and_filter(torch.from_numpy(pairs))
        want = ref_sort_and_filter(pairs.tolist())
        assert [tuple(x) for x in got.tolist()] == want, trial
        labels = rng.integers(-1, hi, 50).astype(np.int64)
        rel = _relabel(got, torch.from_numpy(labels))
        assert rel.tolist() == ref_relabel(want, labels.tolist()), trial


def test_relabel_follows_chains_to_the_smallest_label():
    pairs = torch.tensor([[3, 2], [5, 3], [9, 5], [7, 1]], dtype=torch.int64)
    pairs = _sort_and_filter(pairs)
    labels = torch.tensor([9, 5, 3, 2, 7, 1, 4, -1], dtype=torch.int64)
    assert _relabel(pairs, labels).tolist() == [2, 2, 2, 2, 1, 1, 4, -1]


def test_ghost_distance():
    # DistributedDBSCAN.hpp:77-84: eps for minpts == 2, the float after 2 eps otherwise
    assert _ghost_distance(0.5, 2) == 0.5
    g = np.float32(_ghost_distance(0.5, 5))
    assert g > np.float32(1.0) and g == np.nextafter(np.float32(1.0), np.float32(2.0))
