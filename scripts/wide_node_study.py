"""CPU study for DESIGN.md: how many dependent node loads per query a 4-wide collapse of the LBVH would need,
against the 2-wide Node64 traversal, for the radius and the kNN query of the bench workload (oracle tree, numpy).
A "visit" is one dependent record load: Node64 = an internal node (tests its two children); wide = an internal
node whose internal children are expanded in place (tests up to four grandchildren).
    python scripts/wide_node_study.py [n] [queries]"""
import heapq
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from tests import clouds  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 1500
pts = clouds.filled_box(0x5EED0001, n)
qs = clouds.filled_box(0x5EED0002, n)[:: max(1, n // nq)][:nq]
r = float(clouds.bvh_driver_radius(10))
tree = oracle.Tree(pts, 0)
ex = tree.export()
# reference numbering: leaves 0..n-1 (sorted order), internal n..2n-2, root n; right child = rope of the left child
left = ex["left_child"]
rope_i, rope_l = ex["rope"], ex["leaf_rope"]
boxes = ex["boxes"].astype(np.float64)
order = ex["leaf_index"]
P = pts[order].astype(np.float64)


def children(node):
    l = left[node - n]
    rr = rope_l[l] if l < n else rope_i[l - n]
    return l, rr


def box_d2(node, c):
    if node < n:
        d = P[node] - c
    else:
        b = boxes[node - n]
        d = np.maximum(np.maximum(b[:3] - c, 0.0), c - b[3:])
    return float(d @ d)


def spatial(c, wide):
    visits = 0
    stack = [n]
    while stack:
        x = stack.pop()
        visits += 1
        kids = list(children(x))
        if wide:
            kids = [g for k in kids for g in (children(k) if k >= n else (k,))]
        for k in kids:
            if k >= n and box_d2(k, c) <= r * r:
                stack.append(k)
    return visits


def nearest(c, k, wide):
    visits = 0
    heap = []  # max-heap of (-d2)
    radius2 = np.inf
    stack = [(0.0, n)]
    while stack:
        d, x = stack.pop()
        if not d < radius2:
            continue
        visits += 1
        kids = list(children(x))
        if wide:
            kids = [g for kk in kids for g in (children(kk) if kk >= n else (kk,))]
        cand = sorted((box_d2(kk, c), kk) for kk in kids)
        inner = []
        for dd, kk in cand:
            if not dd < radius2:
                continue
            if kk < n:
                if len(heap) < k:
                    heapq.heappush(heap, -dd)
                else:
                    heapq.heapreplace(heap, -dd)
                if len(heap) == k:
                    radius2 = -heap[0]
            else:
                inner.append((dd, kk))
        for dd, kk in reversed(inner):  # nearest on top of the stack
            stack.append((dd, kk))
    return visits


res = {}
for name, fn in (("radius", lambda c, w: spatial(c, w)), ("knn k=10", lambda c, w: nearest(c, 10, w))):
    two = np.array([fn(c.astype(np.float64), False) for c in qs])
    four = np.array([fn(c.astype(np.float64), True) for c in qs])
    res[name] = (two.mean(), four.mean())
    print("%-9s n=%d: Node64 visits/query %.1f (max %d), 4-wide %.1f (max %d): ratio %.2f" %
          (name, n, two.mean(), two.max(), four.mean(), four.max(), four.mean() / two.mean()))
