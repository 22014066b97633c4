"""Repeated builds compared bit-for-bit against the first one (nodes are a pure function of the input): catches
rare ordering bugs in the hierarchy kernels (shared-memory rounds, global CAS + fence protocol).
    python scripts/stress_build.py [n] [repeats]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import arborx_b200 as abx  # noqa: E402
from tests import clouds  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 3_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 40
space = abx.ExecutionSpace()
bad = 0
for kind, data in (("uniform", clouds.filled_box(7, n)), ("clustered", clouds.gan_tao(5, n))):
    x = torch.from_numpy(data).cuda()
    ref = None
    for it in range(reps):
        bvh = abx.BoundingVolumeHierarchy(space, x)
        lay = bvh.export_reference_layout(space)
        torch.cuda.synchronize()
        if ref is None:
            ref = {k: v.clone() for k, v in lay.items()}
        else:
            for k in ref:
                if not torch.equal(ref[k], lay[k]):
                    bad += 1
                    print("MISMATCH", kind, it, k, int((ref[k] != lay[k]).sum()))
    print(kind, "ok" if not bad else "FAILED", reps, "builds")
print("STRESS", "OK" if not bad else "FAILED")
sys.exit(1 if bad else 0)
