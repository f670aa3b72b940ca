"""2-GPU check of the public API under torchrun (NCCL):  rows all-reduced / queries all-gathered.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/dist_gpu_check.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pandas as pd
import torch
import torch.distributed as dist

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from statdepth_b200 import FunctionalDepth, PointcloudDepth, enable_distributed  # noqa: E402
from statdepth_b200.homogeneity import permutation_test  # noqa: E402

rng = np.random.default_rng(3)
X = pd.DataFrame(rng.standard_normal((37, 2500)).cumsum(0))
P = pd.DataFrame(rng.standard_normal((700, 2)))
F = pd.DataFrame(rng.standard_normal((32, 24)).cumsum(0))
G = pd.DataFrame(rng.standard_normal((32, 24)).cumsum(0) + 3.0)


def run():
    return dict(relax=FunctionalDepth([X], relax=True).values, strict=FunctionalDepth([X], to_compute=list(range(0, 2500, 50))).values,
                l1=PointcloudDepth(P, containment="l1").values, simp=PointcloudDepth(P, containment="simplex").values,
                perm=permutation_test(F, G, B=11, seed=1)["null"])


enable_distributed(True)
multi = run()
enable_distributed(False)
single = run()
ok = all(np.array_equal(multi[k], single[k]) for k in multi)
print("rank", rank, "OK" if ok else "MISMATCH", {k: float(np.abs(multi[k] - single[k]).max()) for k in multi}, flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
