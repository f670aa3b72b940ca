"""CPU, world_size 2, gloo: the multi-GPU sharding logic (rows all-reduced / queries all-gathered)."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import pandas as pd
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from conftest import OracleBackedEngine
    import statdepth_b200._functional as f
    import statdepth_b200._pointcloud as p
    from statdepth_b200 import FunctionalDepth, PointcloudDepth, _dist, enable_distributed
    fake = OracleBackedEngine()
    f.get_engine = lambda device=None: fake
    p.get_engine = lambda device=None: fake

    assert _dist.world() == (0, 1)  # an initialised process group alone does not shard: opt-in
    enable_distributed()
    assert _dist.world() == (rank, world)
    rng = np.random.default_rng(9)
    X = rng.standard_normal((13, 17)).cumsum(0)  # 13 rows: uneven split 7 + 6
    df = pd.DataFrame(X)
    res = dict(
        relax=FunctionalDepth([df], relax=True).values,
        relax_tc=FunctionalDepth([df], relax=True, to_compute=[5, 2, 11]).values,
        strict=FunctionalDepth([df], relax=False).values,
        l1=PointcloudDepth(pd.DataFrame(rng.standard_normal((9, 2))), containment='l1').values,
        blocks=np.array([_dist.block(13, r, world) for r in range(world)]),
    )
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), **res)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_matches_single_process(tmp_path, oracle):
    from math import comb
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    r0 = np.load(tmp_path / "rank0.npz")
    r1 = np.load(tmp_path / "rank1.npz")
    rng = np.random.default_rng(9)
    X = rng.standard_normal((13, 17)).cumsum(0)
    P = rng.standard_normal((9, 2))
    relax = oracle.mbd_counts_all(X).astype(np.float64) / 13 / comb(17, 2)
    strict = oracle.bd_counts(X).astype(np.float64) / comb(17, 2)
    for r in (r0, r1):  # every rank ends with the full, identical result
        np.testing.assert_allclose(r["relax"], relax, rtol=1e-13)
        np.testing.assert_allclose(r["relax_tc"], relax[[5, 2, 11]], rtol=1e-13)
        assert r["strict"].tolist() == strict.tolist()
        np.testing.assert_allclose(r["l1"], oracle.l1_depth(P), rtol=1e-13)
        assert r["blocks"].tolist() == [[0, 7], [7, 13]]
