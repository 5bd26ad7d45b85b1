"""world_size-2 tests of the multi-GPU host logic on the CPU (gloo backend): view / mesh
sharding, the integer all-reduce of label votes and the optional view-slab gather.
The per-rank votes come from the oracle here (no GPU in this container); on the GPU
box test_gpu_parity.py checks the kernels that produce them."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_covers_everything_once():
    from flyby_b200 import dist as fd
    for n in (0, 1, 7, 92, 162, 1000):
        for ws in (1, 2, 3, 4, 8):
            parts = [fd.shard_range(n, r, ws) for r in range(ws)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(ws - 1))
            sizes = [e - b for b, e in parts]
            assert max(sizes) - min(sizes) <= 1
    assert fd.shard_items(list(range(1000)), 3, 8) == list(range(375, 500))  # config 3: 125 meshes per GPU
    with pytest.raises(ValueError):
        fd.shard_range(10, 2, 2)


def _worker(rank, world_size, port, tmpdir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "fly-by-cnn_b200"))
    import torch.distributed as dist
    from flyby_b200 import dist as fd
    from flyby_b200 import synth
    from oracle import oracle as orc
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    try:
        assert fd.world() == (rank, world_size)
        P, F = synth.bumpy_sphere(8, 3)
        U, _, _ = orc.unit_surf(P)
        N = orc.point_normals(U, F)
        views = orc.icosahedron(1, 2.0)[:7]        # 7 views: an uneven split (3 + 4)
        K, res = 5, 24
        b, e = fd.shard_range(len(views), rank, world_size)
        feat, ids = orc.render(U, F, N, views[b:e], 2.0, res, grid=False)
        gidx = np.arange(b * res * res, e * res * res).reshape(ids.shape)
        labels = ((gidx * 7) % K).astype(np.int32)   # a deterministic per-pixel "prediction"
        votes = orc.backproject(ids, labels, len(F), K)
        total = fd.allreduce_votes(votes)
        all_ids = fd.gather_view_slabs(ids, len(views))
        all_feat = fd.gather_view_slabs(feat, len(views))
        np.savez(os.path.join(tmpdir, "rank%d.npz" % rank), votes=total, ids=all_ids, feat=all_feat)
    finally:
        dist.destroy_process_group()


def test_vote_allreduce_and_gather_two_ranks(tmp_path, orc):
    import torch.multiprocessing as mp
    from flyby_b200 import synth
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    P, F = synth.bumpy_sphere(8, 3)
    U, _, _ = orc.unit_surf(P)
    N = orc.point_normals(U, F)
    views = orc.icosahedron(1, 2.0)[:7]
    K, res = 5, 24
    feat, ids = orc.render(U, F, N, views, 2.0, res, grid=False)
    labels = ((np.arange(ids.size).reshape(ids.shape) * 7) % K).astype(np.int32)
    ref = orc.backproject(ids, labels, len(F), K)
    for r in range(2):
        got = np.load(os.path.join(str(tmp_path), "rank%d.npz" % r))
        assert np.array_equal(got["votes"], ref)      # sharded by view == single process, bit for bit
        assert np.array_equal(got["ids"], ids) and np.array_equal(got["feat"], feat)


def test_allreduce_is_identity_without_process_group():
    from flyby_b200 import dist as fd
    v = np.arange(12, dtype=np.int32).reshape(3, 4)
    assert fd.allreduce_votes(v) is v and fd.world() == (0, 1)
    assert fd.gather_view_slabs(v, 3) is v


def test_numa_helpers_without_gpu():
    """near_gpu is a no-op where sysfs does not expose the GPU's node (this container); the cpulist parser is exact."""
    import os
    from flyby_b200 import dist as fdist
    assert fdist._parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert fdist._parse_cpulist("") == set()
    before = os.sched_getaffinity(0)
    with fdist.near_gpu(0) as n:
        assert n.node is None or isinstance(n.node, int)
    assert os.sched_getaffinity(0) == before
