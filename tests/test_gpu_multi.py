"""Multi-GPU tests (need >= 2 GPUs on the box; skipped otherwise): sharding by viewpoint with the
library-side NCCL all-reduce of the label votes, and sharding by mesh.  One process per GPU."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world_size, port, tmpdir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "fly-by-cnn_b200"))
    import torch
    import torch.distributed as dist
    from flyby_b200 import _cabi, labeling, synth
    from flyby_b200 import dist as fd
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world_size, device_id=torch.device("cuda", rank))
    try:
        dev = torch.device("cuda", rank)
        ctx = _cabi.Context(rank)
        fd.init_nccl(ctx)                                   # library-side communicator (unique id over torch)
        # ---- by viewpoint: every rank holds the mesh + LBVH, renders its own views
        P, F = synth.bumpy_sphere(30, 11)
        ctx.set_mesh(P, F)
        ctx.build()
        n = ctx.views_icosahedron(2, 2.0)
        opts = ctx.render_opts()
        res, K = 96, 7
        b, e = fd.shard_range(n, rank, world_size)
        feat = torch.empty((e - b, res, res, 4), dtype=torch.float32, device=dev)
        ids = torch.empty((e - b, res, res), dtype=torch.int32, device=dev)
        ctx.render_into(res, opts, feat, ids, view_begin=b, view_end=e)
        gidx = torch.arange(b * res * res, e * res * res, device=dev, dtype=torch.int64).reshape(ids.shape)
        labels = ((gidx * 5) % K).to(torch.int32)
        voter = labeling.CellVoter(ctx, K, use_torch=True)
        torch.cuda.synchronize()                            # labels were computed on torch's stream
        voter.add(ids, labels)
        voter.reduce()                                      # ncclAllReduce on the context stream
        ctx.synchronize()
        cell_labels = voter.cell_labels(default=33)
        ctx.synchronize()
        all_ids = fd.gather_view_slabs(ids, n)
        # ---- by mesh: rank r owns meshes r, r + world, ... (no exchange at all)
        mine = fd.shard_items(list(range(4)), rank, world_size)
        sums = {}
        for m in mine:
            Pm, Fm = synth.bumpy_sphere(16, 100 + m)
            ctx.set_mesh(Pm, Fm)
            ctx.build()
            fm, im = ctx.render(48, opts, view_end=6)
            sums[m] = (int(im.astype(np.int64).sum()), float(fm.astype(np.float64).sum()))
        np.savez(os.path.join(tmpdir, "rank%d.npz" % rank), votes=voter.votes.cpu().numpy(),
                 labels=cell_labels.cpu().numpy(), ids=all_ids.cpu().numpy(),
                 mesh_ids=np.array(sorted(sums)), mesh_sums=np.array([sums[k] for k in sorted(sums)]))
        ctx.close()
    finally:
        dist.destroy_process_group()


def test_two_gpus_view_sharding_and_vote_allreduce(tmp_path, orc):
    torch = pytest.importorskip("torch")
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from flyby_b200 import _cabi, synth
    port = 29600 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    # single-GPU reference of the same work
    ctx = _cabi.Context(0)
    P, F = synth.bumpy_sphere(30, 11)
    ctx.set_mesh(P, F)
    ctx.build()
    n = ctx.views_icosahedron(2, 2.0)
    opts = ctx.render_opts()
    res, K = 96, 7
    feat, ids = ctx.render(res, opts)
    labels = ((np.arange(ids.size, dtype=np.int64).reshape(ids.shape) * 5) % K).astype(np.int32)
    votes = ctx.backproject(ids, labels, K)
    assert np.array_equal(votes, orc.backproject(ids, labels, len(F), K))
    ref_labels = ctx.votes_argmax(votes, default=33)
    sums = {}
    for m in range(4):
        Pm, Fm = synth.bumpy_sphere(16, 100 + m)
        ctx.set_mesh(Pm, Fm)
        ctx.build()
        fm, im = ctx.render(48, opts, view_end=6)
        sums[m] = (int(im.astype(np.int64).sum()), float(fm.astype(np.float64).sum()))
    ctx.close()
    seen = []
    for r in range(2):
        got = np.load(os.path.join(str(tmp_path), "rank%d.npz" % r))
        assert np.array_equal(got["votes"], votes)          # sharded + all-reduced == single GPU, bit for bit
        assert np.array_equal(got["labels"], ref_labels)
        assert np.array_equal(got["ids"], ids)
        for m, s in zip(got["mesh_ids"], got["mesh_sums"]):
            assert (int(s[0]), float(s[1])) == sums[int(m)]
            seen.append(int(m))
    assert sorted(seen) == [0, 1, 2, 3]
