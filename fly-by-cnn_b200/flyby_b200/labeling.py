"""Universal-labeling back-projection: per-view CNN predictions -> mesh cells / points.

Host API over the vote kernels of csrc/label.cu.  It mirrors what the reference's
predictors do after the network ran:
 * votes[cell, label] += 1 per hit pixel   (src/py/dental_model_seg.py:192-199,
   src/py/predict_v3.py:59-69, src/py/predict.py:154-157)
 * label = argmax over votes, default where nothing voted
   (src/py/predict_v3.py:75-80, src/py/dental_model_seg.py:196-199)
 * cell label -> the cell's first point     (src/py/dental_model_seg.py:208-211)
Background pixels (cell id -1) never vote (the reference indexes row -1 with them).
Work is sharded over ranks by viewpoint; the vote tables are summed with one
integer all-reduce (flyby_b200.dist.allreduce_votes).
"""
import numpy as np

from . import dist as fdist


class CellVoter:
    """Accumulates label votes for the mesh currently held by ``ctx``."""

    def __init__(self, ctx, n_classes, use_torch=None):
        self.ctx = ctx
        self.K = int(n_classes)
        self.n_cells = ctx.n_tris
        if use_torch is None:
            try:
                import torch
                use_torch = torch.cuda.is_available()
            except ImportError:
                use_torch = False
        self.use_torch = use_torch
        if use_torch:
            import torch
            self.votes = torch.zeros((self.n_cells, self.K), dtype=torch.int32, device="cuda:%d" % ctx.device)
            # the zero fill ran on torch's stream; the vote kernels run on the context's: order them once here
            torch.cuda.current_stream(self.votes.device).synchronize()
        else:
            self.votes = np.zeros((self.n_cells, self.K), np.int32)

    def add(self, cell_ids, labels):
        """cell_ids int32 [...], labels int32 [...] (same shape): one vote per hit pixel.  Torch tensors must be
        complete on the context's stream (synchronise the stream that produced them first)."""
        self.ctx.backproject(cell_ids, labels, self.K, votes=self.votes)

    def add_logits(self, cell_ids, logits):
        """logits torch [n, K, H, W] (network output): hard votes from the per-pixel argmax."""
        import torch
        labels = logits.argmax(dim=1).to(torch.int32).contiguous()
        self.add(cell_ids, labels)

    def reduce(self, group=None):
        """Sum the vote tables of all ranks (no-op in a single process)."""
        self.votes = fdist.allreduce_votes(self.votes, ctx=self.ctx, group=group)
        return self.votes

    def cell_labels(self, default=0):
        """int32 [n_cells] argmax label (lowest label wins ties), ``default`` where no pixel voted."""
        out = self.ctx.votes_argmax(self.votes, default=default,
                                    out=self._empty(self.n_cells) if self.use_torch else None)
        if self.use_torch:
            self.ctx.synchronize()  # the result is handed to torch code, which runs on another stream
        return out

    def point_labels(self, default=0):
        """int32 [n_points]: every cell hands its label to its first point (highest cell id wins)."""
        cl = self.cell_labels(default)
        out = self.ctx.cells_to_points(cl, default=default,
                                       out=self._empty(self.ctx.n_verts) if self.use_torch else None)
        if self.use_torch:
            self.ctx.synchronize()
        return out

    def _empty(self, n):
        import torch
        return torch.empty(n, dtype=torch.int32, device=self.votes.device)


def label_mesh_cells(ctx, model, n_classes, res, opts, default=0, batch_views=16, group=None):
    """BASELINE config 4 driver: render this rank's share of the viewpoints, run ``model`` (a torch module
    mapping [n, C, H, W] features to [n, K, H, W] logits) on them, scatter the per-pixel argmax onto the
    cells, all-reduce the votes and return (cell_labels, votes) as torch tensors on the GPU.

    The mesh must be set and built, and the viewpoints set, on ``ctx`` by the caller (every rank holds the
    whole mesh + LBVH; views are sharded contiguously)."""
    import torch
    rank, ws = fdist.world()
    vb, ve = fdist.shard_range(ctx.num_views(), rank, ws)
    dev = torch.device("cuda", ctx.device)
    C = ctx.channels(opts)
    voter = CellVoter(ctx, n_classes, use_torch=True)
    feat = torch.empty((batch_views, res, res, C), dtype=torch.float32, device=dev)
    ids = torch.empty((batch_views, res, res), dtype=torch.int32, device=dev)
    model.eval()
    with torch.no_grad():
        for b in range(vb, ve, batch_views):
            e = min(b + batch_views, ve)
            n = e - b
            ctx.render_into(res, opts, feat[:n], ids[:n], view_begin=b, view_end=e)
            ctx.synchronize()  # the model runs on torch's stream
            logits = model(feat[:n].permute(0, 3, 1, 2).contiguous())
            torch.cuda.current_stream().synchronize()
            voter.add_logits(ids[:n], logits)
    voter.reduce(group)
    return voter.cell_labels(default), voter.votes
