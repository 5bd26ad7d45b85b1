"""The fly-by renderer as an autograd-free torch op with the signature the PyTorch3D-era models of the
reference use (src/py/challenge-teeth/teeth_nets.py:121-153 `render(V, F, CN) -> X, PF`, label gather
`torch.take(YF, PF)` at :166).

    renderer = FlyByRenderer(radius=1.35, subdivision_level=2, image_size=224)
    X, PF = renderer.render(V, F, CN)        # V [B,Nv,3] f32, F [B,Nf,3] int, CN [B,Nv,Cf] f32 (cuda)
    # X  [B, T, Cf + 1, H, W]  per-view images: CN interpolated barycentrically at the hit + depth
    # PF [B, T, 1, H, W] int64 hit face, numbered like PyTorch3D's packed pix_to_face (b * Nf + face), -1 = miss
    y = torch.take(YF, PF) * (PF >= 0)

Differences from the PyTorch3D rasteriser it replaces (SURVEY hazard H4): rays are orthographic from the plane
tangent to the view sphere (the ray-cast engine of fly_by_features.cxx), not a 60-degree pinhole; depth is the
distance from that plane (equal to PyTorch3D's view-space z of the same surface point); background pixels are
`background` in the colour channels and -1 in depth, as HardPhongShader / zbuf leave them.  Everything stays on
the device: the tensors' storage is handed to the C-ABI as is and all kernels run on the current torch stream.
"""
import numpy as np
import torch

from . import _cabi


class FlyByRenderer:
    def __init__(self, radius=1.35, subdivision_level=2, image_size=224, plane_scale=2.0, background=1.0,
                 device=0, spiral=None, turns=4):
        self.device = torch.device("cuda", device)
        self.radius, self.res, self.background = float(radius), int(image_size), float(background)
        with torch.cuda.device(self.device):
            self.stream = torch.cuda.current_stream(self.device)
            handle = self.stream.cuda_stream
            if handle == 0:  # the legacy default stream cannot be passed as "use this stream": make one
                self.stream = torch.cuda.Stream(self.device)
                handle = self.stream.cuda_stream
            self.ctx = _cabi.Context(device, stream=handle)
        if spiral:
            self.n_views = self.ctx.views_spiral(int(spiral), turns, self.radius)
        else:
            self.n_views = self.ctx.views_icosahedron(int(subdivision_level), self.radius)
        # use_z / split_z give [normals, depth]; only the depth channel and the ids of this render are used
        self.opts = self.ctx.render_opts(use_z=1, split_z=1, plane_scale=plane_scale)

    @torch.no_grad()
    def render(self, V, F, CN):
        B, T, H = V.shape[0], self.n_views, self.res
        Cf = CN.shape[-1]
        dev = self.device
        X = torch.empty((B, T, Cf + 1, H, H), dtype=torch.float32, device=dev)
        PF = torch.empty((B, T, 1, H, H), dtype=torch.int64, device=dev)
        cur = torch.cuda.current_stream(dev)
        self.stream.wait_stream(cur)
        with torch.cuda.stream(self.stream):
            ids = torch.empty((T, H, H), dtype=torch.int32, device=dev)
            tt = torch.empty((T, H, H), dtype=torch.float64, device=dev)
            col = torch.empty((T, H, H, Cf), dtype=torch.float32, device=dev)
            for b in range(B):
                verts = V[b].to(torch.float32).contiguous()
                faces = F[b].to(torch.int32).contiguous()
                feats = CN[b].to(torch.float32).contiguous()
                self.ctx.set_mesh(verts, faces, skip_normalise=True)   # the callers normalise their meshes themselves
                self.ctx.build()
                self.ctx.render_into(H, self.opts, None, ids, tt)
                self.ctx.interpolate_features(H, self.opts, ids, feats, zero=self.background, out=col)
                hit = ids >= 0
                depth = torch.where(hit, (tt * (2.0 * self.radius)).to(torch.float32), torch.full_like(col[..., 0], -1.0))
                X[b, :, :Cf] = col.permute(0, 3, 1, 2)
                X[b, :, Cf] = depth
                PF[b, :, 0] = torch.where(hit, ids.to(torch.int64) + b * F.shape[1], torch.full_like(ids, -1, dtype=torch.int64))
        cur.wait_stream(self.stream)
        return X, PF

    def __call__(self, V, F, CN):
        return self.render(V, F, CN)

    def close(self):
        self.ctx.close()


def gather_face_labels(YF, PF):
    """teeth_nets.py:166: y = torch.take(YF, PF) * (PF >= 0) without reading element -1 for the background."""
    safe = PF.clamp_min(0)
    return torch.take(YF, safe).to(torch.int64) * (PF >= 0)
