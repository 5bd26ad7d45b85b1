"""Regenerates tests/golden/flyby_golden.npz from the CPU oracle.

The reference cannot run here (it needs VTK, which is neither vendored nor
installable), and it ships no golden vectors of its own, so these fixtures pin
the ORACLE's behaviour (brute-force nearest hit, no acceleration structure) as a
regression anchor: parity with the reference itself stays "unpinned" (DESIGN.md).
  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "fly-by-cnn_b200"))
from oracle import oracle as orc  # noqa: E402
from flyby_b200 import synth  # noqa: E402


def main():
    P, F = synth.bumpy_sphere(6, seed=42)          # 720 triangles
    U, centre, scale = orc.unit_surf(P)
    N = orc.point_normals(U, F)
    ico = orc.icosahedron(2, 2.0)
    spiral = orc.spiral(8, 4, 4.0)
    res = 24
    f4, ids, t = orc.render(U, F, N, ico, 2.0, res, use_z=1, split_z=1, grid=False, want_t=True)
    f3, ids_s = orc.render(U, F, N, spiral, 4.0, res, plane_scale=2.0, use_z=1, split_z=0, grid=False)
    rays, basis = orc.view_rays(ico[5], 2.0, 5)
    labels = ((np.arange(ids.size).reshape(ids.shape) * 7) % 5).astype(np.int32)
    votes = orc.backproject(ids, labels, len(F), 5)
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "flyby_golden.npz"),
                        P=P, F=F, U=U, centre=centre, scale=scale, N=N, ico=ico, spiral=spiral,
                        feat4=f4, ids=ids, t=t, feat3_spiral=f3, ids_spiral=ids_s, rays_view5=rays,
                        basis_view5=basis, labels=labels, votes=votes,
                        cell_labels=orc.votes_argmax(votes, default=33))
    # ---- rows added after the hot path: label post-processing, barycentric lookup, pinhole views
    rng = np.random.default_rng(7)
    lab = np.argmax(U @ rng.normal(size=(4, 3)).T, axis=1).astype(np.int32)
    flip = rng.random(len(U)) < 0.1
    lab[flip] = rng.integers(-1, 5, flip.sum()).astype(np.int32)
    V = len(U)
    conn, n_regions = orc.pp_connectivity(F, V, lab, 2, 10)
    feats = rng.normal(size=(V, 3)).astype(np.float32)
    pf, pi, pd = orc.render_perspective(U, F, N, ico[:6], 20, 30.0, 0.01, 3.0, normal_encoding=3, grid=False)
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "flyby_golden_ext.npz"),
                        labels_in=lab, islands=orc.pp_remove_islands(F, V, lab, 1, 30, True),
                        connectivity=conn, n_regions=n_regions, dilate=orc.pp_dilate(F, V, lab, 3, 2),
                        erode=orc.pp_erode(F, V, lab, 0, ignore_label=1), erode2=orc.pp_erode(F, V, lab, 2, iterations=2, target=3),
                        feats=feats, interp=orc.interpolate_features(U, F, ico[:8], 2.0, res, ids[:8], feats, zero=-1.0),
                        persp_feat=pf, persp_ids=pi, persp_depth=pd)


if __name__ == "__main__":
    main()
