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


if __name__ == "__main__":
    main()
