"""Deterministic synthetic surfaces for tests and benchmarks.

All meshes are closed, outward-wound triangle surfaces made by *linear*
subdivision of an icosahedron to ``Lm`` segments per edge, giving
``T = 20*Lm**2`` triangles and ``V = 10*Lm**2 + 2`` vertices, then displaced.
They stand in for the dental / cortical surfaces the reference processes
(there is no network here to fetch real data).
"""
import numpy as np

# icosahedron of vtkPlatonicSolidSource (restated; see oracle/flyby_oracle.c O3)
ICO_POINTS = np.array([
    [0.0, 0.0, -1.0], [0.0, 0.894427, -0.447214], [0.850651, 0.276393, -0.447214],
    [0.525731, -0.723607, -0.447214], [-0.525731, -0.723607, -0.447214],
    [-0.850651, 0.276393, -0.447214], [0.0, -0.894427, 0.447214],
    [0.850651, -0.276393, 0.447214], [0.525731, 0.723607, 0.447214],
    [-0.525731, 0.723607, 0.447214], [-0.850651, -0.276393, 0.447214], [0.0, 0.0, 1.0]],
    dtype=np.float64)
ICO_FACES = np.array([
    [0, 5, 4], [0, 4, 3], [0, 3, 2], [0, 2, 1], [0, 1, 5], [5, 1, 9], [5, 9, 10], [5, 10, 4],
    [4, 10, 6], [4, 6, 3], [3, 6, 7], [3, 7, 2], [2, 7, 8], [2, 8, 1], [1, 8, 9], [9, 8, 11],
    [9, 11, 10], [10, 11, 6], [6, 11, 7], [7, 11, 8]], dtype=np.int64)


def icosphere(Lm):
    """Unit sphere by linear subdivision: returns (verts float64 [V,3], tris int32 [T,3])."""
    Lm = int(Lm)
    # lattice (i, j) with i + j <= Lm, row-major in j
    ij = np.array([(i, j) for j in range(Lm + 1) for i in range(Lm + 1 - j)], dtype=np.int64)
    index = -np.ones((Lm + 1, Lm + 1), dtype=np.int64)
    index[ij[:, 0], ij[:, 1]] = np.arange(len(ij))
    up = np.array([(i, j) for j in range(Lm) for i in range(Lm - j)], dtype=np.int64)
    tri_up = np.stack([index[up[:, 0], up[:, 1]], index[up[:, 0] + 1, up[:, 1]],
                       index[up[:, 0], up[:, 1] + 1]], axis=1)
    dn = np.array([(i, j) for j in range(Lm - 1) for i in range(Lm - 1 - j)], dtype=np.int64)
    if len(dn):
        tri_dn = np.stack([index[dn[:, 0] + 1, dn[:, 1]], index[dn[:, 0] + 1, dn[:, 1] + 1],
                           index[dn[:, 0], dn[:, 1] + 1]], axis=1)
        local = np.concatenate([tri_up, tri_dn], axis=0)
    else:
        local = tri_up
    w1 = ij[:, 0:1] / Lm
    w2 = ij[:, 1:2] / Lm
    pts, tris = [], []
    for f, (a, b, c) in enumerate(ICO_FACES):
        A, B, C = ICO_POINTS[a], ICO_POINTS[b], ICO_POINTS[c]
        pts.append(A + w1 * (B - A) + w2 * (C - A))
        tris.append(local + f * len(ij))
    P = np.concatenate(pts, axis=0)
    F = np.concatenate(tris, axis=0)
    P /= np.linalg.norm(P, axis=1, keepdims=True)
    # merge lattice points shared between faces (spacing >> 1e-7 for Lm <= 1000)
    key = np.round(P * 1e7).astype(np.int64)
    _, first, inverse = np.unique(key, axis=0, return_index=True, return_inverse=True)
    inverse = inverse.reshape(-1)
    V = P[first]
    F = inverse[F]
    # make the winding outward
    n = np.cross(V[F[:, 1]] - V[F[:, 0]], V[F[:, 2]] - V[F[:, 0]])
    cen = V[F].mean(axis=1)
    flip = np.einsum('ij,ij->i', n, cen) < 0
    F[flip] = F[flip][:, [0, 2, 1]]
    assert len(V) == 10 * Lm * Lm + 2 and len(F) == 20 * Lm * Lm
    return V, F.astype(np.int32)


def bumpy_sphere(Lm=71, seed=0, amplitude=0.15):
    """C1/C3/C5 family: r = 1 + amplitude * sum_k a_k sin(f_k u_k + phi_k)."""
    V, F = icosphere(Lm)
    rng = np.random.default_rng(seed)
    dirs = rng.normal(size=(6, 3))
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    f = rng.uniform(2.0, 9.0, size=6)
    a = rng.uniform(0.3, 1.0, size=6)
    phi = rng.uniform(0.0, 2 * np.pi, size=6)
    u = V @ dirs.T
    r = 1.0 + amplitude * np.sum(a * np.sin(f * u + phi), axis=1)
    # arbitrary (non-unit) placement so that unit-surf normalisation has work to do
    centre = rng.uniform(-5.0, 5.0, size=3)
    scale = rng.uniform(10.0, 40.0)
    P = (V * r[:, None]) * scale + centre
    return P.astype(np.float32), F


def dental_crown(Lm=100, seed=1):
    """C2: squashed sphere bent into a parabolic arch with 14 Gaussian cusps."""
    V, F = icosphere(Lm)
    rng = np.random.default_rng(seed)
    P = V.copy()
    cusp_x = np.linspace(-0.85, 0.85, 14)
    cusp = np.stack([cusp_x, np.zeros(14), np.full(14, 1.0)], axis=1)
    cusp /= np.linalg.norm(cusp, axis=1, keepdims=True)
    h = rng.uniform(0.15, 0.25, size=14)
    d2 = ((V[:, None, :] - cusp[None, :, :]) ** 2).sum(-1)
    bump = (h * np.exp(-d2 / (2 * 0.08 ** 2))).sum(axis=1)
    P = P * (1.0 + bump)[:, None]
    P[:, 2] *= 0.35
    P[:, 1] *= 0.45
    P[:, 1] += 0.6 * P[:, 0] ** 2
    centre = rng.uniform(-20.0, 20.0, size=3)
    P = P * 30.0 + centre
    return P.astype(np.float32), F


def cube():
    """12-triangle axis-aligned cube [-1,1]^3, outward wound (KAT mesh)."""
    V = np.array([[x, y, z] for z in (-1, 1) for y in (-1, 1) for x in (-1, 1)], dtype=np.float32)
    q = [(0, 2, 3, 1), (4, 5, 7, 6), (0, 1, 5, 4), (2, 6, 7, 3), (0, 4, 6, 2), (1, 3, 7, 5)]
    F = []
    for a, b, c, d in q:
        F += [(a, b, c), (a, c, d)]
    return V, np.array(F, dtype=np.int32)
