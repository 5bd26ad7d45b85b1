"""CPU tests of the mesh-label post-processing oracle (oracle/postproc_oracle.c, a restatement of the
reference's src/py/post_process.py:243-380) on hand-checkable cases, and of the host-side Threshold."""
import numpy as np

from flyby_b200 import synth


def grid_mesh(n):
    """n x n points, pid = j * n + i, every square split along the (i, j) -> (i+1, j+1) diagonal:
    interior points have 6 neighbours: (+-1, 0), (0, +-1), (+1, +1), (-1, -1)."""
    ii, jj = np.meshgrid(np.arange(n), np.arange(n), indexing="xy")
    P = np.stack([ii.ravel(), jj.ravel(), np.zeros(n * n)], 1).astype(np.float32)
    F = []
    for j in range(n - 1):
        for i in range(n - 1):
            a, b, c, d = j * n + i, j * n + i + 1, (j + 1) * n + i + 1, (j + 1) * n + i
            F += [[a, b, c], [a, c, d]]
    return P, np.array(F, np.int32)


def test_connectivity_orders_regions_by_lowest_point_id(orc):
    n = 8
    P, F = grid_mesh(n)
    lab = np.zeros(n * n, np.int32)
    lab[[n * 6 + 5, n * 6 + 6]] = 2            # region found second (lowest pid 53)
    lab[[n * 1 + 1, n * 1 + 2, n * 2 + 2]] = 2  # region found first (lowest pid 9)
    lab[n * 4 + 0] = 2                          # single point, lowest pid 32
    out, k = orc.pp_connectivity(F, n * n, lab, 2, 2)
    assert k == 3
    assert out[n + 1] == out[n + 2] == out[2 * n + 2] == 2 and out[4 * n] == 3 and out[6 * n + 5] == out[6 * n + 6] == 4
    assert (out[lab == 0] == 0).all()
    # (1,1) and (2,2) are joined by the diagonal, (1,2) and (2,1) are not: anti-diagonal points stay apart
    lab2 = np.zeros(n * n, np.int32)
    lab2[[n * 2 + 1, n * 1 + 2]] = 5
    out2, k2 = orc.pp_connectivity(F, n * n, lab2, 5, 10)
    assert k2 == 2 and out2[n * 1 + 2] == 10 and out2[n * 2 + 1] == 11


def test_remove_islands_majority_and_reference_quirk(orc):
    n = 9
    P, F = grid_mesh(n)
    lab = np.full(n * n, 1, np.int32)
    lab[: n * 4] = 3                       # rows 0-3 label 3, rows 4-8 label 1
    isl = [n * 5 + 4, n * 5 + 5]           # a two-point island of label 7 inside label 1
    lab[isl] = 7
    # without ignore_neg1 the reference's RemoveIslands changes nothing (post_process.py:292)
    assert np.array_equal(orc.pp_remove_islands(F, n * n, lab, 7, 10, ignore_neg1=False), lab)
    out = orc.pp_remove_islands(F, n * n, lab, 7, 10, ignore_neg1=True)
    assert (out[isl] == 1).all() and (np.delete(out, isl) == np.delete(lab, isl)).all()
    # big regions are left alone
    assert np.array_equal(orc.pp_remove_islands(F, n * n, lab, 1, 10, ignore_neg1=True), lab)
    # an island on the border of labels 3 and 1: (4,4) touches row 3 at (3,3)? neighbours of (i=4, j=4):
    # (3,4) (5,4) (4,3) (4,5) (5,5) (3,3) -> labels 1 1 3 1 1 3: majority 1
    lab2 = np.full(n * n, 1, np.int32); lab2[: n * 4] = 3; lab2[n * 4 + 4] = 7
    assert orc.pp_remove_islands(F, n * n, lab2, 7, 5, True)[n * 4 + 4] == 1
    # tie (3 neighbours each): the label met first in ascending point id wins -> the row below (label 3)
    lab3 = np.full(n * n, 1, np.int32); lab3[: n * 4] = 3; lab3[n * 4 + 4] = 7; lab3[n * 4 + 3] = 3
    # neighbours of pid 40: 39->3, 41->1, 31->3, 49->1, 50->1, 30->3: 3 vs 3, lowest pid 30 has label 3
    assert orc.pp_remove_islands(F, n * n, lab3, 7, 5, True)[n * 4 + 4] == 3
    # a neighbour label of -1 is never taken
    lab4 = np.full(n * n, -1, np.int32); lab4[n * 4 + 4] = 7
    assert orc.pp_remove_islands(F, n * n, lab4, 7, 5, True)[n * 4 + 4] == 7


def test_dilate_and_erode_sweeps(orc):
    n = 9
    P, F = grid_mesh(n)
    lab = np.zeros(n * n, np.int32)
    c = n * 4 + 4
    lab[c] = 5
    d1 = orc.pp_dilate(F, n * n, lab, 5, iterations=1)
    six = sorted([c - 1, c + 1, c - n, c + n, c + n + 1, c - n - 1])
    assert sorted(np.nonzero(d1 == 5)[0].tolist()) == sorted(six + [c])
    d2 = orc.pp_dilate(F, n * n, lab, 5, iterations=2)
    assert (d2 == 5).sum() == 19
    # dilateOverTarget only grows into `target`
    lab_t = lab.copy(); lab_t[c + 1] = 9
    dt = orc.pp_dilate(F, n * n, lab_t, 5, iterations=1, over_target=True, target=9)
    assert sorted(np.nonzero(dt == 5)[0].tolist()) == [c, c + 1]
    # erode: one sweep peels the rim of the 19-point blob, points take their lowest-id differing neighbour
    e1 = orc.pp_erode(F, n * n, d2, 5, iterations=1)
    assert (e1 == 5).sum() == 7 and sorted(np.nonzero(e1 == 5)[0].tolist()) == sorted(six + [c])
    e_all = orc.pp_erode(F, n * n, d2, 5)
    assert (e_all == 5).sum() == 0 and (e_all == 0).all()
    # ignore_label: a blob surrounded only by the ignored label cannot erode (loop ends without a change)
    assert np.array_equal(orc.pp_erode(F, n * n, d2, 5, ignore_label=0), d2)
    # the first qualifying neighbour in ascending id decides: point c sees c-n-1 (label 1) before c-1 (label 2)
    lab5 = np.full(n * n, 2, np.int32); lab5[c] = 5; lab5[c - n - 1] = 1
    assert orc.pp_erode(F, n * n, lab5, 5, iterations=1)[c] == 1
    assert orc.pp_erode(F, n * n, lab5, 5, iterations=1, target=2)[c] == 2
    assert np.array_equal(orc.pp_relabel(lab5, 2, -1), np.where(lab5 == 2, -1, lab5))


def test_threshold_matches_vtk_all_points_rule():
    from flyby_b200 import post_process, utils
    P, F = grid_mesh(4)
    lab = np.zeros(16, np.int32); lab[[5, 6, 9, 10, 11, 14, 15]] = 4
    s = utils.Surf(P, F, {"ids": lab, "x": P[:, 0].copy()}, {})
    t = post_process.Threshold(s, "ids", 3.5, 4.5)
    kept = [f for f in F.tolist() if all(lab[v] == 4 for v in f)]
    assert len(t.polys) == len(kept) and np.array_equal(t.points[t.polys], P[np.array(kept)])
    assert (t.point_data["ids"] == 4).all() and np.array_equal(t.point_data["x"], t.points[:, 0])
    inv = post_process.Threshold(s, lab, 3.5, 4.5, invert=True)
    assert len(inv.polys) == len(F) - len(kept)
    post_process.ReLabel(s, "ids", 4, -1)
    assert (s.point_data["ids"] == np.where(lab == 4, -1, lab)).all()


def test_oracle_region_sizes_on_a_real_mesh(orc):
    """Regions found by the oracle == connected components by union-find over the same label (numpy check)."""
    P, F = synth.bumpy_sphere(10, 0)
    rng = np.random.default_rng(0)
    lab = (rng.random(len(P)) < 0.55).astype(np.int32) * 2
    out, k = orc.pp_connectivity(F, len(P), lab, 2, 100)
    parent = np.arange(len(P))

    def find(x):
        while parent[x] != x:
            parent[x] = parent[parent[x]]
            x = parent[x]
        return x
    for a, b, c in F:
        for u, v in ((a, b), (b, c), (a, c)):
            if lab[u] == 2 and lab[v] == 2:
                ru, rv = find(u), find(v)
                if ru != rv:
                    parent[max(ru, rv)] = min(ru, rv)
    roots = sorted({find(v) for v in range(len(P)) if lab[v] == 2})
    assert k == len(roots)
    for v in range(len(P)):
        if lab[v] == 2:
            assert out[v] == 100 + roots.index(find(v))
