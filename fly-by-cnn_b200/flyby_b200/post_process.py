"""Mesh-label post-processing with the reference's names and argument meaning
(reference src/py/post_process.py; callers predict_v3.py:93-133, dental_model_seg.py:228-239, predict.py:190-201).

`vtkdata` is a `utils.Surf`; `labels` is an int32 numpy array of one value per point (the reference passes a
vtk array and edits it in place -- so do these functions), or the name of an entry of `vtkdata.point_data`.
The work runs on the GPU through the C-ABI (`flyby_labels_*`, csrc/postproc.cu); the mesh is uploaded once
per surface and reused by consecutive calls, as the predictors call several of these in a row.
"""
import math

import numpy as np

from . import utils

_BOUND = {}  # device -> (id of the Surf whose adjacency the context holds, n_points, n_cells)


def _ctx_for(vtkdata, device=0):
    ctx = utils.default_context(device)
    key = (id(vtkdata), len(vtkdata.points), len(vtkdata.polys))
    if _BOUND.get(device) != key or ctx.n_verts != len(vtkdata.points):
        ctx.set_mesh(np.ascontiguousarray(vtkdata.points, np.float32), np.ascontiguousarray(vtkdata.polys, np.int32),
                     skip_normalise=True)
        ctx.build()
        _BOUND[device] = key
    return ctx


def _labels(vtkdata, labels):
    if isinstance(labels, str):
        labels = vtkdata.point_data[labels]
    if not isinstance(labels, np.ndarray) or labels.dtype != np.int32 or labels.ndim != 1 or not labels.flags.c_contiguous:
        raise TypeError("labels must be a contiguous one-dimensional int32 numpy array (it is edited in place)")
    if len(labels) != len(vtkdata.points):
        raise ValueError("labels has %d entries for %d points" % (len(labels), len(vtkdata.points)))
    return labels


def RemoveIslands(vtkdata, labels, label, min_count, ignore_neg1=False, device=0):
    """post_process.py:285-295.  NB: like the reference, this relabels nothing unless ignore_neg1 is True."""
    lab = _labels(vtkdata, labels)
    _ctx_for(vtkdata, device).labels_remove_islands(lab, int(label), int(min_count), bool(ignore_neg1))


def ConnectivityLabeling(vtkdata, labels, label, start_label, device=0):
    """post_process.py:297-304."""
    lab = _labels(vtkdata, labels)
    return _ctx_for(vtkdata, device).labels_connectivity(lab, int(label), int(start_label))


def ErodeLabel(vtkdata, labels, label, ignore_label=None, iterations=math.inf, target=None, device=0):
    """post_process.py:306-346."""
    lab = _labels(vtkdata, labels)
    it = None if iterations == math.inf else int(iterations)
    _ctx_for(vtkdata, device).labels_erode(lab, int(label), ignore_label, it, target)


def DilateLabel(vtkdata, labels, label, iterations=2, dilateOverTarget=False, target=None, device=0):
    """post_process.py:348-375."""
    lab = _labels(vtkdata, labels)
    _ctx_for(vtkdata, device).labels_dilate(lab, int(label), int(iterations), bool(dilateOverTarget), target)


def ReLabel(surf, labels, label, relabel):
    """post_process.py:377-380 (one vector operation: done on the host array directly)."""
    lab = _labels(surf, labels)
    lab[lab == label] = relabel


def Threshold(vtkdata, labels, threshold_min, threshold_max, invert=False):
    """post_process.py:382-398: vtkThreshold on a point array (a cell passes when all its points are inside
    [min, max]; `invert` negates the decision per cell) followed by vtkGeometryFilter.  Returns a new Surf
    with the kept cells, the points they use in their original order, and the point data restricted to them."""
    lab = vtkdata.point_data[labels] if isinstance(labels, str) else np.asarray(labels)
    inside = (lab >= threshold_min) & (lab <= threshold_max)
    polys = np.asarray(vtkdata.polys)
    keep = inside[polys].all(axis=1)
    if invert:
        keep = ~keep
    kept = polys[keep]
    used = np.zeros(len(vtkdata.points), bool)
    used[kept.reshape(-1)] = True
    remap = np.cumsum(used) - 1
    out = utils.Surf(np.asarray(vtkdata.points)[used].copy(), remap[kept].astype(np.int32),
                     {k: np.asarray(v)[used].copy() for k, v in vtkdata.point_data.items()},
                     {k: np.asarray(v)[keep].copy() for k, v in vtkdata.cell_data.items() if len(v) == len(polys)})
    return out
