"""vtk.util.numpy_support of the container-only vtk stand-in (tests/refshim/vtk)."""
import numpy as np

from .. import _array_for


def vtk_to_numpy(arr):
    # real VTK returns a view shaped (n,) for one component, (n, ncomp) otherwise
    return arr.data[:, 0] if arr.ncomp == 1 else arr.data


def numpy_to_vtk(num_array, deep=0, array_type=None):
    a = np.ascontiguousarray(num_array)
    if a.dtype == np.bool_:
        a = a.astype(np.uint8)
    return _array_for(a.copy() if deep else a)
