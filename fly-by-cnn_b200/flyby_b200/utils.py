"""Host-side mirror of the reference's geometry helpers (reference src/py/utils.py).

Same function names and argument meaning as the reference so that callers of
``fly_by_features`` (``fbf.ReadSurf``, ``fbf.GetUnitSurf``, ``fbf.CreateIcosahedron`` ...,
see reference src/py/predict_v3.py:23-36) keep working.  vtk / itk objects are
replaced by small numpy containers; every numeric step (unit surface, normals,
viewpoints, feature lookup) runs in the CUDA library through the C-ABI -- there
is no CPU implementation behind these functions.
"""
import math

import numpy as np

from . import nrrd, readers
from ._cabi import Context, FlyByError

_CTX = {}


def default_context(device=0):
    """One shared GPU context per device for the free functions below."""
    if device not in _CTX:
        _CTX[device] = Context(device)
    return _CTX[device]


class Surf(readers.MeshData):
    """Stand-in for the vtkPolyData the reference passes around."""

    is_unit = False  # True once GetUnitSurf / ScaleSurf has normalised the points

    def GetNumberOfPoints(self):
        return len(self.points)

    def GetNumberOfCells(self):
        return len(self.polys)

    def GetPoint(self, i):
        return tuple(float(x) for x in self.points[i])

    def GetPoints(self):
        return self.points

    def GetBounds(self):
        lo, hi = self.points.min(0), self.points.max(0)
        return (float(lo[0]), float(hi[0]), float(lo[1]), float(hi[1]), float(lo[2]), float(hi[2]))

    def copy(self):
        s = Surf(self.points.copy(), self.polys.copy(), {k: np.array(v) for k, v in self.point_data.items()},
                 {k: np.array(v) for k, v in self.cell_data.items()})
        s.is_unit = self.is_unit
        return s


class ViewSphere:
    """The viewpoint set (reference: a vtkPolyData whose points are the camera positions)."""

    def __init__(self, points, radius, kind, param):
        self.points = np.ascontiguousarray(points, dtype=np.float32).reshape(-1, 3)
        self.radius = float(radius)
        self.kind = kind      # "icosahedron" | "spiral" | "custom"
        self.param = param

    def GetNumberOfPoints(self):
        return len(self.points)

    def GetPoints(self):
        return self.points

    def GetPoint(self, i):
        return tuple(float(x) for x in self.points[i])


def normalize_vector(x):
    return x / np.linalg.norm(x)


# --------------------------------------------------------------------------- I/O
def ReadSurf(fileName):
    """reference src/py/utils.py:112-194"""
    m = readers.read_mesh(fileName)
    return Surf(m.points, m.polys, m.point_data, m.cell_data)


def WriteSurf(surf, fileName, binary=False):
    """reference src/py/utils.py:196-209 (vtkPolyDataWriter)"""
    readers.write_vtk(fileName, surf.points, surf.polys, surf.point_data, surf.cell_data, binary=binary)


def GetImage(img_np):
    """numpy [.., C] -> vector image (reference src/py/utils.py:704-748 builds an itk.VectorImage[float])."""
    return nrrd.Image(np.asarray(img_np, dtype=np.float32))


def WriteImage(img, fileName, compress=True):
    """itk.ImageFileWriter(...).UseCompressionOn().Update() (reference src/py/fly_by_features.py:419-424)."""
    nrrd.write(fileName, img, compress=compress)


def ReadImage(fName):
    arr, _ = nrrd.read(fName)
    return nrrd.Image(arr)


# --------------------------------------------------------------------------- viewpoints
def CreateIcosahedron(radius, sl=1, dedupe=1, device=0):
    """reference src/py/utils.py:36-50 (+ LinearSubdivisionFilter.py:42-67, utils.py:22-31).

    dedupe=1 (default) reproduces the reference's point list bit for bit, including the duplicate viewpoints its
    float-equality point locator lets through (94 instead of 92 points at sl=3; tests/golden/ref_views.npz was
    computed by the reference's own LinearSubdivisionFilter).  dedupe=0 is the exact lattice (10 sl^2 + 2)."""
    if sl < 1:
        raise ValueError("number of subdivisions must be >= 1")
    ctx = default_context(device)
    ctx.views_icosahedron(int(sl), float(radius), dedupe)
    return ViewSphere(ctx.get_views(), radius, "icosahedron", int(sl))


def CreateSpiral(sphereRadius=4, numberOfSpiralSamples=64, numberOfSpiralTurns=4, device=0):
    """reference src/py/utils.py:52-89"""
    ctx = default_context(device)
    ctx.views_spiral(int(numberOfSpiralSamples), float(numberOfSpiralTurns), float(sphereRadius))
    return ViewSphere(ctx.get_views(), sphereRadius, "spiral", (int(numberOfSpiralSamples), float(numberOfSpiralTurns)))


# --------------------------------------------------------------------------- normalisation
def ScaleSurf(surf, mean_arr=None, scale_factor=None, copy=True, device=0, centre_mode=0, scale_mode=0):
    """reference src/py/utils.py:249-290: centre on the bounding-box middle (or ``mean_arr``), scale by
    1/|bbox_max - centre| (or ``scale_factor``).  Returns (surf, mean_arr, scale_factor).
    ``centre_mode=1`` / ``scale_mode=1`` are the C++ tool's --useCenterOfMass / --useMagnitude
    (src/app/fly_by_features.cxx:397-403, 429-435): centre of mass of the points, 1 / largest point magnitude."""
    out = surf.copy() if copy else surf
    ctx = default_context(device)
    ctx.set_mesh(np.ascontiguousarray(surf.points, np.float32), np.ascontiguousarray(surf.polys, np.int32),
                 centre_mode=centre_mode, scale_mode=scale_mode, translate=mean_arr, scale_factor=scale_factor)
    ctx.build()
    unit, normals, centre, scale = ctx.get_mesh()
    out.points = unit
    out.point_data["Normals"] = normals
    out.is_unit = True
    return out, centre, scale


def GetUnitSurf(surf, mean_arr=None, scale_factor=None, copy=True, device=0, centre_mode=0, scale_mode=0):
    """reference src/py/utils.py:341-343"""
    unit_surf, _, _ = ScaleSurf(surf, mean_arr, scale_factor, copy, device, centre_mode, scale_mode)
    return unit_surf


def ComputeNormals(surf, device=0):
    """reference src/py/utils.py:493-501 (vtkPolyDataNormals, splitting off): adds point data "Normals"."""
    out = surf.copy()
    ctx = default_context(device)
    ctx.set_mesh(np.ascontiguousarray(surf.points, np.float32), np.ascontiguousarray(surf.polys, np.int32),
                 skip_normalise=True)
    ctx.build()
    _, normals, _, _ = ctx.get_mesh()
    out.point_data["Normals"] = normals
    return out


# --------------------------------------------------------------------------- rotations (a13)
def GetTransform(rotationAngle, rotationVector):
    """vtkTransform.RotateWXYZ(angle_degrees, axis) as a 4x4 matrix (reference src/py/utils.py:301-304)."""
    axis = np.asarray(rotationVector, dtype=np.float64)
    n = np.linalg.norm(axis)
    m = np.eye(4)
    if n == 0 or rotationAngle == 0:
        return m
    x, y, z = axis / n
    a = math.radians(rotationAngle)
    c, s = math.cos(a), math.sin(a)
    C = 1 - c
    m[:3, :3] = [[c + x * x * C, x * y * C - z * s, x * z * C + y * s],
                 [y * x * C + z * s, c + y * y * C, y * z * C - x * s],
                 [z * x * C - y * s, z * y * C + x * s, c + z * z * C]]
    return m


def RotateTransform(surf, transform):
    out = surf.copy()
    m = np.asarray(transform, dtype=np.float64)
    out.points = (surf.points.astype(np.float64) @ m[:3, :3].T + m[:3, 3]).astype(np.float32)
    out.point_data.pop("Normals", None)
    return out


def RotateSurf(surf, rotationAngle, rotationVector):
    """reference src/py/utils.py:306-308"""
    return RotateTransform(surf, GetTransform(rotationAngle, rotationVector))


def RotateInverse(surf, rotationAngle, rotationVector):
    """reference src/py/utils.py:310-319"""
    return RotateTransform(surf, np.linalg.inv(GetTransform(rotationAngle, rotationVector)))


def RandomRotation(surf):
    """reference src/py/utils.py:334-338"""
    rotationAngle = np.random.random() * 360.0
    rotationVector = np.random.random(3) * 2.0 - 1.0
    rotationVector = rotationVector / np.linalg.norm(rotationVector)
    return RotateSurf(surf, rotationAngle, rotationVector), rotationAngle, rotationVector


# --------------------------------------------------------------------------- actors
class Actor:
    """What FlyByGenerator renders: a mesh plus which feature to write per pixel."""

    def __init__(self, surf, kind):
        self.surf = surf
        self.kind = kind  # "normals" | "point_id" | "cell_id"


def GetNormalsActor(surf):
    """reference src/py/utils.py:516-578 (normals coded as colours by a shader)."""
    return Actor(surf, "normals")


def GetPointIdMapActor(surf):
    """reference src/py/utils.py:623-637: every cell coloured with the id of its first point."""
    return Actor(surf, "point_id")


def GetCellIdMapActor(surf):
    """reference src/py/utils.py:580-602"""
    return Actor(surf, "cell_id")


def EncodeIdsRGB(ids):
    """id -> (r, g, b) of the reference's base-255 code (src/py/utils.py:604-621); miss (-1) -> (0, 0, 0)."""
    ids = np.asarray(ids, dtype=np.int64)
    rgb = np.zeros(ids.shape + (3,), dtype=np.uint8)
    ok = ids >= 0
    v = ids[ok]
    rgb[ok, 0] = (v % 255) + 1
    rgb[ok, 1] = (v // 255) % 255
    rgb[ok, 2] = (v // (255 * 255)) % 255
    return rgb


def DecodeIdsRGB(rgb):
    """(r, g, b) -> id, -1 for background (reference src/py/utils.py:650-655)."""
    rgb = np.asarray(rgb).astype(np.int64)
    return (rgb[..., 2] * 255 * 255 + rgb[..., 1] * 255 + rgb[..., 0] - 1).astype(np.int32)


def PointFeatureArray(surf, point_features_name):
    """float32 [V, n_components] of a named point-data array, or the coordinates for "coords" / "points"
    (reference src/py/utils.py:668-675)."""
    if point_features_name in ("coords", "points"):
        feats = np.ascontiguousarray(surf.points, dtype=np.float32)
    else:
        if point_features_name not in surf.point_data:
            raise KeyError("point data array %r not found" % point_features_name)
        feats = np.ascontiguousarray(surf.point_data[point_features_name], dtype=np.float32)
    return feats.reshape(len(surf.points), -1)


def ExtractPointFeatures(surf, point_ids_rgb, point_features_name, zero=0, use_multi=True, device=0):
    """reference src/py/utils.py:664-684: per pixel, the named point-data array (or the coordinates for
    "coords"/"points") at the point id coded in the id map; ``zero`` where nothing was hit.
    ``point_ids_rgb`` is the uint8 RGB map ([..., 3]); an int32 array of point ids ([...]) is accepted too."""
    a = np.asarray(point_ids_rgb)
    ids = DecodeIdsRGB(a) if (a.ndim >= 2 and a.shape[-1] == 3 and a.dtype != np.int32) else a.astype(np.int32)
    feats = PointFeatureArray(surf, point_features_name)
    ctx = default_context(device)
    ctx.set_mesh(np.ascontiguousarray(surf.points, np.float32), np.ascontiguousarray(surf.polys, np.int32),
                 skip_normalise=True)
    return ctx.point_features(np.ascontiguousarray(ids), feats, zero=float(zero), mode=1)


__all__ = [n for n in dir() if not n.startswith("_") and n not in ("math", "np", "nrrd", "readers")]
