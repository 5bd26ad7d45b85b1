"""Container-only stand-in for the `vtk` module: just enough of the data-container API
for the reference's *pure-Python* code (src/py/LinearSubdivisionFilter.py, utils.py,
post_process.py, readers.py and slices of predict_v3.py / dental_model_seg.py /
predict.py) to be imported UNMODIFIED from /root/reference and executed, so that its
results become golden vectors (tests/golden/make_ref_golden.py).

TEST INFRASTRUCTURE ONLY.  Nothing here is a VTK algorithm except where stated:

* containers (vtkPoints, vtkIdList, vtkCellArray, vtkPolyData, data arrays) store what
  VTK stores: vtkPoints defaults to float32 (VTK_FLOAT), InsertNextPoint/SetPoint round
  to float32, GetPoint returns the float32 widened to double; point->cell links list
  cells in ascending id (vtkCellLinks::BuildLinks order);
* vtkIncrementalOctreePointLocator is reduced to what the reference uses it for:
  IsInsertedPoint / InsertNextPoint with zero tolerance == exact equality of the
  float32-rounded coordinates (vtkIncrementalOctreePointLocator compares the stored
  floats with (float)x for VTK_FLOAT points);
* vtkPlatonicSolidSource carries the icosahedron tables of vtkPlatonicSolidSource.cxx
  (restated from VTK 9; the uniform solid scale is selectable, see ICOSA_SOLID_SCALE);
* vtkPlaneSource / vtkPolyDataNormals / vtkCellLocator are *oracle-backed*: they call
  oracle/ and therefore pin nothing about VTK itself; they only let reference code
  that sits AROUND those calls run (predict.py:83-139).  Vectors produced through
  them are labelled `vtk_internal=oracle` in the fixtures.
"""
import math

import numpy as np

VTK_FLOAT, VTK_DOUBLE, VTK_INT, VTK_UNSIGNED_CHAR, VTK_ID_TYPE = 10, 11, 6, 3, 12

# vtkPlatonicSolidSource.cxx scales the icosahedron table uniformly before storing the
# points as float32.  The constant is recalled from the VTK 9 source, not read from it;
# after normalize_points it only moves the last float32 bit of a viewpoint and decides
# which float-equality duplicates leak.
ICOSA_SOLID_SCALE = 1.0 / 0.58778524999243


class mutable(object):
    def __init__(self, v):
        self.v = v

    def get(self):
        return self.v

    def set(self, v):
        self.v = v

    def __int__(self):
        return int(self.v)

    def __index__(self):
        return int(self.v)

    def __float__(self):
        return float(self.v)


reference = mutable


class vtkObject(object):
    def Update(self):
        pass

    def Modified(self):
        pass


class vtkDataArray(vtkObject):
    _dtype = np.float64

    def __init__(self):
        self.name = None
        self.ncomp = 1
        self._buf = np.zeros((0, 1), dtype=self._dtype)
        self._n = 0

    # `data` is the [n_tuples, n_components] view of a capacity-doubling buffer
    @property
    def data(self):
        return self._buf[:self._n]

    @data.setter
    def data(self, a):
        a = np.asarray(a)
        self._buf = a.reshape(len(a), -1) if a.ndim != 2 else a
        self._n = len(a)
        self.ncomp = self._buf.shape[1]

    # -- shape
    def SetName(self, n):
        self.name = n

    def GetName(self):
        return self.name

    def SetNumberOfComponents(self, n):
        self.ncomp = int(n)
        self._buf = np.zeros((0, self.ncomp), dtype=self._buf.dtype)
        self._n = 0

    def GetNumberOfComponents(self):
        return self.ncomp

    def SetNumberOfTuples(self, n):
        old = self.data
        buf = np.zeros((int(n), self.ncomp), dtype=old.dtype)
        m = min(len(old), int(n))
        buf[:m] = old[:m]
        self._buf, self._n = buf, int(n)

    def GetNumberOfTuples(self):
        return self._n

    def GetSize(self):
        return self._n * self.ncomp

    def GetNumberOfValues(self):
        return self._n * self.ncomp

    # -- access
    def GetTuple(self, i):
        return tuple(float(x) for x in self.data[int(i)])

    def GetTuple3(self, i):
        return self.GetTuple(i)

    def GetTuple1(self, i):
        return float(self.data[int(i), 0])

    def GetValue(self, i):
        return self.data.reshape(-1)[int(i)].item()

    def SetTuple(self, i, t):
        self.data[int(i)] = np.asarray(t, dtype=np.float64).astype(self._buf.dtype)

    def SetTuple1(self, i, v):
        self.SetTuple(i, (v,))

    def InsertNextTuple(self, t):
        if self._n == len(self._buf):
            grown = np.zeros((max(16, 2 * len(self._buf)), self.ncomp), dtype=self._buf.dtype)
            grown[:self._n] = self._buf[:self._n]
            self._buf = grown
        self._buf[self._n] = np.asarray(t, dtype=np.float64).astype(self._buf.dtype)
        self._n += 1
        return self._n - 1

    def InsertNextTuple3(self, a, b, c):
        return self.InsertNextTuple((a, b, c))

    def InsertNextTuple1(self, a):
        return self.InsertNextTuple((a,))

    def Fill(self, v):
        self.data[...] = v

    def GetRange(self, out=None, comp=0):
        lo = float(self.data[:, comp].min()) if self._n else 1e299
        hi = float(self.data[:, comp].max()) if self._n else -1e299
        if out is not None:
            out[0], out[1] = lo, hi
            return None
        return (lo, hi)

    def DeepCopy(self, other):
        self.name = other.name
        self.data = other.data.copy()

    def __iter__(self):      # np.array(vtk_array) in the reference iterates values
        return iter(self.data.reshape(-1).tolist() if self.ncomp == 1 else self.data.tolist())

    def __len__(self):
        return self._n


class vtkFloatArray(vtkDataArray):
    _dtype = np.float32


class vtkDoubleArray(vtkDataArray):
    _dtype = np.float64


class vtkIntArray(vtkDataArray):
    _dtype = np.int32


class vtkIdTypeArray(vtkDataArray):
    _dtype = np.int64


class vtkUnsignedCharArray(vtkDataArray):
    _dtype = np.uint8


def _array_for(a):
    a = np.asarray(a)
    cls = {np.dtype(np.float32): vtkFloatArray, np.dtype(np.float64): vtkDoubleArray,
           np.dtype(np.int32): vtkIntArray, np.dtype(np.int64): vtkIdTypeArray,
           np.dtype(np.uint8): vtkUnsignedCharArray}[a.dtype]
    out = cls()
    out.data = a.reshape(len(a), 1 if a.ndim == 1 else a.shape[1])
    return out


class vtkPoints(vtkObject):
    def __init__(self):
        self.arr = vtkFloatArray()          # vtkPoints::New() -> VTK_FLOAT
        self.arr.SetNumberOfComponents(3)

    def SetDataTypeToDouble(self):
        self.arr = vtkDoubleArray()
        self.arr.SetNumberOfComponents(3)

    def GetNumberOfPoints(self):
        return self.arr.GetNumberOfTuples()

    def SetNumberOfPoints(self, n):
        self.arr.SetNumberOfTuples(n)

    def GetPoint(self, i):
        return self.arr.GetTuple(i)

    def SetPoint(self, i, *p):
        if len(p) == 1:
            p = p[0]
        self.arr.SetTuple(i, p)

    def InsertNextPoint(self, *p):
        if len(p) == 1:
            p = p[0]
        return self.arr.InsertNextTuple(p)

    def GetData(self):
        return self.arr

    def SetData(self, arr):
        self.arr = arr

    def GetBounds(self):
        d = self.arr.data.astype(np.float64)
        if len(d) == 0:
            return (1.0, -1.0, 1.0, -1.0, 1.0, -1.0)
        lo, hi = d.min(axis=0), d.max(axis=0)
        return (lo[0], hi[0], lo[1], hi[1], lo[2], hi[2])

    def DeepCopy(self, other):
        self.arr = type(other.arr)()
        self.arr.DeepCopy(other.arr)


class vtkIdList(vtkObject):
    def __init__(self):
        self.ids = []

    def GetNumberOfIds(self):
        return len(self.ids)

    def SetNumberOfIds(self, n):
        self.ids = (self.ids + [0] * n)[:n]

    def GetId(self, i):
        return self.ids[i]

    def SetId(self, i, v):
        self.ids[i] = int(v)

    def InsertNextId(self, v):
        self.ids.append(int(v))
        return len(self.ids) - 1

    def Reset(self):
        self.ids = []


class _Cell(vtkObject):
    _n = 0

    def __init__(self):
        self.pids = vtkIdList()
        self.pids.SetNumberOfIds(self._n)

    def GetPointIds(self):
        return self.pids


class vtkTriangle(_Cell):
    _n = 3


class vtkLine(_Cell):
    _n = 2


class vtkVertex(_Cell):
    _n = 1


class vtkCellArray(vtkObject):
    def __init__(self):
        self.cells = []

    def InsertNextCell(self, cell, ids=None):
        if isinstance(cell, _Cell):
            self.cells.append(list(cell.pids.ids))
        elif isinstance(cell, vtkIdList):
            self.cells.append(list(cell.ids))
        else:
            self.cells.append([int(x) for x in ids])
        return len(self.cells) - 1

    def GetNumberOfCells(self):
        return len(self.cells)

    def GetData(self):          # legacy layout: n, id0, id1, ...
        flat = []
        for c in self.cells:
            flat.append(len(c))
            flat.extend(c)
        return _array_for(np.asarray(flat, dtype=np.int64))

    def GetConnectivityArray(self):
        return _array_for(np.asarray([i for c in self.cells for i in c], dtype=np.int64))

    def DeepCopy(self, other):
        self.cells = [list(c) for c in other.cells]


class _Attributes(vtkObject):
    def __init__(self):
        self.arrays = []
        self.scalars = None
        self.normals = None

    def AddArray(self, a):
        for i, b in enumerate(self.arrays):
            if b.name == a.name:
                self.arrays[i] = a
                return i
        self.arrays.append(a)
        return len(self.arrays) - 1

    def GetArray(self, key):
        if isinstance(key, int):
            return self.arrays[key]
        for a in self.arrays:
            if a.name == key:
                return a
        return None

    def GetScalars(self, name=None):
        return self.scalars if name is None else self.GetArray(name)

    def SetScalars(self, a):
        self.scalars = a
        self.AddArray(a)

    def SetActiveScalars(self, name):
        self.scalars = self.GetArray(name)

    def GetNormals(self):
        return self.normals

    def SetNormals(self, a):
        self.normals = a
        self.AddArray(a)

    def GetNumberOfArrays(self):
        return len(self.arrays)

    def DeepCopy(self, other):
        self.arrays = []
        for a in other.arrays:
            b = type(a)()
            b.DeepCopy(a)
            self.arrays.append(b)
        self.scalars = self.GetArray(other.scalars.name) if other.scalars is not None else None
        self.normals = self.GetArray(other.normals.name) if other.normals is not None else None


class vtkPolyData(vtkObject):
    def __init__(self):
        self.points = None
        self.polys = vtkCellArray()
        self.verts = vtkCellArray()
        self.lines = vtkCellArray()
        self.pd = _Attributes()
        self.cd = _Attributes()
        self._links = None

    def SetPoints(self, p):
        self.points = p

    def GetPoints(self):
        return self.points

    def SetPolys(self, c):
        self.polys, self._links = c, None

    def GetPolys(self):
        return self.polys

    def SetVerts(self, c):
        self.verts = c

    def SetLines(self, c):
        self.lines = c

    def GetLines(self):
        return self.lines

    def _cells(self):           # vtkPolyData cell numbering: verts, lines, polys
        return self.verts.cells + self.lines.cells + self.polys.cells

    def GetNumberOfCells(self):
        return len(self.verts.cells) + len(self.lines.cells) + len(self.polys.cells)

    def GetNumberOfPolys(self):
        return len(self.polys.cells)

    def GetNumberOfPoints(self):
        return self.points.GetNumberOfPoints() if self.points is not None else 0

    def GetPoint(self, i):
        return self.points.GetPoint(i)

    def GetBounds(self):
        return self.points.GetBounds()

    def GetCellPoints(self, cid, idlist):
        cid = int(cid)
        for group in (self.verts.cells, self.lines.cells, self.polys.cells):
            if cid < len(group):
                idlist.ids = list(group[cid])
                return
            cid -= len(group)
        raise IndexError("cell id out of range")

    def BuildLinks(self):
        links = [[] for _ in range(self.GetNumberOfPoints())]
        for c, cell in enumerate(self._cells()):
            for p in cell:
                links[p].append(c)
        self._links = links

    def GetPointCells(self, pid, idlist):
        if self._links is None:
            self.BuildLinks()
        idlist.ids = list(self._links[int(pid)])

    def GetPointData(self):
        return self.pd

    def GetCellData(self):
        return self.cd

    def DeepCopy(self, other):
        self.points = vtkPoints()
        self.points.DeepCopy(other.points)
        for k in ("polys", "verts", "lines"):
            c = vtkCellArray()
            c.DeepCopy(getattr(other, k))
            setattr(self, k, c)
        self.pd, self.cd = _Attributes(), _Attributes()
        self.pd.DeepCopy(other.pd)
        self.cd.DeepCopy(other.cd)
        self._links = None


class vtkIncrementalOctreePointLocator(vtkObject):
    """Zero-tolerance point insertion only: exact equality of float32-rounded coordinates."""

    def SetDataSet(self, ds):
        pass

    def BuildLocator(self):
        pass

    def InitPointInsertion(self, points, bounds, est=0):
        self.points = points
        self.index = {}
        return 1

    @staticmethod
    def _key(x):
        return np.asarray(x, dtype=np.float64).astype(np.float32).tobytes()

    def IsInsertedPoint(self, *x):
        if len(x) == 1:
            x = x[0]
        return self.index.get(self._key(x), -1)

    def InsertNextPoint(self, x):
        pid = self.points.InsertNextPoint(x)
        self.index.setdefault(self._key(x), pid)
        return pid

    def InsertUniquePoint(self, x, pid_ref=None):
        pid = self.IsInsertedPoint(x)
        if pid >= 0:
            return 0, pid
        return 1, self.InsertNextPoint(x)


_ICOSA_POINTS = [
    (0.0, 0.0, -1.0), (0.0, 0.894427, -0.447214), (0.850651, 0.276393, -0.447214),
    (0.525731, -0.723607, -0.447214), (-0.525731, -0.723607, -0.447214), (-0.850651, 0.276393, -0.447214),
    (0.0, -0.894427, 0.447214), (0.850651, -0.276393, 0.447214), (0.525731, 0.723607, 0.447214),
    (-0.525731, 0.723607, 0.447214), (-0.850651, -0.276393, 0.447214), (0.0, 0.0, 1.0)]
_ICOSA_VERTS = [
    (0, 5, 4), (0, 4, 3), (0, 3, 2), (0, 2, 1), (0, 1, 5), (5, 1, 9), (5, 9, 10), (5, 10, 4), (4, 10, 6),
    (4, 6, 3), (3, 6, 7), (3, 7, 2), (2, 7, 8), (2, 8, 1), (1, 8, 9), (9, 8, 11), (9, 11, 10), (10, 11, 6),
    (6, 11, 7), (7, 11, 8)]


class vtkPlatonicSolidSource(vtkObject):
    def SetSolidTypeToIcosahedron(self):
        self.solid = "icosahedron"

    def GetOutput(self):
        pts = vtkPoints()
        for p in _ICOSA_POINTS:
            pts.InsertNextPoint([ICOSA_SOLID_SCALE * c for c in p])
        polys = vtkCellArray()
        for f in _ICOSA_VERTS:
            polys.InsertNextCell(None, f)
        out = vtkPolyData()
        out.SetPoints(pts)
        out.SetPolys(polys)
        return out


class vtkPlaneSource(vtkObject):
    """vtkPlaneSource::RequestData restated (oracle-level knowledge of VTK, not a pin):
    point (i,j) = origin + (i/XRes)*(p1-origin) + (j/YRes)*(p2-origin), i fastest, stored float32."""

    def SetOrigin(self, *o):
        self.o = np.asarray(o[0] if len(o) == 1 else o, dtype=np.float64)

    def SetPoint1(self, *o):
        self.p1 = np.asarray(o[0] if len(o) == 1 else o, dtype=np.float64)

    def SetPoint2(self, *o):
        self.p2 = np.asarray(o[0] if len(o) == 1 else o, dtype=np.float64)

    def SetXResolution(self, r):
        self.xr = int(r)

    def SetYResolution(self, r):
        self.yr = int(r)

    def GetOutput(self):
        v1, v2 = self.p1 - self.o, self.p2 - self.o
        pts = vtkPoints()
        rows = []
        for j in range(self.yr + 1):
            tc1 = float(j) / self.yr
            for i in range(self.xr + 1):
                tc0 = float(i) / self.xr
                rows.append([self.o[k] + tc0 * v1[k] + tc1 * v2[k] for k in range(3)])
        pts.arr.data = np.asarray(rows, dtype=np.float64).astype(np.float32)
        out = vtkPolyData()
        out.SetPoints(pts)
        return out


def _tris_of(surf):
    return np.asarray(surf.polys.cells, dtype=np.int32).reshape(-1, 3)


class vtkPolyDataNormals(vtkObject):
    """ORACLE-BACKED (pins nothing about VTK): point normals from oracle.point_normals."""

    def SetInputData(self, s):
        self.inp = s

    def __getattr__(self, name):          # ComputeCellNormalsOn, SplittingOff, ...
        if name.endswith("On") or name.endswith("Off") or name.startswith("Set"):
            return lambda *a: None
        raise AttributeError(name)

    def GetOutput(self):
        from oracle import oracle as orc
        out = vtkPolyData()
        out.DeepCopy(self.inp)
        P = self.inp.points.arr.data.astype(np.float32)
        n = _array_for(orc.point_normals(P, _tris_of(self.inp)))
        n.SetName("Normals")
        out.pd.SetNormals(n)
        return out


class vtkCellLocator(vtkObject):
    """ORACLE-BACKED (pins nothing about VTK): nearest hit from oracle.cast (brute force)."""

    def SetDataSet(self, s):
        self.ds = s

    def BuildLocator(self):
        self.P = self.ds.points.arr.data.astype(np.float32)
        self.T = _tris_of(self.ds)

    def IntersectWithLine(self, p1, p2, tol, t, x, pcoords, subId, cellId):
        from oracle import oracle as orc
        orc.set_tolerance(tol)
        ray = np.concatenate([np.asarray(p1, np.float64), np.asarray(p2, np.float64)])
        ids, tt, xx, ww = orc.cast(self.P, self.T, ray, grid=False)
        orc.set_tolerance(1e-10)
        if ids[0] < 0:
            return 0
        t.set(float(tt[0]))
        x[0], x[1], x[2] = (float(v) for v in xx[0])
        cellId.set(int(ids[0]))
        return 1


class vtkPolyDataWriter(vtkObject):
    """No-op: predict_v3.py:93-133 writes intermediate files between its post-processing calls."""

    def SetFileName(self, n):
        pass

    def SetInputData(self, s):
        pass

    def Write(self):
        return 1
