"""Mesh file readers / writer for the host side (vtk / itk are not available here).

Covers what ``ReadSurf`` of the reference accepts for the fly-by path
(reference src/py/utils.py:112-194): legacy VTK polydata (ASCII and BINARY,
classic ``POLYGONS n size`` cells and the VTK >= 9 ``OFFSETS`` / ``CONNECTIVITY``
form), Wavefront OBJ, STL (binary and ASCII, coincident vertices merged like
vtkSTLReader does by default), OFF (reference src/py/readers.py:15-61), VTK XML
polydata ``.vtp`` (ascii / inline binary / appended raw or base64, optional zlib
blocks, UInt32 or UInt64 headers) and GIFTI ``.gii`` (ASCII, Base64Binary,
GZipBase64Binary), both parsed with the standard library only.
Polygons with more than three vertices are fan-triangulated (the reference's
subdivision filter and ray caster only accept triangles).
"""
import os
import re
import struct

import numpy as np

_VTK_TYPES = {
    "bit": np.uint8, "unsigned_char": np.uint8, "char": np.int8, "unsigned_short": np.uint16,
    "short": np.int16, "unsigned_int": np.uint32, "int": np.int32, "unsigned_long": np.uint64,
    "long": np.int64, "float": np.float32, "double": np.float64, "vtkidtype": np.int64,
    "vtktypeint64": np.int64, "vtktypeint32": np.int32,
}


class MeshData:
    """Plain container: points float32 [V,3], polys int32 [T,3], named point / cell arrays."""

    def __init__(self, points, polys, point_data=None, cell_data=None):
        self.points = np.ascontiguousarray(points, dtype=np.float32).reshape(-1, 3)
        self.polys = np.ascontiguousarray(polys, dtype=np.int32).reshape(-1, 3)
        self.point_data = dict(point_data or {})
        self.cell_data = dict(cell_data or {})


def _fan(polygons):
    tris = []
    for p in polygons:
        for k in range(1, len(p) - 1):
            tris.append((p[0], p[k], p[k + 1]))
    return np.array(tris, dtype=np.int32).reshape(-1, 3)


class _Tokens:
    """Token / binary cursor over a legacy VTK file."""

    def __init__(self, data):
        self.d = data
        self.i = 0

    def eof(self):
        self._skip_ws()
        return self.i >= len(self.d)

    def _skip_ws(self):
        d, n = self.d, len(self.d)
        while self.i < n and d[self.i] in b" \t\r\n":
            self.i += 1

    def line(self):
        j = self.d.find(b"\n", self.i)
        if j < 0:
            j = len(self.d)
        s = self.d[self.i:j]
        self.i = j + 1
        return s.decode("latin-1").strip()

    def word(self):
        self._skip_ws()
        j = self.i
        d, n = self.d, len(self.d)
        while j < n and d[j] not in b" \t\r\n":
            j += 1
        s = d[self.i:j].decode("latin-1")
        self.i = j
        return s

    def peek(self):
        save = self.i
        w = self.word()
        self.i = save
        return w

    def array(self, count, dtype, binary):
        dtype = np.dtype(dtype)
        if binary:
            if self.d[self.i:self.i + 1] == b"\n":
                self.i += 1
            nbytes = count * dtype.itemsize
            a = np.frombuffer(self.d, dtype=dtype.newbyteorder(">"), count=count, offset=self.i)
            self.i += nbytes
            return a.astype(dtype)
        vals = []
        while len(vals) < count:
            vals.append(self.word())
        if dtype.kind == "f":
            return np.array(vals, dtype=np.float64).astype(dtype)
        return np.array([int(float(v)) for v in vals], dtype=dtype)


def read_vtk(path):
    with open(path, "rb") as f:
        tk = _Tokens(f.read())
    header = tk.line()
    if not header.startswith("# vtk DataFile"):
        raise ValueError("%s: not a legacy VTK file" % path)
    tk.line()  # title
    binary = tk.line().upper().startswith("BINARY")
    points = None
    polygons = []
    point_data, cell_data = {}, {}
    target = None
    n_target = 0
    while not tk.eof():
        kw = tk.word().upper()
        if kw == "DATASET":
            kind = tk.word().upper()
            if kind != "POLYDATA":
                raise ValueError("%s: DATASET %s is not supported (POLYDATA only)" % (path, kind))
        elif kw == "POINTS":
            n = int(tk.word())
            dt = _VTK_TYPES[tk.word().lower()]
            points = tk.array(3 * n, dt, binary).reshape(n, 3)
        elif kw in ("POLYGONS", "TRIANGLE_STRIPS", "LINES", "VERTICES"):
            a, b = int(tk.word()), int(tk.word())
            if tk.peek().upper() == "OFFSETS":  # VTK >= 9 (file version 5.x)
                tk.word()
                odt = _VTK_TYPES[tk.word().lower()]
                offsets = tk.array(a, odt, binary)
                assert tk.word().upper() == "CONNECTIVITY"
                cdt = _VTK_TYPES[tk.word().lower()]
                conn = tk.array(b, cdt, binary)
                cells = [conn[offsets[k]:offsets[k + 1]].tolist() for k in range(a - 1)]
            else:
                flat = tk.array(b, np.int32, binary)
                cells, i = [], 0
                for _ in range(a):
                    m = int(flat[i])
                    cells.append(flat[i + 1:i + 1 + m].tolist())
                    i += 1 + m
            if kw == "POLYGONS":
                polygons += cells
            elif kw == "TRIANGLE_STRIPS":
                for s in cells:
                    for k in range(len(s) - 2):
                        polygons.append((s[k], s[k + 1], s[k + 2]) if k % 2 == 0 else (s[k + 1], s[k], s[k + 2]))
        elif kw == "POINT_DATA":
            n_target = int(tk.word())
            target = point_data
        elif kw == "CELL_DATA":
            n_target = int(tk.word())
            target = cell_data
        elif kw == "SCALARS":
            name = tk.word()
            dt = _VTK_TYPES[tk.word().lower()]
            ncomp = 1
            if tk.peek().upper() != "LOOKUP_TABLE":
                ncomp = int(tk.word())
            if tk.peek().upper() == "LOOKUP_TABLE":
                tk.word()
                tk.word()
            a = tk.array(n_target * ncomp, dt, binary)
            target[name] = a.reshape(n_target, ncomp) if ncomp > 1 else a
        elif kw in ("NORMALS", "VECTORS"):
            name = tk.word()
            dt = _VTK_TYPES[tk.word().lower()]
            target[name] = tk.array(n_target * 3, dt, binary).reshape(n_target, 3)
        elif kw == "COLOR_SCALARS":
            name = tk.word()
            ncomp = int(tk.word())
            a = tk.array(n_target * ncomp, np.uint8 if binary else np.float32, binary)
            target[name] = a.reshape(n_target, ncomp)
        elif kw == "TEXTURE_COORDINATES":
            name = tk.word()
            dim = int(tk.word())
            dt = _VTK_TYPES[tk.word().lower()]
            target[name] = tk.array(n_target * dim, dt, binary).reshape(n_target, dim)
        elif kw == "LOOKUP_TABLE":
            tk.word()
            size = int(tk.word())
            tk.array(size * 4, np.uint8 if binary else np.float32, binary)
        elif kw == "FIELD":
            tk.word()
            n_arrays = int(tk.word())
            for _ in range(n_arrays):
                name = tk.word()
                ncomp, ntup = int(tk.word()), int(tk.word())
                dt = _VTK_TYPES[tk.word().lower()]
                a = tk.array(ncomp * ntup, dt, binary)
                if target is not None and ntup == n_target:
                    target[name] = a.reshape(ntup, ncomp) if ncomp > 1 else a
        elif kw == "METADATA":
            while tk.line() != "":
                if tk.eof():
                    break
        else:
            raise ValueError("%s: unsupported legacy VTK keyword %r" % (path, kw))
    if points is None:
        raise ValueError("%s: no POINTS section" % path)
    ntri_src = len(polygons)
    tris = _fan(polygons)
    if len(tris) != ntri_src:
        cell_data = {}  # cell arrays no longer line up after triangulation
    return MeshData(points, tris, point_data, cell_data)


def read_obj(path):
    pts, polys = [], []
    with open(path, "r", errors="replace") as f:
        for line in f:
            if line.startswith("v "):
                p = line.split()
                pts.append((float(p[1]), float(p[2]), float(p[3])))
            elif line.startswith("f "):
                idx = []
                for tok in line.split()[1:]:
                    i = int(tok.split("/")[0])
                    idx.append(i - 1 if i > 0 else len(pts) + i)
                polys.append(idx)
    return MeshData(np.array(pts, dtype=np.float32), _fan(polys))


def read_off(path):
    with open(path, "r", errors="replace") as f:
        toks = f.read().split()
    i = 0
    if toks[0].upper().startswith("OFF"):
        rest = toks[0][3:]
        i = 1
        if rest:  # "OFF123 ..." glued header
            toks.insert(1, rest)
    nv, nf = int(toks[i]), int(toks[i + 1])
    i += 3
    pts = np.array(toks[i:i + 3 * nv], dtype=np.float64).reshape(nv, 3)
    i += 3 * nv
    polys = []
    for _ in range(nf):
        m = int(toks[i])
        polys.append([int(x) for x in toks[i + 1:i + 1 + m]])
        i += 1 + m
    return MeshData(pts, _fan(polys))


def read_stl(path):
    with open(path, "rb") as f:
        data = f.read()
    tri = None
    if len(data) >= 84:
        n = struct.unpack_from("<I", data, 80)[0]
        if 84 + 50 * n == len(data):
            rec = np.frombuffer(data, dtype=np.dtype([("n", "<f4", 3), ("v", "<f4", 9), ("a", "<u2")]),
                                count=n, offset=84)
            tri = rec["v"].reshape(n, 3, 3).astype(np.float32)
    if tri is None:
        nums = re.findall(rb"vertex\s+([-+0-9.eE]+)\s+([-+0-9.eE]+)\s+([-+0-9.eE]+)", data)
        tri = np.array(nums, dtype=np.float64).astype(np.float32).reshape(-1, 3, 3)
    flat = tri.reshape(-1, 3)
    uniq, first, inv = np.unique(flat, axis=0, return_index=True, return_inverse=True)
    # keep first-occurrence order (vtkSTLReader's merging locator inserts in file order)
    order = np.argsort(first)
    rank = np.empty_like(order)
    rank[order] = np.arange(len(order))
    return MeshData(uniq[order], rank[inv.reshape(-1)].reshape(-1, 3))


# --------------------------------------------------------------------------- VTK XML polydata (.vtp)
_XML_TYPES = {"Int8": np.int8, "UInt8": np.uint8, "Int16": np.int16, "UInt16": np.uint16, "Int32": np.int32,
              "UInt32": np.uint32, "Int64": np.int64, "UInt64": np.uint64, "Float32": np.float32, "Float64": np.float64}


def _b64_chars(n_bytes):
    return (n_bytes + 2) // 3 * 4


def _xml_block(raw, header_dt, compressed, base64_text):
    """One data block of a VTK XML file -> bytes.  `raw` is base64 text (inline binary or base64 appended
    data, header and payload encoded separately) or raw bytes (appended raw)."""
    import base64
    import zlib
    hs = header_dt.itemsize
    if base64_text:
        if compressed:
            head = np.frombuffer(base64.b64decode(raw[:_b64_chars(3 * hs)])[:3 * hs], header_dt)
            n_head = _b64_chars((3 + int(head[0])) * hs)
            sizes = np.frombuffer(base64.b64decode(raw[:n_head])[:(3 + int(head[0])) * hs], header_dt)[3:]
            total = int(sizes.sum())
            body = base64.b64decode(raw[n_head:n_head + _b64_chars(total)])
        else:
            n = int(np.frombuffer(base64.b64decode(raw[:_b64_chars(hs)])[:hs], header_dt)[0])
            return base64.b64decode(raw[_b64_chars(hs):_b64_chars(hs) + _b64_chars(n)])[:n]
    else:
        if compressed:
            head = np.frombuffer(raw[:3 * hs], header_dt)
            sizes = np.frombuffer(raw[3 * hs:(3 + int(head[0])) * hs], header_dt)
            body = raw[(3 + int(head[0])) * hs:]
        else:
            n = int(np.frombuffer(raw[:hs], header_dt)[0])
            return bytes(raw[hs:hs + n])
    out, off = [], 0
    for sz in sizes:
        out.append(zlib.decompress(bytes(body[off:off + int(sz)])))
        off += int(sz)
    return b"".join(out)


def read_vtp(path):
    """vtkXMLPolyDataReader for the fly-by path (reference src/py/utils.py:121-125): first piece, points, polys
    (and triangle strips), point / cell data arrays."""
    import xml.etree.ElementTree as ET
    blob = open(path, "rb").read()
    appended_raw = None
    m = re.search(rb"<AppendedData[^>]*encoding\s*=\s*\"raw\"[^>]*>", blob)
    if m:  # raw bytes are not XML: cut them out before parsing
        start = blob.index(b"_", m.end()) + 1
        end = blob.rindex(b"</AppendedData>")
        appended_raw = blob[start:end]
        blob = blob[:start] + blob[end:]
    root = ET.fromstring(blob)
    if root.tag != "VTKFile" or root.get("type") != "PolyData":
        raise ValueError("%s is not a VTK XML PolyData file" % path)
    if root.get("byte_order", "LittleEndian") != "LittleEndian":
        raise NotImplementedError("big-endian .vtp")
    comp = root.get("compressor")
    if comp not in (None, "vtkZLibDataCompressor"):
        raise NotImplementedError("compressor %s (only zlib is supported)" % comp)
    header_dt = np.dtype(_XML_TYPES[root.get("header_type", "UInt32")])
    appended_b64 = None
    ad = root.find("AppendedData")
    if ad is not None and appended_raw is None:
        appended_b64 = (ad.text or "").strip()
        appended_b64 = appended_b64[1:] if appended_b64.startswith("_") else appended_b64

    def array(da):
        dt = np.dtype(_XML_TYPES[da.get("type")])
        fmt = da.get("format", "ascii")
        if fmt == "ascii":
            a = np.array((da.text or "").split(), dtype=np.float64 if dt.kind == "f" else np.int64).astype(dt)
        elif fmt == "binary":
            a = np.frombuffer(_xml_block("".join((da.text or "").split()), header_dt, comp is not None, True), dt)
        elif fmt == "appended":
            off = int(da.get("offset"))
            if appended_raw is not None:
                a = np.frombuffer(_xml_block(appended_raw[off:], header_dt, comp is not None, False), dt)
            else:
                a = np.frombuffer(_xml_block(appended_b64[off:], header_dt, comp is not None, True), dt)
        else:
            raise ValueError("unknown DataArray format %r" % fmt)
        nc = int(da.get("NumberOfComponents", "1"))
        return a.reshape(-1, nc) if nc > 1 else a

    piece = root.find("PolyData").find("Piece")
    pts = array(piece.find("Points").find("DataArray")).reshape(-1, 3)

    def cells(tag):
        el = piece.find(tag)
        if el is None:
            return []
        arrs = {d.get("Name"): array(d) for d in el.findall("DataArray")}
        if "connectivity" not in arrs or len(arrs.get("offsets", [])) == 0:
            return []
        conn, offs = arrs["connectivity"].astype(np.int64), arrs["offsets"].astype(np.int64)
        return [conn[a:b].tolist() for a, b in zip(np.concatenate([[0], offs[:-1]]), offs)]

    polys = _fan(cells("Polys"))
    for strip in cells("Strips"):  # vtkTriangleFilter order: alternate the winding
        for k in range(len(strip) - 2):
            polys.append((strip[k], strip[k + 1], strip[k + 2]) if k % 2 == 0 else (strip[k + 1], strip[k], strip[k + 2]))
    data = {}
    for tag in ("PointData", "CellData"):
        el = piece.find(tag)
        data[tag] = {d.get("Name"): array(d) for d in el.findall("DataArray")} if el is not None else {}
    n_poly_cells = len(cells("Polys"))
    cd = {k: v for k, v in data["CellData"].items() if len(polys) == n_poly_cells and len(v) == n_poly_cells}
    return MeshData(pts, np.array(polys, np.int32).reshape(-1, 3), data["PointData"], cd)


# --------------------------------------------------------------------------- GIFTI (.gii)
_GII_TYPES = {"NIFTI_TYPE_UINT8": np.uint8, "NIFTI_TYPE_INT32": np.int32, "NIFTI_TYPE_FLOAT32": np.float32,
              "NIFTI_TYPE_FLOAT64": np.float64, "NIFTI_TYPE_INT64": np.int64, "NIFTI_TYPE_INT16": np.int16}


def read_gii(path):
    """The pointset and triangle arrays of a GIFTI surface, as nibabel's agg_data('pointset') /
    agg_data('triangle') return them (reference src/py/utils.py:167-191); other arrays with one row per point
    become point data named by their intent."""
    import base64
    import xml.etree.ElementTree as ET
    import zlib
    root = ET.parse(path).getroot()
    pts = tris = None
    extra = {}
    for da in root.iter("DataArray"):
        dt = np.dtype(_GII_TYPES[da.get("DataType")])
        dims = [int(da.get("Dim%d" % k)) for k in range(int(da.get("Dimensionality", "1")))]
        enc = da.get("Encoding")
        text = (da.find("Data").text or "").strip()
        if enc == "ASCII":
            a = np.array(text.split(), dtype=np.float64 if dt.kind == "f" else np.int64).astype(dt)
        elif enc in ("Base64Binary", "GZipBase64Binary"):
            raw = base64.b64decode(text)
            if enc == "GZipBase64Binary":
                raw = zlib.decompress(raw)
            a = np.frombuffer(raw, dt.newbyteorder(">" if da.get("Endian") == "BigEndian" else "<")).astype(dt)
        else:
            raise NotImplementedError("GIFTI encoding %s" % enc)
        a = a.reshape(dims, order="F" if da.get("ArrayIndexingOrder") == "ColumnMajorOrder" else "C")
        intent = da.get("Intent")
        if intent == "NIFTI_INTENT_POINTSET" and pts is None:
            pts = a
        elif intent == "NIFTI_INTENT_TRIANGLE" and tris is None:
            tris = a
        else:
            extra[intent.replace("NIFTI_INTENT_", "").lower()] = a
    if pts is None or tris is None:
        raise ValueError("%s has no pointset / triangle arrays" % path)
    return MeshData(pts, tris, {k: v for k, v in extra.items() if len(v) == len(pts)}, {})


def read_mesh(path):
    ext = os.path.splitext(path)[1].lower()
    if ext == ".vtk":
        return read_vtk(path)
    if ext == ".obj":
        return read_obj(path)
    if ext == ".stl":
        return read_stl(path)
    if ext == ".off":
        return read_off(path)
    if ext == ".vtp":
        return read_vtp(path)
    if ext == ".gii":
        return read_gii(path)
    raise ValueError("unknown mesh extension %r" % ext)


def write_vtk(path, points, polys, point_data=None, cell_data=None, binary=False):
    """Legacy VTK polydata (file version 4.2 cell layout, readable by every VTK)."""
    points = np.asarray(points, dtype=np.float32).reshape(-1, 3)
    polys = np.asarray(polys, dtype=np.int32).reshape(-1, 3)

    def vtk_type(a):
        for name in ("float", "double", "int", "unsigned_char", "long", "short", "unsigned_short", "unsigned_int", "char"):
            if np.dtype(_VTK_TYPES[name]) == a.dtype:
                return name
        return None

    with open(path, "wb") as f:
        def w(s):
            f.write(s.encode("ascii"))

        def dump(a):
            if binary:
                f.write(np.ascontiguousarray(a).astype(a.dtype.newbyteorder(">")).tobytes())
                w("\n")
            else:
                a2 = a.reshape(len(a), -1) if a.ndim > 1 else a.reshape(-1, 1)
                fmt = "%.9g" if a.dtype.kind == "f" else "%d"
                np.savetxt(f, a2, fmt=fmt)

        w("# vtk DataFile Version 4.2\nflyby_b200\n%s\nDATASET POLYDATA\n" % ("BINARY" if binary else "ASCII"))
        w("POINTS %d float\n" % len(points))
        dump(points)
        w("POLYGONS %d %d\n" % (len(polys), 4 * len(polys)))
        cells = np.concatenate([np.full((len(polys), 1), 3, np.int32), polys], axis=1)
        dump(cells)
        for label, n, data in (("POINT_DATA", len(points), point_data), ("CELL_DATA", len(polys), cell_data)):
            if not data:
                continue
            w("%s %d\n" % (label, n))
            for name, a in data.items():
                a = np.asarray(a)
                if vtk_type(a) is None:
                    a = a.astype(np.float32 if a.dtype.kind == "f" else np.int32)
                ncomp = 1 if a.ndim == 1 else a.shape[1]
                w("SCALARS %s %s %d\nLOOKUP_TABLE default\n" % (name.replace(" ", "_"), vtk_type(a), ncomp))
                dump(a)
