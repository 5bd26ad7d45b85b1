"""NRRD vector-image I/O compatible with what the reference writes through ITK.

The reference turns the ``[V, res, res, C]`` numpy result into an
``itk.VectorImage[float, 3]`` (GetImage, reference src/py/utils.py:704-748) and
writes it with ``itk.ImageFileWriter`` + ``UseCompressionOn``
(src/py/fly_by_features.py:419-424).  Its readers reshape purely by element
count (src/py/train_useg_flyby.py:170-174), so the ``sizes`` order below
(component fastest, then x, y, view) is the contract that matters.
"""
import gzip
import zlib

import numpy as np

_TYPES = {
    "float": np.float32, "double": np.float64, "uchar": np.uint8, "unsigned char": np.uint8, "uint8": np.uint8,
    "short": np.int16, "int16": np.int16, "ushort": np.uint16, "unsigned short": np.uint16, "int": np.int32,
    "int32": np.int32, "uint": np.uint32, "unsigned int": np.uint32, "longlong": np.int64, "int64": np.int64,
    "signed char": np.int8, "int8": np.int8, "long long": np.int64, "long long int": np.int64,
    "signed long long": np.int64, "unsigned long long": np.uint64, "ulonglong": np.uint64, "uint64": np.uint64,
    "unsigned long long int": np.uint64, "uint16": np.uint16, "uint32": np.uint32, "signed short": np.int16,
    "short int": np.int16, "signed int": np.int32, "unsigned short int": np.uint16, "int8_t": np.int8,
    "uint8_t": np.uint8, "int16_t": np.int16, "uint16_t": np.uint16, "int32_t": np.int32, "uint32_t": np.uint32,
    "int64_t": np.int64, "uint64_t": np.uint64,
}
_NAMES = {np.dtype(np.float32): "float", np.dtype(np.float64): "double", np.dtype(np.uint8): "unsigned char",
          np.dtype(np.int16): "short", np.dtype(np.uint16): "unsigned short", np.dtype(np.int32): "int",
          np.dtype(np.uint32): "unsigned int", np.dtype(np.int64): "long long", np.dtype(np.int8): "signed char",
          np.dtype(np.uint64): "unsigned long long"}


class Image:
    """What GetImage returns: the pixel array (numpy layout [..., C]) plus spacing."""

    def __init__(self, array, spacing=None, kind="vector"):
        self.array = np.ascontiguousarray(array)
        dim = self.array.ndim - 1
        self.spacing = list(spacing) if spacing is not None else [1.0] * dim
        self.kind = kind  # NRRD kind of the component axis: "vector" (itk.VectorImage), "RGB-color" (itk::RGBPixel)

    def GetNumberOfComponentsPerPixel(self):
        return self.array.shape[-1]

    def GetLargestPossibleRegion(self):
        return tuple(reversed(self.array.shape[:-1]))

    def __array__(self, dtype=None):
        return self.array if dtype is None else self.array.astype(dtype)


def write(path, array, compress=True, spacing=None, kind="vector"):
    """Write ``array`` [d_{n-1}, ..., d_0, C] (numpy order) as a vector image: NRRD sizes = C d_0 ... d_{n-1}."""
    if isinstance(array, Image):
        spacing = array.spacing
        kind = array.kind
        array = array.array
    a = np.ascontiguousarray(array)
    if a.dtype not in _NAMES:
        a = a.astype(np.float32)
    dim = a.ndim - 1          # image dimension (vector component axis excluded)
    sizes = [a.shape[-1]] + list(reversed(a.shape[:-1]))
    if spacing is None:
        spacing = [1.0] * dim
    dirs = []
    for k in range(dim):
        v = ["0"] * dim
        v[k] = repr(float(spacing[k])) if spacing[k] != 1 else "1"
        dirs.append("(" + ",".join(v) + ")")
    hdr = ["NRRD0004", "# Complete NRRD file format specification at:",
           "# http://teem.sourceforge.net/nrrd/format.html",
           "type: " + _NAMES[a.dtype], "dimension: %d" % (dim + 1), "space dimension: %d" % dim,
           "sizes: " + " ".join(str(s) for s in sizes),
           "space directions: none " + " ".join(dirs),
           "kinds: " + kind + " " + " ".join(["domain"] * dim), "endian: little",
           "encoding: " + ("gzip" if compress else "raw"),
           "space origin: (" + ",".join(["0"] * dim) + ")", "", ""]
    payload = a.astype(a.dtype.newbyteorder("<")).tobytes()
    if compress:
        payload = gzip.compress(payload, compresslevel=1)
    with open(path, "wb") as f:
        f.write("\n".join(hdr).encode("ascii"))
        f.write(payload)


def read(path):
    """Read an NRRD written by :func:`write` or by ITK; returns (array in numpy order, header dict)."""
    with open(path, "rb") as f:
        data = f.read()
    end, skip = data.find(b"\n\n"), 2
    crlf = data.find(b"\r\n\r\n")
    if crlf >= 0 and (end < 0 or crlf < end):   # header written with CRLF line ends
        end, skip = crlf, 4
    if end < 0:
        raise ValueError("%s: no NRRD header terminator" % path)
    header = {}
    for line in data[:end].decode("ascii", "replace").replace("\r", "").split("\n")[1:]:
        if line.startswith("#") or ":" not in line:
            continue
        k, v = line.split(":", 1)
        header[k.strip()] = v.strip().lstrip("=").strip()
    payload = data[end + skip:]
    enc = header.get("encoding", "raw").lower()
    if enc in ("gzip", "gz"):
        payload = zlib.decompress(payload, 16 + zlib.MAX_WBITS)
    elif enc != "raw":
        raise ValueError("%s: unsupported NRRD encoding %s" % (path, enc))
    dt = np.dtype(_TYPES[header["type"].lower()])
    if header.get("endian", "little").lower() == "big":
        dt = dt.newbyteorder(">")
    sizes = [int(s) for s in header["sizes"].split()]
    a = np.frombuffer(payload, dtype=dt, count=int(np.prod(sizes))).reshape(list(reversed(sizes)))
    return a.astype(dt.newbyteorder("=")), header
