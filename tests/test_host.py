"""CPU tests of the host side: C-ABI surface, loud failure without a GPU, file formats,
the id codec, rotations, and the reference's command-line option surface."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from flyby_b200 import _cabi, nrrd, readers, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE_FLAGS = {  # reference src/py/fly_by_features.py:432-489 -> default
    "--surf": None, "--dir": None, "--csv": None, "--csv_root_path": "", "--model": None, "--random_rotation": False,
    "--n_rotations": 0, "--scale_factor": None, "--bounds": None, "--translate": None, "--fiberBundle": False,
    "--fiber": False, "--nbFiber": 0.05, "--subject": None, "--extract_components": None, "--norm_shader": 1,
    "--split_z": 0, "--use_z": 1, "--rescale_features": 0, "--property": None, "--point_features": None,
    "--point_features_concat": 0, "--zero": 0, "--view_features": None, "--subdivision": None, "--spiral": None,
    "--turns": 4, "--resolution": 256, "--radius": 4, "--save_rotation": 0, "--visualize": 0, "--out": "out.nrrd",
    "--out_point_id": 0, "--uuid": False, "--ow": 1, "--concatenate": 1, "--verbose": 0,
}


def test_cabi_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "flyby_b200.h")).read()
    declared = re.findall(r"FLYBY_API\s+[\w\s\*]+?\b(flyby_\w+)\s*\(", header)
    assert sorted(declared) == sorted(_cabi.SYMBOLS) and len(declared) == len(set(declared))
    lib = _cabi.load()  # loads without a GPU; no compute call is made
    for s in declared:
        assert hasattr(lib, s), s
    out = subprocess.run(["nm", "-D", "--defined-only", _cabi.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (flyby_\w+)", out))
    assert exported == set(declared)
    assert lib.flyby_abi_version() == _cabi.ABI_VERSION


def test_struct_layouts_match_header():
    import ctypes as C
    assert C.sizeof(_cabi.NormOpts) == 56 and C.sizeof(_cabi.RenderOpts) == 48 and C.sizeof(_cabi.Stats) == 64
    assert _cabi.RenderOpts.tolerance.offset == 40 and _cabi.NormOpts.skip.offset == 48


def test_no_gpu_fails_loudly():
    """The product has no CPU path: without a CUDA device creating a context raises."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(_cabi.FlyByError):
        _cabi.Context(0)


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "fly-by-cnn_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")) or f == "Makefile":
                src = open(os.path.join(dirpath, f), errors="replace").read()
                assert "liborc" not in src and "hostsim" not in src.replace("tests/hostsim", ""), f
                assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), f


def test_vtk_roundtrip_ascii_binary(tmp_path):
    P, F = synth.bumpy_sphere(4, 0)
    pd = {"RegionId": np.arange(len(P), dtype=np.int32) % 7, "Curv": np.linspace(0, 1, len(P)).astype(np.float32)}
    for binary in (False, True):
        path = str(tmp_path / ("m%d.vtk" % binary))
        readers.write_vtk(path, P, F, point_data=pd, cell_data={"cid": np.arange(len(F), dtype=np.int32)}, binary=binary)
        m = readers.read_mesh(path)
        assert np.array_equal(m.polys, F) and np.allclose(m.points, P, rtol=1e-7)
        assert np.array_equal(m.point_data["RegionId"], pd["RegionId"]) and np.allclose(m.point_data["Curv"], pd["Curv"])
        assert np.array_equal(m.cell_data["cid"], np.arange(len(F)))
        if binary:
            assert np.array_equal(m.points, P)


def test_vtk_version5_offsets_connectivity(tmp_path):
    """vtk >= 9 (what the reference's vtk==9.2.2 writes) stores cells as OFFSETS / CONNECTIVITY."""
    path = str(tmp_path / "v5.vtk")
    with open(path, "w") as f:
        f.write("# vtk DataFile Version 5.1\nvtk output\nASCII\nDATASET POLYDATA\nPOINTS 4 float\n"
                "0 0 0 1 0 0 0 1 0 0 0 1\n\nPOLYGONS 3 7\nOFFSETS vtktypeint64\n0 3 7\n"
                "CONNECTIVITY vtktypeint64\n0 1 2 0 1 3 2\n\nPOINT_DATA 4\nNORMALS Normals float\n"
                "0 0 1 0 0 1 0 0 1 0 0 1\n")
    m = readers.read_mesh(path)
    assert len(m.points) == 4 and m.polys.tolist() == [[0, 1, 2], [0, 1, 3], [0, 3, 2]]  # quad fan-triangulated
    assert m.point_data["Normals"].shape == (4, 3)


def test_obj_off_stl(tmp_path):
    V, F = synth.cube()
    obj = tmp_path / "c.obj"
    obj.write_text("".join("v %g %g %g\n" % tuple(p) for p in V) + "".join("f %d//1 %d//1 %d//1\n" % tuple(t + 1) for t in F))
    m = readers.read_mesh(str(obj))
    assert np.array_equal(m.points, V) and np.array_equal(m.polys, F)
    off = tmp_path / "c.off"
    off.write_text("OFF\n8 12 0\n" + "".join("%g %g %g\n" % tuple(p) for p in V) + "".join("3 %d %d %d\n" % tuple(t) for t in F))
    m = readers.read_mesh(str(off))
    assert np.array_equal(m.points, V) and np.array_equal(m.polys, F)
    # binary STL: 36 unmerged corners -> 8 merged points, same geometry
    import struct
    stl = tmp_path / "c.stl"
    with open(stl, "wb") as f:
        f.write(b"\0" * 80 + struct.pack("<I", len(F)))
        for t in F:
            f.write(struct.pack("<12fH", 0, 0, 0, *V[t].reshape(-1), 0))
    m = readers.read_mesh(str(stl))
    assert len(m.points) == 8 and np.array_equal(m.points[m.polys], V[F])
    ascii_stl = tmp_path / "a.stl"
    ascii_stl.write_text("solid c\n" + "".join(
        "facet normal 0 0 0\nouter loop\n" + "".join("vertex %g %g %g\n" % tuple(V[i]) for i in t) + "endloop\nendfacet\n"
        for t in F) + "endsolid c\n")
    m2 = readers.read_mesh(str(ascii_stl))
    assert np.array_equal(m2.points[m2.polys], V[F])


def _vtp_encode(arr, mode, header, compress):
    """One DataArray payload as a VTK XML writer lays it out (header, then data; zlib in blocks)."""
    import base64
    import zlib
    raw = np.ascontiguousarray(arr).tobytes()
    hd = np.dtype(header)
    if compress:
        bs = 64  # tiny blocks so that several are exercised
        blocks = [raw[i:i + bs] for i in range(0, len(raw), bs)] or [b""]
        comp = [zlib.compress(b) for b in blocks]
        head = np.array([len(blocks), bs, len(blocks[-1]) % bs] + [len(c) for c in comp], hd).tobytes()
        body = b"".join(comp)
    else:
        head, body = np.array([len(raw)], hd).tobytes(), raw
    if mode == "raw":
        return head + body
    return base64.b64encode(head).decode() + base64.b64encode(body).decode()


@pytest.mark.parametrize("fmt,header,compress", [("ascii", "uint32", False), ("binary", "uint32", False),
                                                 ("binary", "uint64", True), ("appended-raw", "uint64", False),
                                                 ("appended-raw", "uint32", True), ("appended-base64", "uint64", True)])
def test_vtp_variants(tmp_path, fmt, header, compress):
    P, F = synth.bumpy_sphere(3, 0)
    lab = (np.arange(len(P)) % 5).astype(np.int32)
    arrays = [("Points", "Float32", 3, P.astype(np.float32)), ("connectivity", "Int64", 1, F.astype(np.int64).reshape(-1)),
              ("offsets", "Int64", 1, (np.arange(len(F), dtype=np.int64) + 1) * 3), ("RegionId", "Int32", 1, lab),
              ("cid", "Int32", 1, np.arange(len(F), dtype=np.int32))]
    appended, attr = b"", {}
    for name, typ, nc, a in arrays:
        if fmt == "ascii":
            attr[name] = ('format="ascii"', " ".join(repr(float(x)) if a.dtype.kind == "f" else str(int(x)) for x in a.reshape(-1)))
        elif fmt == "binary":
            attr[name] = ('format="binary"', _vtp_encode(a, "b64", header, compress))
        else:
            blk = _vtp_encode(a, "raw" if fmt == "appended-raw" else "b64", header, compress)
            blk = blk if isinstance(blk, bytes) else blk.encode()
            attr[name] = ('format="appended" offset="%d"' % len(appended), "")
            appended += blk

    def da(name, typ, nc, extra=""):
        return '<DataArray type="%s" Name="%s" NumberOfComponents="%d" %s %s>%s</DataArray>' % (
            typ, name, nc, attr[name][0], extra, attr[name][1])
    head = '<?xml version="1.0"?>\n<VTKFile type="PolyData" version="1.0" byte_order="LittleEndian" header_type="%s"%s>\n' % (
        {"uint32": "UInt32", "uint64": "UInt64"}[header], ' compressor="vtkZLibDataCompressor"' if compress else "")
    body = ('<PolyData><Piece NumberOfPoints="%d" NumberOfVerts="0" NumberOfLines="0" NumberOfStrips="0" NumberOfPolys="%d">'
            '<PointData Scalars="RegionId">%s</PointData><CellData>%s</CellData><Points>%s</Points>'
            '<Verts></Verts><Polys>%s%s</Polys></Piece></PolyData>') % (
        len(P), len(F), da("RegionId", "Int32", 1), da("cid", "Int32", 1), da("Points", "Float32", 3),
        da("connectivity", "Int64", 1), da("offsets", "Int64", 1))
    path = tmp_path / "m.vtp"
    with open(path, "wb") as f:
        f.write((head + body).encode())
        if fmt.startswith("appended"):
            f.write(b'<AppendedData encoding="%s">_' % (b"raw" if fmt == "appended-raw" else b"base64") + appended + b"</AppendedData>")
        f.write(b"</VTKFile>\n")
    m = readers.read_mesh(str(path))
    assert np.array_equal(m.points, P) and np.array_equal(m.polys, F)
    assert np.array_equal(m.point_data["RegionId"], lab) and np.array_equal(m.cell_data["cid"], np.arange(len(F)))


@pytest.mark.parametrize("enc", ["ASCII", "Base64Binary", "GZipBase64Binary"])
def test_gifti_surface(tmp_path, enc):
    import base64
    import zlib
    P, F = synth.bumpy_sphere(3, 1)

    def data(a):
        if enc == "ASCII":
            return " ".join(repr(float(x)) if a.dtype.kind == "f" else str(int(x)) for x in a.reshape(-1))
        raw = a.tobytes()
        return base64.b64encode(zlib.compress(raw) if enc == "GZipBase64Binary" else raw).decode()

    def da(intent, dtype, a):
        return ('<DataArray Intent="%s" DataType="%s" ArrayIndexingOrder="RowMajorOrder" Dimensionality="2" Dim0="%d" Dim1="3" '
                'Encoding="%s" Endian="LittleEndian" ExternalFileName="" ExternalFileOffset=""><MetaData/><Data>%s</Data></DataArray>'
                % (intent, dtype, len(a), enc, data(a)))
    path = tmp_path / "s.surf.gii"
    path.write_text('<?xml version="1.0" encoding="UTF-8"?><GIFTI Version="1.0" NumberOfDataArrays="2"><MetaData/><LabelTable/>'
                    + da("NIFTI_INTENT_POINTSET", "NIFTI_TYPE_FLOAT32", P.astype(np.float32))
                    + da("NIFTI_INTENT_TRIANGLE", "NIFTI_TYPE_INT32", F.astype(np.int32)) + "</GIFTI>")
    m = readers.read_mesh(str(path))
    assert np.array_equal(m.points, P) and np.array_equal(m.polys, F)


def test_nrrd_roundtrip_and_layout(tmp_path):
    a = np.random.default_rng(0).random((5, 6, 7, 4)).astype(np.float32)  # [views, y, x, C]
    for compress in (True, False):
        path = str(tmp_path / ("o%d.nrrd" % compress))
        nrrd.write(path, a, compress=compress)
        b, hdr = nrrd.read(path)
        assert np.array_equal(a, b) and b.dtype == np.float32
        assert hdr["sizes"] == "4 7 6 5" and hdr["dimension"] == "4" and hdr["space dimension"] == "3"
        assert hdr["kinds"] == "vector domain domain domain" and hdr["type"] == "float"
        assert hdr["encoding"] == ("gzip" if compress else "raw")
    # the reference's consumers reshape by element count (train_useg_flyby.py:170-174)
    flat = np.frombuffer(a.tobytes(), np.float32)
    assert np.array_equal(flat.reshape(5, 6, 7, 4), b)
    nrrd.write(str(tmp_path / "one.nrrd"), a[0])  # --concatenate 0: one 2-D vector image per view
    c, hdr = nrrd.read(str(tmp_path / "one.nrrd"))
    assert np.array_equal(c, a[0]) and hdr["sizes"] == "4 7 6" and hdr["space dimension"] == "2"


def test_id_codec_matches_reference_formula():
    from flyby_b200 import utils
    ids = np.array([-1, 0, 1, 253, 254, 255, 256, 65024, 65025, 200000, 16581374], np.int64)
    rgb = utils.EncodeIdsRGB(ids)
    for i, v in enumerate(ids):
        if v < 0:
            assert rgb[i].tolist() == [0, 0, 0]
        else:  # utils.py:616-618
            assert rgb[i].tolist() == [v % 255 + 1, int(v / 255) % 255, int(int(v / 255) / 255) % 255]
    assert np.array_equal(utils.DecodeIdsRGB(rgb), ids.astype(np.int32))  # utils.py:654, background -> -1


def test_rotation_matrix():
    from flyby_b200 import utils
    m = utils.GetTransform(90.0, [0, 0, 2])
    assert np.allclose(m[:3, :3] @ [1, 0, 0], [0, 1, 0]) and np.allclose(m[3], [0, 0, 0, 1])
    m = utils.GetTransform(120.0, [1, 1, 1])
    assert np.allclose(m[:3, :3] @ [1, 0, 0], [0, 1, 0]) and np.isclose(np.linalg.det(m[:3, :3]), 1)
    s = utils.Surf(np.eye(3, dtype=np.float32), np.array([[0, 1, 2]], np.int32))
    r = utils.RotateSurf(s, 37.0, [0.3, -1, 2])
    back = utils.RotateInverse(r, 37.0, [0.3, -1, 2])
    assert np.allclose(back.points, s.points, atol=1e-6)


def test_cli_option_surface_matches_reference():
    from flyby_b200 import fly_by_features as fbf
    parser = fbf.build_parser()
    actions = {o: a for a in parser._actions for o in a.option_strings}
    for flag, default in REFERENCE_FLAGS.items():
        assert flag in actions, flag
        assert actions[flag].default == default, (flag, actions[flag].default)
    args = parser.parse_args("--surf m.vtk --subdivision 2 --resolution 512 --radius 2 --use_z 1 --split_z 1 --out o.nrrd".split())
    assert (args.surf, args.subdivision, args.resolution, args.radius, args.use_z, args.split_z, args.out) == \
        ("m.vtk", 2, 512, 2.0, 1, 1, "o.nrrd")
    for bad in (["--subdivision", "2"], ["--surf", "a", "--dir", "b", "--spiral", "4"],
                ["--surf", "a", "--subdivision", "1", "--spiral", "4"]):
        with pytest.raises(SystemExit):
            parser.parse_args(bad)
    for unsupported in ("--model x", "--fiber 1", "--property p.txt", "--visualize 1"):
        a = parser.parse_args(("--surf m.vtk --subdivision 1 " + unsupported).split())
        with pytest.raises(SystemExit):
            fbf.main(a)


def test_cli_options_of_the_cxx_tool():
    """src/app/fly_by_features.xml:62-154: --planeSpacing / --planeScaleFactor are the C++ spellings of the plane
    options, --scaleFactor DIVIDES (the Python --scale_factor multiplies), --useCenterOfMass / --useMagnitude pick
    the normalisation, --applyRotation / --randomRotation rotate the surface as read, before centring and scaling
    (fly_by_features.cxx:310-347)."""
    from flyby_b200 import fly_by_features as fbf
    from flyby_b200 import utils
    parser = fbf.build_parser()
    base = "--surf m.vtk --subdivision 2 "
    a = parser.parse_args((base + "--planeScaleFactor 2 --planeSpacing 0.5 --scaleFactor 4 --useCenterOfMass 1 "
                                  "--useMagnitude 1").split())
    assert (a.plane_scale, a.plane_spacing, a.useCenterOfMass, a.useMagnitude) == (2.0, 0.5, 1, 1)
    assert fbf.cxx_scale_factor(a) == 0.25
    assert parser.parse_args((base + "--flyByCompose 0 --useOctree 1").split()).flyByCompose == 0
    d = parser.parse_args(base.split())
    assert d.flyByCompose is None and d.concatenate == 1
    assert fbf.cxx_scale_factor(d) is None and (d.scaleFactor, d.rotationAngle, d.rotationVector) == (-1, -1, [1.0, 0.0, 0.0])
    with pytest.raises(SystemExit):
        fbf.cxx_scale_factor(parser.parse_args((base + "--scaleFactor 2 --scale_factor 0.5").split()))
    surf = utils.Surf(np.array([[1, 0, 0], [0, 2, 0], [0, 0, 3]], np.float32), np.array([[0, 1, 2]], np.int32), {}, {})
    assert fbf.cxx_rotation(surf, d) is surf                                   # no rotation asked for
    r = fbf.cxx_rotation(surf, parser.parse_args((base + "--applyRotation 1 --rotationAngle 90 --rotationVector 0,0,1").split()))
    assert np.allclose(r.points, [[0, 1, 0], [-2, 0, 0], [0, 0, 3]], atol=1e-6)  # RotateWXYZ(90, z): x -> y, y -> -x
    rr = fbf.cxx_rotation(surf, parser.parse_args((base + "--randomRotation 1").split()), rng=np.random.default_rng(1))
    assert np.allclose(np.linalg.norm(rr.points, axis=1), [1, 2, 3], atol=1e-5) and not np.allclose(rr.points, surf.points)
    ra = fbf.cxx_rotation(surf, parser.parse_args((base + "--applyRotation 1").split()), rng=np.random.default_rng(2))
    assert np.allclose(ra.points[0, 0], 1.0, atol=1e-6)                        # random angle about the default axis x


def test_flybygenerator_signature_matches_reference():
    import inspect
    from flyby_b200 import fly_by_features as fbf
    p = list(inspect.signature(fbf.FlyByGenerator.__init__).parameters)
    assert p[:7] == ["self", "sphere", "resolution", "visualize", "use_z", "split_z", "rescale_features"]
    g = list(inspect.signature(fbf.FlyByGenerator.getFlyBy).parameters)
    assert g == ["self", "sphere_points", "view_up_points", "focal_points"]
    for name in ("addActor", "removeActor", "removeActors", "getFlyBy"):
        assert hasattr(fbf.FlyByGenerator, name)
    for name in ("ReadSurf", "GetUnitSurf", "ScaleSurf", "CreateIcosahedron", "CreateSpiral", "GetNormalsActor",
                 "GetPointIdMapActor", "ExtractPointFeatures", "GetImage", "ComputeNormals", "RandomRotation"):
        assert hasattr(fbf, name), name  # callers use fbf.ReadSurf etc. (predict_v3.py:23-36)


def test_no_stream_waits_for_an_event_recorded_on_itself():
    """Source lint for the cross-stream hand-offs of the CUDA library: a cudaStreamWaitEvent that directly follows
    the cudaEventRecord of the same event must name a different stream -- recording on the waiting stream orders
    nothing (that slip once let the side stream's point normals race the main stream's unit surface)."""
    import glob
    import re
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "fly-by-cnn_b200", "csrc")
    rec = re.compile(r"cudaEventRecord\(\s*([^,]+?)\s*,\s*([^)]+?)\s*\)")
    wait = re.compile(r"cudaStreamWaitEvent\(\s*([^,]+?)\s*,\s*([^,]+?)\s*,")
    side_wait = re.compile(r"FB_SIDE_WAIT\(\s*c\s*,\s*([^)]+?)\s*\)")  # ctx.h: the side stream waits (a fork, when capturing)
    pairs = 0
    for path in glob.glob(os.path.join(root, "*.cu")):
        lines = open(path).read().splitlines()
        for i, ln in enumerate(lines):
            m, ms = wait.search(ln), side_wait.search(ln)
            if not m and not ms:
                continue
            stream, event = (m.group(1), m.group(2)) if m else ("c->side_stream", ms.group(1))
            for back in lines[max(0, i - 3):i][::-1]:
                r = rec.search(back)
                if r and r.group(1) == event:
                    pairs += 1
                    assert r.group(2) != stream and not (ms and r.group(2) == "ss"), \
                        "%s:%d waits on %s for an event recorded on it" % (path, i + 1, stream)
                    break
    assert pairs >= 4  # unit surface, sort, front buffer, chunked copy-out


def test_nrrd_every_written_type_reads_back_and_crlf_headers(tmp_path):
    """read() must know every type spelling write() emits (int64 -> 'long long'), and CRLF headers."""
    from flyby_b200 import nrrd
    rng = np.random.default_rng(0)
    for dt in nrrd._NAMES:
        a = (rng.random((3, 4, 5, 2)) * 100).astype(dt)
        path = str(tmp_path / ("t_%s.nrrd" % np.dtype(dt).name))
        nrrd.write(path, a)
        b, hdr = nrrd.read(path)
        assert b.dtype == np.dtype(dt) and np.array_equal(a, b), dt
    for name in ("long", "unsigned long long", "uint16", "uint32", "int64_t"):
        assert name in nrrd._TYPES or name == "long"
    a = rng.random((2, 3, 1)).astype(np.float32)
    nrrd.write(str(tmp_path / "lf.nrrd"), a, compress=False)
    raw = open(str(tmp_path / "lf.nrrd"), "rb").read()
    end = raw.find(b"\n\n")
    open(str(tmp_path / "crlf.nrrd"), "wb").write(raw[:end].replace(b"\n", b"\r\n") + b"\r\n\r\n" + raw[end + 2:])
    b, hdr = nrrd.read(str(tmp_path / "crlf.nrrd"))
    assert np.array_equal(a, b) and hdr["encoding"] == "raw"
    # the C++ tool's image type: RGB pixels of doubles with a spacing (fly_by_features.cxx:1047-1068)
    img = nrrd.Image(rng.random((4, 6, 6, 3)), spacing=[0.5, 0.5, 1.0], kind="RGB-color")
    nrrd.write(str(tmp_path / "rgb.nrrd"), img)
    b, hdr = nrrd.read(str(tmp_path / "rgb.nrrd"))
    assert b.dtype == np.float64 and hdr["type"] == "double" and hdr["kinds"].split()[0] == "RGB-color"
    assert hdr["sizes"] == "3 6 6 4" and "(0.5,0,0) (0,0.5,0) (0,0,1)" in hdr["space directions"]


def test_bench_reference_arm_line():
    """`bench.py --impl reference` (the driver's CPU arm: the oracle port of the reference's locator loop, no GPU needed)
    prints one JSON line with the contract's keys, the same metric / unit / config as the CUDA arm, and no GPU work."""
    import json
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.strip().splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "views/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "views/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0 and "workload" in d["config"] and d["n_gpus"] == 1 and d["steps"] == 1
