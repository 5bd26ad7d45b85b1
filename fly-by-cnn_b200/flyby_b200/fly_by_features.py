"""Drop-in for the reference's ``fly_by_features.py`` (reference src/py/fly_by_features.py).

``FlyByGenerator`` keeps the reference's constructor and methods
(src/py/fly_by_features.py:18-154) and ``main`` / the command line keep the
reference's option surface (:156-493).  The renderer behind it is the CUDA ray
caster of this repo (the engine of the reference's C++ tool with ``--useOctree``,
src/app/fly_by_features.cxx:482-852): orthographic rays from the tangent plane at
every viewpoint, nearest hit by LBVH traversal, barycentric normals and depth.

Differences a caller can see (DESIGN.md "Deviations"):
 * depth is the distance from the view plane along the ray (ray-cast engine), not
   the perspective z-buffer of the OpenGL path; images are not flipped by a view-up;
 * normals are float32, not 8-bit quantised; ``normal_encoding`` selects the value
   range: "byte255" (default here, the range the Python reference produces),
   "unit01" (the C++ tool's n*0.5+0.5) or "signed" (rescale_features);
 * options that need an OpenGL window, a Keras model or fibre tubes are rejected.

``FlyByGenerator(..., projection="perspective")`` is the raster-compatible mode: the
pinhole camera of the OpenGL path (view angle 30 degrees, view-up (0,0,-1), rows bottom to top, depth along
the viewing direction, clipping range as vtkRenderer::ResetCameraClippingRange derives it from the bounds,
8-bit colours) traced by the ray caster, for models trained on the reference's raster images.
"""
import argparse
import glob
import os
import sys
import time

import numpy as np

from . import utils
from ._cabi import Context
from .utils import *  # noqa: F401,F403  (the reference re-exports utils through this module)


class FlyByGenerator():
    def __init__(self, sphere=None, resolution=224, visualize=False, use_z=False, split_z=False,
                 rescale_features=False, normal_encoding=None, plane_scale=1.0, plane_spacing=1.0,
                 z_far=None, device=0, projection="orthographic", view_angle=30.0):
        if visualize:
            raise NotImplementedError("visualize needs an interactive VTK window; not available in flyby_b200")
        self.sphere = sphere
        self.visualize = False
        self.resolution = int(resolution)
        self.use_z = bool(use_z)
        self.split_z = bool(split_z)
        self.rescale_features = bool(rescale_features)
        if projection not in ("orthographic", "perspective"):
            raise ValueError("projection must be 'orthographic' or 'perspective'")
        self.projection = projection
        self.view_angle = float(view_angle)
        # perspective = what the OpenGL read-back returns: uint8 colours (rescaled to [-1, 1] on request)
        self.normal_encoding = normal_encoding or ("uint8" if projection == "perspective" else
                                                   ("signed" if rescale_features else "byte255"))
        self.plane_scale = float(plane_scale)
        self.plane_spacing = float(plane_spacing)
        self.z_far = z_far
        self.device = device
        self.actors = []
        self.ctx = Context(device)
        self.last_cell_ids = None  # int32 [V, res, res] hit cell ids of the last getFlyBy (-1 = background)
        self._built_for = None     # (points, polys) the context currently holds a build of
        self._views_for = None
        self._scalars = None       # (float32 [V,S], mode, zero): written by the render behind the image channels
        # vtkPolyDataNormals runs its consistency pass in the reference (utils.py:493-501): "fix" reverses
        # inconsistently wound cells as VTK does; costs nothing on a consistent mesh
        self.consistency = "fix"

    def setPointScalars(self, feats, mode="vertex0", zero=0.0):
        """Per-point arrays [V,S] that getFlyBy appends to every pixel in the same render pass (the reference
        renders a second, point-id image and looks the values up in Python: fly_by_features.py:334-402,
        utils.py:639-684).  mode "vertex0" = the reference's lookup, "barycentric" = the C++ tool's
        interpolation (fly_by_features.cxx:766-793).  None removes them."""
        self._scalars = None if feats is None else (np.ascontiguousarray(feats, np.float32).reshape(len(feats), -1), mode, float(zero))

    # --- reference API ---------------------------------------------------------
    def removeActor(self, actor):
        self.actors = [a for a in self.actors if a is not actor]

    def removeActors(self):
        self.actors = []

    def addActor(self, actor):
        if actor is None:
            return
        if not isinstance(actor, utils.Actor):
            raise TypeError("addActor expects GetNormalsActor / GetPointIdMapActor / GetCellIdMapActor output")
        self.actors.append(actor)

    def _mesh(self):
        if not self.actors:
            raise RuntimeError("getFlyBy: no actor added")
        kinds = {a.kind for a in self.actors}
        if len(kinds) != 1:
            raise NotImplementedError("actors of different kinds cannot be rendered together")
        if len(self.actors) == 1:
            s = self.actors[0].surf
            return s.points, s.polys, self.actors[0].kind
        if "point_id" in kinds or "cell_id" in kinds:
            raise NotImplementedError("id maps are defined for a single surface")
        pts, tris, off = [], [], 0
        for a in self.actors:  # several surfaces in one scene: nearest surface wins, as in a render
            pts.append(a.surf.points)
            tris.append(a.surf.polys + off)
            off += len(a.surf.points)
        return np.concatenate(pts), np.concatenate(tris), self.actors[0].kind

    def getFlyBy(self, sphere_points=None, view_up_points=None, focal_points=None):
        if view_up_points is not None or focal_points is not None:
            raise NotImplementedError("the ray-cast engine looks at the origin with the sphere north as up "
                                      "(src/app/fly_by_features.cxx:656-688); custom view-up / focal points "
                                      "belong to the OpenGL camera path")
        points, polys, kind = self._mesh()
        ctx = self.ctx
        # meshes handed to the generator are already normalised by GetUnitSurf (as in the reference); the upload and
        # the build are skipped when the context still holds this very mesh (same arrays, e.g. n_rotations of views)
        if self._built_for is None or self._built_for[0] is not points or self._built_for[1] is not polys:
            ctx.set_mesh(np.ascontiguousarray(points, np.float32), np.ascontiguousarray(polys, np.int32),
                         skip_normalise=True, consistency=self.consistency)
            ctx.build()
            self._built_for = (points, polys)
        if sphere_points is None:
            if self.sphere is None:
                raise RuntimeError("getFlyBy: no viewpoints")
            sphere_points = self.sphere.points
            radius = self.sphere.radius
        else:
            sphere_points = np.asarray(sphere_points, dtype=np.float32).reshape(-1, 3)
            radius = float(np.linalg.norm(sphere_points[0]))
        ctx.views_custom(sphere_points, radius)
        res = self.resolution
        if self.projection == "perspective":
            return self._fly_by_perspective(points, polys, kind, sphere_points)
        if kind != "normals":
            opts = ctx.render_opts(use_z=0, split_z=0, plane_scale=self.plane_scale, plane_spacing=self.plane_spacing)
            ids = np.empty((len(sphere_points), res, res), np.int32)
            ctx.render_into(res, opts, None, ids)
            self.last_cell_ids = ids
            if kind == "point_id":  # colour = id of the cell's first point (utils.py:604-621)
                pid = np.where(ids >= 0, polys[np.maximum(ids, 0), 0], -1)
                return utils.EncodeIdsRGB(pid)
            return utils.EncodeIdsRGB(ids)
        z_far = self.z_far if self.z_far is not None else radius + 1.0
        z_scale = 1.0
        if self.use_z and (self.rescale_features or not self.split_z):
            z_scale = 1.0 / z_far  # fly_by_features.py:138-146
        opts = ctx.render_opts(use_z=self.use_z, split_z=self.split_z, normal_encoding=self.normal_encoding,
                               plane_scale=self.plane_scale, plane_spacing=self.plane_spacing, z_scale=z_scale)
        if self._scalars is not None:
            ctx.set_point_scalars(self._scalars[0], mode=self._scalars[1], zero=self._scalars[2])
        else:
            ctx.set_point_scalars(None)
        feat, ids = ctx.render(res, opts)
        self.last_cell_ids = ids
        return feat


    def _fly_by_perspective(self, points, polys, kind, sphere_points):
        """fly_by_features.py:85-152 with the ray caster behind the pinhole camera."""
        ctx, res = self.ctx, self.resolution
        n = len(sphere_points)
        lo, hi = points.min(0).astype(np.float64), points.max(0).astype(np.float64)
        feats = []
        ids_all = np.empty((n, res, res), np.int32)
        for i in range(n):  # the clipping range is reset per camera position (:109)
            z_near, z_far = clipping_range(sphere_points[i], (lo[0], hi[0], lo[1], hi[1], lo[2], hi[2]))
            if kind != "normals":
                opts = ctx.render_opts(use_z=0, split_z=0)
                _, ids, _ = ctx.render_perspective(res, opts, self.view_angle, z_near, z_far, i, i + 1)
                ids_all[i] = ids[0]
                continue
            z_scale = 1.0 / z_far if self.use_z and (self.rescale_features or not self.split_z) else 1.0
            opts = ctx.render_opts(use_z=self.use_z, split_z=self.split_z, normal_encoding=self.normal_encoding,
                                   z_scale=z_scale)
            f, ids, _ = ctx.render_perspective(res, opts, self.view_angle, z_near, z_far, i, i + 1)
            if self.rescale_features and self.normal_encoding == "uint8":
                hit = ids[0] >= 0
                if self.use_z and not self.split_z:  # colours were multiplied by z / z_far in the kernel: redo on the host
                    raise NotImplementedError("rescale_features with use_z=1, split_z=0 in perspective mode: "
                                              "pass normal_encoding='signed'")
                f[0][..., :3] = np.where(hit[..., None], 2.0 * (f[0][..., :3] / 255.0) - 1.0, -1.0)  # :119-120
            feats.append(f[0])
            ids_all[i] = ids[0]
        self.last_cell_ids = ids_all
        if kind == "point_id":
            return utils.EncodeIdsRGB(np.where(ids_all >= 0, polys[np.maximum(ids_all, 0), 0], -1))
        if kind != "normals":
            return utils.EncodeIdsRGB(ids_all)
        return np.stack(feats)


def clipping_range(position, bounds, expansion=0.5, near_tolerance=0.001):
    """vtkRenderer::ResetCameraClippingRange(bounds) for a camera at `position` looking at the origin
    (VTK 9 Rendering/Core/vtkRenderer.cxx, restated from memory: distance range of the eight corners of the
    bounding box along the view direction, 1 % + expansion * span breathing room, near at least
    near_tolerance * far; 0.001 is VTK's value for depth buffers deeper than 16 bits)."""
    p = np.asarray(position, np.float64)
    vn = p / np.linalg.norm(p)  # view plane normal: from the focal point (origin) to the camera
    a, b, c = -vn
    d = -(a * p[0] + b * p[1] + c * p[2])
    dist = [a * bounds[i] + b * bounds[2 + j] + c * bounds[4 + k] + d for k in range(2) for j in range(2) for i in range(2)]
    r0, r1 = min(dist), max(max(dist), 1e-18)
    r0 = 0.99 * r0 - (r1 - r0) * expansion
    r1 = 1.01 * r1 + (r1 - r0) * expansion
    if r0 >= r1:
        r0 = 0.01 * r1
    if r0 < near_tolerance * r1:
        r0 = near_tolerance * r1
    return float(r0), float(r1)


def _find_surfaces(args):
    """reference src/py/fly_by_features.py:158-213"""
    filenames = []
    if args.surf:
        filenames.append({"surf": args.surf, "out": args.out})
    else:
        surf_filenames = []
        replace_dir_name = ""
        if args.dir:
            replace_dir_name = args.dir
            normpath = os.path.normpath("/".join([args.dir, '**', '*']))
            for surf in glob.iglob(normpath, recursive=True):
                if os.path.isfile(surf) and True in [ext in surf for ext in [".vtk", ".obj", ".stl", ".off"]]:
                    surf_filenames.append(os.path.realpath(surf))
            replace_dir_name = os.path.realpath(args.dir)
        elif args.csv:
            import pandas
            replace_dir_name = args.csv_root_path
            df = pandas.read_csv(args.csv)
            for _, row in df.iterrows():
                surf_filenames.append(row["surf"])
        for surf in surf_filenames:
            dir_filename = os.path.splitext(surf.replace(replace_dir_name, ''))[0] + ".nrrd"
            out = os.path.normpath("/".join([args.out, dir_filename]))
            if args.uuid:
                import uuid
                out = out.replace(".nrrd", "-" + str(uuid.uuid1()).split('-')[0] + ".nrrd")
            if not os.path.exists(os.path.dirname(out)):
                os.makedirs(os.path.dirname(out))
            if args.ow or not os.path.exists(out):
                filenames.append({"surf": surf, "out": out})
    if args.n_rotations:
        for fobj in filenames.copy():
            for i in range(args.n_rotations):
                out_name, out_ext = os.path.splitext(os.path.normpath(fobj["out"]))
                rot = {"surf": fobj["surf"], "out": out_name + "_rot" + str(i) + out_ext, "rotate": True}
                if args.ow or not os.path.exists(rot["out"]):
                    filenames.append(rot)
    return filenames


def _write(out_np, path, concatenate, suffix="", cxx_spacing=None):
    """reference src/py/fly_by_features.py:404-424.  cxx_spacing = planeSpacing selects the C++ tool's output
    contract instead (src/app/fly_by_features.cxx:1047-1110): itk::Image<itk::RGBPixel<double>, 3> of size
    (res, res, views) with spacing (planeSpacing, planeSpacing, 1) -- `type: double`, component kind RGB-color."""
    def image(a):
        if cxx_spacing is None:
            return utils.GetImage(a)
        from . import nrrd  # values are the float32 results widened to double (the C++ tool computes them in double)
        return nrrd.Image(np.asarray(a, np.float64), spacing=[float(cxx_spacing)] * 2 + [1.0] * (np.ndim(a) - 3),
                          kind="RGB-color")
    if not concatenate:
        if not os.path.exists(path):
            os.makedirs(path)
        for i in range(out_np.shape[0]):
            name = os.path.join(path, str(i) + suffix + ".nrrd")
            print("Writing:", name)
            utils.WriteImage(image(out_np[i]), name)
    else:
        print("Writing:", path)
        utils.WriteImage(image(out_np), path)


def cxx_scale_factor(args):
    """--scale_factor multiplies (src/py/utils.py:285); the C++ tool's --scaleFactor divides
    (src/app/fly_by_features.cxx:448) and -1 means "derive it from the shape" (:426)."""
    sf = getattr(args, "scaleFactor", -1)
    if sf is not None and sf != -1:
        if sf == 0:
            raise SystemExit("--scaleFactor 0: the C++ tool divides by it")
        if args.scale_factor is not None:
            raise SystemExit("--scale_factor and --scaleFactor are two spellings of one setting: give one")
        return 1.0 / sf
    return args.scale_factor


def cxx_rotation(surf, args, rng=None):
    """The C++ tool's --randomRotation / --applyRotation --rotationAngle --rotationVector
    (src/app/fly_by_features.cxx:310-347): the rotation is applied to the surface as read, BEFORE centring and
    scaling (the Python --random_rotation rotates the unit surface, fly_by_features.py:286-287).  Axis of the
    random rotation: a normalised Gaussian vector, angle uniform in [0, 360); --rotationAngle -1 draws the angle."""
    if not (getattr(args, "applyRotation", 0) or getattr(args, "randomRotation", 0)):
        return surf
    rng = rng or np.random.default_rng()
    angle, vec = getattr(args, "rotationAngle", -1), list(getattr(args, "rotationVector", [1.0, 0.0, 0.0]))
    if getattr(args, "randomRotation", 0):
        v = rng.normal(size=3)
        vec = list(v / np.linalg.norm(v))
        angle = float(rng.uniform(0.0, 360.0))
    if angle == -1:
        angle = float(rng.uniform(0.0, 360.0))
    if len(vec) != 3:
        raise SystemExit("--rotationVector needs three components")
    if getattr(args, "verbose", 0):
        print("Apply rotation: angle:", angle, "vector:", vec)
    return utils.RotateSurf(surf, angle, vec)


def main(args):
    for name in ("fiberBundle", "fiber", "property", "view_features", "model", "visualize"):
        if getattr(args, name, None):
            raise SystemExit("--%s is not supported by flyby_b200 (needs VTK tube filters / colour maps / Keras / "
                             "an OpenGL window); see DESIGN.md" % name)
    cxx_spacing = None
    if getattr(args, "flyByCompose", None) is not None:
        # the C++ tool's spelling: its output contract applies (fly_by_features.cxx:802-807,1047-1110) -- one RGB image
        # of doubles per view, pixel = (n * 0.5 + 0.5) * depth, stacked into a volume when flyByCompose is 1
        args.concatenate = 1 if args.flyByCompose else 0
        args.use_z, args.split_z, args.rescale_features = 1, 0, 0
        args.normal_encoding = "unit01"
        cxx_spacing = getattr(args, "plane_spacing", 1.0)
    filenames = _find_surfaces(args)
    device = getattr(args, "device", 0)
    if args.subdivision:
        sphere = utils.CreateIcosahedron(args.radius, args.subdivision, device=device)
    else:
        sphere = utils.CreateSpiral(args.radius, args.spiral, args.turns, device=device)
    kw = dict(plane_scale=getattr(args, "plane_scale", 1.0), plane_spacing=getattr(args, "plane_spacing", 1.0),
              normal_encoding=getattr(args, "normal_encoding", None), device=device)
    if cxx_spacing is not None:
        kw["z_far"] = 1.0  # the C++ tool multiplies by the depth itself, not by depth / z_far
    flyby = FlyByGenerator(sphere, args.resolution, use_z=args.use_z, split_z=args.split_z,
                           rescale_features=args.rescale_features, **kw)
    flyby_features = None
    if args.point_features or args.out_point_id:
        flyby_features = FlyByGenerator(sphere, args.resolution, **kw)

    for fobj in filenames:
        if args.verbose:
            print("Number of sphere points:", sphere.GetNumberOfPoints())
            print("Reading:", fobj["surf"])
        surf = utils.ReadSurf(fobj["surf"])
        surf = cxx_rotation(surf, args)
        surf = utils.GetUnitSurf(surf, args.translate, cxx_scale_factor(args), device=device,
                                 centre_mode=1 if getattr(args, "useCenterOfMass", 0) else 0,
                                 scale_mode=1 if getattr(args, "useMagnitude", 0) else 0)
        if args.random_rotation or fobj.get("rotate"):
            surf, rotationAngle, rotationVector = utils.RandomRotation(surf)
            if args.verbose:
                print("angle:", rotationAngle, "vector:", rotationVector)
            if args.save_rotation:
                out_filename = os.path.splitext(fobj["out"])[0] + "_transform.npy"
                print("Saving rotation:", out_filename)
                np.save(out_filename, utils.GetTransform(rotationAngle, rotationVector).ravel())
        flyby.addActor(utils.GetNormalsActor(surf))
        # --point_features ... --point_features_concat 1: the arrays ride along in the render (vertex-0 lookup, the
        # reference's semantics), instead of a second id-map render + RGB decode + host concatenate
        fused_names = []
        if args.point_features and args.point_features_concat and args.extract_components is None:
            fused_names = list(args.point_features)
            flyby.setPointScalars(np.concatenate([utils.PointFeatureArray(surf, n) for n in fused_names], axis=1),
                                  mode="vertex0", zero=args.zero)
        else:
            flyby.setPointScalars(None)
        out_np = flyby.getFlyBy()
        if args.extract_components is not None:
            out_np = out_np[:, :, :, args.extract_components[0]:args.extract_components[1]]
            print(out_np.shape)
        if flyby_features is not None and (args.out_point_id or not fused_names):
            flyby_features.addActor(utils.GetPointIdMapActor(surf))
            out_point_ids_rgb_np = flyby_features.getFlyBy()
            if args.out_point_id:
                if not args.concatenate:
                    _write(out_point_ids_rgb_np, fobj["out"], 0, "_point_id_map")
                else:
                    name = os.path.splitext(fobj["out"])
                    _write(out_point_ids_rgb_np, name[0] + "_point_id_map" + name[1], 1)
            if args.point_features and not fused_names:
                for point_features_name in args.point_features:
                    print("Extracting:", point_features_name)
                    out_features_np = utils.ExtractPointFeatures(surf, out_point_ids_rgb_np, point_features_name,
                                                                 args.zero, device=device)
                    if args.point_features_concat:
                        out_np = np.concatenate([out_np, out_features_np], axis=-1)
                    elif not args.concatenate:
                        _write(out_features_np, fobj["out"], 0, "_" + point_features_name)
                    else:
                        name = os.path.splitext(fobj["out"])
                        _write(out_features_np, name[0] + "_" + point_features_name + name[1], 1)
            flyby_features.removeActors()
        _write(out_np, fobj["out"], args.concatenate, cxx_spacing=cxx_spacing)
        flyby.removeActors()


def build_parser():
    """Option surface of reference src/py/fly_by_features.py:432-489 (same names, defaults and groups)."""
    parser = argparse.ArgumentParser(description='Fly-by feature extraction on a B200 (drop-in for fly_by_features.py)',
                                     formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    input_group = parser.add_argument_group('Input parameters')
    input_params = input_group.add_mutually_exclusive_group(required=True)
    input_params.add_argument('--surf', type=str, help='Target surface/mesh')
    input_params.add_argument('--dir', type=str, help='Input directory with 3D models')
    input_params.add_argument('--csv', type=str, help='Input csv with column "surf"')
    input_group.add_argument('--csv_root_path', type=str, help='CSV rooth path for replacement', default="")
    input_group.add_argument('--model', type=str, help='(unsupported) directory with a saved Keras model', default=None)
    input_group.add_argument('--random_rotation', type=bool, help='Apply a random rotation', default=False)
    input_group.add_argument('--n_rotations', type=int, help='Number of additional random rotations', default=0)
    input_group.add_argument('--scale_factor', type=float, help='Scale the surface by this vale', default=None)
    input_group.add_argument('--bounds', type=float, nargs="+", help='(unused, as in the reference)', default=None)
    input_group.add_argument('--translate', nargs="+", type=float, help='Center the surface at this point', default=None)
    input_group.add_argument('--fiberBundle', type=bool, help='(unsupported) input is a fiber tract', default=False)
    input_group.add_argument('--fiber', type=bool, help='(unsupported) input surface is a fiber', default=False)
    input_group.add_argument('--nbFiber', type=int, help='(unsupported)', default=0.05)
    input_group.add_argument('--subject', type=str, help='(unsupported)', default=None)

    features_group = parser.add_argument_group('Shape features/property to extract')
    features_group.add_argument('--extract_components', type=int, nargs='+', help='Which components to extract', default=None)
    features_group.add_argument('--norm_shader', type=int, help='(unused, as in the reference)', default=1)
    features_group.add_argument('--split_z', type=int, help='1 to split the z buffer as a separate channel. Otherwise, the normals are scaled by z buffer to create an rgb image.', default=0)
    features_group.add_argument('--use_z', type=int, help='1 to use the z_buffer and compute the depth buffer (distance of the view plane to the shape at every location).', default=1)
    features_group.add_argument('--rescale_features', type=int, help='1 to rescale features (Normals, Depth map) between 0 and 1', default=0)
    features_group.add_argument('--property', type=str, help='(unsupported) input property file', default=None)
    features_group.add_argument('--point_features', nargs='+', type=str, help='Name of array in point data to extract features. If name is coords or points, it extracts the x,y,z coordinates', default=None)
    features_group.add_argument('--point_features_concat', type=int, help='Concatenate point features to the fly_by_features', default=0)
    features_group.add_argument('--zero', type=float, help="Default zero value when extracting properties. This is used when there is no 'collision' with the surface", default=0)
    features_group.add_argument('--view_features', type=str, help='(unsupported) name of array in point data to visualize', default=None)

    sphere_params = parser.add_argument_group('Sampling parameters')
    sphere_params_sampling = sphere_params.add_mutually_exclusive_group(required=True)
    sphere_params_sampling.add_argument('--subdivision', type=int, help='Number of subdivisions for icosahedron')
    sphere_params_sampling.add_argument('--spiral', type=int, help='Number of samples along the spherical spiral')
    sphere_params.add_argument('--turns', type=int, default=4, help='Number of spiral turns')
    sphere_params.add_argument('--resolution', type=int, help='Image resolution', default=256)
    sphere_params.add_argument('--radius', type=float, help='Radius of the sphere for the view points', default=4)

    training_orientation = parser.add_argument_group('training orientation')
    training_orientation.add_argument('--save_rotation', type=int, help="save the orientation transform", default=0)

    visu_params = parser.add_argument_group('Visualize')
    visu_params.add_argument('--visualize', type=int, default=0, help='(unsupported) visualize the sampling')

    output_params = parser.add_argument_group('Output parameters')
    output_params.add_argument('--out', type=str, help='Output filename or directory', default="out.nrrd")
    output_params.add_argument('--out_point_id', type=int, help='Output the point id map. Point ids are encoded in the rgb components', default=0)
    output_params.add_argument('--uuid', type=bool, help='Use uuid to name the outputs', default=False)
    output_params.add_argument('--ow', type=int, help='Overwrite outputs', default=1)
    output_params.add_argument('--concatenate', type=int, help='0 for multiple output files, 1 for single output file', default=1)
    output_params.add_argument('--verbose', type=int, help='Print messages', default=0)

    # options of the reference's C++ tool (src/app/fly_by_features.xml:62-76) and of this implementation
    engine = parser.add_argument_group('Ray-cast engine')
    engine.add_argument('--plane_scale', '--planeScaleFactor', dest='plane_scale', type=float, default=1.0,
                        help='planeScaleFactor: side length of the view plane')
    engine.add_argument('--plane_spacing', '--planeSpacing', dest='plane_spacing', type=float, default=1.0, help='planeSpacing')
    engine.add_argument('--flyByCompose', type=int, default=None,
                        help='C++ tool: 1 = one stacked volume, 0 = one file per view (sets --concatenate)')
    engine.add_argument('--useOctree', type=int, default=0,
                        help='C++ tool: choice of its CPU search structure; accepted, no effect (the results are the same)')
    engine.add_argument('--scaleFactor', type=float, default=-1,
                        help='C++ tool: the centred shape is DIVIDED by this value (-1: derive it from the shape)')
    engine.add_argument('--useCenterOfMass', type=int, default=0, help='C++ tool: centre on the centre of mass of the points')
    engine.add_argument('--useMagnitude', type=int, default=0, help='C++ tool: scale by the largest point magnitude')
    engine.add_argument('--randomRotation', type=int, default=0, help='C++ tool: random rotation before centring and scaling')
    engine.add_argument('--applyRotation', type=int, default=0, help='C++ tool: apply --rotationAngle about --rotationVector before centring and scaling')
    engine.add_argument('--rotationAngle', type=float, default=-1, help='C++ tool: angle in degrees (-1: random)')
    engine.add_argument('--rotationVector', type=lambda s: [float(x) for x in s.split(',')], default=[1.0, 0.0, 0.0],
                        help='C++ tool: rotation axis, comma separated')
    engine.add_argument('--normal_encoding', type=str, default=None, choices=[None, 'byte255', 'unit01', 'signed'],
                        help='value range of the normal channels (default byte255, signed with --rescale_features)')
    engine.add_argument('--device', type=int, default=0, help='CUDA device')
    return parser


if __name__ == '__main__':
    start_time = time.time()
    main(build_parser().parse_args())
    print("FlyBy time:", round((time.time() - start_time) * 1000.0), "ms", file=sys.stderr)
