#!/usr/bin/env python
"""bench.py -- views/s (and Mrays/s) of the fly-by hot path at resolution 512.

Workload (BASELINE.json configs[2], the res-512 configuration the metric is
quoted on, one GPU's share): synthetic 200 000-triangle closed surfaces,
icosahedron subdivision 3 -> 92 views, resolution 512, radius 2, use_z 1,
split_z 1 (C = 4).  One STEP = one mesh through the whole path: normalise to
the unit sphere, point normals, LBVH build, 92 x 512 x 512 rays, feature + cell
id tensors written.  Sharding is by mesh (weak scaling): every rank processes
its own meshes, there is no data-path collective.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

value  : whole-job views/s with the mesh already resident in HBM.
e2e    : same metric through the C-ABI with HOST buffers (pinned): H2D of the
         mesh and D2H of the feature tensor (the reference's --out product)
         inside the timed region; the run that also copies the optional
         hit-cell-id map out is reported beside it (e2e.with_cell_ids).
roofline: the render kernel against the measured HBM peak, algorithmic bytes
          / CUDA-event duration of the kernel.
cpu_baseline: the oracle's vtkCellLocator-style port timed on this box's host
         cores on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "fly-by-cnn_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

LM = 100            # 20*LM^2 = 200 000 triangles
SUBDIVISION = 3     # 92 views
RES = 512
RADIUS = 2.0
PLANE_SCALE = 1.0   # reference default planeScaleFactor (fly_by_features.xml:76)
MESH_POOL = 4       # distinct meshes, cycled; every rank cycles through the SAME four shapes, starting at its own offset
METRIC = "views/s at res 512 (fly-by feature extraction; Mrays/s = views/s * 0.262144)"


def workload_config(n_views):
    return {"workload": "BASELINE configs[2] share of one GPU: 200000-triangle synthetic surfaces, icosahedron "
                        "subdivision 3 (%d views), resolution 512, radius 2, use_z 1, split_z 1, plane scale 1.0; "
                        "one step = one mesh (unit-surf + normals + LBVH build + %d rays + tensor write)"
                        % (n_views, n_views * RES * RES),
            "triangles": 20 * LM * LM, "views": n_views, "resolution": RES, "channels": 4,
            "sharding": "by mesh, no collective; every rank's share is the same four mesh shapes (seeds 0..3) in rotated order",
            "l2": "explicit 256 MiB L2 flush before every step, on the step's stream (inside the timed region of the "
                  "pipelined loop; outside the per-step events of the single-context loop)"}


def algorithmic_bytes(n_rays, C, T, V):
    """DESIGN.md / SURVEY 8(d): compulsory traffic of one render launch."""
    return n_rays * (4 * C + 4) + T * 36 + V * 12 + (2 * T - 1) * 32


class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []
        self.stop_flag = False

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,power.draw,temperature.gpu")
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5)
                parts = [x.strip() for x in out.stdout.strip().split(",")]
                if len(parts) >= 6:
                    self.rows.append(parts)
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(int(r[0]) for r in self.rows if r[0].isdigit())
        mx = max(int(r[1]) for r in self.rows if r[1].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(r[2 + k].lower().startswith("active") for r in self.rows)]
        out = {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": reasons,
               "samples": len(self.rows)}
        try:
            out["power_w_max"] = max(float(r[6]) for r in self.rows if len(r) > 6)
            out["temp_c_max"] = max(float(r[7]) for r in self.rows if len(r) > 7)
        except ValueError:
            pass
        return out


def run_reference(args):
    """The reference's CPU implementation of the path: vtkCellLocator-style bucket grid (VTK default
    bucket rule) + VTK segment/triangle test, restated in oracle/flyby_oracle.c (the reference itself
    needs VTK 9 and cannot be built here), all host threads, on a bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as orc
    from flyby_b200 import synth
    P, F = synth.bumpy_sphere(LM, seed=0)
    U, _, _ = orc.unit_surf(P)
    N = orc.point_normals(U, F)
    views = orc.icosahedron(SUBDIVISION, RADIUS)
    sample_views = 2  # per step
    # all the host threads this process may use: torchrun exports OMP_NUM_THREADS=1 to its workers, which would
    # otherwise silently make this arm single-threaded
    orc.set_threads(_host_threads())
    cores = orc.num_threads()

    def step(k):
        sel = views[(k * sample_views) % len(views):][:sample_views]
        t0 = time.perf_counter()
        orc.render(U, F, N, sel, RADIUS, RES, plane_scale=PLANE_SCALE, grid=True, grid_div=0)
        return time.perf_counter() - t0

    for k in range(args.warmup):
        step(k)
    times = [step(args.warmup + k) for k in range(args.steps)]
    sec_per_view = sum(times) / (args.steps * sample_views)
    value = 1.0 / sec_per_view
    sample = ("%d of %d views per step at res %d on the 200000-triangle mesh, vtkCellLocator default buckets, "
              "OpenMP over rays; unit-surf/normals/grid build not timed" % (sample_views, len(views), RES))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "views/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(len(views)),
            "cpu_baseline": {"value": value, "unit": "views/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "views/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "mrays_per_s": value * RES * RES / 1e6, "gpu_launches": 0}
    emit(line)


def _hbm_peak():
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        return json.load(open(peaks_path))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def _max_over_ranks(torch, dist, world, dev, vals):
    t = torch.tensor(vals, dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t]


def _host_threads():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_baseline_sample(gpu_sample=None):
    """Times the oracle port on a bounded sample; the rays it casts for that are also the parity sample of the
    bench line: gpu_sample = (features, ids) of views 1..4 of mesh seed 0 as the CUDA path rendered them."""
    from oracle import oracle as orc
    from flyby_b200 import synth
    orc.set_threads(_host_threads())
    P, F = synth.bumpy_sphere(LM, seed=0)
    U, _, _ = orc.unit_surf(P)
    N = orc.point_normals(U, F)
    views = orc.icosahedron(SUBDIVISION, RADIUS)
    out = {}
    parity = None
    for name, div, nv in (("vtk_buckets", 0, 4), ("tuned_grid64", 64, 8)):
        orc.render(U, F, N, views[:1], RADIUS, RES, plane_scale=PLANE_SCALE, grid=True, grid_div=div)
        t0 = time.perf_counter()
        fo, io = orc.render(U, F, N, views[1:1 + nv], RADIUS, RES, plane_scale=PLANE_SCALE, grid=True, grid_div=div)
        out[name] = nv / (time.perf_counter() - t0)
        if name == "vtk_buckets" and gpu_sample is not None:
            gf, gi = gpu_sample
            hit = io >= 0
            parity = {"rays": int(io.size), "hits": int(hit.sum()), "id_mismatches": int((gi != io).sum()),
                      "grazing_edge_disagreements": int(((gi != io) & (gi >= 0) & hit).sum()),
                      "features_bit_equal": bool(np.array_equal(gf, fo)),
                      "max_rel_err_depth": float(np.max(np.abs(gf[..., 3][hit] - fo[..., 3][hit]) / fo[..., 3][hit])) if hit.any() else 0.0,
                      "max_abs_err_normals": float(np.max(np.abs(gf[..., :3] - fo[..., :3]))),
                      "sample": "views 1..4 of the first bench mesh at res 512 (the rays the CPU baseline is timed on), "
                                "CUDA path vs oracle; tie-break smallest t, then lowest cell id"}
    cores = orc.num_threads()
    orc.set_threads(1)  # the reference's own loop is serial: one view, one thread
    t0 = time.perf_counter()
    orc.render(U, F, N, views[1:2], RADIUS, RES, plane_scale=PLANE_SCALE, grid=True, grid_div=0)
    single = 1.0 / (time.perf_counter() - t0)
    orc.set_threads(cores)
    return {"value": out["vtk_buckets"], "unit": "views/s", "cores": cores, "kind": "port", "parity_sample": parity,
            "single_thread_views_per_s": single,
            "sample": "4 views at res 512 of the bench mesh, oracle port of vtkCellLocator (VTK default bucket "
                      "rule) + VTK triangle test, OpenMP over rays, ray loop only",
            "tuned_grid64_views_per_s": out["tuned_grid64"]}


def run_ours(args):
    import torch
    import torch.distributed as dist
    from flyby_b200 import _cabi, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the fly-by path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    # an explicit (non-default) stream: the context launches every kernel on it and
    # the timing events are recorded on the same stream
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    ctx = _cabi.Context(local, stream=stream.cuda_stream)
    assert stream.cuda_stream != 0
    n_views = ctx.views_icosahedron(SUBDIVISION, RADIUS)
    opts = ctx.render_opts(use_z=1, split_z=1, plane_scale=PLANE_SCALE)
    C = 4
    n_rays = n_views * RES * RES

    # Render time depends on the shape (seeds 12..15 take 5 % longer than 0..3), and the job's time is the MAX over
    # ranks: with four private meshes per rank the scaling curve measured which rank drew the hardest four, not the
    # system.  Every rank now processes the same four shapes in rotated order (K steps = K / 4 passes over the pool),
    # so the shares have identical composition, as a 125-mesh share of BASELINE configs[2] has on average.
    pool = [synth.bumpy_sphere(LM, seed=k) for k in range(MESH_POOL)]
    meshes = [pool[(k + rank) % MESH_POOL] for k in range(MESH_POOL)]
    dev_meshes = [(torch.from_numpy(P).to(dev), torch.from_numpy(F).to(dev)) for P, F in meshes]
    torch.cuda.synchronize()  # uploads ran on torch's default stream; the contexts work on their own streams
    # pinned host buffers are allocated (first-touched) on the NUMA node the GPU hangs off: with several ranks the
    # copy-outs then do not cross sockets.  The affinity is restored afterwards (the CPU arm uses all cores).
    from flyby_b200 import dist as fdist
    numa = fdist.near_gpu(local)
    numa.__enter__()
    pin_meshes = [(torch.from_numpy(P).pin_memory(), torch.from_numpy(F).pin_memory()) for P, F in meshes]
    feat = torch.empty((n_views, RES, RES, C), dtype=torch.float32, device=dev)
    ids = torch.empty((n_views, RES, RES), dtype=torch.int32, device=dev)
    # two sets of pinned host outputs: the streaming call fills one while the other is being consumed
    feat_hosts = [torch.empty((n_views, RES, RES, C), dtype=torch.float32).pin_memory() for _ in range(2)]
    ids_hosts = [torch.empty((n_views, RES, RES), dtype=torch.int32).pin_memory() for _ in range(2)]
    feat_host, ids_host = feat_hosts[0], ids_hosts[0]
    for hb in feat_hosts + ids_hosts:
        hb.zero_()  # touch every page while the affinity holds
    numa.__exit__(None, None, None)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    T, V = len(meshes[0][1]), len(meshes[0][0])

    # FLYBY_BUILD_DEFERRED: the build's index / winding checks run on the device as always, but the host does not
    # wait for their verdict -- it arrives with the next synchronising call (ctx.synchronize() below raises on a bad
    # mesh).  The host thus queues mesh n+1 while mesh n renders; `sync_build` times the same loop with the
    # synchronising build for comparison.
    def step_resident(k, deferred=True):
        P, F = dev_meshes[k % MESH_POOL]
        ctx.set_mesh(P, F, deferred=deferred)
        ctx.build()
        ctx.render_into(RES, opts, feat, ids)

    def step_resident_sync(k):
        step_resident(k, deferred=False)

    # The e2e product is what the reference's entry point returns / writes with --out: the views x res x res x C
    # feature tensor (fly_by_features.py:63-154).  The hit-cell-id map is a second, optional output (needed only by
    # the labeling back-projection); the run that also copies it out is reported beside the headline.
    def step_e2e(k, with_ids=False):
        P, F = pin_meshes[k % MESH_POOL]
        ctx.set_mesh(P.numpy(), F.numpy(), deferred=True)
        ctx.build()
        # D2H inside, returns after the copy
        ctx.render_into(RES, opts, feat_host.numpy(), ids_host.numpy() if with_ids else None)

    def step_e2e_ids(k):
        step_e2e(k, True)

    def step_e2e_stream(k):
        """the same through flyby_render_async: returns when queued; this mesh's copy-out overlaps the next mesh"""
        P, F = pin_meshes[k % MESH_POOL]
        ctx.set_mesh(P.numpy(), F.numpy(), deferred=True)
        ctx.build()
        ctx.render_into(RES, opts, feat_hosts[k % 2].numpy(), None, wait=False)

    # the packed contract (flyby_render_packed, normal_encoding "uint8": what the reference's Python class holds after
    # its RGB read-back): uint8 colours + float32 depth, 7 bytes per pixel across PCIe instead of 16
    opts_u8 = ctx.render_opts(use_z=1, split_z=1, plane_scale=PLANE_SCALE, normal_encoding="uint8")
    pk_hosts = None

    def step_e2e_packed(k):
        P, F = pin_meshes[k % MESH_POOL]
        ctx.set_mesh(P.numpy(), F.numpy(), deferred=True)
        ctx.build()
        rgb, z = pk_hosts[k % 2]
        ctx.render_packed_into(RES, opts_u8, rgb.numpy(), z.numpy(), None)

    def timed_stream(steps, warmup, step_e2e_stream=step_e2e_stream):
        """K streamed steps timed as a whole (every step's H2D, device work and D2H lie inside the region)"""
        for k in range(warmup):
            step_e2e_stream(k)
        ctx.synchronize()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        t0 = time.perf_counter()
        for k in range(steps):
            flush.zero_()
            step_e2e_stream(warmup + k)
        ctx.synchronize()  # both streams: every host buffer is complete
        wall_ms = (time.perf_counter() - t0) * 1e3
        e1.record(stream)
        e1.synchronize()
        barrier()
        return max(wall_ms, e0.elapsed_time(e1)) / steps

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(step_fn, steps, warmup, collect_render=False):
        for k in range(warmup):
            step_fn(k)
        barrier()
        total_ms, render_ms, trace_ms = 0.0, 0.0, 0.0
        for k in range(steps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            step_fn(warmup + k)
            e1.record(stream)
            e1.synchronize()
            total_ms += e0.elapsed_time(e1)
            if collect_render:
                st_ = ctx.stats()
                render_ms += st_["ms_render"]
                trace_ms += st_["ms_trace"]
        barrier()
        return total_ms / steps, render_ms / steps, trace_ms / steps

    sampler = ClockSampler(local)
    sampler.start()
    l0 = ctx.launch_count()
    ms_step, ms_render, ms_trace = timed(step_resident, args.steps, args.warmup, collect_render=True)
    launches = ctx.launch_count() - l0
    st = ctx.stats()
    ctx.synchronize()  # (reads the deferred builds' verdict: raises if a mesh was bad)
    ms_step_sync, _, _ = timed(step_resident_sync, max(4, min(args.steps, 10)), 3)
    ms_e2e, _, _ = timed(step_e2e, max(2, min(args.steps, 8)), 3)
    ms_e2e_ids, _, _ = timed(step_e2e_ids, max(2, min(args.steps, 8)), 2)
    ms_e2e_stream = timed_stream(max(4, args.steps), 3)  # K meshes as a whole: fill and drain of the pipeline included

    # The headline `value`: two contexts (two streams) on this GPU alternate meshes, so that the small latency-bound
    # build kernels of mesh n+1 overlap the render of mesh n.  Exactly K meshes, timed as a whole with CUDA events
    # between a barrier + synchronize on both sides; every step is preceded by the L2 flush on its own stream.
    stream2 = torch.cuda.Stream(device=dev)
    ctx2 = _cabi.Context(local, stream=stream2.cuda_stream)
    ctx2.views_icosahedron(SUBDIVISION, RADIUS)
    feat2, ids2 = torch.empty_like(feat), torch.empty_like(ids)
    pair = [(ctx, stream, feat, ids), (ctx2, stream2, feat2, ids2)]

    # FLYBY_SHARE_GPU: the exact-resolve kernels of one context are four persistent CTAs per SM (dispatched at once, registers
    # to spare), so that the other context's build and the start of its candidate search run beside them, not behind their last wave
    opts_share = ctx.render_opts(use_z=1, split_z=1, plane_scale=PLANE_SCALE, share_gpu=True)

    def step_two(k):
        c_, s_, f_, i_ = pair[k % 2]
        P, F = dev_meshes[k % MESH_POOL]
        with torch.cuda.stream(s_):
            flush.zero_()
        c_.set_mesh(P, F, deferred=True)
        c_.build()
        c_.render_into(RES, opts_share, f_, i_)

    def timed_pipelined(step_fn, warm):
        """exactly K steps as a whole: CUDA events on the two streams between barriers; returns (ms per step, launches)"""
        for k in range(warm):
            step_fn(k)
        barrier()
        l0_ = ctx.launch_count() + ctx2.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        stream2.wait_event(e0)
        for k in range(args.steps):
            step_fn(k)
        stream.wait_stream(stream2)
        e1.record(stream)
        e1.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        n_launch = ctx.launch_count() + ctx2.launch_count() - l0_
        ctx.synchronize()   # deferred builds: a bad mesh would raise here
        ctx2.synchronize()
        barrier()
        return ms, n_launch

    k_two = args.steps

    def last_two_digests():
        """int64 sums of the ids and of the feature bit patterns the last two steps of a pipelined loop left behind"""
        return [(int(pair[k % 2][3].to(torch.int64).sum()), int(pair[k % 2][2].view(torch.int32).to(torch.int64).sum()))
                for k in (k_two - 2, k_two - 1) if k >= 0]

    ms_two_calls, launches_calls = timed_pipelined(step_two, max(4, args.warmup))
    dig_calls = last_two_digests()

    # The same loop with every step replayed as a CUDA graph (flyby_capture_begin / _end / flyby_graph_launch): one
    # launch per mesh instead of ~50, the side-stream work as parallel branches of the graph.  One graph per
    # (context, pool mesh): a graph re-reads the device buffers its captured flyby_set_mesh named.
    graphs = {}
    for ci, (c_, s_, f_, i_) in enumerate(pair):
        for m in range(MESH_POOL):
            P, F = dev_meshes[m]
            c_.set_mesh(P, F, deferred=True)   # warm-up of this very step: buffers get their size
            c_.build()
            c_.render_into(RES, opts, f_, i_)
            c_.synchronize()
            with c_.capture() as cap:
                c_.set_mesh(P, F, deferred=True)
                c_.build()
                c_.render_into(RES, opts, f_, i_)
            graphs[(ci, m)] = cap.graph

    def step_two_graph(k):
        c_, s_, f_, i_ = pair[k % 2]
        with torch.cuda.stream(s_):
            flush.zero_()
        c_.graph_launch(graphs[(k % 2, k % MESH_POOL)])

    # the same pipelined graph loop with FLYBY_SHADE_TOLERANCE (opt-in contract: ids exact, depth / normals within
    # 1e-5 relative), reported beside the headline as tolerance_mode.views_per_s
    opts_tol_two = ctx.render_opts(use_z=1, split_z=1, plane_scale=PLANE_SCALE, shade_tolerance=True)
    graphs_tol = {}
    if not args.no_extras:
        for ci, (c_, s_, f_, i_) in enumerate(pair):
            for m in range(MESH_POOL):
                P, F = dev_meshes[m]
                c_.set_mesh(P, F, deferred=True)
                c_.build()
                c_.render_into(RES, opts_tol_two, f_, i_)
                c_.synchronize()
                with c_.capture() as cap:
                    c_.set_mesh(P, F, deferred=True)
                    c_.build()
                    c_.render_into(RES, opts_tol_two, f_, i_)
                graphs_tol[(ci, m)] = cap.graph

        def step_two_graph_tol(k):
            c_, s_, f_, i_ = pair[k % 2]
            with torch.cuda.stream(s_):
                flush.zero_()
            c_.graph_launch(graphs_tol[(k % 2, k % MESH_POOL)])

        ms_two_tol, _ = timed_pipelined(step_two_graph_tol, max(4, args.warmup))
        opts_tol_share = ctx.render_opts(use_z=1, split_z=1, plane_scale=PLANE_SCALE, shade_tolerance=True, share_gpu=True)

        def step_two_tol(k):
            c_, s_, f_, i_ = pair[k % 2]
            P, F = dev_meshes[k % MESH_POOL]
            with torch.cuda.stream(s_):
                flush.zero_()
            c_.set_mesh(P, F, deferred=True)
            c_.build()
            c_.render_into(RES, opts_tol_share, f_, i_)

        ms_two_tol_calls, _ = timed_pipelined(step_two_tol, max(4, args.warmup))
        tol_both = _max_over_ranks(torch, dist, world, dev, [ms_two_tol, ms_two_tol_calls])
        ms_two_tol = min(tol_both)
    else:
        ms_two_tol = 0.0

    ms_two_graph, launches_graph = timed_pipelined(step_two_graph, max(4, args.warmup))
    dig_graph = last_two_digests()  # (these tensors are compared bit for bit with a serial context's below)
    # the headline is the faster of the two ways to drive the pipeline, decided on the max over ranks of each
    both = _max_over_ranks(torch, dist, world, dev, [ms_two_graph, ms_two_calls])
    use_graph = both[0] <= both[1]
    ms_two, launches_two = (ms_two_graph, launches_graph) if use_graph else (ms_two_calls, launches_calls)
    sampler.stop_flag = True
    sampler.join(timeout=2)

    # Not timed: the tensors the last two pipelined steps left behind must equal what a strictly serial context
    # (no side stream, a synchronisation after every call) computes for the same meshes -- the overlapped build /
    # render / initialisation of the timed loops must not change a single bit.
    os.environ["FLYBY_NO_SIDE_TREE"] = "1"
    ctx3 = _cabi.Context(local, stream=stream.cuda_stream)
    del os.environ["FLYBY_NO_SIDE_TREE"]
    ctx3.views_icosahedron(SUBDIVISION, RADIUS)
    feat3, ids3 = torch.empty_like(feat), torch.empty_like(ids)
    verified = dig_calls == dig_graph  # the host-launched FLYBY_SHARE_GPU loop and the graph replays left the same tensors
    for k in (k_two - 2, k_two - 1):
        if k < 0:
            continue
        _, _, f_, i_ = pair[k % 2]
        P, F = dev_meshes[k % MESH_POOL]
        ctx3.set_mesh(P, F)
        ctx3.synchronize()
        ctx3.build()
        ctx3.synchronize()
        ctx3.render_into(RES, opts, feat3, ids3)
        ctx3.synchronize()
        torch.cuda.synchronize()
        verified = verified and bool(torch.equal(ids3, i_)) and bool(torch.equal(feat3.view(torch.int32), f_.view(torch.int32)))
    ctx3.close()
    for (ci, m), g in list(graphs.items()) + list(graphs_tol.items()):
        pair[ci][0].graph_destroy(g)
    ctx2.close()
    if not verified:
        raise SystemExit("bench.py: the pipelined loop's outputs differ from a serial context's -- numbers withheld")

    # ---- records beside the headline (bench_extra.py): plane scale 2.0, tolerance shading, the node's D2H ceiling,
    # BASELINE configs[4] sharded by viewpoint and configs[3] with the NCCL vote all-reduce.  Not part of `value`.
    import bench_extra as bx
    extras = {}
    if not args.no_extras:
        P0, F0 = dev_meshes[0]
        ms_r1, ms_t1, st1 = bx.render_variant(torch, ctx, stream, flush, (P0, F0), RES, opts, feat, ids)
        feat_exact, ids_exact = feat.clone(), ids.clone()
        opts_tol = ctx.render_opts(use_z=1, split_z=1, plane_scale=PLANE_SCALE, shade_tolerance=True)
        ms_rt, ms_tt, _ = bx.render_variant(torch, ctx, stream, flush, (P0, F0), RES, opts_tol, feat, ids)
        tol = bx.tolerance_errors(torch, feat_exact, ids_exact, feat, ids)
        del feat_exact, ids_exact
        opts2 = ctx.render_opts(use_z=1, split_z=1, plane_scale=2.0)
        ms_r2, ms_t2, st2 = bx.render_variant(torch, ctx, stream, flush, (P0, F0), RES, opts2, feat, ids)
        vals = _max_over_ranks(torch, dist, world, dev, [ms_r1, ms_t1, ms_rt, ms_tt, ms_r2, ms_t2, ms_two_tol])
        alg1 = algorithmic_bytes(n_rays, C, T, V)
        extras["plane_scale_2"] = {
            "render_ms": vals[4], "candidate_search_ms": vals[5], "exact_resolve_ms": vals[4] - vals[5],
            "render_mrays_per_s": n_rays / (vals[4] * 1e-3) / 1e6, "views_per_s_render_only": n_views / (vals[4] * 1e-3),
            "hit_fraction": st2["n_hits"] / max(1, st2["n_rays"]),
            "roofline_frac": alg1 / (vals[4] * 1e-3) / 1e9 / _hbm_peak()[0],
            "beside": {"plane_scale": 1.0, "render_ms": vals[0], "hit_fraction": st1["n_hits"] / max(1, st1["n_rays"])},
            "note": "planeScaleFactor 2.0 shows the whole unit surface (SURVEY 8d O5); 1.0 is the reference default the "
                    "headline uses (the unit surface puts the bounding-box corner on the unit sphere, so the mesh is ~1.15 wide: a 1 x 1 "
                    "window is 91 % covered, a 2 x 2 window 27 %)"}
        extras["tolerance_mode"] = dict(tol, render_ms=vals[2], candidate_search_ms=vals[3],
                                        exact_resolve_ms=vals[2] - vals[3], exact_render_ms=vals[0],
                                        views_per_s_render_only=n_views / (vals[2] * 1e-3),
                                        ms_per_step=vals[6], views_per_s=world * n_views / (vals[6] * 1e-3),
                                        note="FLYBY_SHADE_TOLERANCE (opt-in; the headline uses the exact chain)")
        d2h_bytes = n_rays * 4 * C
        slow, total = bx.copy_ceiling(torch, dist, world, dev, d2h_bytes, feat_hosts[0])
        extras["copy_ceiling"] = {"d2h_gb_s_slowest_rank": slow, "d2h_gb_s_all_ranks": total, "bytes": d2h_bytes,
                                  "views_per_s": world * n_views / (d2h_bytes / (slow * 1e9)),
                                  "how": "every rank: one plain cudaMemcpyAsync of the e2e payload into its pinned buffer, all "
                                         "ranks at the same time, best of 6"}
        pk_hosts = [(torch.empty((n_views, RES, RES, 3), dtype=torch.uint8).pin_memory(),
                     torch.empty((n_views, RES, RES), dtype=torch.float32).pin_memory()) for _ in range(2)]
        ms_pk = timed_stream(max(4, args.steps), 3, step_e2e_packed)
        # what the packed buffers hold against the float tensor of the same options (last mesh of the loop)
        k_last = 3 + max(4, args.steps) - 1
        Pl, Fl = dev_meshes[k_last % MESH_POOL]
        ctx.set_mesh(Pl, Fl); ctx.build(); ctx.render_into(RES, opts_u8, feat, None); ctx.synchronize()
        rgb_l, z_l = pk_hosts[k_last % 2]
        pk_ok = bool(torch.equal(feat[..., :3].cpu(), rgb_l.to(torch.float32))) and bool(torch.equal(feat[..., 3].cpu(), z_l))
        ms_pk = _max_over_ranks(torch, dist, world, dev, [ms_pk])[0]
        extras["e2e_packed"] = {"views_per_s": world * n_views / (ms_pk * 1e-3), "ms_per_step": ms_pk,
                                "d2h_bytes_per_step": n_rays * 7, "h2d_bytes_per_step": V * 12 + T * 12,
                                "equals_float_tensor": pk_ok,
                                "note": "flyby_render_packed, normal_encoding uint8 (the 8-bit colours FlyByGenerator.getFlyBy holds "
                                        "after its RGB read-back) + float32 depth: 7 bytes per pixel instead of 16, lossless for that "
                                        "contract; same streamed loop as e2e (pinned H2D of the mesh, build, render, chunked D2H). "
                                        "Beside the headline, which moves the float tensor of the C++ tool's encoding"}
        del pk_hosts
        del feat, ids
        torch.cuda.empty_cache()
        extras["by_view"] = bx.by_view(torch, dist, _cabi, synth, fdist, world, rank, local, dev, stream, barrier)
        try:
            extras["votes_allreduce"] = bx.votes_allreduce(torch, dist, _cabi, synth, fdist, world, rank, local, dev, stream,
                                                           barrier)
        except (_cabi.FlyByError, ImportError) as ex:  # NCCL or torchvision missing on this box
            extras["votes_allreduce"] = {"unavailable": str(ex)[:200]}

    t = torch.tensor([ms_step, ms_e2e, ms_e2e_stream, ms_two, ms_e2e_ids, ms_step_sync, ms_two_calls, ms_two_graph], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step_max, ms_e2e_max, ms_e2e_stream_max, ms_two_max, ms_e2e_ids_max, ms_step_sync_max, ms_two_calls_max, ms_two_graph_max = (float(x) for x in t)
    # per-rank view of the headline loop: the job's ms_per_step is the MAX over ranks, i.e. the slowest GPU of the node
    per_rank = [{"rank": rank, "gpu": local, "ms_per_step": ms_two, "ms_per_step_graph_replay": ms_two_graph,
                 "ms_per_step_host_launches": ms_two_calls, "render_ms": ms_render, **sampler.summary()}]
    if world > 1:
        box = [None] * world
        dist.all_gather_object(box, per_rank[0])
        per_rank = box

    if rank == 0:
        peak, peak_src = _hbm_peak()
        alg = algorithmic_bytes(n_rays, C, T, V)
        # the render is a short pipeline of kernels (candidate search: raster_kernel; exact decision:
        # raster_resolve1/2, spill_*); the roofline is taken over the pipeline, CUDA events on its stream
        achieved = alg / (ms_render * 1e-3) / 1e9
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "render_traffic.json")
        if os.path.exists(tpath):
            traffic = json.load(open(tpath)).get("dram_bytes_per_render")
        views_per_s = world * n_views / (ms_two_max * 1e-3)
        # headline e2e: the faster of the two public calls (blocking flyby_render / streaming flyby_render_async);
        # both move every step's mesh H2D and its feature tensor D2H inside the timed region
        ms_e2e_best = min(ms_e2e_max, ms_e2e_stream_max)
        e2e_views = world * n_views / (ms_e2e_best * 1e-3)
        cpu = None
        if world == 1:
            c4 = _cabi.Context(local, stream=stream.cuda_stream)
            c4.set_mesh(*dev_meshes[0])
            c4.build()
            c4.views_icosahedron(SUBDIVISION, RADIUS)
            cpu = cpu_baseline_sample(c4.render(RES, opts, view_begin=1, view_end=5))
            c4.close()
        line = {"metric": METRIC, "value": views_per_s, "unit": "views/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_two_max, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64 (hit decision, depth, normals) over f32 traversal",
                "data": "synthetic", "config": workload_config(n_views),
                "mrays_per_s": views_per_s * RES * RES / 1e6,
                "render_ms": ms_render, "candidate_search_ms": ms_trace, "exact_resolve_ms": ms_render - ms_trace,
                "render_mrays_per_s": n_rays / (ms_render * 1e-3) / 1e6,
                "build_ms": {k: st[k] for k in ("ms_normalise", "ms_normals", "ms_sort", "ms_tree")},
                "hit_fraction": st["n_hits"] / max(1, st["n_rays"]),
                "pipelining": "two contexts (two streams) per GPU alternate meshes, K meshes timed as a whole, max over ranks; "
                              "driven two ways, the faster is the headline: " +
                              ("[headline] " if not use_graph else "") + "host launches (~51 per mesh) with FLYBY_SHARE_GPU -- the "
                              "exact-resolve kernels of one context are 4 persistent CTAs per SM and the build of the other "
                              "context's next mesh runs beside them; " + ("[headline] " if use_graph else "") + "one CUDA-graph replay of set_mesh + build "
                              "+ render per mesh (flyby_graph_launch), where the two graphs overlap only in the build",
                "headline_loop": "graph_replay" if use_graph else "host_launches_share_gpu",
                "ms_per_step_graph_replay": ms_two_graph_max,
                "ms_per_step_host_launches": ms_two_calls_max,
                "ms_per_step_without_graph": ms_two_calls_max,
                "per_rank": per_rank,
                "gpu_launches_without_graph": launches_calls,
                "single_context_ms_per_step": ms_step_max,
                "single_context_sync_build_ms_per_step": ms_step_sync_max,
                "build_checks": "FLYBY_BUILD_DEFERRED in the timed loops: the index / winding checks run on the device in "
                                "every step, their verdict is read by the synchronising call after the loop; "
                                "single_context_sync_build_ms_per_step is the same loop with the synchronising build",
                "single_context_note": "one context, every step timed on its own between device synchronisations "
                                       "(render_ms, build_ms and the roofline come from this loop)",
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                             "frac": achieved / peak, "traffic": traffic,
                             "kernel": "render pipeline: raster_kernel (largest), raster_resolve1_kernel, raster_resolve2_kernel, spill_*_kernel",
                             "algorithmic_bytes_per_launch": alg, "peak_source": peak_src,
                             "note": "algorithmic bytes = compulsory output + mesh + BVH bytes of one render (SURVEY 8d); the pipeline is bound by float32/float64 instruction issue and L2 atomic latency, not by HBM; the 0.05 ms front-buffer initialisation of the next render runs on the side stream behind the resolve kernels and lies outside this duration (it is inside ms_per_step)"},
                "e2e": {"value": e2e_views, "unit": "views/s", "ms_per_step": ms_e2e_best,
                        "h2d_bytes_per_step": V * 12 + T * 12, "d2h_bytes_per_step": n_rays * 4 * C,
                        "mode": ("flyby_render_async" if ms_e2e_stream_max < ms_e2e_max else "blocking flyby_render")
                                + " per mesh: pinned H2D of the mesh, build, render in view chunks, chunked D2H of the "
                                  "feature tensor (what fly_by_features.py --out holds) into pinned host memory; "
                                  "PCIe-bound: copy_ceiling_frac = this value / what N concurrent plain pinned D2H copies of the "
                                  "same payload reach on this node (copy_ceiling)",
                        "host_buffers_numa_node": numa.node,
                        "copy_ceiling_frac": (e2e_views / extras["copy_ceiling"]["views_per_s"]) if "copy_ceiling" in extras else None,
                        "blocking_ms_per_step": ms_e2e_max,
                        "streamed_ms_per_step": ms_e2e_stream_max,
                        "streamed_note": "flyby_render_async (copy-out of mesh n overlaps upload/build/render of mesh "
                                         "n+1), K meshes timed as a whole incl. pipeline drain",
                        "with_cell_ids": {"views_per_s": world * n_views / (ms_e2e_ids_max * 1e-3),
                                          "ms_per_step": ms_e2e_ids_max,
                                          "d2h_bytes_per_step": n_rays * (4 * C + 4),
                                          "note": "blocking call that also copies out the int32 hit-cell-id map "
                                                  "(input of the labeling back-projection)"}},
                "outputs_verified": "ids and features of the last two pipelined steps are bit-equal to a serial "
                                    "context's (no side stream, synchronised after every call); the host-launched FLYBY_SHARE_GPU "
                                    "loop and the graph replays left tensors with equal digests",
                "gpu_launches": launches_two, "gpu_launches_single_context_loop": launches,
                **extras,
                "clocks": sampler.summary()}
        if cpu is not None:
            line["parity_sample"] = cpu.pop("parity_sample")
            line["cpu_baseline"] = cpu
        emit(line)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


_json_out = None


def emit(obj):
    """The one JSON line, on the process's original stdout."""
    out = _json_out if _json_out is not None else sys.stdout
    out.write(json.dumps(obj) + "\n")
    out.flush()


def main():
    # stdout carries the JSON line only: whatever a library prints there (NCCL's version banner under
    # NCCL_DEBUG=VERSION, for one) is sent to stderr by pointing file descriptor 1 at it
    global _json_out
    sys.stdout.flush()
    _json_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-extras", dest="no_extras", action="store_true",
                    help="skip the records beside the headline (plane scale 2, tolerance mode, configs[3] / [4])")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
