"""Where does the streamed end-to-end step spend its time?  python profiles/dbg_stream.py"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "fly-by-cnn_b200"))
import numpy as np, torch
from flyby_b200 import _cabi, synth
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream)
ctx = _cabi.Context(0, stream=stream.cuda_stream)
n = ctx.views_icosahedron(3, 2.0)
opts = ctx.render_opts(use_z=1, split_z=1)
meshes = [synth.bumpy_sphere(100, seed=k) for k in range(4)]
pin = [(torch.from_numpy(P).pin_memory(), torch.from_numpy(F).pin_memory()) for P, F in meshes]
fh = [torch.empty((n, 512, 512, 4), dtype=torch.float32).pin_memory() for _ in range(2)]
ih = [torch.empty((n, 512, 512), dtype=torch.int32).pin_memory() for _ in range(2)]
def step(k, wait):
    P, F = pin[k % 4]
    t0 = time.perf_counter()
    ctx.set_mesh(P.numpy(), F.numpy()); t1 = time.perf_counter()
    ctx.build(); t2 = time.perf_counter()
    ctx.render_into(512, opts, fh[k % 2].numpy(), ih[k % 2].numpy(), wait=wait); t3 = time.perf_counter()
    return (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for mode, K, do_flush in (("blocking", 12, False), ("streamed", 12, False), ("streamed", 30, False), ("streamed", 60, False), ("streamed", 30, True), ("blocking", 30, True)):
    for k in range(3): step(k, mode == "blocking")
    ctx.synchronize(); torch.cuda.synchronize()
    t = time.perf_counter(); rows = []
    for k in range(K):
        if do_flush: flush.zero_()
        rows.append(step(3 + k, mode == "blocking"))
    ctx.synchronize()
    ms = (time.perf_counter() - t) * 1e3
    print(mode, K, "flush" if do_flush else "", "%.2f ms/step" % (ms / K), "host ms per call (set_mesh, build, render):", np.round(np.mean(rows, 0), 2), "last", np.round(rows[-1], 2))
# blocking call with only one of the outputs: effective copy bandwidth per buffer
P, F = pin[0]
ctx.set_mesh(P.numpy(), F.numpy()); ctx.build()
dfeat = torch.empty((n, 512, 512, 4), dtype=torch.float32, device=dev); dids = torch.empty((n, 512, 512), dtype=torch.int32, device=dev)
for name, args in (("device outputs", (dfeat, dids)), ("features to host", (fh[0].numpy(), None)), ("ids to host", (None, ih[0].numpy())), ("both to host", (fh[0].numpy(), ih[0].numpy()))):
    for _ in range(2):
        ctx.render_into(512, opts, *args); ctx.synchronize()
    t = time.perf_counter()
    for _ in range(5):
        ctx.render_into(512, opts, *args); ctx.synchronize()
    ms = (time.perf_counter() - t) * 1e3 / 5
    nb = (0 if args[0] is None or args[0] is dfeat else fh[0].numel() * 4) + (0 if args[1] is None or args[1] is dids else ih[0].numel() * 4)
    print("%-18s %.2f ms  (%.0f MB to host, %.1f GB/s if all of it were copy time)" % (name, ms, nb / 1e6, nb / ms / 1e6 if nb else 0))
