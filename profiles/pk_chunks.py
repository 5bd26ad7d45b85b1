import os, sys, time
ROOT='/root/repo'
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "fly-by-cnn_b200"))
import torch
from flyby_b200 import _cabi, synth
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream)
ctx = _cabi.Context(0, stream=stream.cuda_stream)
n = ctx.views_icosahedron(3, 2.0)
meshes = [synth.bumpy_sphere(100, seed=k) for k in range(4)]
pin = [(torch.from_numpy(P).pin_memory(), torch.from_numpy(F).pin_memory()) for P, F in meshes]
opts = ctx.render_opts(use_z=1, split_z=1, normal_encoding="uint8")
bufs = [(torch.empty((n,512,512,3),dtype=torch.uint8).pin_memory(), torch.empty((n,512,512),dtype=torch.float32).pin_memory()) for _ in range(2)]
def step(k):
    P, F = pin[k % 4]
    ctx.set_mesh(P.numpy(), F.numpy(), deferred=True); ctx.build()
    rgb, z = bufs[k % 2]
    ctx.render_packed_into(512, opts, rgb.numpy(), z.numpy(), None)
for k in range(4): step(k)
ctx.synchronize()
for rep in range(3):
    t0 = time.perf_counter()
    K = 20
    for k in range(K): step(k)
    ctx.synchronize()
    print("chunks", os.environ.get("FLYBY_PK_CHUNKS"), "ms per mesh %.3f" % ((time.perf_counter() - t0) * 1e3 / K), flush=True)
