"""Run the tiled warp kernel many times and compare every launch with the flat gather kernel's
result for the same frames (bit-exact expected).  Usage: stress_determinism.py [launches] [workload]"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import blinky_b200 as bb
from bench import WORKLOADS

n = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 300
W, H, PS, globe, lens, zoom, rubix = WORKLOADS[sys.argv[2] if len(sys.argv) > 2 and not sys.argv[2].startswith("--") else "4k-cube-panini"]
F = 16
fe = bb.Fisheye(device=0, palette=bb.synthetic_palette())
for c in (f"f_globe {globe}", f"f_lens {lens}", zoom): fe.command(c)
fe.set_rubix(rubix)
fe.build_lensmap(W, H, PS, threads=0)
P = fe.numplates
gen = torch.Generator(device="cuda").manual_seed(1000)
d_faces = torch.randint(0, 256, (F, P, PS, PS), dtype=torch.uint8, device="cuda", generator=gen)
ref = torch.zeros((F, H, W), dtype=torch.uint8, device="cuda")
out = torch.zeros((F, H, W), dtype=torch.uint8, device="cuda")
st = torch.cuda.current_stream().cuda_stream
fe.set_kernel(1); fe.warp(d_faces, ref, nframes=F, stream=st); torch.cuda.synchronize()
fe.set_kernel(0)
bad_launches = 0
# --slow-stores: the kernel writes straight into mapped pinned host memory (stores as slow as
# NVLink/PCIe peers make them), which stretches the time a consumer warp spends between releasing a
# ring stage and finishing its stores
slow = "--slow-stores" in sys.argv
h_out = fe.alloc_pinned(F * H * W) if slow else None
side = torch.cuda.Stream()
junk_a = torch.empty(256 << 20, dtype=torch.uint8, device="cuda"); junk_b = torch.empty_like(junk_a)
for i in range(n):
    out.fill_(0xEE)
    if i % 2:  # concurrent copy traffic on another stream perturbs the timing
        with torch.cuda.stream(side): junk_b.copy_(junk_a)
    if slow:
        h_out[:] = 0xEE
        fe.warp(d_faces, int(h_out.ctypes.data), nframes=F, stream=st)
        torch.cuda.synchronize()
        out.copy_(torch.from_numpy(h_out.reshape(F, H, W)))
    else:
        fe.warp(d_faces, out, nframes=F, stream=st)
    torch.cuda.synchronize()
    if not torch.equal(out, ref):
        bad_launches += 1
        d = (out != ref).nonzero()
        k = d[0].tolist()
        print(f"launch {i}: {d.shape[0]} bytes differ, first at frame {k[0]} y {k[1]} x {k[2]}: got {int(out[tuple(k)])} want {int(ref[tuple(k)])}", flush=True)
print(json.dumps({"launches": n, "bad_launches": bad_launches, "kernel": fe.last_kernel}))
