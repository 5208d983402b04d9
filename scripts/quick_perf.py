"""Developer scratch benchmark (not the contract bench; see bench.py): times the warp
kernel alone on synthetic 4K frames and prints achieved algorithmic GB/s."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import blinky_b200 as bb


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--w", type=int, default=3840)
    ap.add_argument("--h", type=int, default=2160)
    ap.add_argument("--ps", type=int, default=2048)
    ap.add_argument("--globe", default="cube")
    ap.add_argument("--lens", default="panini")
    ap.add_argument("--zoom", default="f_fov 180")
    ap.add_argument("--rubix", action="store_true")
    ap.add_argument("--frames", type=int, default=16)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--threads", type=int, default=bb.usable_cpus())
    ap.add_argument("--cold", action="store_true", help="single frame, L2 flushed before every launch")
    ap.add_argument("--kernel", type=int, default=0)
    a = ap.parse_args()
    fe = bb.Fisheye(device=0, palette=bb.synthetic_palette())
    fe.command(f"f_globe {a.globe}")
    fe.command(f"f_lens {a.lens}")
    fe.command(a.zoom)
    if a.rubix:
        fe.command("f_rubix")
    t0 = time.time()
    fe.build_lensmap(a.w, a.h, a.ps, a.threads)
    t_build = time.time() - t0
    fe.set_kernel(a.kernel)
    P, ps2 = fe.numplates, a.ps * a.ps
    npix = a.w * a.h
    M = fe.mapped_pixels
    gen = torch.Generator(device="cuda").manual_seed(1000)
    faces = torch.randint(0, 256, (a.frames, P, a.ps, a.ps), dtype=torch.uint8, device="cuda", generator=gen)
    out = torch.zeros((a.frames, a.h, a.w), dtype=torch.uint8, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    nf = 1 if a.cold else a.frames
    for _ in range(3):
        fe.warp(faces, out, nframes=nf, stream=st)
    torch.cuda.synchronize()
    times = []
    for _ in range(a.iters):
        if a.cold:
            flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fe.warp(faces, out, nframes=nf, stream=st)
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1) * 1e-3)
    t = float(np.median(times))
    alg = (5 * npix + M) * nf
    print(json.dumps(dict(lens=a.lens, globe=a.globe, zoom=a.zoom, rubix=a.rubix, w=a.w, h=a.h, ps=a.ps, frames=nf,
                          cold=a.cold, build_s=round(t_build, 2), mapped_frac=round(M / npix, 4),
                          ms=round(t * 1e3, 4), us_per_frame=round(t * 1e6 / nf, 2),
                          mpix_s=round(npix * nf / t / 1e6, 1), alg_gbs=round(alg / t / 1e9, 1),
                          frac_of_6485=round(alg / t / 1e9 / 6485.5, 3), kernel=fe.last_kernel,
                          min_ms=round(min(times) * 1e3, 4), plan=fe.plan_summary)))


if __name__ == "__main__":
    main()
