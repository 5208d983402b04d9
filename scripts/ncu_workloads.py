"""Run under `ncu --profile-from-start off`: one profiled launch (FRAMES frames, or a single frame) of the warp
for each named bench workload, in order; writes the launch order to gpurun_out/ncu_workloads_order.json so
that scripts/ncu_traffic.py can attribute the captured kernels.

    ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/r2_all \
        python scripts/ncu_workloads.py [--frames 16] [workload ...]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import blinky_b200 as bb  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=16)
    ap.add_argument("--order", default="gpurun_out/ncu_workloads_order.json")
    ap.add_argument("workloads", nargs="*")
    a = ap.parse_args()
    names = a.workloads or ["4k-cube-panini"] + bench.SECONDARY
    order = []
    fe = bb.Fisheye(device=0, palette=bb.synthetic_palette())
    st = torch.cuda.current_stream().cuda_stream
    for name in names:
        W, H, PS = bench.WORKLOADS[name][:3]
        bench.setup_workload(fe, name)
        P = fe.numplates
        gen = torch.Generator(device="cuda").manual_seed(2000)
        d_faces = torch.randint(0, 256, (a.frames, P, PS, PS), dtype=torch.uint8, device="cuda", generator=gen)
        d_out = torch.zeros((a.frames, H, W), dtype=torch.uint8, device="cuda")
        for _ in range(2):
            fe.warp(d_faces, d_out, nframes=a.frames, stream=st)
        torch.cuda.synchronize()
        n0 = fe.launch_count
        torch.cuda.profiler.start()
        fe.warp(d_faces, d_out, nframes=a.frames, stream=st)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        order.append({"workload": name, "frames": a.frames, "launches": fe.launch_count - n0, "kernel": fe.last_kernel,
                      "plan": fe.plan_summary})
        del d_faces, d_out
    fe.close()
    os.makedirs(os.path.dirname(a.order) or ".", exist_ok=True)
    json.dump(order, open(a.order, "w"), indent=1)


if __name__ == "__main__":
    main()
