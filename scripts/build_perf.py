"""Lensmap build time: GPU (translated lens, NVRTC) vs interpreter on all usable CPUs.
Usage: python scripts/build_perf.py [W H PS]"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import blinky_b200 as bb

W, H, PS = (int(a) for a in sys.argv[1:4]) if len(sys.argv) >= 4 else (3840, 2160, 2048)
fe = bb.Fisheye(device=0, palette=bb.synthetic_palette())
print(f"usable cpus {bb.usable_cpus()}  size {W}x{H} ps {PS}")
LENSES = [a for a in sys.argv[4:] if not a.startswith("--")] or ["panini", "stereographic", "equirect", "hammer", "fisheye1", "mollweide", "vandergrinten", "winkeltripel", "eckert4", "quincuncial", "cube"]
for lens in LENSES:
    fe.command("f_globe cube"); fe.command(f"f_lens {lens}")
    t = time.time(); fe.build_lensmap(W, H, PS, threads=0); t_first = time.time() - t
    info = fe.build_info
    a = fe.lensmap_packed().copy()
    fe.command(f"f_lens {lens}")
    t = time.time(); fe.build_lensmap(W, H, PS, threads=0); t_again = time.time() - t
    row = f"{lens:14s} gpu first {t_first*1e3:8.1f} ms  again {t_again*1e3:8.1f} ms  [{info}]"
    if "--host" in sys.argv:
        fe.command(f"f_lens {lens}")
        t = time.time(); fe.build_lensmap(W, H, PS, threads=-1); t_host = time.time() - t
        same = np.array_equal(a, fe.lensmap_packed())
        row += f"  host {t_host*1e3:8.1f} ms  identical={same}"
    print(row, flush=True)
