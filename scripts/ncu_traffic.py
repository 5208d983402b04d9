"""Turns an .ncu-rep captured with scripts/ncu_workloads.py into (a) profiles/traffic_r2.json — measured DRAM
bytes per launch and per workload, what bench.py's roofline.traffic / dram_frac read — and (b) one text
summary per workload under profiles/.

    python scripts/ncu_traffic.py gpurun_out/r2_all.ncu-rep gpurun_out/ncu_workloads_order.json [tag]
"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scripts"))
from ncu_summary import WANT  # noqa: E402


def num(s):
    try:
        return float(s.replace(",", ""))
    except ValueError:
        return 0.0


def scale(value, unit):
    u = unit.strip().lower()
    return value * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "ns": 1e-3, "us": 1, "usecond": 1, "ms": 1e3, "msecond": 1e3,
                    "nsecond": 1e-3}.get(u, 1)


def main():
    rep, order_path = sys.argv[1], sys.argv[2]
    tag = sys.argv[3] if len(sys.argv) > 3 else "r2"
    order = json.load(open(order_path))
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, body = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    db_path = os.path.join(ROOT, "profiles", "traffic_r2.json")
    try:
        db = json.load(open(db_path))
    except Exception:  # noqa: BLE001
        db = {}
    db["_note"] = ("dram__bytes_read.sum + dram__bytes_write.sum per launch (all kernels of the launch), ncu --set full "
                   "--clock-control none, cache control all (cold L2), captured with scripts/ncu_workloads.py")
    k = 0
    for o in order:
        mine = body[k:k + o["launches"]]
        k += o["launches"]
        rd = wr = dur = 0.0
        lines = [f"# {o['workload']}  ({o['frames']} frames per launch)  {o['kernel']}", f"# plan: {o['plan']}", f"# source: {os.path.basename(rep)}"]
        kernels = []
        for r in mine:
            rd += scale(num(r[col["dram__bytes_read.sum"]]), units[col["dram__bytes_read.sum"]])
            wr += scale(num(r[col["dram__bytes_write.sum"]]), units[col["dram__bytes_write.sum"]])
            dur += scale(num(r[col["gpu__time_duration.sum"]]), units[col["gpu__time_duration.sum"]])
            name = r[col["Kernel Name"]]
            kernels.append(name)
            lines.append(f"== {name}")
            for w in WANT:
                if w in col:
                    lines.append(f"{w:82s} {r[col[w]]:>18s} {units[col[w]]}")
        db[o["workload"] if o["frames"] != 1 else o["workload"] + ":1frame"] = {
            "dram_bytes_per_launch": int(rd + wr), "dram_read": int(rd), "dram_write": int(wr), "frames": o["frames"],
            "ncu_duration_us": round(dur, 2), "kernels": kernels, "source": f"profiles/{tag}_{o['workload']}.txt"}
        suffix = "" if o["frames"] != 1 else "_1frame"
        open(os.path.join(ROOT, "profiles", f"{tag}_{o['workload']}{suffix}.txt"), "w").write("\n".join(lines) + "\n")
        print(o["workload"], o["frames"], int(rd + wr), round(dur, 2), kernels)
    json.dump(db, open(db_path, "w"), indent=1)


if __name__ == "__main__":
    main()
