"""Developer helper: dump the SASS of one kernel from build/warp_device.o and report the size of its innermost
loops that contain shared-memory byte loads (the ring kernel's per-frame loop).
usage: python scripts/sass_loop.py <substring of the mangled name> [--print]"""
import re
import subprocess
import sys

obj = "blinky_b200/build/warp_device.o"
want = sys.argv[1]
names = subprocess.run(["cuobjdump", "-elf", obj], stdout=subprocess.PIPE, text=True).stdout
funcs = sorted(set(re.findall(r"\.text\.(\S+)", names)))
for fn in funcs:
    if want not in fn:
        continue
    sass = subprocess.run(["cuobjdump", "-sass", "-fun", fn, obj], stdout=subprocess.PIPE, text=True).stdout
    ins = []
    for line in sass.splitlines():
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    addr_index = {a: i for i, (a, _) in enumerate(ins)}
    print(fn[-60:], "instructions:", len(ins), "LDL/STL:", sum(1 for _, t in ins if "LDL" in t or "STL" in t))
    # backward branches = loops
    loops = []
    for i, (a, t) in enumerate(ins):
        m = re.search(r"BRA(?:\.U)?(?:\.ANY)?\s+(?:\S+,\s*)?(0x[0-9a-f]+)", t)
        if m:
            tgt = int(m.group(1), 16)
            if tgt <= a and tgt in addr_index:
                loops.append((addr_index[tgt], i))
    for lo, hi in loops:
        body = ins[lo:hi + 1]
        n_lds = sum(1 for _, t in body if "LDS.U8" in t)
        if n_lds >= 32 and not any(l2 > lo and h2 < hi and sum(1 for _, t in ins[l2:h2 + 1] if "LDS.U8" in t) >= 32 for l2, h2 in loops):
            kinds = {}
            for _, t in body:
                op = t.split()[1] if t.startswith("@") else t.split()[0]
                op = op.split(".")[0]
                kinds[op] = kinds.get(op, 0) + 1
            print(f"  loop {ins[lo][0]:#x}..{ins[hi][0]:#x}: {len(body)} instructions, LDS.U8 {n_lds};", dict(sorted(kinds.items(), key=lambda x: -x[1])))
            if "--print" in sys.argv:
                for a, t in body:
                    print(f"    {a:#06x} {t}")
