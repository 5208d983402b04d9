"""Drift of the lensmap under the reference's shipping flags: engine/Makefile:270-277 compiles fisheye.c with
-ffast-math, whose floating-point results are compiler-dependent; the parity oracle (oracle/_ref/libblinky_ref.so)
is the same source WITHOUT -ffast-math.  This script builds every inverse/forward lens with both compiled
references (same scripts, same minilua) and counts the lensmap entries that differ.  The warp itself (integer
byte moves) is identical under any flags; only the map can drift.
    python scripts/fastmath_drift.py [W H PS] > profiles/r2_fastmath_drift.txt"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

LENSES = ["cube", "cubestereo", "cylinder", "debug", "eckert1", "eckert4", "eckert5", "equirect", "fahey", "fisheye1", "fisheye2",
          "gallstereo", "gins8", "gumby", "hammer", "kavrayskiy7", "larrivee", "mercator", "miller", "mollweide", "panini",
          "polyconic", "quincuncial", "rectilinear", "sinusoidal", "stereographic", "vandergrinten", "wagner6", "winkel1",
          "winkel2", "winkeltripel"]


_fast = None


def drift(lenses=LENSES, size=(320, 200, 128), globe="cube", out=sys.stdout, exact=None):
    global _fast
    import blinky_b200 as bb
    from oracle.pyoracle import RefOracle

    pal = bb.synthetic_palette()
    if exact is None:
        exact = RefOracle.get(bb.SCRIPT_DIR, pal)
    if _fast is None:  # F_Init can run once per loaded library
        _fast = RefOracle(bb.SCRIPT_DIR, pal, fastmath=True)
    fast = _fast
    W, H, PS = size
    rows = []
    print(f"# lensmap entries that differ between the -O2 reference and the -O2 -ffast-math reference "
          f"({W}x{H}, {globe} {PS}^2; gcc {os.popen('gcc -dumpversion').read().strip()})", file=out)
    print("# lens            map      pixels  differ  mapped<->unmapped  max |texel dx|,|dy|  scale equal", file=out)
    for lens in lenses:
        res = []
        for R in (exact, fast):
            R.set_screen(W, H)
            R.command(f"f_globe {globe}")
            R.command(f"f_lens {lens}")
            rc = R.build(W, H, PS)
            idx, tint = R.lensmap()
            res.append((rc, idx.copy(), tint.copy(), R.scale))
        (rc0, i0, t0, s0), (rc1, i1, t1, s1) = res
        diff = i0 != i1
        flips = int(((i0 < 0) != (i1 < 0)).sum())
        both = diff & (i0 >= 0) & (i1 >= 0)
        dx = dy = 0
        if both.any():
            a, b = i0[both] % (PS * PS), i1[both] % (PS * PS)
            same_plate = (i0[both] // (PS * PS)) == (i1[both] // (PS * PS))
            if same_plate.any():
                dx = int(np.abs(a % PS - b % PS)[same_plate].max())
                dy = int(np.abs(a // PS - b // PS)[same_plate].max())
        rows.append((lens, int(diff.sum()), flips, dx, dy))
        print(f"{lens:16s} {'ok' if rc0 == 0 and rc1 == 0 else 'fail':7s} {W * H:8d} {int(diff.sum()):7d} {flips:10d} {dx:14d},{dy:<6d} {s0 == s1}", file=out)
    return rows


if __name__ == "__main__":
    size = tuple(int(a) for a in sys.argv[1:4]) if len(sys.argv) >= 4 else (320, 200, 128)
    drift(size=size)
