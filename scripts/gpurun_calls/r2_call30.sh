#!/bin/bash
mkdir -p gpurun_out
{
python - <<'PY'
import sys; sys.argv=['x']
exec(open('scripts/pcie_probe.py').read().split("if len(sys.argv) > 1")[0])
for a,b in [(64<<20,0),(0,64<<20),(64<<20,64<<20),(10<<20,8<<20),(10<<20,0),(0,8<<20)]:
    h,d=bw(a,b)
    print(f"H2D {a>>20:3d} MiB + D2H {b>>20:3d} MiB per round: H2D {h:6.1f} GB/s  D2H {d:6.1f} GB/s", flush=True)
PY
for cfg in "3 1" "4 1" "6 1" "3 0" "2 1" "8 1"; do set -- $cfg; BLINKY_HOST_SLOTS=$1 BLINKY_E2E_BATCH=$2 python scripts/pcie_probe.py e2e | sed "s/^/slots=$1 batch=$2 /"; done
} > gpurun_out/r2_c30_pcie.log 2>&1
cat gpurun_out/r2_c30_pcie.log
( timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "host" 2>&1 | tail -3 )
