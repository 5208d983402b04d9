#!/bin/bash
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-100
( timeout 200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "c1 or native or ragged or batches" 2>&1 | tail -2 )
