#!/bin/bash
# ncu of the dieted ring kernel: as shipped, and with every memory operation of the BOX path removed (lab 23)
mkdir -p gpurun_out
export BLINKY_SPLIT_PERCENT=0
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/r2_c18_ship \
  python scripts/ncu_workloads.py --frames 16 --order gpurun_out/r2_c18_order_ship.json 4k-cube-panini > gpurun_out/r2_c18_ncu.log 2>&1
export BLINKY_B200_LIB=$PWD/blinky_b200/libblinky_b200_lab.so
BLINKY_LAB=23 timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/r2_c18_lab23 \
  python scripts/ncu_workloads.py --frames 16 --order gpurun_out/r2_c18_order_lab23.json 4k-cube-panini >> gpurun_out/r2_c18_ncu.log 2>&1
tail -4 gpurun_out/r2_c18_ncu.log
