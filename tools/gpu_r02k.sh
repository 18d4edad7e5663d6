#!/bin/bash
out=gpurun_out; tag=r02k
for cl in 1 0; do for mb in 64 0; do echo "CLOSERLESS=$cl L2MB=$mb"; CCQP_SYNC_CLOSERLESS=$cl CCQP_L2_RESIDENT_MB=$mb timeout 300 python tools/ab_apgd.py 2>&1 | tail -1; done; done
for cl in 1 0; do echo "n=4096 CLOSERLESS=$cl"; CCQP_SYNC_CLOSERLESS=$cl timeout 300 python tools/bench_n4096.py 2>&1 | tail -1; done
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,temperature.gpu --format=csv
