#!/bin/bash
# round-2 GPU call 1: baseline tests, precision table, conv1 chunk A/B
set -u
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/c1_smi.txt 2>&1
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > $O/c1_pytest.log
(timeout 900 python tools/diag_precision.py 2>&1 | tail -60) > $O/c1_precision.log
for mb in 0 40 64 100 128; do
  echo "DGP_CONV1_CHUNK_MB=$mb" >> $O/c1_chunk_ab.log
  DGP_CONV1_CHUNK_MB=$mb timeout 300 python bench.py --steps 20 --no-train --no-cpu-baseline 2>>$O/c1_chunk_ab.err | tail -1 >> $O/c1_chunk_ab.log
done
tail -3 $O/c1_pytest.log
