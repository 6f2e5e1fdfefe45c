#!/bin/bash
# Regenerates the raw material of profiles/ on a GPU box (run through gpurun; outputs under gpurun_out/).
#   bash tools/profile_round.sh r01
set -u
R=${1:-r01}
O=gpurun_out
mkdir -p $O
# 1. bench line (not under a profiler)
timeout 600 python bench.py 2>$O/bench_${R}.err | tail -1 > $O/bench_${R}_1gpu.json
# 2. bandwidth-kernel microbenchmarks
timeout 300 python tools/bench_kernels.py > $O/${R}_kernels_microbench.jsonl 2>$O/bench_kernels.err
# 3. launch list of ONE step (61 launches: 60 of ours + one torch copy), with DRAM bytes
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
  -s 183 -c 61 --csv --log-file $O/${R}_launches_dram.csv python bench.py --steps 1 --warmup 3 > $O/ncu_launch.log 2>&1
# 4. --set full on representative GEMM layers (index within the 54 GEMM launches of a step; 3 warm-up steps skipped)
for idx in 3 4 26 45 46; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_gemm -s $((162 + idx)) -c 1 \
    -o $O/${R}_gemm_$idx python bench.py --steps 1 --warmup 3 > $O/ncu_gemm_$idx.log 2>&1
done
# 5. --set full on the bandwidth kernels at GPU-filling sizes
timeout 300 ncu --set full --clock-control none --import-source on -k regex:softargmax -s 4 -c 2 -o $O/${R}_softargmax \
  python tools/prof_softargmax.py 4096 94 104 4 nodlc > $O/ncu_sa.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:potentials -s 2 -c 1 -o $O/${R}_potentials \
  python tools/bench_kernels.py > $O/ncu_pot.log 2>&1
ls -la $O
