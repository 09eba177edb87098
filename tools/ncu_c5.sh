#!/bin/bash
# ncu --set full capture of the general kernel draining a C5 batch (2^16 requests) + per-line tables (GPU box).
N=${1:-65536}
ncu --set full --import-source on --clock-control none -k regex:check_kernel -s 7 -c 1 -o gpurun_out/c5_general -f \
    python bench.py --workload C5 --requests $N --steps 1 --warmup 3 --batches-per-step 1 --no-e2e --no-cpu --no-verify --no-secondary > gpurun_out/ncu_c5.log 2>&1
python tools/ncu_summary.py gpurun_out/c5_general.ncu-rep > gpurun_out/c5_general_ncu_full.json 2>>gpurun_out/ncu_c5.log
python tools/ncu_lines.py gpurun_out/c5_general.ncu-rep 70 > gpurun_out/c5_general_lines.txt 2>>gpurun_out/ncu_c5.log
ncu -i gpurun_out/c5_general.ncu-rep --page source --csv --print-source=cuda > gpurun_out/c5_general_cuda_source.csv 2>>gpurun_out/ncu_c5.log
tail -3 gpurun_out/ncu_c5.log; head -c 600 gpurun_out/c5_general_ncu_full.json
