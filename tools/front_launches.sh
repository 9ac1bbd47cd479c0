#!/bin/bash
# ncu launch list of the front executor's factor kernels (per-launch time, DRAM bytes, pipe activity) for config 5:
#   tools/front_launches.sh TAG [ac|tran|dc] [lanes] [kernel regex, default krb_front_factor]
set -u
cd "$(dirname "$0")/.."
TAG=${1:-fl}; MODE=${2:-ac}; LANES=${3:-256}; RX=${4:-krb_front_factor}
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__shared_mem_per_block_dynamic
timeout 1200 ncu --metrics $M --clock-control none -k regex:$RX --csv --log-file gpurun_out/${TAG}_${MODE}_factor_launches.csv \
    python tools/measure_mesh.py --grid 224 --lanes $LANES --modes $MODE --steps 2 --reps 1 > gpurun_out/${TAG}_${MODE}_ncu.log 2>&1
echo "ncu rc=$?"; wc -l gpurun_out/${TAG}_${MODE}_factor_launches.csv
