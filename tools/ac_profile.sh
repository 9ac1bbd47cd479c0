#!/bin/bash
# ncu --set full of the complex block chain kernel (BASELINE config 3):  tools/ac_profile.sh TAG
set -u
cd "$(dirname "$0")/.."
TAG=${1:-acprof}
mkdir -p gpurun_out
B="python bench.py --workload rc_ac --steps 1 --warmup 1 --no-e2e --no-cpu --no-others"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:krb_chain_block -s 1 -c 1 -o gpurun_out/${TAG}_block_ac -f $B > gpurun_out/${TAG}_ncu.log 2>&1; echo "ncu rc=$?"
python tools/ncu_summary.py gpurun_out/${TAG}_block_ac.ncu-rep gpurun_out/${TAG}_block_ac_ncu_summary.json "BASELINE config 3: rc_ladder_5000 AC, 131072 lanes c64, block-partitioned chain kernel" "$B" 131072
ncu -i gpurun_out/${TAG}_block_ac.ncu-rep --page source --csv > gpurun_out/${TAG}_block_ac_sass.csv 2>/dev/null; echo "sass rc=$?"
