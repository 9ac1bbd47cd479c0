#!/bin/bash
# ncu evidence for the specialised lane kernel (vm_jit.cpp) against the interpreter:  tools/jit_profile.sh TAG
# (1) launch lists (time + executed warp instructions per launch) of bench.py --workload rectifier_tran with KRB_VM_JIT=1 / 0,
# (2) one --set full capture of the specialised step kernel, summarised by tools/ncu_summary.py.
set -u
cd "$(dirname "$0")/.."
TAG=${1:-jitprof}
mkdir -p gpurun_out
B="python bench.py --workload rectifier_tran --steps 1 --warmup 1 --no-e2e --no-cpu --no-others"
for jit in 1 0; do
  KRB_VM_JIT=$jit timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,launch__registers_per_thread --clock-control none -k regex:"krb_jit_vm|krb_lane_vm" -c 8 --csv \
    --log-file gpurun_out/${TAG}_launches_jit${jit}.csv $B > gpurun_out/${TAG}_ncu_jit${jit}.log 2>&1; echo "ncu launches jit=$jit rc=$?"
done
# the step kernel is every second launch (OP kernel first): skip 1, capture 1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"krb_jit_vm" -s 1 -c 1 -o gpurun_out/${TAG}_jit_full -f $B > gpurun_out/${TAG}_ncu_full.log 2>&1; echo "ncu full rc=$?"
python tools/ncu_summary.py gpurun_out/${TAG}_jit_full.ncu-rep gpurun_out/${TAG}_jit_ncu_summary.json "circuits/rectifier transient, 1000 BE steps x 16384 Monte-Carlo lanes, NR per step (specialised lane kernel)" "KRB_VM_JIT=1 $B" 16384 && echo summary ok
grep -v "^==" gpurun_out/${TAG}_launches_jit1.csv | tail -12; grep -v "^==" gpurun_out/${TAG}_launches_jit0.csv | tail -12
