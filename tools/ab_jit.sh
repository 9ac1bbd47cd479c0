#!/bin/bash
# Specialised lane kernels (vm_jit.cpp) against the interpreter inside ONE gpurun call:  tools/ab_jit.sh TAG
# runs the lane-VM GPU tests, then the nonlinear bench workload with KRB_VM_JIT=1 (default) and KRB_VM_JIT=0.
set -u
cd "$(dirname "$0")/.."
TAG=${1:-jit}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_jit.py -m gpu -x -q > gpurun_out/${TAG}_tests.log 2>&1; echo "jit tests rc=$?"; tail -15 gpurun_out/${TAG}_tests.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_accuracy.py tests/test_gpu_result_io.py tests/test_gpu_select.py tests/test_c_abi.py -m gpu -x -q >> gpurun_out/${TAG}_tests.log 2>&1; echo "parity rc=$?"; tail -4 gpurun_out/${TAG}_tests.log
for jit in 1 0; do
  KRB_VM_JIT=$jit timeout 300 python bench.py --workload rectifier_tran --steps 5 --warmup 3 --no-e2e --no-cpu --no-others > gpurun_out/${TAG}_rect_jit${jit}.json 2> gpurun_out/${TAG}_rect_jit${jit}.err
  python - "$jit" gpurun_out/${TAG}_rect_jit${jit}.json <<'PY'
import json, sys
jit, f = sys.argv[1:3]
try:
    d = json.loads([l for l in open(f).read().strip().splitlines() if l.startswith('{')][-1])
    print('KRB_VM_JIT=%s rectifier_tran ms/step %8.3f value %.4g status %s parity %s' % (jit, d['ms_per_step'], d['value'], d['lanes_with_status'], d.get('parity_spot_check')))
except Exception as e:
    print(jit, 'ERROR', e, open(f.replace('.json', '.err')).read()[-800:])
PY
done
