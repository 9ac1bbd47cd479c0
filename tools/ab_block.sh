#!/bin/bash
# A/B of block-chain-kernel variants inside ONE gpurun call (same box, same clocks):  tools/ab_block.sh TAG [ENV=VAL ...]
# runs the block-kernel tests once, then bench.py for configs 2, 3 and the Monte-Carlo ladder with the default environment and with
# every given ENV=VAL setting; prints kernel_ms of each.
set -u
cd "$(dirname "$0")/.."
TAG=${1:-ab}; shift
mkdir -p gpurun_out
if [ "${SKIP_TESTS:-0}" != "1" ]; then
  timeout 900 python -m pytest tests/test_gpu_chain_block.py tests/test_gpu_accuracy.py -m gpu -x -q > gpurun_out/${TAG}_tests.log 2>&1; echo "block tests rc=$?"; tail -3 gpurun_out/${TAG}_tests.log
  timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "config2 or config3 or executors_agree or chain_monte" >> gpurun_out/${TAG}_tests.log 2>&1; echo "parity subset rc=$?"; tail -2 gpurun_out/${TAG}_tests.log
fi
run() {  # name, env...
  local name=$1; shift
  for wl in ${WORKLOADS:-ladder_dc rc_ac ladder_mc}; do
    env "$@" timeout 300 python bench.py --workload $wl --steps 20 --warmup 3 --no-e2e --no-cpu --no-others > gpurun_out/${TAG}_${name}_${wl}.json 2> gpurun_out/${TAG}_${name}_${wl}.err
    python - "$name" "$wl" gpurun_out/${TAG}_${name}_${wl}.json <<'PY'
import json, sys
name, wl, f = sys.argv[1:4]
try:
    d = json.loads([l for l in open(f).read().strip().splitlines() if l.startswith('{')][-1])
    r = d['roofline']
    print('%-22s %-10s ms/step %8.4f  kernel_ms %8.4f  value %.4g  frac %.3f  status %s  parity %s' % (name, wl, d['ms_per_step'], r['kernel_ms'], d['value'], r['frac'], d['lanes_with_status'], (d.get('parity_spot_check') or {}).get('max_rel_err_vs_oracle')))
except Exception as e:
    print(name, wl, 'ERROR', e, open(f.replace('.json', '.err')).read()[-600:])
PY
  done
}
run default KRB_DUMMY=1
for kv in "$@"; do run "${kv//[^A-Za-z0-9_=]/_}" "$kv"; done
