#!/bin/bash
# A/B of front-executor variants inside ONE gpurun call:  tools/ab_front.sh TAG [ENV=VAL ...]
set -u
cd "$(dirname "$0")/.."
TAG=${1:-abf}; shift
mkdir -p gpurun_out
if [ "${SKIP_TESTS:-0}" != "1" ]; then
  timeout 900 python -m pytest tests/test_gpu_front.py tests/test_gpu_config5.py tests/test_gpu_accuracy.py tests/test_gpu_select.py -m gpu -x -q > gpurun_out/${TAG}_tests.log 2>&1; echo "front tests rc=$?"; tail -3 gpurun_out/${TAG}_tests.log
fi
run() {
  local name=$1; shift
  for wl in ${WORKLOADS:-mesh_ac mesh_tran}; do
    env "$@" KRB_FRONT_DEBUG=1 timeout 600 python bench.py --workload $wl --steps 2 --warmup 3 --no-e2e --no-cpu --no-others > gpurun_out/${TAG}_${name}_${wl}.json 2> gpurun_out/${TAG}_${name}_${wl}.err
    python - "$name" "$wl" gpurun_out/${TAG}_${name}_${wl}.json <<'PY'
import json, sys
name, wl, f = sys.argv[1:4]
try:
    d = json.loads([l for l in open(f).read().strip().splitlines() if l.startswith('{')][-1])
    r = d['roofline']
    print('%-22s %-10s ms/step %9.3f  value %.4g  frac %.3f  status %s  parity %s' % (name, wl, d['ms_per_step'], d['value'], r['frac'], d['lanes_with_status'], d.get('parity_spot_check')))
    ph = [l.strip() for l in open(f.replace('.json', '.err')) if l.startswith('[krb front]')]
    for l in ph[-6:]: print('      ', l)
except Exception as e:
    print(name, wl, 'ERROR', e, open(f.replace('.json', '.err')).read()[-800:])
PY
  done
}
run default KRB_DUMMY=1
for kv in "$@"; do run "${kv//[^A-Za-z0-9_=]/_}" "$kv"; done
