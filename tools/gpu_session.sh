#!/bin/bash
# One GPU session: the -m gpu tests, smoke(), the default bench line, the reference arm, the launch list and one full ncu
# capture of the headline kernel.  Everything lands in gpurun_out/ (scratch); summaries are copied into profiles/ by hand.
#   gpurun --timeout 2400 -- 'bash tools/gpu_session.sh r2a'
set -u
cd "$(dirname "$0")/.."
TAG=${1:-r2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
if [ "${SKIP_TESTS:-0}" != "1" ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_tests.log 2>&1; echo "pytest -m gpu rc=$?"; tail -5 gpurun_out/${TAG}_tests.log
  timeout 300 python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${TAG}_smoke.log
fi
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; echo "bench ref rc=$?"
python - "$TAG" <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.loads([l for l in open(f'gpurun_out/{tag}_bench.json').read().strip().splitlines() if l.startswith('{')][-1])
    r = d['roofline']
    print('headline: ms/step %.4f  value %.4g  kernel_ms %.4f  frac %.3f  frac_alg %.2f  e2e %s  parity %s' % (
        d['ms_per_step'], d['value'], r['kernel_ms'], r['frac'], r['frac_alg_model'], (d.get('e2e') or {}).get('value'), d.get('parity_spot_check')))
    print('e2e:', {k: v for k, v in (d.get('e2e') or {}).items() if k in ('value', 'd2h_GBps', 'link_ceiling_GBps', 'frac_of_link_ceiling', 'selected')})
    if d.get('gather'): print('gather:', d['gather'])
    for o in d.get('other_configs', []):
        if 'error' in o: print('  ', o['name'], 'ERROR', o['error']); continue
        if 'roofline' not in o: print('   %-16s %s' % (o['name'], {k: v for k, v in o.items() if k not in ('workload', 'name')})); continue
        comp = o['roofline'].get('compute')
        print('   %-16s ms/step %10.3f  value %.4g  frac %.3f%s  status %d  parity %s' % (o['name'], o['ms_per_step'], o['value'], o['roofline']['frac'],
              ('  fp64 %.3f' % comp['frac']) if comp else '', o['lanes_with_status'], o.get('parity_spot_check')))
except Exception as e:
    print('bench parse error', e); print(open(f'gpurun_out/{tag}_bench.err').read()[-1500:])
PY
if [ "${SKIP_NCU:-0}" != "1" ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${TAG}_launches.csv \
      python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-others > gpurun_out/${TAG}_ncu_launches.log 2>&1; echo "ncu launch list rc=$?"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:krb_chain_block -s 3 -c 1 -o gpurun_out/${TAG}_block_dc -f \
      python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-others > gpurun_out/${TAG}_ncu_dc.log 2>&1; echo "ncu full (dc) rc=$?"
  python tools/ncu_summary.py gpurun_out/${TAG}_block_dc.ncu-rep gpurun_out/${TAG}_block_dc_ncu_summary.json \
      "BASELINE config 2: resistor_ladder_5000 DC sweep, 65536 lanes, block-partitioned chain kernel" \
      "ncu --set full --clock-control none --import-source on -k regex:krb_chain_block -s 3 -c 1 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-others" 65536
fi
