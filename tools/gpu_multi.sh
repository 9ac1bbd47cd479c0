#!/bin/bash
# Multi-GPU bench line exactly as the driver launches it:  gpurun --gpus N --timeout 1800 -- 'bash tools/gpu_multi.sh N TAG'
set -u
cd "$(dirname "$0")/.."
N=${1:-2}; TAG=${2:-r2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/${TAG}_topo_n${N}.txt 2>&1
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 10 --warmup 3 ${BENCH_ARGS:-} \
    > gpurun_out/${TAG}_bench_n${N}.json 2> gpurun_out/${TAG}_bench_n${N}.err; echo "bench n$N rc=$?"
python - "$TAG" "$N" <<'PY'
import json, sys
tag, n = sys.argv[1:3]
try:
    d = json.loads([l for l in open(f'gpurun_out/{tag}_bench_n{n}.json').read().strip().splitlines() if l.startswith('{')][-1])
    print('n_gpus', d['n_gpus'], 'ms/step %.4f value %.4g' % (d['ms_per_step'], d['value']))
    print('e2e:', {k: v for k, v in (d.get('e2e') or {}).items() if k in ('value', 'd2h_GBps', 'link_ceiling_GBps', 'frac_of_link_ceiling', 'selected', 'host_cpus_rank0')})
    print('gather:', d.get('gather'))
    for o in d.get('other_configs', []):
        if 'roofline' not in o: print('   %-16s %s' % (o['name'], {k: v for k, v in o.items() if k not in ('workload', 'name')})); continue
        print('   %-16s ms/step %10.3f  value %.4g  frac %.3f  status %d' % (o['name'], o['ms_per_step'], o['value'], o['roofline']['frac'], o['lanes_with_status']))
except Exception as e:
    print('parse error', e); print(open(f'gpurun_out/{tag}_bench_n{n}.err').read()[-2500:])
PY
