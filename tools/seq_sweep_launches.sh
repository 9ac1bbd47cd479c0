#!/bin/bash
# ncu launch list of the sequential DC sweep (krb_solve_dc_sweep_sequential) through the plain-C client tests/c_abi/seq_sweep.c:
# circuits/diode_iv_curve, V1 0 -> 2 V in 101 points, 3 Monte-Carlo lanes of R1.  Output: gpurun_out/seq_sweep_launches.csv
set -e
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
gcc -std=c11 -O1 -I include tests/c_abi/seq_sweep.c -o /tmp/seq_sweep -L krets_b200 -lkrets_b200 -Wl,-rpath,"$PWD/krets_b200" -lm
python -c "import json; open('/tmp/diode.cir','w').write(json.load(open('tests/golden/golden.json'))['netlists']['diode_iv_curve'])"
timeout 60 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/seq_sweep_launches.csv \
    /tmp/seq_sweep /tmp/diode.cir V1 0 2 0.02 R1 1000 900 1500 > gpurun_out/seq_sweep_ncu.json
python tools/launch_summary.py gpurun_out/seq_sweep_launches.csv 2>/dev/null | head -20 || true
