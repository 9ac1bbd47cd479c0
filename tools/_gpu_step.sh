ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1k.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_bench_k.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:krb_chain -s 3 -c 1 -o gpurun_out/prof_chain_r1k python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_full_k.log 2>&1
ncu -i gpurun_out/prof_chain_r1k.ncu-rep --page raw --csv > gpurun_out/prof_chain_r1k_raw.csv
ls -la gpurun_out/prof_chain_r1k.ncu-rep
