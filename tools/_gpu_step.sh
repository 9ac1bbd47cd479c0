python __graft_entry__.py smoke 2>&1 | tail -2
python bench.py --steps 5 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_r1j.log; cut -c1-400 gpurun_out/bench_r1j.log
for w in rc_ac dual_rc_tran mesh_ac mesh_tran; do python bench.py --workload $w --steps 3 --warmup 3 2>&1 | tail -1 | tee -a gpurun_out/bench_workloads_r1.jsonl | cut -c1-700; done
