set -u
cd /root/repo
B="python bench.py --workload ladder_mc --steps 1 --warmup 1 --no-e2e --no-cpu --no-others"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:krb_block_stamp_rows -s 1 -c 1 -o gpurun_out/r4t_stamp -f $B > gpurun_out/r4t_ncu.log 2>&1; echo "ncu rc=$?"
python tools/ncu_summary.py gpurun_out/r4t_stamp.ncu-rep gpurun_out/r4t_stamp_ncu_summary.json "ladder 5000 Monte-Carlo (all resistors per lane), 32768 lanes: pre-stamping kernel" "$B" 32768
ncu -i gpurun_out/r4t_stamp.ncu-rep --page source --csv > gpurun_out/r4t_stamp_sass.csv 2>/dev/null; echo "sass rc=$?"
