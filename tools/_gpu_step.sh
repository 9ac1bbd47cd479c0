python -m pytest tests/test_gpu_front.py -x -q 2>&1 | tail -15
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/front_launches_a.csv python tools/measure_mesh.py --grid 224 --lanes 32 --steps 2 --modes tran --reps 1 > gpurun_out/front_ncu_a.log 2>&1
tail -2 gpurun_out/front_ncu_a.log
