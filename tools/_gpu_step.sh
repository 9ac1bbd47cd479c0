python -m pytest tests/test_gpu_front.py -x -q 2>&1 | tail -15
export KRB_FRONT_DEBUG=1
python tools/measure_mesh.py --grid 224 --lanes 128 --steps 4 --modes tran,ac --reps 1 2>&1 | tail -16
