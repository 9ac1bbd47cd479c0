python -m pytest tests/test_gpu_front.py -x -q 2>&1 | tail -2
export KRB_FRONT_DEBUG=1
python tools/measure_mesh.py --grid 224 --lanes 512 --steps 4 --modes tran --reps 1 2>&1 | grep -v "  end  " | tail -5 | cut -c1-300
