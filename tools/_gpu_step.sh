python -m pytest tests -x -q -m gpu 2>&1 | tail -4
python tools/measure_configs.py 2>&1 | tail -2
