python tools/frei_time.py
python -m pytest tests/test_gpu_frei.py -x -q -m gpu 2>&1 | tail -2
