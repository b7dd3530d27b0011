python -m pytest tests/test_facade_gpu.py -x -q -m gpu 2>&1 | tail -12
