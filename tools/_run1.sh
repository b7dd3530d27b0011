python -m pytest tests/test_dropins_gpu.py -x -q -m gpu 2>&1 | tail -8
