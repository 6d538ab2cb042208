python -m pytest tests/test_gpu_seeds.py -m gpu -x -q -s 2>&1 | tail -25
