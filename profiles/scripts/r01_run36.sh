python bench.py --steps 60 --warmup 5 --sweep 2x2,4x2,4x3,4x4,6x3,8x3 --sweep-device 2>&1 | grep sweep | sed "s/^/dev /"
python bench.py --steps 60 --warmup 5 --sweep 4x2,4x4,8x4 2>&1 | grep sweep | sed "s/^/e2e /"
