set -x
nproc; lscpu | grep -E "Model name|Thread|Core|Socket" 
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err; tail -c 3000 gpurun_out/bench_a.json
python bench.py --steps 20 --warmup 5 --groups 8 > gpurun_out/bench_g8.json 2>> gpurun_out/bench_a.err
python bench.py --steps 20 --warmup 5 --groups 32 > gpurun_out/bench_g32.json 2>> gpurun_out/bench_a.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --groups 1 > gpurun_out/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"pyr_down|fast_cells|fast_select|image_align|search_points" -s 54 -c 9 -o gpurun_out/prof_r01 python bench.py --steps 2 --warmup 3 --groups 1 > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
