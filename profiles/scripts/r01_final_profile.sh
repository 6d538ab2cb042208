# final round-1 evidence: ncu launch list + ncu --set full of every kernel of the path (one context: no overlap)
TAG=${1:-h}
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --groups 1 --threads 1 > gpurun_out/ncu_launch_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"upload|pyr_|fast_cells|fast_select|image_align|search_seq|seq_prep|seq_post" -s 37 -c 24 -o gpurun_out/prof_$TAG python bench.py --steps 2 --warmup 3 --groups 1 --threads 1 > gpurun_out/ncu_full_$TAG.log 2>&1
ncu --set full --clock-control none -k regex:"seed_update" -c 2 -o gpurun_out/prof_seed_$TAG python -m pytest tests/test_gpu_seeds.py -m gpu -x -q -k "C2-0-3" > gpurun_out/ncu_seed_$TAG.log 2>&1
ls -la gpurun_out | tail -6
