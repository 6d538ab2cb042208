# bench (both arms) + ncu launch list + ncu --set full of every kernel of the path; outputs under gpurun_out/
TAG=${1:-x}
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
tail -c 3500 gpurun_out/bench_$TAG.json; tail -c 1200 gpurun_out/bench_ref_$TAG.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --groups 1 > gpurun_out/ncu_launch_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"upload|pyr_|fast_cells|fast_select|image_align|search_points" -s 42 -c 14 -o gpurun_out/prof_$TAG python bench.py --steps 2 --warmup 3 --groups 1 > gpurun_out/ncu_full_$TAG.log 2>&1
ls -la gpurun_out | tail -12
