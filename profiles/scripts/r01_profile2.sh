# ncu launch list + ncu --set full of the resident-sequence chain (one context so launches do not overlap)
TAG=${1:-x}
PAT=${2:-"seq_|image_align|search_seq|fast_cells|fast_select|pyr_|upload"}
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --groups 1 --threads 1 > gpurun_out/ncu_launch_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"$PAT" -s 36 -c 12 -o gpurun_out/prof_$TAG python bench.py --steps 2 --warmup 3 --groups 1 --threads 1 > gpurun_out/ncu_full_$TAG.log 2>&1
ls -la gpurun_out | tail -8
