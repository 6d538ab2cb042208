ncu --set full --clock-control none --import-source on -k regex:"fast_cells" -s 3 -c 1 -o gpurun_out/prof_fast python bench.py --steps 2 --warmup 3 --groups 1 --threads 1 > gpurun_out/ncu_fast.log 2>&1
ls -la gpurun_out/prof_fast.ncu-rep
