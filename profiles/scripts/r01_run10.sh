ncu --set full --clock-control none --import-source on -k regex:"pyr_|fast_cells" -s 8 -c 3 -o gpurun_out/prof_b python bench.py --steps 2 --warmup 3 --groups 1 > gpurun_out/ncu_full_b.log 2>&1
