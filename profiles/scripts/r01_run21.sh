ncu --set full --clock-control none --import-source on -k regex:"image_align" -s 6 -c 2 -o gpurun_out/prof_align python bench.py --steps 2 --warmup 3 --groups 1 --threads 1 > gpurun_out/ncu_align.log 2>&1
ls -la gpurun_out/*.ncu-rep
