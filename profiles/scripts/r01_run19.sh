for st in 2 3 4; do for ctas in 1 2 3; do for pf in 2 3; do
SDVLB_UPLOAD_STREAMS=$st SDVLB_UPLOAD_CTAS=$ctas python bench.py --steps 60 --warmup 5 --prefetch $pf --sweep 8x4,16x4 2>&1 | grep sweep | sed "s/^/streams=$st ctas=$ctas pf=$pf /"
done; done; done | tee gpurun_out/sweep_upload3.txt
