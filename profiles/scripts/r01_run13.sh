# bulk-copy (TMA) upload vs load/store upload: correctness + e2e throughput + timeline
python -m pytest tests/test_gpu_tracker.py -m gpu -x -q 2>&1 | tail -3
for mode in bulk ldg; do for ctas in 1 2 3 6; do
SDVLB_UPLOAD=$mode SDVLB_UPLOAD_CTAS=$ctas python bench.py --steps 60 --warmup 5 --sweep 8x4,16x4 2>&1 | grep sweep | sed "s/^/$mode ctas=$ctas /"
done; done | tee gpurun_out/sweep_upload.txt
python profiles/scripts/timeline.py 8 4 64 2 > gpurun_out/tl_8_2b.txt 2>&1
python profiles/scripts/timeline_summary.py gpurun_out/timeline_8_2.json > gpurun_out/tls_8_2b.txt 2>&1
tail -12 gpurun_out/tls_8_2b.txt
