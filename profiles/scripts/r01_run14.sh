# shared upload streams + bounded PCIe depth: correctness + e2e throughput + timeline
python -m pytest tests/test_gpu_tracker.py -m gpu -x -q 2>&1 | tail -3
for st in 1 2; do for ctas in 2 4 8 16; do
SDVLB_UPLOAD_STREAMS=$st SDVLB_UPLOAD_CTAS=$ctas python bench.py --steps 60 --warmup 5 --sweep 8x4,16x4 2>&1 | grep sweep | sed "s/^/bulk streams=$st ctas=$ctas /"
done; done | tee gpurun_out/sweep_upload2.txt
SDVLB_UPLOAD=ldg SDVLB_UPLOAD_STREAMS=1 python bench.py --steps 60 --warmup 5 --sweep 8x4,16x4 2>&1 | grep sweep | sed "s/^/ldg streams=1 ctas=1 /" | tee -a gpurun_out/sweep_upload2.txt
python profiles/scripts/timeline.py 8 4 64 2 > gpurun_out/tl_8_2c.txt 2>&1
python profiles/scripts/timeline_summary.py gpurun_out/timeline_8_2.json > gpurun_out/tls_8_2c.txt 2>&1
tail -12 gpurun_out/tls_8_2c.txt
