# does the DURATION of the upload (not the PCIe traffic) explain the e2e loss?  HBM-resident frames, upload kernel padded
for ns in 0 7000 14000; do
SDVLB_UPLOAD_MIN_NS_PER_FRAME=$ns python bench.py --steps 60 --warmup 5 --sweep 8x4 --sweep-device 2>&1 | grep sweep | sed "s/^/pad_ns_per_frame=$ns /"
done | tee gpurun_out/upload_pad.txt
