for m in none ce ce_small; do python profiles/scripts/pcie_noise.py $m 2>&1 | tail -1; done | tee gpurun_out/pcie_noise.txt
for st in 1 2 4; do
SDVLB_UPLOAD_STREAMS=$st python bench.py --steps 60 --warmup 5 --e2e-upload dma --sweep 8x4,16x4 2>&1 | grep sweep | sed "s/^/dma streams=$st /"
done | tee gpurun_out/sweep_dma.txt
