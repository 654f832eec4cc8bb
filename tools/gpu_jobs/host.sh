cd /root/repo; mkdir -p gpurun_out
python tools/host_overhead.py > gpurun_out/host_overhead_full.txt 2>&1
cat gpurun_out/host_overhead_full.txt | head -60
