cd /root/repo; mkdir -p gpurun_out
python tools/trace_tail.py 1.0 8 > gpurun_out/trace_tail_k8.txt 2>&1
python tools/trace_tail.py 1.0 1 > gpurun_out/trace_tail_k1.txt 2>&1
cat gpurun_out/trace_tail_k8.txt gpurun_out/trace_tail_k1.txt
