cd /root/repo; mkdir -p gpurun_out
timeout 400 python tests/soak_gpu.py 6000 > gpurun_out/soak_fft.log 2>&1; echo "rc $?" >> gpurun_out/soak_fft.log
SPL_ENGINE=umma timeout 400 python tests/soak_gpu.py 6000 > gpurun_out/soak_umma.log 2>&1; echo "rc $?" >> gpurun_out/soak_umma.log
tail -n 3 gpurun_out/soak_fft.log; tail -n 6 gpurun_out/soak_umma.log
