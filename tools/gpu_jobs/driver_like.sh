cd /root/repo; mkdir -p gpurun_out
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --impl reference --gpus 2 --steps 20 --warmup 3 ) > gpurun_out/drv_ref_2.out 2> gpurun_out/drv_ref_2.err
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3 ) > gpurun_out/drv_ours_2.out 2> gpurun_out/drv_ours_2.err
grep -c '^{' gpurun_out/drv_ref_2.out gpurun_out/drv_ours_2.out
grep '^{' gpurun_out/drv_ref_2.out | cut -c1-200; grep '^{' gpurun_out/drv_ours_2.out | cut -c1-260
tail -n 3 gpurun_out/drv_ref_2.err; tail -n 3 gpurun_out/drv_ours_2.err
