cd /root/repo; mkdir -p gpurun_out
B="python bench.py --steps 32 --warmup 8 --no-cpu-baseline --no-secondary --e2e-steps 16"
$B > gpurun_out/r2_plain.json 2> gpurun_out/r2_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches.csv $B > gpurun_out/ncu_ll.log 2>&1
for wl in aishell hkust libri; do
ncu --set full --clock-control none --import-source on -k regex:fbank_warp -s 4 -c 2 -o gpurun_out/r2_prof_fft_$wl -f python bench.py --workload $wl --steps 32 --warmup 8 --no-cpu-baseline --no-secondary --e2e-steps 8 > gpurun_out/ncu_full_$wl.log 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:post_kernel -s 2 -c 1 -o gpurun_out/r2_prof_post -f python bench.py --steps 32 --warmup 8 --no-cpu-baseline --no-secondary --e2e-steps 8 > gpurun_out/ncu_full_post.log 2>&1
for k in 1 2 4 8 16; do python bench.py --steps 64 --warmup 8 --no-cpu-baseline --no-secondary --batches-per-launch $k 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('K',r['batches_per_launch'],'value %.2fM A us/batch %.2f frac %.4f e2e %.3fM e2e16 %.3fM copy %.3fM'%(d['value']/1e6,r['us_per_batch'],r['frac'],d['e2e']['value']/1e6,d['e2e_int16']['value']/1e6,d['copy_control']['value']/1e6))"; done | tee gpurun_out/r2_ksweep.txt
python tools/host_overhead.py 2>&1 | head -8 | tee gpurun_out/r2_host_overhead.txt
tail -3 gpurun_out/r2_plain.err; head -c 600 gpurun_out/r2_plain.json
