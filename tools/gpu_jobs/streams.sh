cd /root/repo; mkdir -p gpurun_out
for s in 1 2 3; do
python bench.py --no-cpu-baseline --no-secondary --e2e-steps 8 --streams $s 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][0]); print('streams', d['harness']['streams'], 'value %.3fM ms/step %.5f A %.2f' % (d['value']/1e6, d['ms_per_step'], d['roofline']['us_per_batch']))"
done | tee gpurun_out/streams.txt
