cd /root/repo; mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/pytest_q.log 2>&1; echo "pytest rc $?" >> gpurun_out/pytest_q.log
python bench.py --no-cpu-baseline --no-secondary --e2e-steps 16 > gpurun_out/bench_q1.json 2> gpurun_out/bench_q.err
python bench.py --no-cpu-baseline --no-secondary --e2e-steps 16 --dither 0 > gpurun_out/bench_q0.json 2>> gpurun_out/bench_q.err
tail -n 4 gpurun_out/pytest_q.log
python - <<'PY'
import json
for f in ("gpurun_out/bench_q1.json","gpurun_out/bench_q0.json"):
    for l in open(f):
        if l.startswith("{"):
            d=json.loads(l); print(f, "value %.3fM A %.2f us frac %.4f single-batch %.1f us" % (d["value"]/1e6, d["roofline"]["us_per_batch"], d["roofline"]["frac"], d["single_batch_launch"]["us_per_step"]))
PY
