cd /root/repo; mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/pytest_dith1.log 2>&1; echo "pytest rc $?" >> gpurun_out/pytest_dith1.log
python bench.py --no-cpu-baseline --no-secondary --e2e-steps 16 > gpurun_out/bench_dith1.json 2> gpurun_out/bench_dith1.err
python bench.py --no-cpu-baseline --no-secondary --e2e-steps 16 --dither 0 > gpurun_out/bench_dith0.json 2>> gpurun_out/bench_dith1.err
tail -5 gpurun_out/pytest_dith1.log
python - <<'PY'
import json
for f in ("gpurun_out/bench_dith1.json","gpurun_out/bench_dith0.json"):
    for l in open(f):
        if l.startswith("{"):
            d=json.loads(l); print(f, d["value"], d["roofline"]["us_per_batch"], d["roofline"]["frac"])
PY
