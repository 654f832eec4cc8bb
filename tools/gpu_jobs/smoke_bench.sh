cd /root/repo; mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" > gpurun_out/smoke_t.log 2>&1; echo "smoke rc $?" >> gpurun_out/smoke_t.log
( time python bench.py ) > gpurun_out/bench_default4.json 2> gpurun_out/bench_default4.err
tail -5 gpurun_out/smoke_t.log; tail -4 gpurun_out/bench_default4.err
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/bench_default4.json") if l.startswith("{")][0])
print(d["value"], d["roofline"]["us_per_batch"], d["roofline"]["frac"], d["single_batch_launch"], d["e2e"]["value"], d["copy_control"]["value"])
PY
