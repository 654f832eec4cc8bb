cd /root/repo; mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu3.log 2>&1; echo "pytest rc $?" >> gpurun_out/pytest_gpu3.log
python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" > gpurun_out/smoke3.log 2>&1; echo "smoke rc $?" >> gpurun_out/smoke3.log
( time python bench.py ) > gpurun_out/bench_default3.json 2> gpurun_out/bench_default3.err
( time python bench.py --impl reference ) > gpurun_out/bench_ref3.json 2> gpurun_out/bench_ref3.err
tail -3 gpurun_out/pytest_gpu3.log; tail -2 gpurun_out/smoke3.log; tail -4 gpurun_out/bench_default3.err; tail -4 gpurun_out/bench_ref3.err
