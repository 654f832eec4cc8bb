cd /root/repo; mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/pytest_t.log 2>&1; echo "pytest rc $?" >> gpurun_out/pytest_t.log
python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" > gpurun_out/smoke_t.log 2>&1; echo "smoke rc $?" >> gpurun_out/smoke_t.log
tail -30 gpurun_out/pytest_t.log; tail -4 gpurun_out/smoke_t.log
