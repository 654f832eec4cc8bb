cd /root/repo; mkdir -p gpurun_out
nvidia-smi -L | wc -l > gpurun_out/s8_ngpu.txt
python -m pytest tests/test_gpu_multi.py -m gpu -q > gpurun_out/s8_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/s8_pytest.log
run() {  # n, tag, extra args
  n=$1; tag=$2; shift 2
  if [ "$n" = 1 ]; then
    timeout 600 python bench.py --gpus 1 --no-cpu-baseline --no-secondary "$@" 2> gpurun_out/s8_${tag}_1.err | grep '^{"metric' > gpurun_out/s8_${tag}_1.json
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) bench.py --gpus $n --no-cpu-baseline --no-secondary "$@" 2> gpurun_out/s8_${tag}_$n.err | grep '^{"metric' > gpurun_out/s8_${tag}_$n.json
  fi
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/s8_${tag}_$n.json").read())
    e = d.get("e2e") or {}
    c = d.get("copy_control") or {}
    print("${tag} N=$n value %.3fM ms/step %.4f e2e %.3fM copy %.3fM" % (d["value"] / 1e6, d["ms_per_step"], (e.get("value") or 0) / 1e6, (c.get("value") or 0) / 1e6), {k: d[k] for k in ("epoch",) if k in d})
except Exception as ex:
    print("${tag} N=$n FAILED", ex)
PY
}
for n in 1 2 4 8; do run $n aishell; done
for n in 1 2 4 8; do run $n epoch --workload epoch; done
for n in 1 2 4 8; do run $n librishard --workload libri --shard; done
tail -3 gpurun_out/s8_pytest.log
