# 2-GPU run: N-rank ingest vs the oracle (pytest), then the default bench under torchrun (value = pull + compute + gather)
set -x
mkdir -p gpurun_out
nvidia-smi topo -m 2>/dev/null | head -12
python -m pytest tests/test_gpu_multirank.py -m gpu -x -q 2>&1 | tail -8
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 --no-extra > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err
tail -c 1500 gpurun_out/r02_bench_n2.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02_bench_n2.json"))
print("N=2 value %.4g ms/step %.3f" % (d["value"], d["ms_per_step"]))
print("ingest", {k: v for k, v in d["ingest"].items() if k not in ("how",)})
print("e2e", d["e2e"])
PY
