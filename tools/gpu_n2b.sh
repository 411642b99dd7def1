set -x
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 --no-extra --no-e2e > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err
tail -c 600 gpurun_out/r02_bench_n2.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02_bench_n2.json"))
print("N=2 value %.4g ms/step %.3f" % (d["value"], d["ms_per_step"]))
print("ingest", {k: v for k, v in d["ingest"].items() if k not in ("how",)})
PY
for m in 0 1 2; do MWW_LIVE_MODE=$m python tools/live_time.py f32 60 2>&1 | tail -1; done
python -m pytest tests/test_gpu_parity.py -m gpu -q -k "live" 2>&1 | tail -3
