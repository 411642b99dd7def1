# 8-GPU run of the default multi-GPU bench line (auto ingest share), e2e included, side legs for the other configs skipped
set -x
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 8 --steps 5 --warmup 3 --no-extra > gpurun_out/r02_bench_n8_auto.json 2> gpurun_out/r02_bench_n8_auto.err
tail -c 1500 gpurun_out/r02_bench_n8_auto.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02_bench_n8_auto.json"))
i = d["ingest"]
print("N=8 value %.4g ms/step %.3f launches %s" % (d["value"], d["ms_per_step"], d["gpu_launches"]))
print({k: v for k, v in i.items() if k not in ("how", "tile_timeline_ms", "nccl_serial")})
print("e2e", d.get("e2e"))
for r, t in enumerate(i["tile_timeline_ms"]["ranks"]):
    print("rank", r, "tiles", len(t), "last", t[-1] if t else None)
PY
