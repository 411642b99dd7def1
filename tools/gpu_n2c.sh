# 2-GPU check of the uneven-share ingest: the N-rank oracle test, then bench.py with the default rule and with a forced share of 0
set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_multirank.py -m gpu -x -q 2>&1 | tail -8
for sh in auto 0; do
  extra=""; [ "$sh" = "0" ] && extra="--ingest-share 0"
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 3 --warmup 3 --no-extra --no-e2e $extra > gpurun_out/r02_bench_n2_share_$sh.json 2> gpurun_out/r02_bench_n2_share_$sh.err
  tail -c 1200 gpurun_out/r02_bench_n2_share_$sh.err
  python - <<PY
import json
d = json.load(open("gpurun_out/r02_bench_n2_share_$sh.json"))
i = d["ingest"]
print("share=$sh N=2 value %.4g ms/step %.3f launches %s" % (d["value"], d["ms_per_step"], d["gpu_launches"]))
print({k: v for k, v in i.items() if k not in ("how", "tile_timeline_ms", "nccl_serial")})
print("timeline rank1 first tiles", i["tile_timeline_ms"]["ranks"][1][:3], "n tiles", [len(r) for r in i["tile_timeline_ms"]["ranks"]])
PY
done
