# 8-GPU runs of the multi-GPU bench line: ingest mode x ingest-rank share (side legs and e2e skipped), per-rank tile timelines inside
set -x
mkdir -p gpurun_out
for cfg in "staged 0" "staged 1" "direct 0"; do
  set -- $cfg
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 8 --steps 4 --warmup 3 --no-extra --no-e2e --no-cpu --ingest-mode $1 --ingest-share $2 > gpurun_out/r02_bench_n8_$1_share$2.json 2> gpurun_out/r02_bench_n8_$1_share$2.err
  tail -c 400 gpurun_out/r02_bench_n8_$1_share$2.err | grep -v Warning
  python - <<PY
import json
d = json.load(open("gpurun_out/r02_bench_n8_$1_share$2.json"))
i = d["ingest"]
print("mode=$1 share=$2 N=8 value %.4g ms/step %.3f launches %s per-rank %s pull-only %.2f presharded %.2f equal %s %s" % (
    d["value"], d["ms_per_step"], d["gpu_launches"], [round(x, 1) for x in i["per_rank_ms_per_step"]], i["egress_floor_ms"],
    i["ms_per_step_presharded"], i["every_rank_block_equals_resident_path"], i["every_rank_pulled_audio_equals_own_synthesis"]))
t = i["tile_timeline_ms"]["ranks"]
print("  tiles per rank", [len(r) for r in t], "rank1 first/last tile", (t[1][0], t[1][-1]) if t[1] else None)
PY
done
