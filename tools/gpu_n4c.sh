# 4-GPU A/B of the ingest mode (side legs and e2e skipped)
set -x
mkdir -p gpurun_out
for cfg in "direct 0" "staged 8"; do
  set -- $cfg
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29524 bench.py --gpus 4 --steps 4 --warmup 3 --no-extra --no-e2e --no-cpu --ingest-mode $1 --tiles $2 > gpurun_out/r02_bench_n4_$1.json 2> gpurun_out/r02_bench_n4_$1.err
  tail -c 300 gpurun_out/r02_bench_n4_$1.err | grep -v Warning
  python - <<PY
import json
d = json.load(open("gpurun_out/r02_bench_n4_$1.json"))
i = d["ingest"]
print("mode=$1 N=4 value %.4g ms/step %.3f launches %s per-rank %s pull-only %.2f presharded %.2f equal %s %s" % (
    d["value"], d["ms_per_step"], d["gpu_launches"], [round(x, 1) for x in i["per_rank_ms_per_step"]], i["egress_floor_ms"],
    i["ms_per_step_presharded"], i["every_rank_block_equals_resident_path"], i["every_rank_pulled_audio_equals_own_synthesis"]))
PY
done
