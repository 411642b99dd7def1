set -x
mkdir -p gpurun_out
for tc in 1 0; do
if [ $tc = 0 ]; then export MWW_NO_TC=1; else unset MWW_NO_TC; fi
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 5 --warmup 3 --no-extra --no-e2e > gpurun_out/r02_bench_n8_tc$tc.json 2> gpurun_out/r02_bench_n8_tc$tc.err
tail -c 300 gpurun_out/r02_bench_n8_tc$tc.err
python - <<PY
import json
d = json.load(open("gpurun_out/r02_bench_n8_tc$tc.json"))
i = d["ingest"]
print("tc=$tc N=8 value %.4g ms/step %.3f presharded %.2f pull-only %.2f per-rank" % (d["value"], d["ms_per_step"], i["ms_per_step_presharded"], i["egress_floor_ms"]), [round(x, 1) for x in i["per_rank_ms_per_step"]])
PY
done
