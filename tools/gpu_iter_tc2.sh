set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q 2>&1 | grep -E "^E  |FAILED|passed|failed" | head -40
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --no-extra > gpurun_out/bench_tc.json 2> gpurun_out/bench_tc.err; tail -c 600 gpurun_out/bench_tc.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_tc.json"))
print("tc value %.4g ms %.3f" % (d["value"], d["ms_per_step"]), {k: round(v["ms_per_step"], 3) for k, v in d["kernels"].items()})
PY
B="python bench.py --streams 8192 --steps 1 --warmup 1 --no-e2e --no-cpu --no-extra"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:nn_f32_clip_tc_kernel -s 1 -c 1 -o gpurun_out/nn_tc2 -f $B > gpurun_out/ncu_nn_tc.log 2>&1; tail -2 gpurun_out/ncu_nn_tc.log
