set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_live_v2c.json 2> gpurun_out/bench_live_v2c.err; tail -c 500 gpurun_out/bench_live_v2c.err
python - <<PY
import json
d = json.load(open("gpurun_out/bench_live_v2c.json"))
for k in ("live", "live_int8"):
    print("v2c", k, "ms/call %.4f" % d[k]["ms_per_call"], {a: round(b, 4) for a, b in d[k]["kernels_ms_per_call"].items()}, "hbm frac %.3f" % d[k]["roofline"]["frac"])
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:nn_f32_live2_kernel -s 6 -c 1 -o gpurun_out/nn_live2c -f python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_live2b.log 2>&1; tail -2 gpurun_out/ncu_live2b.log
