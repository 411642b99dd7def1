# one GPU iteration on the frontend kernel: parity tests, the default bench line, one ncu --set full capture of the clip frontend kernel
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -6
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_iter.json 2> gpurun_out/bench_iter.err; tail -c 1500 gpurun_out/bench_iter.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_iter.json"))
print("value %.4g ms %.3f" % (d["value"], d["ms_per_step"]), {k: round(v["ms_per_step"], 3) for k, v in d["kernels"].items()})
print("e2e", d["e2e"]["ms_per_step"], "int8", d["int8"]["ms_per_step"], "live", d["live"]["ms_per_call"], d["live"]["roofline"]["frac"], "live_int8", d["live_int8"]["ms_per_call"], d["live_int8"]["roofline"]["frac"])
print("feat", {k: (round(v["ms"], 3), "%.3g" % v["windows_per_s"]) for k, v in d["features_only"].items()})
PY
B="python bench.py --streams 8192 --steps 1 --warmup 1 --no-e2e --no-cpu --no-extra"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k1_spectral_kernel -s 1 -c 1 -o gpurun_out/k1_iter -f $B > gpurun_out/ncu_k1_iter.log 2>&1; tail -3 gpurun_out/ncu_k1_iter.log
