set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for occ in 3 4; do
MWW_K1_OCC=$occ python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_occ$occ.json 2> gpurun_out/bench_occ$occ.err; tail -c 500 gpurun_out/bench_occ$occ.err
python - <<PY
import json
d = json.load(open("gpurun_out/bench_occ$occ.json"))
print("occ $occ value %.4g ms %.3f" % (d["value"], d["ms_per_step"]), {k: round(v["ms_per_step"], 3) for k, v in d["kernels"].items()})
print("feat", {k: (round(v["ms"], 3), "%.3g" % v["windows_per_s"]) for k, v in d["features_only"].items()})
PY
done
B="python bench.py --streams 8192 --steps 1 --warmup 1 --no-e2e --no-cpu --no-extra"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k1_spectral_kernel -s 1 -c 1 -o gpurun_out/k1_occ4 -f $B > gpurun_out/ncu_k1_occ4.log 2>&1; tail -2 gpurun_out/ncu_k1_occ4.log
