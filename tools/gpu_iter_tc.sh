set -x
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o /tmp/tcgen05_probe tools/probes/tcgen05_pointwise_probe.cu 2>&1 | tail -3
timeout 120 /tmp/tcgen05_probe 2>&1 | tee gpurun_out/tcgen05_probe.txt
timeout 600 python -m pytest tests -m gpu -q 2>&1 | grep -E "^E  |FAILED|passed|failed" | head -40
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --no-extra > gpurun_out/bench_tc.json 2> gpurun_out/bench_tc.err; tail -c 600 gpurun_out/bench_tc.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_tc.json"))
print("tc value %.4g ms %.3f" % (d["value"], d["ms_per_step"]), {k: round(v["ms_per_step"], 3) for k, v in d["kernels"].items()})
PY
MWW_NO_TC=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --no-extra 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('no-tc value %.4g ms %.3f' % (d['value'], d['ms_per_step']), {k: round(v['ms_per_step'], 3) for k, v in d['kernels'].items()})"
