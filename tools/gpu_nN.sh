# N-GPU bench (N = $1): value = pull + compute + gather; e2e with NUMA-local pinned buffers
N=$1
set -x
mkdir -p gpurun_out
nvidia-smi topo -m 2>/dev/null | grep -E "^GPU" | cut -c1-120
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps 5 --warmup 3 --no-extra > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err
tail -c 800 gpurun_out/r02_bench_n$N.err
python - <<PY
import json
d = json.load(open("gpurun_out/r02_bench_n$N.json"))
print("N=$N value %.4g ms/step %.3f" % (d["value"], d["ms_per_step"]))
print("ingest", {k: v for k, v in d["ingest"].items() if k not in ("how",)})
print("e2e", {k: d["e2e"][k] for k in ("value", "ms_per_step", "h2d_gbs_per_gpu", "host_buffers_numa_node")})
PY
