# e2e A/B: pinned audio buffer plain vs write-combined
mkdir -p gpurun_out
for wc in 1 0; do
  timeout 60 python bench.py --no-extra --no-cpu --steps 4 --warmup 3 --e2e-wc $wc > gpurun_out/bench_e2e_wc$wc.json 2> gpurun_out/bench_e2e_wc$wc.err
  python - <<PY
import json
d = json.load(open("gpurun_out/bench_e2e_wc$wc.json"))
e = d["e2e"]
print("wc=$wc e2e %.4g ms/step %.2f h2d %.2f GB/s checksum_ok %s" % (e["value"], e["ms_per_step"], e["h2d_gbs_per_gpu"], e["checksum_matches_device_path"]))
PY
done
