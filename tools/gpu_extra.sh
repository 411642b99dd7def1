set -x
mkdir -p gpurun_out
B="python bench.py --streams 8192 --steps 1 --warmup 1 --no-e2e --no-cpu --no-extra"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:nn_f32_clip_tc_kernel -s 1 -c 1 -o gpurun_out/nn_tc -f $B > gpurun_out/ncu_nn.log 2>&1
bash tools/sanitize.sh
