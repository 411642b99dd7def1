set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | grep -E "^E  |FAILED|passed|failed" | head -30
for m in 0 1 2; do MWW_LIVE_MODE=$m python tools/live_time.py f32 60 2>&1 | tail -1; done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:nn_f32_live2_kernel -s 3 -c 1 -o gpurun_out/nn_live2e -f python tools/live_time.py f32 8 > gpurun_out/ncu_live2d.log 2>&1; tail -2 gpurun_out/ncu_live2d.log
