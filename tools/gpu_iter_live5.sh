#!/bin/bash
# live v3 (bulk-copy ring stages): parity, then timing of variants 1 / 3 and v3's two halves alone, then one ncu capture
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -m gpu -k "live_rings or reset" 2>&1 | tail -5
for v in 1 3; do MWW_LIVE_VARIANT=$v timeout 300 python tools/live_time.py f32 60 2>&1 | tail -1; done
for m in 1 2; do MWW_LIVE_VARIANT=3 MWW_LIVE_MODE=$m timeout 300 python tools/live_time.py f32 60 2>&1 | tail -1; done
MWW_LIVE_VARIANT=3 timeout 300 ncu --set full --clock-control none --import-source on -k regex:nn_f32_live3 -s 6 -c 1 -o gpurun_out/nn_live3 -f python tools/live_time.py f32 8 > gpurun_out/ncu_live3.log 2>&1
ls -la gpurun_out/nn_live3*
