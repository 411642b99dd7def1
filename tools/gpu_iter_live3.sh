set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -30
for m in 0 1 2; do MWW_LIVE_MODE=$m python tools/live_time.py f32 60 2>&1 | tail -1; done
MWW_LIVE_V1=1 python tools/live_time.py f32 60 2>&1 | tail -1
python tools/live_time.py int8 60 2>&1 | tail -1
