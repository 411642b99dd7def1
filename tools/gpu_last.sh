set -x
mkdir -p gpurun_out
timeout 200 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
timeout 100 python tools/live_time.py f32 60 2>&1 | tail -1
