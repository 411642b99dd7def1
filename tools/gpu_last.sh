# last check of the round: full GPU suite, then the int8 clip / live timings with the feature-quantisation table
mkdir -p gpurun_out
timeout 60 python -m pytest tests -m gpu -q -x 2>&1 | tail -2
timeout 40 python bench.py --model int8 --no-e2e --no-cpu --no-extra --steps 3 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('int8 value %.4g ms %.2f nn %.2f' % (d['value'], d['ms_per_step'], d['kernels']['mixednet']['ms_per_step']))"
timeout 30 python tools/live_time.py int8 40 2>&1 | tail -1
