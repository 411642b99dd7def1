# Round-end measurement set on ONE B200 (run under gpurun from the repo root); outputs land in gpurun_out/ and are
# summarised into profiles/ with tools/ncu_summary.py / tools/ncu_phases.py afterwards.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
# launch list of the bench command: ONLY the library's kernels (without the name filter the first 400 launches are the torch
# kernels that synthesise the input audio and the list never reaches the timed region -- what happened to the r01 file)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:k1_|k2_|nn_|carry_|fill_' -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-extra > gpurun_out/launches_bench.log 2>&1
B="python bench.py --streams 8192 --steps 1 --warmup 1 --no-e2e --no-cpu --no-extra"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k1_spectral_kernel -s 1 -c 1 -o gpurun_out/k1 -f $B > gpurun_out/ncu_k1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:nn_f32_clip_tc_kernel -s 1 -c 1 -o gpurun_out/nn_tc -f $B > gpurun_out/ncu_nn.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:nn_f32_live -s 6 -c 1 -o gpurun_out/nn_live -f python tools/live_time.py f32 8 > gpurun_out/ncu_live.log 2>&1
ls -la gpurun_out/ | tail -15
timeout 300 ncu --set full --clock-control none --import-source on -k regex:nn_i8_clip_kernel -s 1 -c 1 -o gpurun_out/nn_i8 -f python bench.py --model int8 --streams 8192 --steps 1 --warmup 1 --no-e2e --no-cpu --no-extra > gpurun_out/ncu_i8.log 2>&1
for v in 1 3; do MWW_LIVE_VARIANT=$v python tools/live_time.py f32 60 2>&1 | tail -1; done | tee gpurun_out/live_time.txt
python tools/live_time.py int8 60 2>&1 | tail -1 | tee -a gpurun_out/live_time.txt
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.draw --format=csv | tee gpurun_out/smi.txt
