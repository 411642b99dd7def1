# Round-end measurement set on ONE B200 (run under gpurun from the repo root); outputs land in gpurun_out/ and are
# summarised into profiles/ with tools/ncu_summary.py / tools/ncu_phases.py afterwards.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -2
timeout 600 python bench.py --steps 10 --warmup 3 --live 480 --features > gpurun_out/bench_n1_f32.json 2> gpurun_out/bench_n1_f32.err
timeout 600 python bench.py --steps 10 --warmup 3 --model int8 --live 480 > gpurun_out/bench_n1_int8.json 2> gpurun_out/bench_n1_int8.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
# launch list of the bench command: ONLY the library's kernels (without the name filter the first 400 launches are the torch
# kernels that synthesise the input audio and the list never reaches the timed region -- what happened to the r01 file)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:k1_|k2_|nn_|carry_|fill_state' -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/launches_bench.log 2>&1
B="python bench.py --streams 8192 --steps 1 --warmup 1 --no-e2e --no-cpu"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k1_spectral_kernel -s 1 -c 1 -o gpurun_out/k1 -f $B > gpurun_out/ncu_k1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:nn_f32_clip_kernel -s 1 -c 1 -o gpurun_out/nn_f32 -f $B > gpurun_out/ncu_nn.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:nn_f32_live_kernel -s 6 -c 1 -o gpurun_out/nn_live -f python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --live 480 > gpurun_out/ncu_live.log 2>&1
ls -la gpurun_out/ | tail -15
