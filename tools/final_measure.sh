set -x
python -m pytest tests -m gpu -q 2>&1 | tail -2
timeout 600 python bench.py --steps 10 --warmup 3 --live 480 --features > gpurun_out/bench_n1_f32.json 2> gpurun_out/bench_n1_f32.err
timeout 600 python bench.py --steps 10 --warmup 3 --model int8 > gpurun_out/bench_n1_int8.json 2> gpurun_out/bench_n1_int8.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/launches_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:nn_f32_live_kernel -s 6 -c 1 -o gpurun_out/nn_live -f python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --live 480 > gpurun_out/ncu_live.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k1_spectral_packed_kernel -s 6 -c 1 -o gpurun_out/k1_packed -f python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --live 480 > gpurun_out/ncu_k1p.log 2>&1
ls -la gpurun_out/ | tail -15
