set -x
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o /tmp/tcgen05_probe tools/probes/tcgen05_pointwise_probe.cu 2>&1 | tail -3
timeout 120 /tmp/tcgen05_probe 2>&1 | tee gpurun_out/tcgen05_probe.txt
