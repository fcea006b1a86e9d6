#!/usr/bin/env bash
# builds and runs the tcgen05 probe on the GPU box:  gpurun -- tools/probes/run_mma_probe.sh
set -e
cd "$(dirname "$0")"
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o /tmp/mma_probe mma_probe.cu ../../multimodal_image_transformer_b200/libb200decoder.so -Xlinker -rpath -Xlinker "$(cd ../../multimodal_image_transformer_b200 && pwd)" 2>&1 | tail -3
/tmp/mma_probe
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o /tmp/ex2_probe ex2_probe.cu 2>&1 | tail -3
/tmp/ex2_probe
