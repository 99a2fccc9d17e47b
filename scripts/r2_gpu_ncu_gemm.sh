#!/bin/bash
# ncu --set full of the 2-CTA GEMM (bench shapes) and of the tcgen05 convolution launches of a VQ decode
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none -k regex:gemm_tc5x2_kernel -s 25 -c 3 -o gpurun_out/prof_gemm_x2 python scripts/bench_gemm.py > gpurun_out/ncu_gemm_x2.log 2>&1
B=4 timeout 600 ncu --set full --clock-control none -k regex:gemm_tc5_kernel -s 330 -c 8 -o gpurun_out/prof_conv_tc5 python scripts/vision_once.py > gpurun_out/ncu_conv_tc5.log 2>&1
ls -la gpurun_out/prof_gemm_x2.ncu-rep gpurun_out/prof_conv_tc5.ncu-rep
