#!/bin/bash
# ncu --set full of the final TMA GEMM kernels (complex128 4096^3 NN / CN, float64 8192^3 NN / TN)
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tma_kernel -s 6 -c 2 -f -o gpurun_out/prof_gemm_c_final python scratch/gemm_shapes.py > gpurun_out/ncu_gc.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tma_real -s 6 -c 2 -f -o gpurun_out/prof_gemm_r_final python scratch/gemm_shapes_real.py > gpurun_out/ncu_gr.log 2>&1
for f in prof_gemm_c_final prof_gemm_r_final; do ncu -i gpurun_out/$f.ncu-rep --page raw --csv > gpurun_out/$f.raw.csv 2>/dev/null; done
rm -f gpurun_out/*.ncu-rep; ls -la gpurun_out/prof_gemm_*final*; tail -2 gpurun_out/ncu_gc.log gpurun_out/ncu_gr.log
