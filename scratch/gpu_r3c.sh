#!/bin/bash
# Jacobi round variants: Gram in the 4-product form, own buffer for the cluster partials (one cluster barrier fewer)
mkdir -p gpurun_out
timeout 300 python scratch/jac_time.py 2>&1 | head -3
for v in gram4m gpart both; do
  TNB_LIB_PATH=scratch/exp/libtnb_$v.so timeout 300 python scratch/jac_time.py 2>&1 | head -3
  TNB_LIB_PATH=scratch/exp/libtnb_$v.so timeout 600 python -m pytest tests -m gpu -q -k "svd or fullsize or batched" 2>&1 | tail -2
done
