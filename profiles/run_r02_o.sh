#!/bin/bash
# r02 call O: ncu --set full of kmer_window_kernel and finish_rows_kernel (296 windows x 50 kbp)
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"kmer_window_kernel" -c 1 -o gpurun_out/o_density_full -f python profiles/run_density_c5.py 296 1 > gpurun_out/o_ncu.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/o_density_full.ncu-rep
