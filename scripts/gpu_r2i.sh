#!/bin/bash
mkdir -p gpurun_out
python scripts/diag_dense.py 5 4 > gpurun_out/r2i_diag_tiled.txt 2>&1; cat gpurun_out/r2i_diag_tiled.txt
VORS_NO_TILED=1 python scripts/diag_dense.py 5 4 > gpurun_out/r2i_diag_generic.txt 2>&1; cat gpurun_out/r2i_diag_generic.txt
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:k_align -s 9 -c 1 -o gpurun_out/r2i_prof_tiled python bench.py --steps 2 --warmup 9 --no-cpu-baseline --no-parity > gpurun_out/r2i_ncu.log 2>&1; tail -2 gpurun_out/r2i_ncu.log
VORS_NO_TILED=1 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:k_align -s 9 -c 1 -o gpurun_out/r2i_prof_generic python bench.py --steps 2 --warmup 9 --no-cpu-baseline --no-parity > gpurun_out/r2i_ncu2.log 2>&1; tail -2 gpurun_out/r2i_ncu2.log
