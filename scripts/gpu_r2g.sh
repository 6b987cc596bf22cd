#!/bin/bash
mkdir -p gpurun_out
python scripts/diag_dense.py 100000 > gpurun_out/r2g_diag_tiled.txt 2>&1; cat gpurun_out/r2g_diag_tiled.txt
VORS_NO_TILED=1 python scripts/diag_dense.py 100000 > gpurun_out/r2g_diag_generic.txt 2>&1; cat gpurun_out/r2g_diag_generic.txt
