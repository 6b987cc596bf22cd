#!/bin/bash
# Builds variants of libvors_b200.so that differ only in align_kernel.cu compile-time knobs, for A/B timing on the GPU box
# (scripts/ab_bench.sh).  usage: build_variants.sh name1:"-DVORS_X=.. -DVORS_Y=.." name2:"..." ...
set -e
cd "$(dirname "$0")/../visual-odometry-rs_b200"
make -s all
mkdir -p lib_variants build/variants
NVCC=/usr/local/cuda/bin/nvcc
ARCH="-gencode arch=compute_100a,code=sm_100a"
FLAGS="-O3 -std=c++17 -lineinfo $ARCH -Xcompiler -fPIC,-Wall,-Wno-unused-function -Xptxas -v --expt-relaxed-constexpr"
for spec in "$@"; do
  name="${spec%%:*}"; defs="${spec#*:}"
  $NVCC $FLAGS $defs -c csrc/align_kernel.cu -o build/variants/$name.o 2> build/variants/$name.log
  OTHERS="build/image_kernels.o build/dso_kernels.o build/engine.o"
  if [[ "$defs" == *VORS_STAGE_CHUNKS* || "$defs" == *VORS_TEX* ]]; then  # shared with the keyframe kernels and the engine
    OTHERS=""
    for f in image_kernels dso_kernels engine; do
      $NVCC $FLAGS $defs -c csrc/$f.cu -o build/variants/${name}_$f.o 2> /dev/null
      OTHERS="$OTHERS build/variants/${name}_$f.o"
    done
  fi
  $NVCC -shared $ARCH -o lib_variants/$name.so $OTHERS build/variants/$name.o
  echo "$name: $(grep -A2 'k_alignILb0ELb0ELb1' build/variants/$name.log | grep -E 'registers|spill' | tr '\n' ' ' | sed -E 's/ptxas info    ://g')"
done
