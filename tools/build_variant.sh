#!/bin/bash
# Build an experimental variant of libwfb200.so: tools/build_variant.sh <name> [-DWFB_LPR_...=...]
# (objects of the other sources are reused from csrc/_obj; run `python -m waveformanalysis_b200.build` first)
set -e
name=$1; shift
cd "$(dirname "$0")/../waveformanalysis_b200"
mkdir -p ../gpurun_out/variants
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr "$@" -c csrc/fused_lpr.cu -o /tmp/fused_lpr_$name.o
objs=$(ls csrc/_obj/*.o | grep -v fused_lpr.o)
nvcc -shared -o variants_$name.so $objs /tmp/fused_lpr_$name.o -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC
echo waveformanalysis_b200/variants_$name.so
