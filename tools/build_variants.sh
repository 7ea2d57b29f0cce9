#!/bin/bash
# Build experimental variants of libmsda_b200.so (same ABI) into uvhand_b200/lib/variants/ for A/B timing:
#   tools/build_variants.sh name "-DMSDA_FWD_MINB=5 -DMSDA_BWD_WAVES=4" [name2 "flags2" ...]
# then  MSDA_B200_LIB=uvhand_b200/lib/variants/name.so python tools/quick_bench.py
set -e
cd "$(dirname "$0")/.."
mkdir -p uvhand_b200/lib/variants
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -shared -Xcompiler -fPIC $flags \
    -o uvhand_b200/lib/variants/$name.so uvhand_b200/csrc/msda_b200.cu &
done
wait
ls -la uvhand_b200/lib/variants
