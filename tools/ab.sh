#!/bin/bash
# A/B timing of library variants on the GPU box: tools/ab.sh "workloads" variant[:opts] ...
wl=$1; shift
for v in "$@"; do
  name=${v%%:*}; opts=""; [[ "$v" == *:* ]] && opts=${v#*:}
  lib=uvhand_b200/lib/variants/$name.so; [ "$name" = default ] && lib=uvhand_b200/lib/libmsda_b200.so
  MSDA_B200_LIB=$lib python tools/quick_bench.py --iters 30 --workloads "$wl" --opts "$opts" ${CHECK:+--check} 2>&1 | grep -v "^lib" | sed "s/^/[$v] /"
done
