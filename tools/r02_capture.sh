#!/bin/bash
# Final round-2 evidence in one GPU call (run under gpurun from the repo root); everything lands in gpurun_out/.
set -x
O=gpurun_out
python bench.py --steps 20 --warmup 5 > $O/r02_bench_final.json 2> $O/r02_bench_final.err
python bench.py --impl reference --steps 5 --warmup 1 > $O/r02_bench_final_reference_arm.json 2>> $O/r02_bench_final.err
python bench.py --steps 20 --warmup 5 --extended --no-cpu-baseline > $O/r02_bench_extended.json 2>> $O/r02_bench_final.err
# launch list of the bench command (shares of the step), after the same command exited 0 without ncu
CMD="python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --main-only"
$CMD > $O/r02_launch_plain.json 2> $O/r02_launch_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_bench_launches.csv $CMD > $O/r02_launch_ncu.log 2>&1
# full captures of the shipped kernels at the DETR-encoder shape (fp32 and bf16 value)
P="python tools/prof_run.py --workload c2_detr_encoder --steps 2"
$P > $O/r02_prof_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:d32 -s 2 -c 2 -o $O/r02_c2_f32 $P > $O/r02_prof_ncu.log 2>&1
$P --kind bf16 > $O/r02_prof_plain_bf16.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:d32 -s 2 -c 2 -o $O/r02_c2_bf16 $P --kind bf16 > $O/r02_prof_ncu_bf16.log 2>&1
# the slab forward at the ARCTIC encoder shape (BASELINE configs[2])
P3="python tools/prof_run.py --workload c3_arctic_encoder --steps 2"
$P3 > $O/r02_prof_plain_c3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:slab -s 1 -c 1 -o $O/r02_c3_f32 $P3 > $O/r02_prof_ncu_c3.log 2>&1
tail -n 2 $O/r02_prof_ncu_c3.log
tail -n 2 $O/r02_prof_ncu.log $O/r02_prof_ncu_bf16.log $O/r02_launch_ncu.log
