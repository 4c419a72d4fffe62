#!/bin/bash
# One GPU call for the gradient's evidence, into gpurun_out/: timings of both sweeps (node patterns / families), the
# per-launch device times of the node-pattern sweep (events between the launches, no profiler), the ncu launch list of one
# gradient, and one ncu --set full capture (with source) of its largest launch (phase A at the root).
tag=${1:-r02}
python tools/grad_time.py 2> gpurun_out/gradient_time_$tag.txt
python tools/grad_debug.py 2>&1 | grep "padjoint\|gradient:" > gpurun_out/gradient_launches_$tag.txt
GRAD_SPECS="BLG_GRAD_PAT=1" ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:padj_\|adjoint_params\|tables_dual\|grad_reduce --csv \
    --log-file gpurun_out/gradient_ncu_launches_$tag.csv python tools/grad_time.py > gpurun_out/gradient_ncu1_$tag.log 2>&1
GRAD_SPECS="BLG_GRAD_PAT=1" ncu --set full --clock-control none --import-source on -k regex:padj_parent -c 1 -o gpurun_out/gradient_prof_$tag -f \
    python tools/grad_time.py > gpurun_out/gradient_ncu2_$tag.log 2>&1
cat gpurun_out/gradient_time_$tag.txt; wc -l gpurun_out/gradient_ncu_launches_$tag.csv; ls -la gpurun_out/gradient_prof_$tag.ncu-rep
