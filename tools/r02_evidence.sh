#!/bin/bash
# One GPU call that refreshes round 2's evidence into gpurun_out/: bench line, ncu launch list of the same command,
# DRAM bytes of every pruning launch of a sweep, and one ncu --set full capture (with source) of the largest launch
# (the root level's tile kernel).
tag=${1:-r02}
BARGS="--steps 3 --warmup 3 --no-cpu-baseline --no-mcmc --leaf50-families 0"
python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
tail -c 300 gpurun_out/bench_$tag.json
K='regex:ptile_kernel|gsweep_kernel|root_pattern_kernel|root_kernel|root_ragged_kernel|eps_kernel|tables_kernel|root_tables_kernel|sweep_kernel'
ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" --csv --log-file gpurun_out/launches_$tag.csv python bench.py $BARGS > gpurun_out/ncu1_$tag.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k "$K" --csv --log-file gpurun_out/traffic_$tag.csv python bench.py $BARGS > gpurun_out/ncu3_$tag.log 2>&1
# the root level's tile launch is the last ptile_kernel<false> launch of a sweep: 21 tile launches per sweep, sweeps 0..5 resident
ncu --set full --clock-control none --import-source on -k regex:ptile_kernel --launch-skip 83 --launch-count 1 -o gpurun_out/prof_$tag -f python bench.py $BARGS > gpurun_out/ncu2_$tag.log 2>&1
wc -l gpurun_out/launches_$tag.csv gpurun_out/traffic_$tag.csv; ls -la gpurun_out/prof_$tag.ncu-rep
