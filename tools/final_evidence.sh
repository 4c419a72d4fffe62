#!/bin/bash
# One GPU call that refreshes the round's evidence: picks the sweep block shape (BLG_WPB) by A/B, then the full
# bench line, the ncu launch list and one ncu --set full capture of sweep_kernel, all into gpurun_out/.
tag=${1:-r01b}
ms() { python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-mcmc 2>/dev/null | tail -1 | python -c "import sys,json;print(json.loads(sys.stdin.readline())['ms_per_step'])"; }
if [ -z "$SKIP_WPB" ]; then
  A=$(BLG_WPB=4 ms); B=$(BLG_WPB=2 ms)
  best=4; python -c "import sys; sys.exit(0 if float('$B') < 0.985*float('$A') else 1)" && best=2
  echo "wpb4 $A ms  wpb2 $B ms  -> BLG_WPB=$best" | tee gpurun_out/wpb_$tag.log
  export BLG_WPB=$best
fi
python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
tail -c 400 gpurun_out/bench_$tag.json
K='regex:sweep_kernel|root_kernel|eps_kernel|tables_kernel|root_tables_kernel|fp64_peak_kernel|pascal_kernel|prune_node_kernel'
ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" --csv --log-file gpurun_out/launches_$tag.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-mcmc > gpurun_out/ncu1_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s 3 -c 1 -o gpurun_out/prof_$tag -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-mcmc > gpurun_out/ncu2_$tag.log 2>&1
wc -l gpurun_out/launches_$tag.csv; ls -la gpurun_out/prof_$tag.ncu-rep
