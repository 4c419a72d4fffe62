#!/bin/bash
# A/B of prebuilt libraries on one box: tools/ab_bench.sh name1 name2 ...   (expects _ab/<name>.so)
for rep in 1 2; do
  for v in "$@"; do
    cp _ab/$v.so beluga.jl_b200/libbeluga_b200.so
    touch beluga.jl_b200/libbeluga_b200.so
    echo -n "$v: "
    python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-mcmc 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print(round(d['ms_per_step'],3), 'ms', d['logpdf'])"
  done
done
