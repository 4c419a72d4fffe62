#!/bin/bash
# code size per function of the built library (device helpers are listed inside their kernel's section)
cuobjdump -sass "${1:-beluga.jl_b200/libbeluga_b200.so}" 2>/dev/null | awk '/Function :/ {name=$3} /^ +\/\*[0-9a-f]+\*\/ / {cnt[name]++} END {for (n in cnt) printf "%7d B  %s\n", cnt[n]*16, n}' | sed 's/_ZN4[57]_[A-Za-z_0-9]*bd68bf2f//g' | cut -c1-120 | sort -k3
