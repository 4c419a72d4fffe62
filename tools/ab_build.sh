#!/bin/bash
# tools/ab_build.sh <name> [git-rev]  : build the working tree's csrc (or the csrc of <git-rev>) into _ab/<name>.so
set -e
name=$1; rev=$2
src=beluga.jl_b200/csrc
if [ -n "$rev" ]; then
  rm -rf _ab/src/$name && mkdir -p _ab/src/$name
  for f in $(git ls-tree --name-only $rev $src/); do git show $rev:$f > _ab/src/$name/$(basename $f); done
  src=_ab/src/$name
fi
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC -ccbin /usr/bin/g++ -I include -o _ab/$name.so $src/beluga_b200.cu
ls -la _ab/$name.so
