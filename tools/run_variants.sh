#!/bin/bash
# usage: tools/run_variants.sh <n> <only> <variant .so files...>: runs the K2T split microbenchmark with each library
# variant (only = all | quads | pairs+diag | "" for the three of them)
n=$1; only=$2; shift; shift
cp qibochem_b200/libqibochem_b200.so /tmp/orig.so
for v in "$@"; do
  cp "$v" qibochem_b200/libqibochem_b200.so
  echo "== $v"
  timeout 300 python tools/microbench_k2_split.py "$n" 100000 $only 2>&1 | tail -3
done
cp /tmp/orig.so qibochem_b200/libqibochem_b200.so
