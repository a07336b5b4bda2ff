#!/bin/bash
# Builds libpasio_b200.so variants for A/B runs on one box: tools/build_variants.sh NAME "EXTRA FLAGS" [NAME "FLAGS" ...]
# -> build_variants/libNAME.so (git-ignored; travels with gpurun); select with PASIO_B200_LIB=build_variants/libNAME.so
set -e
cd "$(dirname "$0")/.."
mkdir -p build_variants
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  make -C pasio_b200/csrc clean > /dev/null
  make -C pasio_b200/csrc -j8 EXTRA="$flags" > build_variants/$name.build.log 2>&1
  cp pasio_b200/libpasio_b200.so build_variants/lib$name.so
  echo "$name: $flags"; grep -A1 "exact_pruned_kernelILb1ELb1E" build_variants/$name.build.log | grep -o "[0-9]* bytes spill stores" | head -1
done
make -C pasio_b200/csrc clean > /dev/null
make -C pasio_b200/csrc -j8 > /dev/null 2>&1
