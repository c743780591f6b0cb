#!/bin/bash
# Builds variant libraries into variants/ (git-ignored, travels with gpurun) for A/B runs with NTR_B200_LIB.
# usage: tools/build_variants.sh name1="-DFLAG=.." name2="-DFLAG=.."    e.g.  sm64="-DNTR_SINGLE_MAILBOX=64"
set -e
cd "$(dirname "$0")/.."
mkdir -p variants
for spec in "$@"; do
  name="${spec%%=*}"; flags="${spec#*=}"
  make -C ntracer_b200/csrc -j"$(nproc)" OBJDIR=/tmp/ntr_obj_$name TARGET="$PWD/variants/libntr_$name.so" EXTRA="$flags" >/dev/null
  echo "variants/libntr_$name.so  ($flags)"
done
