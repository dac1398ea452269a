#!/bin/bash
# A/B helper: link tools/variants/libcurie_<name>.so from the current objects, with spectrum.cu taken from git
# revision <rev> (its kernels only; the C ABI must be unchanged).  Select it with CB_LIB_PATH at run time.
#   bash tools/build_variant.sh <rev> <name>
set -e
REV=$1; NAME=$2
ROOT=$(cd "$(dirname "$0")/.." && pwd)
TMP=$(mktemp -d)
git -C "$ROOT" archive "$REV" pycurious_b200/csrc include | tar -x -C "$TMP"
(cd "$TMP/pycurious_b200/csrc" && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC \
    -I../../include -I. -c spectrum.cu -o "$TMP/spectrum_variant.o")
mkdir -p "$ROOT/tools/variants"
OBJS=$(ls "$ROOT"/pycurious_b200/csrc/build/*.o | grep -v '/spectrum.o')
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o "$ROOT/tools/variants/libcurie_$NAME.so" $OBJS "$TMP/spectrum_variant.o" -lcudart
rm -rf "$TMP"
ls -la "$ROOT/tools/variants/libcurie_$NAME.so"
