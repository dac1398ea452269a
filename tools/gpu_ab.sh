#!/bin/bash
# A/B on one box: bench stage figures with the in-tree library and with tools/variants/libcurie_<name>.so
for i in 1 2; do
  echo "== new"; bash tools/gpu_exp_fit.sh | sed 's/, .fit_nfev[^}]*}//'
  echo "== $1"; CB_LIB_PATH=$PWD/tools/variants/libcurie_$1.so bash tools/gpu_exp_fit.sh | sed 's/, .fit_nfev[^}]*}//'
done
