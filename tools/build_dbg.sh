#!/bin/bash
# Debug build of the CUDA library with per-phase cycle counters in the lean kernel (never shipped / never loaded by the package):
#   tools/build_dbg.sh && GSM_LIB_PATH=$PWD/build/libgsm_b200_dbg.so python tools/tune.py C4 20000 2000 256x2
set -e
cd "$(dirname "$0")/.."
mkdir -p build
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -ccbin /usr/bin/g++ -DGSM_PHASE_TIMING $GSM_DBG_FLAGS \
  -shared -o ${GSM_DBG_OUT:-build/libgsm_b200_dbg.so} gammaspikemodel_b200/csrc/gsm_kernels.cu gammaspikemodel_b200/csrc/gsm_capi.cu gammaspikemodel_b200/csrc/gsm_treelog.cpp
