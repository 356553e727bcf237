#!/bin/bash
# builds tools/diag_probe against the in-tree library (run from the repository root, after `make -C bot7_b200/csrc`)
set -e
cd "$(dirname "$0")/.."
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DB7_DIAG_STAMPS -Ibot7_b200/csrc -Iinclude tools/diag_probe.cu -o tools/diag_probe \
  -Lbot7_b200 -lbot7_b200 -Xlinker -rpath -Xlinker '$ORIGIN/../bot7_b200'
