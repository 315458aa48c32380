#!/bin/bash
# Dev tool: build diagnostic variants of the library (timing only -- their results are WRONG on purpose)
# into tools/diag_libs/ (git-ignored). Usage: tools/diag_build.sh NAME "-DFLAG1 -DFLAG2" [NAME2 "..."] ...
set -e
cd "$(dirname "$0")/.."
mkdir -p tools/diag_libs
FLAGS=$(python -c "import vlp_b200; from vlp_b200 import _build; print(' '.join(_build.NVCC_FLAGS)); print(' '.join(_build.sources()))")
NV=$(echo "$FLAGS" | sed -n 1p); SRC=$(echo "$FLAGS" | sed -n 2p)
while [ $# -ge 2 ]; do
  ( nvcc $NV -DVLP_PROFILE_WAITS $2 -o tools/diag_libs/lib_$1_prof.so $SRC && echo "built $1" ) &
  shift 2
done
wait
