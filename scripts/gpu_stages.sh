#!/bin/bash
# pipeline-depth experiment for the INT8 digit GEMM: default ring depth, then forced 2 and 3 stages
echo "== default"; timeout 200 python scripts/prof_i8.py 16384 65536 7 nocheck morton 2>&1 | grep "i8 S"
for st in 2 3; do
  ALGP_NVCC_EXTRA="-DI8_STAGES_OVERRIDE=$st" python algp_b200/build.py --force > /dev/null
  echo "== stages $st"; timeout 200 python scripts/prof_i8.py 16384 65536 7 nocheck morton 2>&1 | grep "i8 S"
done
python algp_b200/build.py --force > /dev/null
