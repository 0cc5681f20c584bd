#!/bin/bash
# where the masked INT8 digit GEMM spends its cycles: library variants built with -DI8_PROFILE sum clock64()
# intervals per phase over all CTAs (build: ALGP_NVCC_EXTRA="-DI8_PROFILE [-DI8_EXP_NOLOAD -DI8_EXP_NOMMA]" python
# algp_b200/build.py --force, copied to algp_b200/lib/variants/)
cp algp_b200/lib/libalgp_b200.so /tmp/lib_default.so
for v in algp_b200/lib/variants/lib_PROFILE*.so; do
  cp $v algp_b200/lib/libalgp_b200.so
  echo "== $v"
  timeout 250 python scripts/prof_i8_phases.py 2>&1 | tee -a gpurun_out/i8_phases.log
done
cp /tmp/lib_default.so algp_b200/lib/libalgp_b200.so
