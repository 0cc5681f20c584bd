#!/bin/bash
cp algp_b200/lib/libalgp_b200.so /tmp/lib_default.so
echo "== default (4 stages)"; timeout 200 python scripts/prof_i8.py 16384 65536 7 nocheck morton 2>&1 | grep "i8 S"
for v in algp_b200/lib/variants/*.so; do
  cp $v algp_b200/lib/libalgp_b200.so; touch algp_b200/lib/libalgp_b200.so
  echo "== $v"; timeout 200 python scripts/prof_i8.py 16384 65536 7 nocheck morton 2>&1 | grep "i8 S"
done
cp /tmp/lib_default.so algp_b200/lib/libalgp_b200.so
