#!/bin/bash
mkdir -p gpurun_out
python scripts/prof_fit.py --n 4096 > gpurun_out/prof_fit4096_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_fit4096.csv python scripts/prof_fit.py --n 4096 --reps 1 > gpurun_out/ncu_fit.log 2>&1
echo "fit ncu rc=$?"
python scripts/prof_score.py > gpurun_out/prof_score_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:score_sets_k8 -c 2 -o gpurun_out/prof_score python scripts/prof_score.py > gpurun_out/ncu_score.log 2>&1
echo "score ncu rc=$?"
python scripts/prof_fit.py --n 16384 --side 128 --reps 2 > gpurun_out/prof_fit16384_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:kbuild_kernel -c 2 -o gpurun_out/prof_kbuild python scripts/prof_fit.py --n 16384 --side 128 --reps 1 > gpurun_out/ncu_kbuild.log 2>&1
echo "kbuild ncu rc=$?"
timeout 600 python -m pytest tests/test_gpu_api.py -m gpu -q --tb=short -p no:cacheprovider --timeout 300 > gpurun_out/pytest_api.log 2>&1
echo "api rc=$?"
cat gpurun_out/prof_fit4096_plain.log gpurun_out/prof_score_plain.log gpurun_out/prof_fit16384_plain.log; tail -3 gpurun_out/pytest_api.log
