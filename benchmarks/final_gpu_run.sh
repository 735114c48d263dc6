#!/bin/bash
# One-box validation used at the end of round 1: full GPU test suite, the bench line in both sweep modes, the secondary sweeps,
# the ncu captures kept under profiles/ (full set of the field-mode sweep kernel, launch list of bench.py).
# Run under gpurun from the repo root:  gpurun --timeout 1500 -- 'bash benchmarks/final_gpu_run.sh'
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -5) > gpurun_out/final_tests.log 2>&1
(timeout 300 python bench.py 2>&1 | tail -1) > gpurun_out/final_bench.json 2>gpurun_out/final_bench.err
(timeout 200 python bench.py --no-cpu-baseline --sweep-mode classic --steps 30 2>&1 | tail -1) > gpurun_out/final_bench_classic.json 2>/dev/null
(timeout 200 python benchmarks/annealer_sweep.py --sizes 128,256,512,1024,2048,4096 2>&1 | tail -8) > gpurun_out/sweep_auto.log 2>&1
(timeout 200 python benchmarks/replicas.py --replicas-per-gpu 512 2>&1 | tail -3) > gpurun_out/replicas_auto.log 2>&1
if [ "$1" != "--no-ncu" ]; then
timeout 300 ncu --set full --clock-control none --import-source on -k regex:denseSweep -s 8 -c 1 -o gpurun_out/r1_field_sweep_final -f \
    python bench.py --steps 4 --warmup 10 --no-cpu-baseline > gpurun_out/ncu_final.log 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r1_field_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1
fi
cat gpurun_out/final_tests.log
head -c 2500 gpurun_out/final_bench.json; echo
head -c 400 gpurun_out/final_bench_classic.json; echo
cat gpurun_out/sweep_auto.log gpurun_out/replicas_auto.log
