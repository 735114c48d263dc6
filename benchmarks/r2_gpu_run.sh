#!/bin/bash
# One-box validation and evidence run of round 2 (one B200):  gpurun --timeout 2400 -- 'bash benchmarks/r2_gpu_run.sh'
# Full GPU test suite, the bench line, the harness in the reference's protocol, and the ncu captures summarised under profiles/.
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6) > gpurun_out/r2_tests.log 2>&1
(timeout 400 python bench.py 2>gpurun_out/r2_bench.err | tail -1) > gpurun_out/r2_bench_n1.json
(timeout 300 python benchmarks/annealer_sweep.py --duration 1.0 --csv gpurun_out/r2_harness 2>&1 | tail -20) > gpurun_out/r2_harness.log 2>&1
(timeout 300 python benchmarks/perf_shapes.py --cpu 2>&1 | tail -10) > gpurun_out/r2_perf_shapes.log 2>&1
(timeout 100 python benchmarks/chain_profile.py --steps 30 2>&1 | tail -40) > gpurun_out/r2_chain_profile.log 2>&1
(timeout 200 python benchmarks/bench_paths.py --what bipartite 2>&1 | tail -4) > gpurun_out/r2_bipartite.log 2>&1
(timeout 200 python benchmarks/bench_paths.py --what energy 2>&1 | tail -4) > gpurun_out/r2_energy.log 2>&1
if [ "$1" != "--no-ncu" ]; then
timeout 300 ncu --set full --clock-control none --import-source on -k regex:denseSweep -s 10 -c 1 -o gpurun_out/r2_field_sweep -f \
    python benchmarks/chain_profile.py --steps 3 --warmup 12 > gpurun_out/ncu_field.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:sweepTables -s 10 -c 1 -o gpurun_out/r2_tables -f \
    python benchmarks/chain_profile.py --steps 3 --warmup 12 > gpurun_out/ncu_tables.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:denseSweep -s 4 -c 1 -o gpurun_out/r2_classic_sweep -f \
    python benchmarks/chain_profile.py --mode classic --steps 2 --warmup 5 > gpurun_out/ncu_classic.log 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --quick --steps 10 --warmup 3 > gpurun_out/ncu_list.log 2>&1
timeout 200 ncu --set full --clock-control none -k regex:bgFlipFused -s 6 -c 1 -o gpurun_out/r2_bg_flip -f \
    python benchmarks/bench_paths.py --what bipartite > gpurun_out/ncu_bgflip.log 2>&1
timeout 200 ncu --set full --clock-control none -k regex:tcSpinGemm -s 6 -c 1 -o gpurun_out/r2_bg_gemm -f \
    python benchmarks/bench_paths.py --what bipartite > gpurun_out/ncu_bggemm.log 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_bipartite_launches.csv \
    python benchmarks/bench_paths.py --what bipartite > gpurun_out/ncu_bglist.log 2>&1
fi
cat gpurun_out/r2_tests.log
head -c 1500 gpurun_out/r2_bench_n1.json; echo
tail -3 gpurun_out/r2_bench.err
cat gpurun_out/r2_harness.log gpurun_out/r2_perf_shapes.log gpurun_out/r2_bipartite.log gpurun_out/r2_energy.log
