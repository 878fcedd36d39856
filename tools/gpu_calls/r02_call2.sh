#!/bin/bash
# round-2 GPU call 2: f16f8 cell unit tests, rollout goldens, A/B bench of the class decoder format
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity_gpu.py -q -x -k "cell_f16f8 or cell_golden or rollout_golden or onehot or fanout" -s 2>&1 | tail -25 > gpurun_out/c2_tests.log
cat gpurun_out/c2_tests.log
MVB_F16F8=0 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/c2_bench_bf16.json 2> gpurun_out/c2_bench_bf16.err
MVB_F16F8=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/c2_bench_f16f8.json 2> gpurun_out/c2_bench_f16f8.err
tail -c 1500 gpurun_out/c2_bench_bf16.json; echo; tail -c 1500 gpurun_out/c2_bench_f16f8.json
