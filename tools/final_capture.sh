# Round-end capture: GPU tests, smoke, the three bench lines, launch lists and ncu --set full of the top kernels.
set -x
timeout -s KILL 900 python -m pytest tests -q -m gpu 2>&1 | tail -3 > gpurun_out/pytest_gpu.txt
python __graft_entry__.py --smoke 2>&1 | tail -1 > gpurun_out/smoke.txt
python bench.py > gpurun_out/c4.json 2> gpurun_out/c4.err
python bench.py --workload c3 > gpurun_out/c3.json 2> gpurun_out/c3.err
python bench.py --workload c5 > gpurun_out/c5.json 2> gpurun_out/c5.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/ref.json 2> gpurun_out/ref.err
L="--metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv"
ncu $L --log-file gpurun_out/launches_c4.csv python tools/profile_step.py --workload c4 --global-batch 512 > /dev/null 2>&1
ncu $L --log-file gpurun_out/launches_c5.csv python tools/profile_train.py --batch 128 > /dev/null 2>&1
F="--set full --clock-control none --import-source on --profile-from-start off"
ncu $F -k regex:cell_fwd -s 10 -c 1 -o gpurun_out/cell_final -f python tools/profile_step.py --workload c4 --global-batch 512 > /dev/null 2>&1
ncu $F -k "regex:gnn_kernel|head_kernel|beam_step" -s 6 -c 4 -o gpurun_out/aux_final -f python tools/profile_step.py --workload c4 --global-batch 512 > /dev/null 2>&1
python tools/gpu_time_aux.py 2>&1 | tail -1 > gpurun_out/aux_time.txt
cat gpurun_out/pytest_gpu.txt gpurun_out/smoke.txt gpurun_out/aux_time.txt
