# 4/8-GPU check of the driver's launch line (run under gpurun --gpus 8)
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
$T --nproc-per-node 8 --master-port 29511 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/s8_c4.json 2> gpurun_out/s8_c4.err
$T --nproc-per-node 4 --master-port 29512 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/s4_c4.json 2> gpurun_out/s4_c4.err
$T --nproc-per-node 8 --master-port 29513 bench.py --gpus 8 --workload c5 --steps 3 --warmup 3 > gpurun_out/s8_c5.json 2> gpurun_out/s8_c5.err
for f in s8_c4 s4_c4 s8_c5; do tail -c 300 gpurun_out/$f.err; python -c "
import json
d=json.loads(open('gpurun_out/$f.json').read().strip().splitlines()[-1]); print('$f', d['n_gpus'], d['value'], d['e2e']['value'], d['ms_per_step'], d['config'])"; done
