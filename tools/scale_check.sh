# 2/4/8-GPU check of the driver's launch line (run under gpurun --gpus 8)
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/s1_c4.json 2> gpurun_out/s1_c4.err
for n in 2 4 8; do
$T --nproc-per-node $n --master-port 2951$n bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/s${n}_c4.json 2> gpurun_out/s${n}_c4.err
done
$T --nproc-per-node 8 --master-port 29519 bench.py --gpus 8 --workload c5 --steps 3 --warmup 3 > gpurun_out/s8_c5.json 2> gpurun_out/s8_c5.err
for f in s1_c4 s2_c4 s4_c4 s8_c4 s8_c5; do tail -c 200 gpurun_out/$f.err | grep -i "error\|Traceback"; python -c "
import json
d=json.loads(open('gpurun_out/$f.json').read().strip().splitlines()[-1]); print('$f', d['n_gpus'], round(d['value'],1), round(d['e2e']['value'],1), round(d['ms_per_step'],1), d['clocks']['sm_mhz'])"; done
