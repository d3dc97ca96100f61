#!/bin/bash
# round-2 final scaling lines on one 8-GPU box: strong scaling (65536 channels sharded) at N = 8, 4, 2 with the weak value alongside
cd /root/repo
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
$TR --nproc-per-node 8 --master-port 29521 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2f_n8.json 2> gpurun_out/r2f_n8.err
CUDA_VISIBLE_DEVICES=0,2,4,6 $TR --nproc-per-node 4 --master-port 29522 bench.py --gpus 4 --steps 10 --warmup 3 --no-cpu --no-side > gpurun_out/r2f_n4.json 2> gpurun_out/r2f_n4.err
CUDA_VISIBLE_DEVICES=0,2 $TR --nproc-per-node 2 --master-port 29523 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu --no-side > gpurun_out/r2f_n2.json 2> gpurun_out/r2f_n2.err
python - <<'PY'
import json
for f in ['r2f_n8','r2f_n4','r2f_n2']:
    try:
        d=json.loads([l for l in open('gpurun_out/%s.json'%f) if l.startswith('{"metric')][-1])
        e=d['e2e']
        print(f, 'value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'weak', d.get('weak') and round(d['weak']['value']), 'e2e', round(e['value']), 'i16', round(e['int16_iq']['value']), 'pageable', round(e['pageable']['value']))
    except Exception as ex:
        print(f, 'failed', ex)
PY
