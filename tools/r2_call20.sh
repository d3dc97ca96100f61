#!/bin/bash
cd /root/repo
nvidia-smi topo -m > gpurun_out/r2_topo.txt 2>&1
numactl -H > gpurun_out/r2_numa.txt 2>&1 || lscpu | grep -i numa > gpurun_out/r2_numa.txt
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
$TR --nproc-per-node 8 --master-port 29511 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2_n8_full.json 2> gpurun_out/r2_n8_full.err
LQB_BENCH_NO_PIN=1 $TR --nproc-per-node 8 --master-port 29512 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu --no-side > gpurun_out/r2_n8_nopin.json 2> gpurun_out/r2_n8_nopin.err
CUDA_VISIBLE_DEVICES=0,1,2,3 $TR --nproc-per-node 4 --master-port 29513 bench.py --gpus 4 --steps 5 --warmup 3 --no-cpu --no-side > gpurun_out/r2_n4_0123.json 2> gpurun_out/r2_n4_0123.err
CUDA_VISIBLE_DEVICES=0,2,4,6 $TR --nproc-per-node 4 --master-port 29514 bench.py --gpus 4 --steps 5 --warmup 3 --no-cpu --no-side > gpurun_out/r2_n4_0246.json 2> gpurun_out/r2_n4_0246.err
python - <<'PY'
import json
for f in ['r2_n8_full','r2_n8_nopin','r2_n4_0123','r2_n4_0246']:
    try:
        d=json.loads([l for l in open('gpurun_out/%s.json'%f) if l.startswith('{"metric')][-1])
        e=d['e2e']
        print(f, 'value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'weak', d.get('weak') and round(d['weak']['value']), 'e2e', round(e['value']), 'i16', round(e['int16_iq']['value']), 'pageable', round(e['pageable']['value']))
        print('   per rank', [(r['pinned_h2d_gbs'], r['first_cpu']) for r in e['per_rank']], e['host_binding'])
    except Exception as ex:
        print(f, 'failed', ex)
PY
