#!/bin/bash
for rs in 0 8 16; do
  FG_DP_RESERVE_SMS=$rs timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29911 bench.py --gpus 2 --steps 40 --warmup 5 --no-secondary 2>gpurun_out/dp2_r$rs.err | tail -1 > gpurun_out/dp2_r$rs.json
  python -c "import json; d=json.load(open('gpurun_out/dp2_r$rs.json')); print('reserve', $rs, d['value'], d['ms_per_step'])"
done
FG_DP_OVERLAP=0 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29911 bench.py --gpus 2 --steps 40 --warmup 5 --no-secondary 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('serial', d['value'], d['ms_per_step'])"
