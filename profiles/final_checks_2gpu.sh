#!/bin/bash
# 2-GPU checks of the final tree: DP parity tests, the 2-GPU bench line, and where NCCL's log lands
if [ "$1" != "benchonly" ]; then timeout 400 python -m pytest tests/test_gpu_dp.py -m gpu -q 2>&1 | tail -3 | cut -c1-200; fi
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29911 bench.py --gpus 2 --steps 40 --warmup 5 --no-secondary > gpurun_out/r2_scale_n2.json 2> gpurun_out/r2_scale_n2.err
python -c "import json; d=json.loads(open('gpurun_out/r2_scale_n2.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['n_gpus'])"
echo "stdout lines: $(wc -l < gpurun_out/r2_scale_n2.json)  NCCL INFO lines on stderr: $(grep -c 'NCCL INFO' gpurun_out/r2_scale_n2.err)"
grep -m2 "nranks" gpurun_out/r2_scale_n2.err | cut -c1-170
