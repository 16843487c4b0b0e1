#!/bin/bash
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 | cut -c1-300
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err
python -c "import json; d=json.loads(open('gpurun_out/r2_bench_final.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['clocks'], d['cpu_baseline'])"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_ref_final.json 2>/dev/null; tail -c 700 gpurun_out/r2_bench_ref_final.json
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2_launches_final.csv python profiles/prof_step.py 2 > gpurun_out/prof_final.log 2>&1; tail -1 gpurun_out/prof_final.log
