#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_bn_apply_gpu.py -x -q > gpurun_out/r02_bn_apply_tests.log 2>&1; echo "bn apply tests rc=$?"; tail -3 gpurun_out/r02_bn_apply_tests.log
python -m pytest tests/test_api_gpu.py tests/test_step_gpu.py -x -q > gpurun_out/r02_api_step_tests.log 2>&1; echo "api+step tests rc=$?"; tail -3 gpurun_out/r02_api_step_tests.log
python tools/bn_bench.py --out gpurun_out/r02_bn_bench.json > gpurun_out/r02_bn_bench.log 2>&1; tail -2 gpurun_out/r02_bn_bench.log
python bench.py --steps 100 --warmup 10 --mnbn --no-cpu-baseline > gpurun_out/r02_bench_n1_mnbn.json 2> gpurun_out/r02_bench_n1_mnbn.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02_bench_n1_mnbn.json').read().splitlines() if l.startswith('{')][-1])
print('ms/step', d['ms_per_step'], 'img/s', d.get('img_per_s'), 'train', {k:v for k,v in d['train'].items() if k!='model'})
print('mnbn', d['mnbn'])
PY
