#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_bn_apply_gpu.py tests/test_master_gpu.py tests/test_hooks_gpu.py -q > gpurun_out/r02_n1c_tests.log 2>&1; echo "bn/master/hooks tests rc=$?"; tail -4 gpurun_out/r02_n1c_tests.log
python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/r02_bench_n1_train.json 2> gpurun_out/r02_bench_n1_train.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02_bench_n1_train.json').read().splitlines() if l.startswith('{')][-1])
print('ms/step', d['ms_per_step'], 'img/s', d.get('img_per_s'))
print({k:v for k,v in d['train'].items() if k!='model'})
PY
