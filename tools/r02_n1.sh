#!/bin/bash
# round-2 N=1 validation: new step tests, API tests, bench with and without the one-launch step
mkdir -p gpurun_out
python -m pytest tests/test_step_gpu.py -x -q > gpurun_out/r02_step_tests.log 2>&1; echo "step tests rc=$?"
tail -3 gpurun_out/r02_step_tests.log
python -m pytest tests/test_api_gpu.py -x -q > gpurun_out/r02_api_tests.log 2>&1; echo "api tests rc=$?"
tail -3 gpurun_out/r02_api_tests.log
python bench.py --steps 200 --warmup 20 > gpurun_out/r02_bench_n1_step.json 2> gpurun_out/r02_bench_n1_step.err; echo "bench rc=$?"
python bench.py --steps 200 --warmup 20 --no-step --no-train --no-cpu-baseline > gpurun_out/r02_bench_n1_nostep.json 2> gpurun_out/r02_bench_n1_nostep.err
python bench.py --steps 100 --warmup 20 --workload seq2seq --no-train --no-cpu-baseline > gpurun_out/r02_bench_n1_seq2seq.json 2> gpurun_out/r02_bench_n1_seq2seq.err
python bench.py --steps 200 --warmup 20 --allreduce-dtype float16 --no-train --no-cpu-baseline > gpurun_out/r02_bench_n1_f16.json 2>/dev/null
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02_bench_n1*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        r=d['roofline']
        print(f.split('/')[-1], 'ms/step %.4f'%d['ms_per_step'], 'value %.0f'%d['value'], r['kernel'], 'us %.1f'%r['us_per_launch'], 'frac %.3f'%r['frac'], 'stepfrac8 %.3f'%r['step_frac_of_nominal_8TBs'], 'parity', (d.get('parity') or {}).get('ok'), 'host_enq %.1f'%d['host_enqueue_us_per_step'], 'img/s', d.get('img_per_s'))
    except Exception as e:
        print(f, 'ERR', e)
PY
