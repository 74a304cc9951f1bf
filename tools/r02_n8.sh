#!/bin/bash
# round-2 N-GPU run (default 8): NVLink status, tests of the three transports, step sweep, then the
# bench (best step tuning found, separate launches, fp16 + MNBN, seq2seq Adam) and the size sweep.
N=${1:-8}
mkdir -p gpurun_out
export CHAINER_B200_PEER_TIMEOUT_S=60
nvidia-smi nvlink --status -i 0 > gpurun_out/r02_nvlink_status.txt 2>&1
nvidia-smi topo -m >> gpurun_out/r02_nvlink_status.txt 2>&1
head -8 gpurun_out/r02_nvlink_status.txt
timeout 1500 python -m pytest tests/test_multi_gpu.py -q -rA -k "test_multi_gpu_path[$N-" > gpurun_out/r02_multi_gpu_n$N.log 2>&1; echo "multi-gpu tests rc=$?"
grep -E "^PASSED|^FAILED|^SKIPPED|passed|failed" gpurun_out/r02_multi_gpu_n$N.log | head -20
CFG="step=0;reducers=16;reducers=32;reducers=48;reducers=64;reducers=96;reducers=128;reducers=32,unroll=8;reducers=48,unroll=8;reducers=64,unroll=8;reducers=96,unroll=8;reducers=64,unroll=2;reducers=96,unroll=2;reducers=128,unroll=2;reducers=64,tile_elems=32768;reducers=64,tile_elems=8192;reducers=64,unroll=8,tile_elems=32768;reducers=64,ctas_per_sm=5;reducers=48,unroll=8,ctas_per_sm=3"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 1000)) \
  tools/step_sweep.py --multicast on --configs "$CFG" --out gpurun_out/r02_step_sweep_n${N}_on.json 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM"
BEST=$(python - <<PY
import json
d=json.load(open('gpurun_out/r02_step_sweep_n${N}_on.json'))
rows=[r for r in d['rows'] if not r['config'].startswith('step=0')]
print(min(rows,key=lambda r:r['ms_per_step'])['config'])
PY
)
echo "best step tuning: $BEST"
run() {
  name=$1; shift
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 1000)) \
    bench.py --gpus $N "$@" > gpurun_out/r02_bench_n${N}_$name.json 2> gpurun_out/r02_bench_n${N}_$name.err
  echo "bench $name rc=$?"
}
run step --steps 100 --warmup 10 --step-tuning "$BEST"
run nostep --steps 100 --warmup 10 --no-step --no-train --no-e2e
run f16_mnbn --steps 100 --warmup 10 --allreduce-dtype float16 --mnbn --no-train --no-e2e --step-tuning "$BEST"
run seq2seq --steps 40 --warmup 5 --workload seq2seq --no-train --no-e2e --step-tuning "$BEST"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 1000)) \
  tools/size_sweep.py --dtypes float32,float16 --optimizers momentum_sgd --out gpurun_out/r02_size_sweep_n$N.json > gpurun_out/r02_size_sweep_n$N.log 2>&1
echo "size sweep rc=$?"
python - <<'PY'
import json,glob,os
for f in sorted(glob.glob('gpurun_out/r02_bench_n*_*.json')):
    try:
        d=json.loads([l for l in open(f).read().splitlines() if l.startswith('{')][-1])
        if d['n_gpus'] < 4: continue
        r=d['roofline']; a=d.get('allreduce') or {}
        print(f.split('/')[-1], 'ms/step %.4f'%d['ms_per_step'], 'value %.0f'%d['value'], r['kernel'], 'us %.1f'%r['us_per_launch'], 'parity', (d.get('parity') or {}).get('ok'), (d.get('parity') or {}).get('mode','')[:10], 'allreduce us %.1f wire %.0f bus %.0f'%(a.get('us',0),a.get('wire_gbs',0),a.get('bus_gbs',0)), 'img/s', d.get('img_per_s'), 'mnbn', (d.get('mnbn') or {}).get('us_per_step'))
    except Exception as e:
        print(f, 'ERR', e)
PY
