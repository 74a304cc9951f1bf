#!/bin/bash
# round-2 multi-GPU validation at N ranks: tests of the three transports (full -rA log), then the
# bench with the one-launch step (default transport), the other transport, and the separate launches
N=${1:-2}
mkdir -p gpurun_out
export CHAINER_B200_PEER_TIMEOUT_S=60
timeout 1500 python -m pytest tests/test_multi_gpu.py -x -q -rA -k "test_multi_gpu_path[$N-" > gpurun_out/r02_multi_gpu_n$N.log 2>&1; echo "multi-gpu tests rc=$?"
grep -E "PASSED|FAILED|SKIPPED|passed|failed|ONE-LAUNCH|MNBN stat" gpurun_out/r02_multi_gpu_n$N.log | head -20
run() {
  name=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 1000)) \
    bench.py --gpus $N --steps 100 --warmup 10 "$@" > gpurun_out/r02_bench_n${N}_$name.json 2> gpurun_out/r02_bench_n${N}_$name.err
  echo "bench $name rc=$?"
}
run step
if [ "$N" = "2" ]; then run step_mc --multicast on --no-train --no-e2e; else run step_p2p --multicast off --no-train --no-e2e; fi
run nostep --no-step --no-train --no-e2e
for t in "reducers=16" "reducers=64" "tile_elems=8192" "tile_elems=32768" "unroll=8" "unroll=2" "ctas_per_sm=2" "reducers=64,unroll=8" "tile_elems=8192,reducers=64"; do
  run "step_$t" --no-train --no-e2e --no-parity --step-tuning "$t"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02_bench_n%s_*.json' % __import__('os').environ.get('NN','*'))):
    try:
        d=json.loads([l for l in open(f).read().splitlines() if l.startswith('{')][-1])
        r=d['roofline']; a=d.get('allreduce') or {}
        print(f.split('/')[-1], 'ms/step %.4f'%d['ms_per_step'], 'value %.0f'%d['value'], r['kernel'], 'us %.1f'%r['us_per_launch'], 'parity', (d.get('parity') or {}).get('ok'), (d.get('parity') or {}).get('mode','')[:10], 'allreduce us %.1f wire %.0f'%(a.get('us',0),a.get('wire_gbs',0)), d['impl_detail']['transport'], 'img/s', d.get('img_per_s'))
    except Exception as e:
        print(f, 'ERR', e)
PY
