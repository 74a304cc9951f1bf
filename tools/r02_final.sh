#!/bin/bash
# Round-2 evidence run at N GPUs (tools/r02_final.sh N): tests with a full -rA log, the default
# bench line (what the driver runs), the separate-launch comparison, seq2seq / Adam, the
# BASELINE config 5 size sweep.  Everything lands in gpurun_out/ and is copied to profiles/.
N=${1:-1}
mkdir -p gpurun_out
export CHAINER_B200_PEER_TIMEOUT_S=60
if [ "$N" = "1" ]; then
  timeout 1500 python -m pytest tests -m gpu -q -rA > gpurun_out/r02_final_gpu_tests_n1.log 2>&1; echo "gpu tests rc=$?"
  grep -E "passed|failed" gpurun_out/r02_final_gpu_tests_n1.log | tail -2
  python __graft_entry__.py smoke > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r02_smoke.log
  LAUNCH="python"
else
  timeout 1500 python -m pytest tests/test_multi_gpu.py -q -rA -k "test_multi_gpu_path[$N-" > gpurun_out/r02_final_multi_gpu_n$N.log 2>&1; echo "multi-gpu tests rc=$?"
  grep -E "^PASSED|^FAILED|^SKIPPED|passed|failed" gpurun_out/r02_final_multi_gpu_n$N.log | head
  LAUNCH="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
fi
run() {
  name=$1; shift
  timeout 900 $LAUNCH $PORT bench.py --gpus $N "$@" > gpurun_out/r02_final_bench_n${N}_$name.json 2> gpurun_out/r02_final_bench_n${N}_$name.err
  echo "bench $name rc=$?"
}
port() { if [ "$N" != "1" ]; then PORT="--master-port $((29500 + RANDOM % 1000))"; else PORT=""; fi; }
port; run default --steps 200 --warmup 20
port; run nostep --steps 100 --warmup 10 --no-step --no-train --no-e2e --no-config3 --no-cpu-baseline
if [ "$N" = "1" ] || [ "$N" = "8" ]; then
  port; run seq2seq --steps 40 --warmup 5 --workload seq2seq --no-train --no-e2e --no-cpu-baseline
fi
port
if [ "$N" = "1" ]; then
  timeout 900 python tools/size_sweep.py --out gpurun_out/r02_final_size_sweep_n1.json > gpurun_out/r02_final_size_sweep_n1.log 2>&1
else
  timeout 900 $LAUNCH $PORT tools/size_sweep.py --dtypes float32,float16 --optimizers momentum_sgd --out gpurun_out/r02_final_size_sweep_n$N.json > gpurun_out/r02_final_size_sweep_n$N.log 2>&1
fi
echo "size sweep rc=$? rows=$(grep -c 'GB/s/GPU' gpurun_out/r02_final_size_sweep_n$N.log)"
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/r02_final_bench_n${N}_*.json')):
    try:
        d=json.loads([l for l in open(f).read().splitlines() if l.startswith('{')][-1])
        r=d['roofline']; a=d.get('allreduce') or {}; c3=d.get('config3') or {}
        print(f.split('/')[-1], 'ms/step %.4f'%d['ms_per_step'], 'value %.0f'%d['value'], r['kernel'], 'parity', (d.get('parity') or {}).get('ok'), (d.get('parity') or {}).get('mode','')[:9],
              'allreduce us %.1f'%a.get('us',0), 'img/s', d.get('img_per_s'), 'e2e', (d.get('e2e') or {}).get('ms_per_step'),
              'fp16', (c3.get('float16_buffer') or {}).get('ms_per_step'), 'mnbn', (c3.get('mnbn') or {}).get('us_per_step'), 'cpu', (d.get('cpu_baseline') or {}).get('ms_per_step'))
    except Exception as e:
        print(f, 'ERR', e)
PY
