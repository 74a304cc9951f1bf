#!/bin/bash
# quick N-GPU check of the latest changes: (multi-)GPU tests + default bench line
N=${1:-2}
mkdir -p gpurun_out
export CHAINER_B200_PEER_TIMEOUT_S=30
export BENCH_WATCHDOG_S=200
if [ "$N" = "1" ]; then
  timeout 300 python -m pytest tests -m gpu -q -x > gpurun_out/r02_check_n1.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02_check_n1.log
  timeout 260 python bench.py --steps 100 --warmup 10 --legs-deadline-s 120 > gpurun_out/r02_check_bench_n1.json 2> gpurun_out/r02_check_bench_n1.err; echo "bench rc=$?"
else
  timeout 150 python -m pytest tests/test_multi_gpu.py -q -rA -k "test_multi_gpu_path[$N-peer-memory]" > gpurun_out/r02_check_multi_gpu_n$N.log 2>&1; echo "multi-gpu tests rc=$?"
  grep -E "^PASSED|^FAILED|passed|failed|Error|MIXED" gpurun_out/r02_check_multi_gpu_n$N.log | head
  timeout 260 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29911 bench.py --gpus $N --steps 100 --warmup 10 --legs-deadline-s 120 > gpurun_out/r02_check_bench_n$N.json 2> gpurun_out/r02_check_bench_n$N.err; echo "bench rc=$?"
fi
tail -5 gpurun_out/r02_check_bench_n$N.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r02_check_bench_n$N.json').read().splitlines() if l.startswith('{')][-1])
print('ms/step', d['ms_per_step'], 'parity', d['parity']['ok'], 'img/s', d.get('img_per_s'), 'legs_error', d.get('legs_error'))
print('config3', json.dumps(d.get('config3'))[:1200])
PY
