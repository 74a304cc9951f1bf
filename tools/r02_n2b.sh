#!/bin/bash
N=2
mkdir -p gpurun_out
export CHAINER_B200_PEER_TIMEOUT_S=60
timeout 900 python -m pytest tests/test_multi_gpu.py -q -rA -k "test_multi_gpu_path[$N-" > gpurun_out/r02_multi_gpu_n$N.log 2>&1; echo "multi-gpu tests rc=$?"
grep -E "^PASSED|^FAILED|^SKIPPED|passed|failed" gpurun_out/r02_multi_gpu_n$N.log | head
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29711 \
  tools/size_sweep.py --dtypes float32,float16 --optimizers momentum_sgd,adam --out gpurun_out/r02_size_sweep_n$N.json > gpurun_out/r02_size_sweep_n$N.log 2>&1
echo "size sweep rc=$?"; grep -c "GB/s/GPU" gpurun_out/r02_size_sweep_n$N.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29712 \
  bench.py --gpus $N --steps 100 --warmup 10 --mnbn > gpurun_out/r02_bench_n${N}_final.json 2> gpurun_out/r02_bench_n${N}_final.err; echo "bench rc=$?"
# compute-sanitizer: memcheck over every collective kernel at N = 2 (peer-memory, then multicast), racecheck at N = 1
export CHAINER_B200_PEER_TIMEOUT_S=600
timeout 600 python -m torch.distributed.run --no-python --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29713 \
  compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_run.py > gpurun_out/r02_sanitizer_memcheck_n2_peer.log 2>&1; echo "memcheck peer rc=$?"
CHAINER_B200_MULTICAST=1 timeout 600 python -m torch.distributed.run --no-python --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29714 \
  compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_run.py > gpurun_out/r02_sanitizer_memcheck_n2_mc.log 2>&1; echo "memcheck mc rc=$?"
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_run.py > gpurun_out/r02_sanitizer_racecheck_n1.log 2>&1; echo "racecheck n1 rc=$?"
grep -h "ERROR SUMMARY\|RACECHECK SUMMARY\|SANITIZE RANK" gpurun_out/r02_sanitizer_*.log
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02_bench_n2_final.json').read().splitlines() if l.startswith('{')][-1])
print('ms/step', d['ms_per_step'], 'parity', d['parity']['ok'], d['parity']['mode'][:12], 'img/s', d.get('img_per_s'), 'mnbn', d['mnbn']['us_per_step'], d['mnbn']['launches_per_step'], 'allreduce', d['allreduce']['us'])
PY
