#!/bin/bash
# tools/multi_gpu_eval2.sh N -- multicast validation + chunked-pipeline variants
N=${1:-8}
mkdir -p gpurun_out
[ -n "$SKIP_TESTS" ] || { echo "== tests"; timeout 600 python -m pytest tests/test_multi_gpu.py -q -m gpu -k "${N}-multicast or ${N}-nccl" > gpurun_out/tests_n${N}.log 2>&1; tail -${TAIL:-30} gpurun_out/tests_n${N}.log; }
for f in "" "--mc-chunk-mb 52" "--mc-chunk-mb 34" $EXTRA; do
echo "== bench $f"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29902 bench.py --gpus $N --steps 100 --warmup 10 --no-cpu-baseline --no-e2e $f 2>&1 | grep '^{"metric' | tee gpurun_out/bench_n${N}_mc$(echo $f | tr -d ' -').json | python -c "
import sys,json
l=json.loads(sys.stdin.read()); print(l['config']['allreduce_impl'], 'ms/step %.4f'%l['ms_per_step'], 'value %.0f'%l['value'], l.get('allreduce'))"
done
