#!/bin/bash
# One-shot multi-GPU evaluation: tools/multi_gpu_eval.sh N  (tests, allreduce microbench, bench with both transports)
N=${1:-8}
mkdir -p gpurun_out
echo "== tests"; timeout 900 python -m pytest tests/test_multi_gpu.py -x -q -m gpu -k "path and $N" 2>&1 | tail -4
echo "== p2p_bench"; P2P_SIZES=25557096,173300800 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29901 tools/p2p_bench.py 2>&1 | grep -E "^n=|p2p|worst" | head -20
for f in "" "--no-p2p" "--p2p-chunk-mb 0"; do
echo "== bench $f"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29902 bench.py --gpus $N --steps 200 --warmup 10 --no-cpu-baseline --no-e2e $f 2>&1 | grep '^{"metric' | tee gpurun_out/bench_n${N}_$(echo $f | tr -d ' -').json | python -c "
import sys,json
l=json.loads(sys.stdin.read()); print(l['config']['allreduce_impl'], 'ms/step %.4f'%l['ms_per_step'], 'value %.0f'%l['value'], l.get('allreduce'))"
done
