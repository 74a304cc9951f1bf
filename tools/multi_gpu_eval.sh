#!/bin/bash
# One-shot multi-GPU evaluation: tools/multi_gpu_eval.sh N  (tests, allreduce microbench, bench with both transports)
N=${1:-8}
mkdir -p gpurun_out
echo "== tests"; timeout 900 python -m pytest tests/test_multi_gpu.py -q -m gpu -k "${N}-multicast or ${N}-nccl" 2>&1 | tail -${TAIL:-40}
echo "== p2p_bench"; P2P_SIZES=25557096,173300800 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29901 tools/p2p_bench.py 2>&1 | grep -E "^n=|p2p|worst|multicast" | head -40
for f in "" "--multicast off" "--allreduce-dtype float32" "--allreduce-dtype float32 --multicast off"; do
echo "== bench $f"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29902 bench.py --gpus $N --steps 100 --warmup 10 --no-cpu-baseline --no-e2e $f 2>&1 | grep '^{"metric' | tee gpurun_out/bench_n${N}_$(echo $f | tr -d ' -').json | python -c "
import sys,json
l=json.loads(sys.stdin.read()); print(l['config']['allreduce_impl'], 'ms/step %.4f'%l['ms_per_step'], 'value %.0f'%l['value'], l.get('allreduce'))"
done
if [ "$N" = 8 ]; then
for f in "" "--multicast off"; do
echo "== bench N=4 $f"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29903 bench.py --gpus 4 --steps 100 --warmup 10 --no-cpu-baseline --no-e2e $f 2>&1 | grep '^{"metric' | tee gpurun_out/bench_n4_$(echo $f | tr -d ' -').json | python -c "
import sys,json
l=json.loads(sys.stdin.read()); print(l['config']['allreduce_impl'], 'ms/step %.4f'%l['ms_per_step'], 'value %.0f'%l['value'], l.get('allreduce'))"
done
fi
