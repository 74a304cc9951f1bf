#!/bin/bash
# chunk / CTA sweep of the peer-memory pipeline: tools/n2_p2p_sweep.sh [N]
N=${1:-2}
port=29810
for chunk in 0 16 32 64; do for ctas in 0 32 64 148; do
port=$((port+1))
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N --steps 200 --warmup 10 --no-cpu-baseline --no-e2e --p2p-chunk-mb $chunk --p2p-ctas $ctas 2>&1 | grep '^{"metric' | python -c "
import sys,json
l=json.loads(sys.stdin.read()); print('chunk_mb', $chunk, 'ctas', $ctas, 'ms/step %.4f'%l['ms_per_step'], 'value %.0f'%l['value'], 'AR us %.1f bus %.0f'%(l['allreduce']['us'], l['allreduce']['bus_gbs']))"
done; done
