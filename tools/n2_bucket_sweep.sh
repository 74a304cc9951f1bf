#!/bin/bash
# bucket-size sweep of the NCCL pipeline at N GPUs (default 2): tools/n2_bucket_sweep.sh [N]
N=${1:-2}
port=29610
for b in 8 16 32 64 200; do
port=$((port+1))
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N --steps 200 --warmup 10 --no-cpu-baseline --no-e2e --bucket-mb $b 2>&1 | grep '^{"metric' | python -c "
import sys,json
l=json.loads(sys.stdin.read()); print('bucket_mb', $b, 'ms/step %.4f'%l['ms_per_step'], 'value %.0f'%l['value'], 'AR us %.1f bus %.0f'%(l['allreduce']['us'], l['allreduce']['bus_gbs']), 'upd us %.1f pack us %.1f'%(l['roofline']['us_per_launch'], l['roofline']['pack_us']))"
done
