#!/bin/bash
N=${1:-4}
mkdir -p gpurun_out
export CHAINER_B200_PEER_TIMEOUT_S=60
timeout 1500 python -m pytest tests/test_multi_gpu.py -x -q -rA -k "test_multi_gpu_path[$N-" > gpurun_out/r02_multi_gpu_n$N.log 2>&1; echo "multi-gpu tests rc=$?"
grep -E "PASSED|FAILED|SKIPPED|passed|failed|Error" gpurun_out/r02_multi_gpu_n$N.log | head -20
bash tools/r02_sweep.sh $N on
CFG="step=0;reducers=64;reducers=128;reducers=192;reducers=256"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 1000)) \
  tools/step_sweep.py --multicast off --configs "$CFG" --out gpurun_out/r02_step_sweep_n${N}_off.json 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM"
