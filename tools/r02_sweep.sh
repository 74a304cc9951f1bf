#!/bin/bash
N=${1:-2}
MC=${2:-auto}
mkdir -p gpurun_out
CFG="step=0;reducers=32;reducers=64;reducers=96;reducers=128;reducers=192;reducers=256;reducers=128,tile_elems=32768;reducers=128,tile_elems=65536;reducers=192,tile_elems=32768;reducers=64,tile_elems=32768;reducers=256,tile_elems=32768;reducers=128,tile_elems=8192;reducers=128,unroll=8;reducers=64,unroll=8;reducers=128,unroll=2;reducers=128,ctas_per_sm=3;reducers=128,ctas_per_sm=5"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 1000)) \
  tools/step_sweep.py --multicast $MC --configs "$CFG" --out gpurun_out/r02_step_sweep_n${N}_${MC}.json 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM"
