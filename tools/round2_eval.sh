#!/bin/bash
# What round 1 could not measure (GPU budget).  Usage:
#   gpurun --timeout 900 -- 'bash tools/round2_eval.sh 1'
#   gpurun --gpus N --timeout 900 -- 'bash tools/round2_eval.sh N'      (N = 2, 4, 8)
N=${1:-1}
mkdir -p gpurun_out
run() {  # run <tag> <script> [args...]: python or torchrun depending on N
  tag=$1; shift
  if [ "$N" = 1 ]; then timeout 600 python "$@"
  else timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) "$@"; fi 2>&1 | tail -${TAIL:-25}
}
echo "== default bench line (with e2e) at N=$N"
run bench bench.py --gpus $N | grep '^{' | tee gpurun_out/r02_bench_n${N}.json | cut -c1-300
echo "== seq2seq (Adam) at N=$N"
run seq2seq bench.py --gpus $N --workload seq2seq --no-cpu-baseline --steps 30 --warmup 5 | grep '^{' | tee gpurun_out/r02_bench_seq2seq_n${N}.json | cut -c1-300
echo "== seq2seq, 99% of embedding-gradient rows zero"
run seq2seq0 bench.py --gpus $N --workload seq2seq --no-cpu-baseline --no-e2e --steps 30 --warmup 5 --zero-embedding-rows 0.99 | grep '^{' | tee gpurun_out/r02_bench_seq2seq_sparse_n${N}.json | cut -c1-200
echo "== fp16 allreduce buffer (config 3)"
run fp16 bench.py --gpus $N --allreduce-dtype float16 --no-cpu-baseline --no-e2e --steps 100 --warmup 10 | grep '^{' | tee gpurun_out/r02_bench_fp16_n${N}.json | cut -c1-200
echo "== hooks / rule family cost"
TAIL=14 run hooks tools/hooks_bench.py --out gpurun_out/r02_hooks_bench_n${N}.json
echo "== size sweep (config 5)"
TAIL=12 run sweep tools/size_sweep.py --out gpurun_out/r02_size_sweep_n${N}.json --max-mb 1024
