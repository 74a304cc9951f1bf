#!/bin/bash
# round-2 ncu evidence (N = 1): launch list of a short default bench run, full capture of the step
# kernel (MomentumSGD and Adam), the BN kernels and the master kernel.
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_launches_bench.csv \
  python bench.py --steps 20 --warmup 5 --no-e2e --no-train --no-cpu-baseline --no-parity > gpurun_out/r02_launches_bench.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:walk_kernel -s 10 -c 2 -o gpurun_out/r02_prof_step_sgd -f \
  python bench.py --steps 10 --warmup 5 --no-e2e --no-train --no-cpu-baseline --no-parity > gpurun_out/r02_prof_step_sgd.log 2>&1
echo "ncu step sgd rc=$?"
ncu --set full --clock-control none --import-source on -k regex:walk_kernel -s 10 -c 2 -o gpurun_out/r02_prof_step_adam -f \
  python bench.py --steps 10 --warmup 5 --optimizer adam --no-e2e --no-train --no-cpu-baseline --no-parity > gpurun_out/r02_prof_step_adam.log 2>&1
echo "ncu step adam rc=$?"
ncu --set full --clock-control none -k regex:bn_ -c 12 -o gpurun_out/r02_prof_bn -f \
  python tools/bn_probe.py > gpurun_out/r02_prof_bn.log 2>&1
echo "ncu bn rc=$?"
for f in r02_prof_step_sgd r02_prof_step_adam r02_prof_bn; do
  ncu -i gpurun_out/$f.ncu-rep --page raw --csv > gpurun_out/$f.raw.csv 2>/dev/null
done
ls -la gpurun_out/*.ncu-rep
