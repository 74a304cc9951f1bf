"""BASELINE config 5: mean_grad+update over packed buffers of 1 KiB ... 1 GiB
(x4 steps, plus the ResNet-50 and seq2seq sizes), fp32 / fp16 / bf16 buffers, two
tensor-list shapes (one tensor; the ResNet-50 size histogram scaled to the byte
count), MomentumSGD and Adam, at N GPUs (run under torchrun for N > 1).

    python tools/size_sweep.py [--out gpurun_out/size_sweep.json] [--max-mb 1024]

Each point is the median CUDA-event time of the public-API step
(`create_multi_node_optimizer(...).update()`): pack -> allreduce -> fused update.
Reported: us/step, algorithmic GB/s and fraction of the measured HBM peak; for
N > 1 also the allreduce bus bandwidth of that buffer.  Sizes < ~1 MB are
latency-bound: read the us column.
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import torch.distributed as dist
    ap = argparse.ArgumentParser()
    ap.add_argument('--out', default='gpurun_out/size_sweep.json')
    ap.add_argument('--max-mb', type=float, default=1024)
    ap.add_argument('--dtypes', default='float32,float16,bfloat16')
    ap.add_argument('--optimizers', default='momentum_sgd,adam')
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', '0')))
    if world > 1:
        dist.init_process_group('gloo', rank=rank, world_size=world)
    import chainer_b200
    from chainer_b200 import workloads
    from chainer_b200.core.link import link_from_named_arrays
    peak = 6462.1
    try:
        peak = json.load(open('MEASURED_PEAKS.json'))['hbm_gbs']
    except Exception:
        pass
    sizes = []
    b = 1024
    while b <= args.max_mb * (1 << 20):
        sizes.append(b)
        b *= 4
    sizes += [102228384, 693203200]
    sizes = sorted(set(s for s in sizes if s <= args.max_mb * (1 << 20)))
    rows = []
    for adt in args.dtypes.split(','):
        bsz = 4 if adt == 'float32' else 2
        comm = chainer_b200.create_communicator(
            'pure_nccl', allreduce_grad_dtype={'float32': np.float32, 'float16': np.float16,
                                               'bfloat16': 'bfloat16'}[adt])
        for nbytes in sizes:
            n = max(nbytes // bsz, 4) // 4 * 4
            for shape_kind in ('single', 'resnet50_hist'):
                if shape_kind == 'single':
                    plist = [('/w', (n,))]
                else:
                    if n < 162 * 64:
                        continue
                    plist = workloads.scaled_histogram(n)
                counts = [int(np.prod(s)) for _, s in plist]
                total = sum(counts)
                for opt_name in args.optimizers.split(','):
                    p_arena = torch.randn(total, device='cuda') * 0.05
                    g_arenas = [torch.randn(total, device='cuda') * 1e-2 for _ in range(2)]
                    offs = np.concatenate([[0], np.cumsum(counts)])
                    views = lambda a: [a[offs[i]:offs[i + 1]] for i in range(len(counts))]  # noqa: E731
                    model = link_from_named_arrays([(nm, v) for (nm, _), v in zip(plist, views(p_arena))])
                    params = [p for _, p in sorted(model.namedparams())]
                    gv = [views(a) for a in g_arenas]
                    actual = chainer_b200.MomentumSGD(lr=0.01) if opt_name == 'momentum_sgd' \
                        else chainer_b200.Adam()
                    opt = chainer_b200.create_multi_node_optimizer(actual, comm)
                    opt.setup(model)

                    def step(k):
                        for p, g in zip(params, gv[k % 2]):
                            p.grad = g
                        opt.update()
                    step(0)
                    for k in range(5):
                        step(k)
                    torch.cuda.synchronize()
                    if world > 1:
                        dist.barrier()
                    reps = 30 if total * 4 < (64 << 20) else 12
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for k in range(reps):
                        step(k)
                    e1.record()
                    torch.cuda.synchronize()
                    us = e0.elapsed_time(e1) * 1e3 / reps
                    if world > 1:
                        t = torch.tensor([us], dtype=torch.float64)
                        dist.all_reduce(t, op=dist.ReduceOp.MAX)
                        us = float(t.item())
                    bpe = (4 + bsz) + (bsz + 16 + (8 if opt_name == 'adam' else 0) + 4)
                    gbs = total * bpe / us / 1e3
                    row = dict(dtype=adt, packed_bytes=total * bsz, n_elems=total, tensors=len(counts),
                               shape=shape_kind, optimizer=opt_name, n_gpus=world, us=us, gbs=gbs,
                               frac=gbs / peak, bytes_per_elem=bpe)
                    rows.append(row)
                    if rank == 0:
                        print('%-8s %12d B %-13s %-12s N=%d  %9.1f us  %7.1f GB/s/GPU (%.2f)' % (
                            adt, total * bsz, shape_kind, opt_name, world, us, gbs, gbs / peak), flush=True)
                    del model, opt, actual, params, gv, g_arenas, p_arena
                    torch.cuda.empty_cache()
        comm.finalize()
    if rank == 0:
        os.makedirs(os.path.dirname(args.out) or '.', exist_ok=True)
        json.dump(dict(peak=peak, n_gpus=world, rows=rows,
                       note='us = whole step through the public API (host enqueue ~45 us/step is a floor '
                            'for small sizes); gbs per GPU'), open(args.out, 'w'), indent=1)


if __name__ == '__main__':
    main()
