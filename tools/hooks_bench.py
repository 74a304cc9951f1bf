"""Cost of the fused hooks and of the other rules, through the public API (N = 1 or
under torchrun): ResNet-50 gradient list, device time per step (CUDA events).

    python tools/hooks_bench.py [--out gpurun_out/hooks_bench.json]

Rows: MomentumSGD plain / +WeightDecay / +GradientClipping / +both / +loss scale,
SGD, CorrectedMomentumSGD, NesterovAG, Adam plain / +both hooks, and the unfused
reference sequence ([WeightDecay, GradientClipping] order) for comparison.
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
    ap.add_argument('--out', default='gpurun_out/hooks_bench.json')
    ap.add_argument('--steps', type=int, default=50)
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', '0')))
    if world > 1:
        dist.init_process_group('gloo', rank=rank, world_size=world)
    import chainer_b200
    from chainer_b200 import optimizer_hooks as H
    from chainer_b200 import workloads
    from chainer_b200.core.link import link_from_named_arrays
    peak = 6462.1
    try:
        peak = json.load(open('MEASURED_PEAKS.json'))['hbm_gbs']
    except Exception:
        pass
    plist = workloads.resnet50()
    counts = [int(np.prod(s)) for _, s in plist]
    n = sum(counts)
    offs = np.concatenate([[0], np.cumsum(counts)])
    cases = [
        ('momentum_sgd', lambda: chainer_b200.MomentumSGD(lr=0.01), [], None, 32),
        ('momentum_sgd + wd', lambda: chainer_b200.MomentumSGD(lr=0.01), ['wd'], None, 32),
        ('momentum_sgd + clip', lambda: chainer_b200.MomentumSGD(lr=0.01), ['clip'], None, 36),
        ('momentum_sgd + clip + wd', lambda: chainer_b200.MomentumSGD(lr=0.01), ['clip', 'wd'], None, 36),
        ('momentum_sgd + wd + loss scale', lambda: chainer_b200.MomentumSGD(lr=0.01), ['wd'], 128.0, 32),
        ('momentum_sgd, [wd, clip] (unfused reference sequence)',
         lambda: chainer_b200.MomentumSGD(lr=0.01), ['wd', 'clip'], None, None),
        ('sgd', lambda: chainer_b200.SGD(lr=0.01), [], None, 24),
        ('corrected_momentum_sgd', lambda: chainer_b200.CorrectedMomentumSGD(lr=0.01), [], None, 32),
        ('nesterov_ag', lambda: chainer_b200.NesterovAG(lr=0.01), [], None, 32),
        ('adam', lambda: chainer_b200.Adam(), [], None, 40),
        ('adam + clip + wd', lambda: chainer_b200.Adam(), ['clip', 'wd'], None, 44),
    ]
    rows = []
    for name, mk, hooks, ls, bpe in cases:
        comm = chainer_b200.create_communicator('pure_nccl')
        p_arena = torch.randn(n, device='cuda') * 0.05
        g_arenas = [torch.randn(n, device='cuda') * 1e-2 for _ in range(2)]
        views = lambda a: [a[offs[i]:offs[i + 1]] for i in range(len(counts))]  # noqa: E731
        model = link_from_named_arrays([(nm, v) for (nm, _), v in zip(plist, views(p_arena))])
        params = [p for _, p in sorted(model.namedparams())]
        gv = [views(a) for a in g_arenas]
        actual = mk()
        opt = chainer_b200.create_multi_node_optimizer(actual, comm)
        opt.setup(model)
        for h in hooks:
            opt.add_hook(H.WeightDecay(1e-4) if h == 'wd' else H.GradientClipping(1.0))
        if ls is not None:
            actual.loss_scaling(scale=ls)

        def step(k):
            for p, g in zip(params, gv[k % 2]):
                p.grad = g
                p._loss_scale = ls
            opt.update()
        for k in range(6):
            step(k)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for k in range(args.steps):
            step(k)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / args.steps
        if world > 1:
            t = torch.tensor([us], dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            us = float(t.item())
        row = dict(case=name, us_per_step=us, n_gpus=world, fused=comm._fused_plan is not None)
        if bpe is not None:
            row['gbs'] = n * bpe / us / 1e3
            row['frac_of_measured_peak'] = row['gbs'] / peak
        rows.append(row)
        if rank == 0:
            print('%-55s %9.1f us/step  %s' % (name, us, '%.0f GB/s (%.2f)' % (
                row['gbs'], row['frac_of_measured_peak']) if bpe else '(launch-bound)'), flush=True)
        comm.finalize()
        del model, opt, actual, params, gv, g_arenas, p_arena
        torch.cuda.empty_cache()
    if rank == 0:
        os.makedirs(os.path.dirname(args.out) or '.', exist_ok=True)
        json.dump(dict(peak=peak, n_gpus=world, rows=rows,
                       note='bytes/elem: 32 MomentumSGD with write-back (+4 for the norm pass), '
                            '24 SGD, 40 Adam'), open(args.out, 'w'), indent=1)


if __name__ == '__main__':
    main()
