"""Launch each stream kernel a few times at a BASELINE workload, for ncu.

    ncu --set full --clock-control none --import-source on -k regex:walk_kernel \
        -o gpurun_out/prof python tools/prof_kernels.py [--kinds pack,sgd,adam] [--reps 2]
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


class P(object):
    def __init__(self, data, grad):
        self.data, self.grad = data, grad


def main():
    import torch
    from chainer_b200 import _lib, workloads
    from chainer_b200.communicators import _memory_utility as mu
    ap = argparse.ArgumentParser()
    ap.add_argument('--workload', default='resnet50')
    ap.add_argument('--kinds', default='pack,sgd,sgd_wg,adam')
    ap.add_argument('--reps', type=int, default=2)
    ap.add_argument('--buf', default='float32')
    ap.add_argument('--set', action='append', default=[], help='tuning key=value')
    args = ap.parse_args()
    lib = _lib.get()
    for kv in args.set:
        k, v = kv.split('=')
        lib.gp_set_tuning(k.encode(), int(v))
    plist = workloads.WORKLOADS[args.workload]()
    sizes = [int(np.prod(s)) for _, s in plist]
    n = sum(sizes)
    bdt = {'float32': torch.float32, 'float16': torch.float16, 'bfloat16': torch.bfloat16}[args.buf]
    bid = {'float32': 7, 'float16': 6, 'bfloat16': 9}[args.buf]
    torch.manual_seed(0)
    grads = [torch.randn(k, device='cuda') * 1e-2 for k in sizes]
    data = [torch.randn(k, device='cuda') * 0.05 for k in sizes]
    m = [torch.zeros(k, device='cuda') for k in sizes]
    v = [torch.zeros(k, device='cuda') for k in sizes]
    buf = torch.zeros(n, dtype=bdt, device='cuda')
    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
    params = [P(d, g) for d, g in zip(data, grads)]
    pd_sgd = mu.ParamsData(params, 'grad', False, extra_ptrs=[(d, [x]) for d, x in zip(data, v)])
    pd_adam = mu.ParamsData(params, 'grad', False,
                            extra_ptrs=[(d, [x, y]) for d, x, y in zip(data, m, v)])
    torch.cuda.synchronize()
    bnp = {'float32': np.float32, 'float16': np.float16, 'bfloat16': 'bfloat16'}[args.buf]
    hint = pd_sgd.layout_hint(bnp)
    for kind in args.kinds.split(','):
        for _ in range(args.reps):
            flush.zero_()        # evict L2 between launches (a vectorized fill, not walk_kernel)
            if kind == 'pack':
                lib.gp_pack(buf.data_ptr(), bid, pd_sgd.d_csum, pd_sgd.d_segs, len(sizes), 0, n, 1.0, hint, 0)
            elif kind == 'unpack':
                lib.gp_unpack_scale(buf.data_ptr(), bid, pd_sgd.d_csum, pd_sgd.d_segs, len(sizes), 0, n,
                                    0.125, hint, 0)
            elif kind in ('sgd', 'sgd_wg'):
                lib.gp_unpack_momentum_sgd(buf.data_ptr(), bid, pd_sgd.d_csum, pd_sgd.d_segs,
                                           len(sizes), 0, n, 0.125, 0.01, 0.9,
                                           1 if kind == 'sgd_wg' else 0, hint, 0)
            elif kind in ('adam', 'adam_wg'):
                lib.gp_unpack_adam(buf.data_ptr(), bid, pd_adam.d_csum, pd_adam.d_segs, len(sizes), 0,
                                   n, 0.125, 1e-3, 0.1, 0.001, 1e-8, 1.0, 0.0, 0.0, 0.0, 0,
                                   1 if kind == 'adam_wg' else 0, hint, 0)
            torch.cuda.synchronize()
    print('done', args.kinds)


if __name__ == '__main__':
    main()
