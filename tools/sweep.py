"""Tuning sweep of the stream kernels at a BASELINE workload (run on the GPU box).

    python tools/sweep.py [--workload resnet50] [--buf float32] [--out gpurun_out/sweep.json]

Times gp_pack, gp_unpack_scale, gp_unpack_momentum_sgd (write_grad 0/1) and
gp_unpack_adam with CUDA events over rotating data sets (so that nothing stays
in L2 between iterations) for every (threads, unroll, ctas_per_sm, persistent)
combination, and prints GB/s against the measured copy peak.
"""
import argparse
import itertools
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


class P(object):
    def __init__(self, data, grad):
        self.data, self.grad = data, grad


class Buf(object):
    def __init__(self, t):
        self.t = t

    def ptr(self):
        return self.t.data_ptr()


def main():
    import torch
    from chainer_b200 import _lib, workloads
    from chainer_b200.communicators import _memory_utility as mu
    ap = argparse.ArgumentParser()
    ap.add_argument('--workload', default='resnet50')
    ap.add_argument('--buf', default='float32')
    ap.add_argument('--iters', type=int, default=20)
    ap.add_argument('--sets', type=int, default=3)
    ap.add_argument('--out', default='gpurun_out/sweep.json')
    ap.add_argument('--quick', action='store_true')
    ap.add_argument('--bulk-sweep', action='store_true', help='sweep the TMA-staged kernels (tile, stages)')
    ap.add_argument('--bulk-debug', action='store_true', help='experiment: skip stores / arithmetic')
    args = ap.parse_args()
    lib = _lib.get()
    plist = workloads.WORKLOADS[args.workload]()
    sizes = [int(np.prod(s)) for _, s in plist]
    n = sum(sizes)
    bdt = {'float32': torch.float32, 'float16': torch.float16, 'bfloat16': torch.bfloat16}[args.buf]
    bid = {'float32': 7, 'float16': 6, 'bfloat16': 9}[args.buf]
    bsz = 4 if args.buf == 'float32' else 2
    sets = []
    for s in range(args.sets):
        torch.manual_seed(s)
        grads = [torch.randn(k, device='cuda') * 1e-2 for k in sizes]
        data = [torch.randn(k, device='cuda') * 0.05 for k in sizes]
        m = [torch.zeros(k, device='cuda') for k in sizes]
        v = [torch.zeros(k, device='cuda') for k in sizes]
        buf = torch.zeros(n, dtype=bdt, device='cuda')
        params = [P(d, g) for d, g in zip(data, grads)]
        pd_sgd = mu.ParamsData(params, 'grad', False, extra_ptrs=[(d, [x]) for d, x in zip(data, v)])
        pd_adam = mu.ParamsData(params, 'grad', False,
                                extra_ptrs=[(d, [x, y]) for d, x, y in zip(data, m, v)])
        sets.append(dict(buf=buf, pd_sgd=pd_sgd, pd_adam=pd_adam, keep=(grads, data, m, v, params)))
    torch.cuda.synchronize()

    bnp = {'float32': np.float32, 'float16': np.float16, 'bfloat16': 'bfloat16'}[args.buf]

    def run(kind, st):
        hint = st['pd_sgd'].layout_hint(bnp)
        if kind == 'pack':
            lib.gp_pack(st['buf'].data_ptr(), bid, st['pd_sgd'].d_csum, st['pd_sgd'].d_segs,
                        len(sizes), 0, n, 1.0, hint, 0)
        elif kind == 'unpack':
            lib.gp_unpack_scale(st['buf'].data_ptr(), bid, st['pd_sgd'].d_csum, st['pd_sgd'].d_segs,
                                len(sizes), 0, n, 0.125, hint, 0)
        elif kind in ('sgd', 'sgd_wg'):
            lib.gp_unpack_momentum_sgd(st['buf'].data_ptr(), bid, st['pd_sgd'].d_csum,
                                       st['pd_sgd'].d_segs, len(sizes), 0, n, 0.125, 0.01, 0.9,
                                       1 if kind == 'sgd_wg' else 0, hint, 0)
        elif kind in ('adam', 'adam_wg'):
            lib.gp_unpack_adam(st['buf'].data_ptr(), bid, st['pd_adam'].d_csum, st['pd_adam'].d_segs,
                               len(sizes), 0, n, 0.125, 1e-3, 0.1, 0.001, 1e-8, 1.0, 0.0, 0.0, 0.0,
                               0, 1 if kind == 'adam_wg' else 0, hint, 0)

    bytes_per_elem = {'pack': 4 + bsz, 'unpack': bsz + 4, 'sgd': bsz + 16, 'sgd_wg': bsz + 20,
                      'adam': bsz + 24, 'adam_wg': bsz + 28}
    peak = 6462.1
    try:
        peak = json.load(open('MEASURED_PEAKS.json'))['hbm_gbs']
    except Exception:
        pass

    def time_kind(kind):
        for w in range(3):
            run(kind, sets[w % len(sets)])
        torch.cuda.synchronize()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
               for _ in range(args.iters)]
        for i in range(args.iters):
            evs[i][0].record()
            run(kind, sets[i % len(sets)])
            evs[i][1].record()
        torch.cuda.synchronize()
        ts = sorted(a.elapsed_time(b) * 1e3 for a, b in evs)
        return ts[len(ts) // 2], ts[0]

    if args.quick:
        lib.gp_set_tuning(b'bulk', 0)
        grid = [(256, 0, 0, 0), (128, 2, 0, 0), (128, 4, 0, 0), (256, 2, 0, 0), (256, 4, 0, 0), (384, 2, 0, 0),
                (512, 2, 0, 0), (512, 4, 0, 0), (256, 2, 16, 1)]
    else:
        grid = list(itertools.product([128, 256, 512], [1, 2, 4], [2, 4, 8, 16], [1])) + \
            list(itertools.product([128, 256, 512], [1, 2, 4], [0], [0]))
    results = []
    kinds = ['pack', 'unpack', 'sgd', 'sgd_wg', 'adam', 'adam_wg']
    if args.bulk_debug:
        kinds = ['sgd', 'adam']
        lib.gp_set_tuning(b'bulk', 1)
        for dbg, chunk in [(3, 8192), (3, 4096), (3, 2048), (3, 1024), (3, 512), (0, 4096), (0, 2048),
                           (0, 1024), (0, 512)]:
            for tile, stages, ctas in [(2048, 4, 1), (4096, 4, 1), (2048, 3, 2), (1024, 6, 2)]:
                lib.gp_set_tuning(b'bulk_chunk', chunk)
                lib.gp_set_tuning(b'bulk_tile', tile)
                lib.gp_set_tuning(b'bulk_stages', stages)
                lib.gp_set_tuning(b'bulk_ctas', ctas)
                lib.gp_set_tuning(b'bulk_debug', dbg)
                msg = 'debug%d chunk%4d T%4d S%d C%d |' % (dbg, chunk, tile, stages, ctas)
                for kind in kinds:
                    med, best = time_kind(kind)
                    msg += ' %s %6.1fus' % (kind, med)
                print(msg, flush=True)
        lib.gp_set_tuning(b'bulk_debug', 0)
        return
    if args.bulk_sweep:
        kinds = ['sgd', 'sgd_wg', 'adam', 'adam_wg']
        combos = [(0, 2048, 4, 1)] + [(1, t, s, c) for c in (1, 2) for t in (1024, 2048, 4096)
                                      for s in (3, 4, 6, 8)]
        for bulk, tile, stages, ctas in combos:
            lib.gp_set_tuning(b'bulk', bulk)
            lib.gp_set_tuning(b'bulk_tile', tile)
            lib.gp_set_tuning(b'bulk_stages', stages)
            lib.gp_set_tuning(b'bulk_ctas', ctas)
            row = dict(bulk=bulk, tile=tile, stages=stages, ctas=ctas)
            msg = 'bulk%d T%4d S%d C%d |' % (bulk, tile, stages, ctas)
            for kind in kinds:
                med, best = time_kind(kind)
                gbs = bytes_per_elem[kind] * n / med / 1e3
                row[kind] = dict(us=med, best_us=best, gbs=gbs, frac=gbs / peak)
                msg += ' %s %6.1fus %4.0f (%.2f)' % (kind, med, gbs, gbs / peak)
            print(msg, flush=True)
            results.append(row)
        grid = []
    print('workload %s n_elems %d buf %s peak %.1f GB/s' % (args.workload, n, args.buf, peak))
    for threads, unroll, ctas, persistent in grid:
        lib.gp_set_tuning(b'threads', threads)
        lib.gp_set_tuning(b'unroll', unroll)
        lib.gp_set_tuning(b'ctas_per_sm', ctas)
        lib.gp_set_tuning(b'persistent', persistent)
        row = dict(threads=threads, unroll=unroll, ctas_per_sm=ctas, persistent=persistent)
        msg = 't%3d u%d c%2d p%d |' % (threads, unroll, ctas, persistent)
        for kind in kinds:
            med, best = time_kind(kind)
            gbs = bytes_per_elem[kind] * n / med / 1e3
            row[kind] = dict(us=med, best_us=best, gbs=gbs, frac=gbs / peak)
            msg += ' %s %6.1fus %4.0f (%.2f)' % (kind, med, gbs, gbs / peak)
        print(msg, flush=True)
        results.append(row)
    # plain copy reference on the same box: torch copy_ of the packed size
    a = torch.empty(n, device='cuda')
    b = torch.empty(n, device='cuda')
    for _ in range(3):
        b.copy_(a)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        b.copy_(a)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / 10
    print('torch copy_ of %d floats: %.1f us = %.0f GB/s' % (n, us, 8 * n / us / 1e3))
    os.makedirs(os.path.dirname(args.out) or '.', exist_ok=True)
    json.dump(dict(workload=args.workload, n_elems=n, buf=args.buf, peak=peak, results=results,
                   copy_us=us), open(args.out, 'w'), indent=1)


if __name__ == '__main__':
    main()
