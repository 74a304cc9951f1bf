"""Host (Python) cost of one fused `update()` call, measured WITHOUT a GPU: the
library is replaced by a double whose kernels return immediately, so what remains is
exactly the enqueue path of the product (plan lookup, table cache, bookkeeping, ctypes
argument marshalling is not included).  ResNet-50 layout by default.

    python tools/host_overhead.py [--profile]
"""
import argparse
import cProfile
import os
import pstats
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--profile', action='store_true')
    ap.add_argument('--optimizer', default='momentum_sgd')
    ap.add_argument('--steps', type=int, default=3000)
    ap.add_argument('--arrays', default='torch', choices=['torch', 'numpy'],
                    help='torch (CPU tensors: the pointer route a GPU run takes) or numpy arrays')
    ap.add_argument('--scale-elems', type=int, default=400000,
                    help='total elements (the size histogram of ResNet-50 scaled down; host cost '
                         'does not depend on it)')
    args = ap.parse_args()
    os.environ['CHAINER_B200_TEST_BACKEND'] = '1'     # this tool measures the host path on a double
    import chainer_b200
    from chainer_b200 import _lib, workloads
    from chainer_b200.core.link import link_from_named_arrays
    from tests import fake_lib

    class NullLib(fake_lib.FakeLib):
        def _noop(self, *a):
            return 0
    for name in ('gp_pack', 'gp_unpack_scale', 'gp_unpack_momentum_sgd', 'gp_unpack_adam',
                 'gp_unpack_momentum_sgd_hooked', 'gp_unpack_adam_hooked', 'gp_unpack_sgd_family',
                 'gp_sqnorm', 'gp_scale'):
        setattr(NullLib, name, NullLib._noop)
    lib = NullLib()
    _lib.set_backend_for_testing(lib)
    plist = workloads.scaled_histogram(args.scale_elems)
    rng = np.random.default_rng(0)
    if args.arrays == 'torch':
        import torch
        model = link_from_named_arrays(
            [(n, torch.from_numpy(rng.standard_normal(s).astype(np.float32))) for n, s in plist])
        params = [p for _, p in sorted(model.namedparams())]
        grads = [[torch.zeros_like(p.data) for p in params] for _ in range(2)]
    else:
        model = link_from_named_arrays(
            [(n, rng.standard_normal(s).astype(np.float32)) for n, s in plist])
        params = [p for _, p in sorted(model.namedparams())]
        grads = [[np.zeros_like(p.data) for p in params] for _ in range(2)]
    comm = chainer_b200.create_communicator('pure_nccl')
    actual = chainer_b200.MomentumSGD() if args.optimizer == 'momentum_sgd' else chainer_b200.Adam()
    opt = chainer_b200.create_multi_node_optimizer(actual, comm)
    opt.setup(model)

    def step(k):
        for p, g in zip(params, grads[k % 2]):
            p.grad = g
        opt.update()
    for k in range(10):
        step(k)
    if args.profile:
        pr = cProfile.Profile()
        pr.enable()
        for k in range(args.steps):
            step(k)
        pr.disable()
        pstats.Stats(pr).sort_stats('cumulative').print_stats(22)
        return
    # best of several short runs: the container shares its cores
    t_set = t_all = 1e9
    per = max(args.steps // 10, 100)
    for _ in range(10):
        t0 = time.perf_counter()
        for k in range(per):
            for p, g in zip(params, grads[k % 2]):
                p.grad = g
        t_set = min(t_set, (time.perf_counter() - t0) / per)
        t0 = time.perf_counter()
        for k in range(per):
            step(k)
        t_all = min(t_all, (time.perf_counter() - t0) / per)
    print('%d parameters: set grads %.1f us, update() %.1f us per step (host only)' % (
        len(params), t_set * 1e6, (t_all - t_set) * 1e6))


if __name__ == '__main__':
    main()
