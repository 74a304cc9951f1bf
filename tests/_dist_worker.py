"""Worker of tests/test_distributed_cpu.py: one rank of a world_size-N gloo
launch (python -m torch.distributed.run).  Runs the reference's multi-rank
checks against the product's host logic with the oracle-backed library double.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import chainer_b200  # noqa: E402
from chainer_b200 import config  # noqa: E402
from oracle import gradpath as og  # noqa: E402
from tests import fake_lib  # noqa: E402
from tests.test_host_logic import ExampleModel, _fill_grads  # noqa: E402


def main():
    fake, _ = fake_lib.install()
    comm = chainer_b200.create_communicator('pure_nccl')
    rank, size = comm.rank, comm.size
    assert size == int(os.environ['WORLD_SIZE']) and rank == int(os.environ['RANK'])
    assert comm.intra_size == size and comm.inter_size == 1 and comm.intra_rank == rank
    comm.bucket_bytes = 64          # force several buckets: 16 floats each... (>= 1024 elems min)

    # check_bcast_data (test_communicator.py:242-249)
    model = ExampleModel()
    model.a.W.data[...] = rank
    model.b.W.data[...] = rank + 1
    model.c.b.data[...] = rank + 2
    comm.bcast_data(model)
    np.testing.assert_array_equal(model.a.W.data, 0 * np.ones((3, 2), np.float32))
    np.testing.assert_array_equal(model.b.W.data, 1 * np.ones((4, 3), np.float32))
    np.testing.assert_array_equal(model.c.b.data, 2 * np.ones((5,), np.float32))

    # check_multi_node_mean_grad (test_communicator.py:252-269), twice
    for _ in range(2):
        _fill_grads(model, rank)
        comm.multi_node_mean_grad(model)
        base = (size - 1.0) / 2
        np.testing.assert_allclose(model.a.W.grad, (base + 0) * np.ones((3, 2)), rtol=1e-4, atol=1e-5)
        np.testing.assert_allclose(model.b.W.grad, (base + 1) * np.ones((4, 3)), rtol=1e-4, atol=1e-5)
        np.testing.assert_allclose(model.c.b.grad, (base + 2) * np.ones((5,)), rtol=1e-4, atol=1e-5)

    # check_multi_node_mean_grad_empty_half (test_communicator.py:289-319)
    _fill_grads(model, rank)
    if rank % 2 == 1:
        model.c.b.grad = None
    comm.multi_node_mean_grad(model, zero_fill=True)
    v = sum(i + 2 for i in range(size) if i % 2 == 0) / float(size)
    np.testing.assert_allclose(model.c.b.grad, v * np.ones((5,)), rtol=1e-4, atol=1e-5)

    # debug mode: shape agreement passes, NaN is detected on every rank
    config.set_debug(True)
    _fill_grads(model, rank)
    comm.multi_node_mean_grad(model)
    if rank == 0:
        model.a.W.grad[0, 0] = np.nan
    try:
        comm.multi_node_mean_grad(model)
        raise AssertionError('divergence not detected')
    except ValueError as e:
        assert 'diverged' in str(e)
    config.set_debug(False)

    # multi-node optimizer with a bigger model: several allreduce buckets, fused update
    rng = np.random.default_rng(7)                    # identical initial params on all ranks
    shapes = [(300, 40), (40,), (5000,), (64,), (3, 3, 3, 3), (2049,)]
    arrays = [('/p%d/W' % i, (rng.standard_normal(s) * 0.05).astype(np.float32))
              for i, s in enumerate(shapes)]
    from chainer_b200.core.link import link_from_named_arrays
    net = link_from_named_arrays(arrays)
    ref = [a.copy() for _, a in arrays]
    vs = [np.zeros_like(a) for a in ref]
    actual = chainer_b200.MomentumSGD(lr=0.01, momentum=0.9)
    opt = chainer_b200.create_multi_node_optimizer(actual, comm)
    opt.setup(net)
    comm.bucket_bytes = 8192                          # 2048 floats per bucket
    opt.update()                                      # broadcast
    for step in range(3):
        all_grads = []
        for r in range(size):
            grng = np.random.default_rng(1000 * (step + 1) + r)
            all_grads.append([(grng.standard_normal(s) * 1e-2).astype(np.float32) for s in shapes])
        for (name, p), g in zip(sorted(net.namedparams()), all_grads[rank]):
            p.grad = g.copy()
        fake.calls[:] = []
        opt.update()
        names = [c[0] for c in fake.calls]
        n_elems = sum(int(np.prod(s)) for s in shapes)
        assert names.count('gp_nccl_allreduce') == len(comm._bucket_bounds(n_elems, 4)) - 1 > 1
        assert names.count('gp_pack') == names.count('gp_nccl_allreduce')
        assert names.count('gp_unpack_momentum_sgd') == names.count('gp_nccl_allreduce')
        mean = og.multi_node_mean_grad(all_grads, np.float32)
        for (name, p), q, v, g in zip(sorted(net.namedparams()), ref, vs, mean):
            og.momentum_sgd_update(q, g, v, 0.01, 0.9)
            np.testing.assert_allclose(p.data, q, rtol=1e-6, atol=1e-8)
            np.testing.assert_allclose(p.grad, g, rtol=1e-6, atol=1e-9)
        assert actual.t == step + 1

    # ---- BASELINE configs[0]: MNIST MLP (784-1000-1000-10, 6 tensors, 1,796,010 elements),
    #      Adam defaults, gradients ~ N(0, 1e-2) from default_rng(1000 + rank) ----
    from chainer_b200 import workloads
    from chainer_b200.core.link import link_from_named_arrays as _from_arrays
    mlp = workloads.mnist_mlp()
    assert sum(int(np.prod(sh)) for _, sh in mlp) == 1796010 and len(mlp) == 6
    prng = np.random.default_rng(7)
    mlp_p = [(prng.standard_normal(sh) * 0.05).astype(np.float32) for _, sh in mlp]
    mlp_net = _from_arrays([(nm, a.copy()) for (nm, _), a in zip(mlp, mlp_p)])
    adam = chainer_b200.Adam()
    mlp_opt = chainer_b200.create_multi_node_optimizer(adam, comm)
    mlp_opt.setup(mlp_net)
    comm.bucket_bytes = 256 << 20
    mlp_opt.update()
    mlp_st = [dict(m=np.zeros_like(a), v=np.zeros_like(a)) for a in mlp_p]
    for step in range(1, 4):
        all_g = [[(np.random.default_rng(1000 + r + 10 * step).standard_normal(sh) * 1e-2)
                  .astype(np.float32) for _, sh in mlp] for r in range(size)]
        for (_, p), g in zip(sorted(mlp_net.namedparams()), all_g[rank]):
            p.grad = g.copy()
        mlp_opt.update()
        mean = og.multi_node_mean_grad(all_g, np.float32)
        for (_, p), q, g, st in zip(sorted(mlp_net.namedparams()), mlp_p, mean, mlp_st):
            og.adam_update_gpu(q, g, st['m'], st['v'], step)
            np.testing.assert_allclose(p.grad, g, rtol=1e-6, atol=1e-9)
            # 3 ranks: gloo's summation order differs from the oracle's rank order by an
            # ulp of the mean; Adam divides by sqrt(v) + eps, so the few elements with
            # |g| ~ eps amplify that (d step / d g ~ alpha / eps).  The mean is judged
            # strictly above, the step with the amplified bound.
            np.testing.assert_allclose(p.data, q, rtol=1e-6, atol=2e-7 if size == 2 else 2e-4)
    assert adam.t == 3

    # ---- fused GradientClipping + WeightDecay across ranks: the norm is that of the
    #      MEAN gradient, every rank derives the same rate ----
    from chainer_b200 import optimizer_hooks as H
    from chainer_b200.core.link import link_from_named_arrays
    shapes2 = [(300,), (17, 5), (2048,)]
    rng2 = np.random.default_rng(3)
    host_p = [(rng2.standard_normal(s) * 0.05).astype(np.float32) for s in shapes2]
    net2 = link_from_named_arrays([('/q%d' % i, a.copy()) for i, a in enumerate(host_p)])
    actual2 = chainer_b200.MomentumSGD(lr=0.01, momentum=0.9)
    opt2 = chainer_b200.create_multi_node_optimizer(actual2, comm)
    opt2.setup(net2)
    opt2.add_hook(H.GradientClipping(0.02))
    opt2.add_hook(H.WeightDecay(0.01))
    opt2.update()
    vs2 = [np.zeros_like(a) for a in host_p]
    for step in range(2):
        all_g = [[(np.random.default_rng(500 * (step + 1) + r).standard_normal(s) * 1e-2)
                  .astype(np.float32) for s in shapes2] for r in range(size)]
        for (_, p), g in zip(sorted(net2.namedparams()), all_g[rank]):
            p.grad = g.copy()
        fake.calls[:] = []
        opt2.update()
        names = [c[0] for c in fake.calls]
        assert names.count('gp_sqnorm') == 1 and 'gp_unpack_momentum_sgd_hooked' in names
        assert names.index('gp_sqnorm') > max(i for i, n in enumerate(names) if n == 'gp_nccl_allreduce')
        mean = og.multi_node_mean_grad(all_g, np.float32)
        rate = og.gradient_clipping_hook(mean, 0.02)
        assert rate < 1
        for (_, p), q, v, g in zip(sorted(net2.namedparams()), host_p, vs2, mean):
            og.weight_decay_hook(q, g, 0.01)
            og.momentum_sgd_update(q, g, v, 0.01, 0.9)
            np.testing.assert_allclose(p.data, q, rtol=2e-6, atol=1e-8)
            np.testing.assert_allclose(p.grad, g, rtol=2e-6, atol=1e-9)

    # ---- AllreducePersistent (chainermn/extensions/allreduce_persistent.py): float
    #      persistents become their mean over ranks, integers are left alone ----
    from chainer_b200.core import link as L
    from chainer_b200.extensions import AllreducePersistent

    class FakeBN(L.Link):
        def __init__(self, c, dtype):
            super(FakeBN, self).__init__()
            self.add_persistent('avg_mean', np.full((c,), rank + 1, dtype=dtype))
            self.add_persistent('avg_var', np.arange(c, dtype=dtype) * (rank + 1))
            self.add_persistent('N', 7 + rank)

    class Net(L.Chain):
        def __init__(self):
            super(Net, self).__init__()
            with self.init_scope():
                self.bn1 = FakeBN(5, np.float32)
                self.bn2 = FakeBN(3, np.float16)
                self.bn3 = FakeBN(4, np.float64)

    net3 = Net()
    fake.calls[:] = []
    AllreducePersistent(net3, comm)()
    names = [c[0] for c in fake.calls]
    assert names.count('gp_nccl_allreduce') == 2          # one per buffer dtype, not per array
    mean_rank = (size + 1) / 2.0
    for bn, c in ((net3.bn1, 5), (net3.bn2, 3), (net3.bn3, 4)):
        np.testing.assert_allclose(bn.avg_mean, np.full((c,), mean_rank), rtol=1e-3)
        np.testing.assert_allclose(bn.avg_var, np.arange(c) * mean_rank, rtol=1e-3)
        assert bn.N == 7 + rank
    assert net3.bn2.avg_mean.dtype == np.float16 and net3.bn3.avg_var.dtype == np.float64

    comm.finalize()
    print('RANK %d OK' % rank, flush=True)


if __name__ == '__main__':
    main()
