"""Worker of tests/test_multi_gpu.py: one rank (one GPU) of a torchrun launch
with the REAL library and NCCL.  Mirrors the reference's multi-rank tests
(test_communicator.py:242-319, test_multi_node_optimizer.py:51-110,
links_tests/test_batch_normalization.py:54-186)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def t(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def main():
    rank = int(os.environ['RANK'])
    world = int(os.environ['WORLD_SIZE'])
    torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', rank)))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    import chainer_b200
    from chainer_b200 import config
    from chainer_b200.core.link import link_from_named_arrays
    from oracle import gradpath as og

    comm = chainer_b200.create_communicator('pure_nccl')
    assert comm.size == world and comm.rank == rank and comm.intra_size == world

    # ---- peer-memory allreduce kernel vs the oracle (rank-order sums): bit-exact ----
    comm._init_comms()
    if os.environ.get('CHAINER_B200_P2P', '1') != '0':
        assert comm._p2p is not None, 'peer-memory allreduce should be active on one box'
        from chainer_b200.communicators._memory_utility import DeviceMemory
        from chainer_b200 import _lib
        lib = _lib.get()
        for dt_name, tdt, odt in (('float32', torch.float32, np.float32),
                                  ('float16', torch.float16, np.float16),
                                  ('bfloat16', torch.bfloat16, og.BF16),
                                  ('float64', torch.float64, np.float64)):
            for n in (1, 7, 4096, 100003, 3000000):
                isz = 8 if dt_name == 'float64' else (4 if dt_name == 'float32' else 2)
                mem = DeviceMemory()
                mem.assign(n * isz)
                comm._p2p.ensure(mem)
                parts = [og.cast(np.random.default_rng(50 + r).standard_normal(n) * (r + 1), odt)
                         for r in range(world)]
                mine = t(parts[rank]).to(tdt)
                lib.gp_memcpy_async(mem.ptr(), mine.data_ptr(), n * isz, 2, 0)
                comm._p2p.allreduce(odt, 0, n, None)
                out = torch.empty(n, dtype=tdt, device='cuda')
                lib.gp_memcpy_async(out.data_ptr(), mem.ptr(), n * isz, 2, 0)
                torch.cuda.synchronize()
                want = og.allreduce_sum(parts, odt)
                got = out.float().cpu().numpy() if dt_name == 'bfloat16' else out.cpu().numpy()
                assert np.array_equal(got.view(np.uint8), np.asarray(want).view(np.uint8)), \
                    ('p2p allreduce', dt_name, n)
        # ---- NVSwitch-multicast allreduce kernel (csrc/gp_mc.cu) vs the oracle ----
        if os.environ.get('CHAINER_B200_MULTICAST') == '1':
            p2p = comm._p2p
            if not p2p.multicast_supported:
                print('MULTICAST UNSUPPORTED: device attribute', flush=True)
            else:
                alloc = p2p.mc_allocate(3000000 * 4)
                if alloc is None:
                    print('MULTICAST UNSUPPORTED: %s' % p2p.multicast_error, flush=True)
                else:
                    for dt_name, tdt, odt in (('float32', torch.float32, np.float32),
                                              ('float16', torch.float16, np.float16),
                                              ('bfloat16', torch.bfloat16, og.BF16)):
                        isz = 4 if dt_name == 'float32' else 2
                        for n in (1, 7, 4096, 100003, 3000000):
                            parts = [og.cast(np.random.default_rng(50 + r).standard_normal(n) * (r + 1), odt)
                                     for r in range(world)]
                            mine = t(parts[rank]).to(tdt)
                            lib.gp_memcpy_async(alloc.ptr, mine.data_ptr(), n * isz, 2, 0)
                            p2p.mc_allreduce(odt, 0, n, None)
                            out = torch.empty(n, dtype=tdt, device='cuda')
                            lib.gp_memcpy_async(out.data_ptr(), alloc.ptr, n * isz, 2, 0)
                            torch.cuda.synchronize()
                            want = np.asarray(og.allreduce_sum(parts, odt)).astype(np.float64)
                            got = out.double().cpu().numpy()
                            mag = np.sum([np.abs(np.asarray(q, dtype=np.float64)) for q in parts], axis=0)
                            if dt_name == 'float32' and world == 2:
                                assert np.array_equal(got, want), ('mc allreduce', dt_name, n)
                            else:
                                # the switch's order of addition is not the oracle's rank
                                # order; 16-bit buffers accumulate in fp32 (one rounding)
                                eps = 2e-7 if dt_name == 'float32' else (1e-3 if dt_name == 'float16' else 8e-3)
                                assert np.all(np.abs(got - want) <= eps * world * mag + 1e-30), \
                                    ('mc allreduce', dt_name, n, float(np.max(np.abs(got - want))))
                            # every rank holds the same bits
                            mx = out.clone().view(torch.uint8).to(torch.int32).cpu()
                            mn = mx.clone()
                            dist.all_reduce(mx, op=dist.ReduceOp.MAX)
                            dist.all_reduce(mn, op=dist.ReduceOp.MIN)
                            assert bool((mx == mn).all()), ('mc allreduce replicas differ', dt_name, n)
                    print('MULTICAST KERNEL OK', flush=True)
        # one-shot small allreduce (MNBN statistics messages): bit-exact, incl. var
        for C in (1, 3, 64, 257, 2048, 4096):
            vals = [np.random.default_rng(90 + r).standard_normal(2 * C).astype(np.float32)
                    for r in range(world)]
            for use_var in (0, C):
                src = t(vals[rank])
                out = torch.empty(2 * C, device='cuda')
                comm._p2p.allreduce_small(src.data_ptr(), out.data_ptr(), 2 * C, use_var,
                                          1.0 / world, None)
                torch.cuda.synchronize()
                want = og.scale_buffer(og.allreduce_sum(vals, np.float32), np.float32, 1.0 / world)
                if use_var:
                    want = want.copy()
                    want[C:] = want[C:] - np.square(want[:C])
                assert np.array_equal(out.cpu().numpy().view(np.uint32), want.view(np.uint32)), \
                    ('small allreduce', C, use_var)
        # latency of one MNBN statistics exchange: NCCL allreduce + scale vs one-shot kernel
        from chainer_b200 import nccl as _nccl
        C = 256
        src = torch.randn(2 * C, device='cuda')
        out = torch.empty(2 * C, device='cuda')

        def lat(fn, reps=200):
            for _ in range(10):
                fn()
            torch.cuda.synchronize()
            dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                fn()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) * 1e3 / reps

        def via_nccl():
            comm.nccl_comm.allReduce(src.data_ptr(), out.data_ptr(), 2 * C, 7, _nccl.NCCL_SUM, 0)
            lib.gp_scale(out.data_ptr(), 7, 2 * C, 1.0 / world, 0)
            lib.gp_bn_finish_mean_var(out.data_ptr(), 7, C, 1.0, out.data_ptr() + 4 * C, 0)

        def via_p2p():
            comm._p2p.allreduce_small(src.data_ptr(), out.data_ptr(), 2 * C, C, 1.0 / world, None)
        a_us, b_us = lat(via_nccl), lat(via_p2p)
        if rank == 0:
            print('MNBN statistics exchange (2C = %d floats, %d ranks): NCCL allreduce + scale + var '
                  '%.1f us, one-shot peer-memory kernel %.1f us' % (2 * C, world, a_us, b_us), flush=True)
        dist.barrier()

    # ---- bcast_data + mean_grad analytic vectors -----------------------------
    plist = [('/a/W', (3, 2)), ('/a/b', (3,)), ('/b/W', (4, 3)), ('/b/b', (4,)), ('/c/b', (5,))]
    fills = [0, 0, 1, 1, 2]
    model = link_from_named_arrays([(n, t(np.full(s, rank + k, np.float32)))
                                    for (n, s), k in zip(plist, fills)])
    comm.bcast_data(model)
    for (_, p), k in zip(sorted(model.namedparams()), fills):
        assert torch.equal(p.data, torch.full_like(p.data, k))
    for adt in (None, np.float16, 'bfloat16'):
        comm.set_config('allreduce_grad_dtype', adt)
        for _ in range(2):
            for (_, p), k in zip(sorted(model.namedparams()), fills):
                p.grad = torch.full_like(p.data, rank + k)
            comm.multi_node_mean_grad(model)
            base = (world - 1.0) / 2
            for (_, p), k in zip(sorted(model.namedparams()), fills):
                np.testing.assert_allclose(p.grad.cpu().numpy(), base + k, rtol=1e-2 if adt else 1e-6)
    comm.set_config('allreduce_grad_dtype', None)
    # zero_fill on odd ranks (test_communicator.py:289-319)
    for (_, p), k in zip(sorted(model.namedparams()), fills):
        p.grad = torch.full_like(p.data, rank + k)
    last = sorted(model.namedparams())[-1][1]
    if rank % 2 == 1:
        last.grad = None
    comm.multi_node_mean_grad(model, zero_fill=True)
    v = sum(i + 2 for i in range(world) if i % 2 == 0) / float(world)
    np.testing.assert_allclose(last.grad.cpu().numpy(), v, rtol=1e-6)

    # ---- fused multi-node optimizer vs oracle, bucketed -----------------------
    from chainer_b200 import workloads
    wl = workloads.scaled_histogram(600000)
    # every configuration runs twice: the separate launches (pack, allreduce, fused update:
    # use_step False) and the one-launch step (csrc/gp_step.cu; small tiles so that the
    # 0.6 M elements are ~150 tiles spread over reducers and workers)
    from chainer_b200 import _lib as _l
    _l.get().gp_step_set_tuning(b'tile_elems', 4096)
    step_launches = 0
    # (the element count changes between the configurations -- fewer tiles, more tiles, a
    # larger buffer: the step's per-tile words must not depend on history)
    for opt_name, adt, tol, use_step, n_target in (
            ('momentum_sgd', None, 1e-6, False, 600000), ('adam', None, 1e-6, False, 600000),
            ('momentum_sgd', np.float16, 2e-3, False, 600000),
            ('momentum_sgd', None, 1e-6, True, 600000), ('adam', None, 1e-6, True, 20000),
            ('momentum_sgd', np.float16, 2e-3, True, 900000),
            ('adam', 'bfloat16', 1.6e-2, True, 150000)):
        wl = workloads.scaled_histogram(n_target)
        comm.use_step = use_step
        comm.set_config('allreduce_grad_dtype', adt)
        comm.bucket_bytes = 256 << 10                 # several buckets
        # Adam runs: several pipelined chunks on the peer-memory / multicast paths too
        comm.p2p_chunk_bytes = (256 << 10) if opt_name == 'adam' else 0
        comm.mc_chunk_bytes = (256 << 10) if opt_name == 'adam' else None
        rng = np.random.default_rng(7)
        host_p = [(rng.standard_normal(s) * 0.05).astype(np.float32) for _, s in wl]
        m = link_from_named_arrays([(n, t(a + rank)) for (n, _), a in zip(wl, host_p)])
        actual = chainer_b200.MomentumSGD(lr=0.01) if opt_name == 'momentum_sgd' else chainer_b200.Adam()
        opt = chainer_b200.create_multi_node_optimizer(actual, comm)
        opt.setup(m)
        opt.update()                                  # broadcast: rank 0's values everywhere
        st = [dict(m=np.zeros_like(a), v=np.zeros_like(a)) for a in host_p]
        for step in range(1, 4):
            all_g = [[(np.random.default_rng(1000 * step + r).standard_normal(a.shape) * 1e-2)
                      .astype(np.float32) for a in host_p] for r in range(world)]
            for (_, p), g in zip(sorted(m.namedparams()), all_g[rank]):
                p.grad = t(g)
            calls0 = _l.get().launches
            opt.update()
            torch.cuda.synchronize()
            if use_step and comm._p2p is not None:
                assert _l.get().launches - calls0 == 1, 'the step should be ONE launch'
                step_launches += 1
            mean = og.multi_node_mean_grad(all_g, np.float32 if adt is None
                                           else (og.BF16 if adt == 'bfloat16' else adt))
            exact = adt is None and (world == 2 or (comm._p2p is not None
                                                    and not comm._mc_active(comm.gpu_buffer_a)))
            for i, ((name, p), q, g, s) in enumerate(zip(sorted(m.namedparams()), host_p, mean, st)):
                if opt_name == 'momentum_sgd':
                    og.momentum_sgd_update(q, g, s['v'], 0.01, 0.9)
                else:
                    og.adam_update_gpu(q, g, s['m'], s['v'], step)
                got = p.data.cpu().numpy()
                got_g = p.grad.cpu().numpy()
                if exact:
                    # 2-term sums are order-free; the peer-memory kernel adds in rank
                    # order like the oracle: bit-exact for every world size
                    assert np.array_equal(got, q), (opt_name, name, step)
                    np.testing.assert_allclose(got_g, g, rtol=tol, atol=2e-8)
                    continue
                # NCCL and the NVSwitch add in their own order: the mean differs from
                # the rank-order oracle by the rounding of the partial sums, bounded by
                # (N-1) * eps * sum_r |g_r| / N  (eps of the allreduce dtype; every
                # partial sum of a 16-bit ring is rounded to 16 bits)
                gmag = np.sum([np.abs(all_g[r][i]) for r in range(world)], axis=0) / world
                eps = 1.2e-7 if adt is None else (8e-3 if adt == 'bfloat16' else 1e-3)
                # (absolute floor: half the spacing of float16 subnormals per addition)
                gerr = world * eps * gmag + (1e-12 if adt is None else world * 6e-8)
                bad = np.abs(got_g - g) > gerr
                assert not bad.any(), (opt_name, name, step, 'grad', float(np.abs(got_g - g).max()))
                if opt_name == 'momentum_sgd':
                    # v = 0.9 v - lr g; p += v: the velocity inherits 0.9 of its old
                    # deviation plus lr * gerr of THIS step's gradients, the parameter
                    # accumulates the velocity deviations of all steps (+ its rounding)
                    s['verr'] = 0.9 * s.get('verr', 0.0) + 0.01 * gerr
                    s['perr'] = s.get('perr', 0.0) + s['verr']
                    perr = 1.5 * s['perr'] + 4e-7 * np.abs(q) + 1e-9
                    assert not (np.abs(got - q) > perr).any(), \
                        (opt_name, name, step, float(np.abs(got - q).max()))
                else:
                    # Adam divides by sqrt(v) + eps, so elements with |g| ~ eps amplify
                    # the difference (d step / d g ~ alpha / eps): allow the amplified
                    # step error; the mean gradient is judged strictly above
                    # (a 16-bit mean can flip the sign of a gradient within its rounding
                    # error: such an element moves by up to 2 * alpha per step)
                    np.testing.assert_allclose(got, q, rtol=tol, atol=2e-4 if adt is None else 8e-3)
            assert actual.t == step
        # every rank holds identical parameters
        flat = torch.cat([p.data.reshape(-1) for _, p in sorted(m.namedparams())])
        ref = flat.clone().cpu()
        parts = [torch.zeros_like(ref) for _ in range(world)]
        dist.all_gather(parts, ref)
        for prt in parts:
            assert torch.equal(prt, ref)
    comm.set_config('allreduce_grad_dtype', None)
    comm.use_step = None
    if comm._p2p is not None:
        assert step_launches == 12
        print('ONE-LAUNCH STEP OK (%s transport)' % (
            'multicast' if comm._mc_active(comm.gpu_buffer_a) else 'peer-memory'), flush=True)

    # ---- debug mode across ranks ---------------------------------------------
    config.set_debug(True)
    for (_, p), k in zip(sorted(model.namedparams()), fills):
        p.grad = torch.full_like(p.data, rank + k)
    comm.multi_node_mean_grad(model)
    if rank == 0:
        sorted(model.namedparams())[0][1].grad[0, 0] = float('nan')
    try:
        comm.multi_node_mean_grad(model)
        raise AssertionError('divergence not detected')
    except ValueError as e:
        assert 'diverged' in str(e)
    config.set_debug(False)

    # ---- MNBN: multi-worker == single worker on the global batch ---------------
    from chainer_b200.links import MultiNodeBatchNormalization
    z = np.load(os.path.join(ROOT, 'tests', 'golden', 'mnbn.npz'))
    if world == 2:
        nb = 4
        x = t(z['x'][rank * nb:(rank + 1) * nb]).requires_grad_(True)
        C = z['gamma'].size
        bn = MultiNodeBatchNormalization(C, comm)
        bn.gamma.data.copy_(torch.from_numpy(z['gamma']))
        bn.beta.data.copy_(torch.from_numpy(z['beta']))
        bn.gamma.data.requires_grad_(True)
        bn.beta.data.requires_grad_(True)
        y = bn(x)
        y.backward(t(z['gy'][rank * nb:(rank + 1) * nb]))
        np.testing.assert_allclose(y.detach().cpu().numpy(), z['mn|%d|y' % rank], rtol=1e-4, atol=1e-5)
        np.testing.assert_allclose(x.grad.cpu().numpy(), z['mn|%d|gx' % rank], rtol=1e-3, atol=1e-5)
        np.testing.assert_allclose(bn.gamma.data.grad.cpu().numpy(), z['mn|%d|ggamma' % rank],
                                   rtol=1e-4, atol=1e-5)
        np.testing.assert_allclose(bn.beta.data.grad.cpu().numpy(), z['mn|%d|gbeta' % rank],
                                   rtol=1e-4, atol=1e-5)
        np.testing.assert_allclose(bn.avg_mean.cpu().numpy(), z['mn|%d|avg_mean' % rank],
                                   rtol=1e-4, atol=1e-6)
        np.testing.assert_allclose(bn.avg_var.cpu().numpy(), z['mn|%d|avg_var' % rank],
                                   rtol=1e-4, atol=1e-6)
    # ---- the statistics + exchange kernel replays from a CUDA graph (device-side epochs) ----
    if comm._p2p is not None:
        from chainer_b200.functions.batch_normalization import _NcclImpl
        impl = _NcclImpl(comm)
        gen = torch.Generator(device='cuda')
        gen.manual_seed(300 + rank)
        xg = torch.randn(4, 48, 6, 6, device='cuda', generator=gen)
        gam = torch.ones(48, device='cuda')
        m_e, v_e = impl.get_mean_and_var(None, gam, xg)                 # eager, stream 0
        torch.cuda.synchronize()
        side = torch.cuda.Stream()
        impl.stream = side.cuda_stream
        with torch.cuda.stream(side):
            impl.get_mean_and_var(None, gam, xg)
        side.synchronize()
        dist.barrier()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=side):
            m_g, v_g = impl.get_mean_and_var(None, gam, xg)
        for _ in range(3):
            graph.replay()
        torch.cuda.synchronize()
        assert torch.equal(m_g, m_e) and torch.equal(v_g, v_e), 'graph replay of the MNBN exchange'
        impl.stream = 0
        impl.get_mean_and_var(None, gam, xg)                             # and eager again after it
        torch.cuda.synchronize()
        dist.barrier()

    # ---- layers of different width share the communicator's ONE statistics workspace:
    #      every fused statistics + exchange call stays right (and the ranks stay in step)
    if comm._p2p is not None:
        from chainer_b200.functions.batch_normalization import _NcclImpl
        impl = _NcclImpl(comm)
        gen = torch.Generator(device='cuda')
        gen.manual_seed(900 + rank)
        shapes = [(32, 64, 56, 56), (32, 128, 28, 28), (32, 256, 14, 14), (32, 512, 28, 28),
                  (32, 1024, 14, 14), (32, 2048, 7, 7)]
        data = {}
        for s in shapes:
            x = torch.randn(*s, device='cuda', generator=gen) + 0.1 * rank
            gy = torch.randn(*s, device='cuda', generator=gen) * 1e-3
            st = torch.stack([x.double().mean(dim=(0, 2, 3)),
                              (x.double() ** 2).mean(dim=(0, 2, 3))]).cpu()
            dist.all_reduce(st)                     # gloo: host tensors
            st = (st / world).cuda()
            m, v = st[0], st[1] - st[0] * st[0]
            inv = torch.rsqrt(v + 2e-5)
            bw = torch.stack([gy.double().sum(dim=(0, 2, 3)),
                              (gy.double() * (x.double() - m[None, :, None, None])
                               * inv[None, :, None, None]).sum(dim=(0, 2, 3))]).cpu()
            dist.all_reduce(bw)
            bw = (bw / world).cuda()
            data[s] = (x, gy, torch.ones(s[1], device='cuda'), m, v, inv, bw)
        impl.get_mean_and_var(None, data[shapes[-1]][2], data[shapes[-1]][0])   # final size
        for order in (shapes, shapes[::-1], shapes):
            for s in order:
                x, gy, gamma, m, v, inv, bw = data[s]
                mean, var = impl.get_mean_and_var(None, gamma, x)
                torch.testing.assert_close(mean.double(), m, rtol=1e-5, atol=2e-6, msg=str(s))
                torch.testing.assert_close(var.double(), v, rtol=1e-5, atol=2e-6, msg=str(s))
                gbeta, ggamma = impl.get_ggamma_and_gbeta_from_x(None, gamma, gy, x, m.float(),
                                                                 inv.float())
                torch.testing.assert_close(gbeta.double(), bw[0], rtol=1e-4, atol=1e-5, msg=str(s))
                torch.testing.assert_close(ggamma.double(), bw[1], rtol=1e-4, atol=1e-5, msg=str(s))
                # every rank holds the same bits
                both = torch.stack([mean, var]).cpu()
                lo, hi = both.clone(), both.clone()
                dist.all_reduce(lo, op=dist.ReduceOp.MIN)
                dist.all_reduce(hi, op=dist.ReduceOp.MAX)
                assert torch.equal(lo, hi), 'MNBN statistics differ between ranks %s' % (s,)
        torch.cuda.synchronize()
        dist.barrier()
        if rank == 0:
            print('MNBN MIXED-WIDTH WORKSPACE OK')

    # ---- AllreducePersistent: running statistics become their mean over the ranks ----
    from chainer_b200.core import link as L
    from chainer_b200.extensions import AllreducePersistent

    class _Net(L.Chain):
        def __init__(self):
            super(_Net, self).__init__()
            with self.init_scope():
                self.bn1 = MultiNodeBatchNormalization(64, comm)
                self.bn2 = MultiNodeBatchNormalization(5, comm)

    net = _Net()
    net.bn1.avg_mean.fill_(float(rank))
    net.bn1.avg_var.copy_(torch.arange(64, device='cuda', dtype=torch.float32) * (rank + 1))
    net.bn2.avg_mean.fill_(1.5)
    net.bn1.N = 3 + rank
    AllreducePersistent(net, comm)()
    torch.cuda.synchronize()
    mr = (world - 1) / 2.0
    np.testing.assert_allclose(net.bn1.avg_mean.cpu().numpy(), np.full(64, mr), rtol=1e-6)
    np.testing.assert_allclose(net.bn1.avg_var.cpu().numpy(), np.arange(64) * (mr + 1), rtol=1e-6)
    np.testing.assert_allclose(net.bn2.avg_mean.cpu().numpy(), np.full(5, 1.5), rtol=1e-6)
    assert net.bn1.N == 3 + rank

    if os.environ.get('CHAINER_B200_MULTICAST') == '1' and comm._p2p is not None \
            and comm._p2p.multicast_supported:
        assert comm._mc_active(comm.gpu_buffer_a), 'the public API should have used the multicast path'
        print('MULTICAST PATH OK', flush=True)
    comm.finalize()
    print('GPU RANK %d OK' % rank, flush=True)


if __name__ == '__main__':
    main()
