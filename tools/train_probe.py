"""Where the host time of one training step goes (torchvision resnet50 fwd/bwd + the
gradient path with FRESH gradient arrays every step, as after cleargrads()).

    python tools/train_probe.py
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import numpy as np
    import torch
    import torchvision
    import chainer_b200
    from chainer_b200.core.link import link_from_named_arrays
    torch.manual_seed(7)
    net = torchvision.models.resnet50(weights=None).cuda()
    net.train()
    x = torch.randn(32, 3, 224, 224, device='cuda').to(memory_format=torch.channels_last)
    y = torch.randint(0, 1000, (32,), device='cuda')
    named = [(nm.replace('.', '/'), p) for nm, p in net.named_parameters()]
    comm = chainer_b200.create_communicator('pure_nccl')
    model = link_from_named_arrays([('/' + nm, p.data) for nm, p in named])
    plink = [p for _, p in sorted(model.namedparams())]
    tparam = [p for _, p in sorted((('/' + nm), p) for nm, p in named)]
    opt = chainer_b200.create_multi_node_optimizer(chainer_b200.MomentumSGD(lr=0.01), comm)
    opt.setup(model)
    T = {k: [] for k in ('clear', 'fwd_bwd_enqueue', 'assign', 'update_enqueue', 'sync', 'total')}

    def one(rec):
        t0 = time.perf_counter()
        for tp in tparam:
            tp.grad = None
        t1 = time.perf_counter()
        with torch.autocast('cuda', dtype=torch.bfloat16):
            loss = torch.nn.functional.cross_entropy(net(x), y)
        loss.backward()
        t2 = time.perf_counter()
        for lp, tp in zip(plink, tparam):
            lp.grad = tp.grad
        t3 = time.perf_counter()
        opt.update()
        t4 = time.perf_counter()
        torch.cuda.synchronize()
        t5 = time.perf_counter()
        if rec:
            for k, v in zip(('clear', 'fwd_bwd_enqueue', 'assign', 'update_enqueue', 'sync', 'total'),
                            (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4, t5 - t0)):
                T[k].append(v)
    for _ in range(8):
        one(False)
    for _ in range(30):
        one(True)
    for k, v in T.items():
        print('%-16s median %8.1f us   max %8.1f us' % (k, 1e6 * np.median(v), 1e6 * max(v)))
    # without the per-step synchronize: what the bench's train leg times
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for upd in (False, True):
        for _ in range(5):
            for tp in tparam:
                tp.grad = None
            with torch.autocast('cuda', dtype=torch.bfloat16):
                loss = torch.nn.functional.cross_entropy(net(x), y)
            loss.backward()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e0.record()
        for _ in range(20):
            for tp in tparam:
                tp.grad = None
            with torch.autocast('cuda', dtype=torch.bfloat16):
                loss = torch.nn.functional.cross_entropy(net(x), y)
            loss.backward()
            if upd:
                for lp, tp in zip(plink, tparam):
                    lp.grad = tp.grad
                opt.update()
        e1.record()
        t_enq = time.perf_counter() - t0
        torch.cuda.synchronize()
        print('update=%s: device %.3f ms/step, host enqueue %.3f ms/step' % (
            upd, e0.elapsed_time(e1) / 20, 1e3 * t_enq / 20))
    comm.finalize()


if __name__ == '__main__':
    main()
