"""Sweep of the one-launch step's tuning knobs inside ONE process group (run under torchrun).

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/step_sweep.py \
        [--configs "reducers=64;reducers=128,tile_elems=32768;..."] [--multicast on|off|auto]

Every configuration: 5 warm-up + 30 timed steps of the ResNet-50 workload (fp32, MomentumSGD)
through create_multi_node_optimizer(...).update(); device time, max over ranks.
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--configs', default='')
    ap.add_argument('--multicast', default='auto')
    ap.add_argument('--allreduce-dtype', default='float32')
    ap.add_argument('--workload', default='resnet50')
    ap.add_argument('--optimizer', default='momentum_sgd')
    ap.add_argument('--steps', type=int, default=30)
    ap.add_argument('--out', default='')
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', rank)))
    os.environ.setdefault('CHAINER_B200_PEER_TIMEOUT_S', '60')
    if world > 1:
        dist.init_process_group('gloo', rank=rank, world_size=world)
    import chainer_b200
    from chainer_b200 import _lib, workloads
    from chainer_b200.core.link import link_from_named_arrays
    lib = _lib.get()
    adt = {'float32': np.float32, 'float16': np.float16, 'bfloat16': 'bfloat16'}[args.allreduce_dtype]
    comm = chainer_b200.create_communicator('pure_nccl', allreduce_grad_dtype=adt)
    if args.multicast != 'auto':
        comm.use_multicast = args.multicast == 'on'
    plist = workloads.WORKLOADS[args.workload]()
    sizes = [int(np.prod(s)) for _, s in plist]
    n = sum(sizes)
    gen = torch.Generator(device='cuda')
    gen.manual_seed(7)
    p_arena = torch.randn(n, device='cuda', generator=gen) * 0.05
    gen.manual_seed(1000 + rank)
    g_arenas = [torch.randn(n, device='cuda', generator=gen) * 1e-2 for _ in range(2)]
    offs = np.concatenate([[0], np.cumsum(sizes)])
    views = lambda a: [a[offs[i]:offs[i + 1]] for i in range(len(sizes))]  # noqa: E731
    model = link_from_named_arrays([(nm, v) for (nm, _), v in zip(plist, views(p_arena))])
    params = [p for _, p in sorted(model.namedparams())]
    gv = [views(a) for a in g_arenas]
    actual = chainer_b200.MomentumSGD(lr=0.01, momentum=0.9) if args.optimizer == 'momentum_sgd' \
        else chainer_b200.Adam()
    opt = chainer_b200.create_multi_node_optimizer(actual, comm)
    opt.setup(model)

    def step(k):
        for p, g in zip(params, gv[k % 2]):
            p.grad = g
        opt.update()
    step(0)
    rows = []
    defaults = dict(tile_elems=16384, reducers=0, unroll=8, ctas_per_sm=4)
    configs = [c for c in args.configs.split(';') if c] or ['']
    for cfg in configs:
        kv = dict(defaults)
        use_step = True
        for item in [x for x in cfg.split(',') if x]:
            k, v = item.split('=')
            if k == 'step':
                use_step = v != '0'
            else:
                kv[k] = int(v)
        for k, v in kv.items():
            lib.gp_step_set_tuning(k.encode(), v)
        comm.use_step = use_step
        for k in range(5):
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
        ms = e0.elapsed_time(e1) / args.steps
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        rows.append({'config': cfg or 'default', 'ms_per_step': ms})
        if rank == 0:
            print('%-60s %.4f ms/step' % (cfg or 'default', ms), flush=True)
    if rank == 0 and args.out:
        json.dump({'n_gpus': world, 'workload': args.workload, 'allreduce_dtype': args.allreduce_dtype,
                   'multicast': args.multicast, 'rows': rows}, open(args.out, 'w'), indent=1)
    comm.finalize()


if __name__ == '__main__':
    main()
