"""A small end-to-end run of every collective kernel for compute-sanitizer (memcheck /
racecheck): the one-launch step (peer-memory and, with CHAINER_B200_MULTICAST=1, multicast
transport), the separate pack / allreduce / update launches, the MNBN statistics with the
in-kernel exchange and the BN apply kernels, on tiny tensors.  1 rank or N ranks (torchrun).

    python -m torch.distributed.run --no-python --nproc-per-node 2 --master-addr 127.0.0.1 \
        compute-sanitizer --tool memcheck --error-exitcode 1 python tools/sanitize_run.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', rank)))
    os.environ.setdefault('CHAINER_B200_PEER_TIMEOUT_S', '240')      # the sanitizer is slow
    if world > 1:
        dist.init_process_group('gloo', rank=rank, world_size=world)
    import chainer_b200
    from chainer_b200 import _lib, workloads
    from chainer_b200.core.link import link_from_named_arrays
    from chainer_b200.links import MultiNodeBatchNormalization
    lib = _lib.get()
    lib.gp_step_set_tuning(b'tile_elems', 4096)
    comm = chainer_b200.create_communicator('pure_nccl')
    wl = workloads.scaled_histogram(40000)
    for use_step in (True, False):
        for opt_name in ('momentum_sgd', 'adam'):
            comm.use_step = use_step
            rng = np.random.default_rng(7)
            model = link_from_named_arrays(
                [(n, torch.from_numpy((rng.standard_normal(s) * 0.05).astype(np.float32)).cuda())
                 for n, s in wl])
            actual = chainer_b200.MomentumSGD(lr=0.01) if opt_name == 'momentum_sgd' \
                else chainer_b200.Adam()
            opt = chainer_b200.create_multi_node_optimizer(actual, comm)
            opt.setup(model)
            for step in range(3):
                for _, p in sorted(model.namedparams()):
                    p.grad = torch.randn_like(p.data) * 1e-2
                opt.update()
            torch.cuda.synchronize()
    bn = MultiNodeBatchNormalization(16, comm)
    x = torch.randn(4, 16, 6, 6, device='cuda', requires_grad=True)
    y = bn(x)
    y.backward(torch.randn_like(y))
    torch.cuda.synchronize()
    transport = 'single rank' if world == 1 else (
        'multicast' if comm._mc_active(comm.gpu_buffer_a) else 'peer memory')
    comm.finalize()
    print('SANITIZE RANK %d DONE (transport: %s)' % (rank, transport), flush=True)


if __name__ == '__main__':
    main()
