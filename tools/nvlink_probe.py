"""Do the NVSwitch-multicast reduction and peer-memory traffic share one NVLink ceiling?
Times gp_mc_allreduce (multicast buffer) and gp_p2p_allreduce (IPC buffer) alone and
CONCURRENTLY on two streams (run under torchrun, N >= 2).
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    rank = int(os.environ['RANK'])
    world = int(os.environ['WORLD_SIZE'])
    torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', rank)))
    os.environ.setdefault('CHAINER_B200_PEER_TIMEOUT_S', '60')
    dist.init_process_group('gloo', rank=rank, world_size=world)
    import chainer_b200
    from chainer_b200 import _lib
    from chainer_b200 import device as dev
    from chainer_b200.communicators._memory_utility import DeviceMemory
    lib = _lib.get()
    comm = chainer_b200.create_communicator('pure_nccl')
    comm._init_comms()
    p2p = comm._p2p
    # a second handle with its OWN flag words: the two kernels run concurrently and must
    # not share the in-kernel barrier state
    from chainer_b200.communicators import _p2p as _p2p_mod
    p2p_b = _p2p_mod.PeerAllreduce(comm.mpi_comm)
    n = 25557096
    mem = DeviceMemory()
    mem.assign(n * 4)
    p2p_b.ensure(mem)
    lib.gp_memset_async(mem.ptr(), 0, n * 4, 0)
    alloc = p2p.mc_allocate(n * 4)
    assert alloc is not None, p2p.multicast_error
    s1, s2 = dev.Stream(non_blocking=True), dev.Stream(non_blocking=True)

    def timeit(fns, reps=10):
        """fns: list of (callable(stream), stream); all enqueued per rep; returns us per rep"""
        for _ in range(3):
            for f, s in fns:
                f(s)
        torch.cuda.synchronize()
        dist.barrier()
        evs = []
        for f, s in fns:
            e0, e1 = dev.Event(timing=True), dev.Event(timing=True)
            evs.append((e0, e1, s))
        for e0, _, s in evs:
            e0.record(s)
        for _ in range(reps):
            for f, s in fns:
                f(s)
        for _, e1, s in evs:
            e1.record(s)
        torch.cuda.synchronize()
        out = []
        for e0, e1, _ in evs:
            t = torch.tensor([e0.elapsed_ms(e1) * 1e3 / reps], dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            out.append(float(t.item()))
        return out

    def mc(s):
        p2p.mc_allreduce(np.float32, 0, n, s)

    def pp(s):
        p2p_b.allreduce(np.float32, 0, n, s)

    def mc_half(s):
        p2p.mc_allreduce(np.float32, 0, n // 2 // 4096 * 4096, s)

    def pp_half(s):
        p2p_b.allreduce(np.float32, 0, n // 2 // 4096 * 4096, s)
    S = n * 4
    res = {}
    res['mc alone'] = timeit([(mc, s1)])
    res['p2p alone'] = timeit([(pp, s2)])
    res['mc + p2p concurrent (full buffers each)'] = timeit([(mc, s1), (pp, s2)])
    res['mc half alone'] = timeit([(mc_half, s1)])
    res['p2p half alone'] = timeit([(pp_half, s2)])
    res['mc half + p2p half concurrent'] = timeit([(mc_half, s1), (pp_half, s2)])
    for ctas in (16, 32, 64, 128):
        lib.gp_mc_set_tuning(ctas, 256, 4)
        res['mc alone ctas=%d' % ctas] = timeit([(mc, s1)])
    lib.gp_mc_set_tuning(32, 256, 4)
    if rank == 0:
        mcw = S * (world + 1.0) / world
        ppw = S * 2.0 * (world - 1) / world
        for k, v in res.items():
            print('%-45s %s us' % (k, ' / '.join('%.1f' % x for x in v)))
        print('wire bytes per direction: mc %.1f MB, p2p %.1f MB' % (mcw / 1e6, ppw / 1e6))
        a, b = res['mc + p2p concurrent (full buffers each)']
        print('concurrent: (%.1f + %.1f MB) / max(%.1f, %.1f us) = %.0f GB/s on the wire' % (
            mcw / 1e6, ppw / 1e6, a, b, (mcw + ppw) / max(a, b) / 1e3))
        print('alone: mc %.0f GB/s, p2p %.0f GB/s' % (mcw / res['mc alone'][0] / 1e3, ppw / res['p2p alone'][0] / 1e3))
    p2p.mc_release()
    p2p_b.destroy()
    comm.finalize()


if __name__ == '__main__':
    main()
