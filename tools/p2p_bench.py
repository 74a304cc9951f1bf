"""Microbenchmark of the peer-memory allreduce kernel vs NCCL (run under torchrun).

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/p2p_bench.py
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
    dist.init_process_group('gloo', rank=rank, world_size=world)
    import chainer_b200
    from chainer_b200 import _lib, nccl
    from chainer_b200.communicators._memory_utility import DeviceMemory
    lib = _lib.get()
    comm = chainer_b200.create_communicator('pure_nccl')
    comm._init_comms()
    sizes = [int(x) for x in os.environ.get('P2P_SIZES', '25557096,173300800,1048576,65536').split(',')]
    for n in sizes:
        mem = DeviceMemory()
        mem.assign(n * 4)
        comm._p2p.ensure(mem)
        lib.gp_memset_async(mem.ptr(), 0, n * 4, 0)

        def timeit(fn, reps=20):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                fn()
            e1.record()
            torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1) * 1e3 / reps], dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())

        us_nccl = timeit(lambda: comm.nccl_comm.allReduce(mem.ptr(), mem.ptr(), n, 7, nccl.NCCL_SUM, 0))
        # NCCL with a symmetric-registered buffer (ncclMemAlloc + ncclCommWindowRegister)
        us_sym = None
        try:
            if os.environ.get('P2P_NCCL_SYMMETRIC') != '1':
                raise RuntimeError('skipped (set P2P_NCCL_SYMMETRIC=1)')
            import ctypes
            sp = ctypes.c_void_p()
            nb = (n * 4 + (2 << 20) - 1) // (2 << 20) * (2 << 20)
            lib.gp_nccl_mem_alloc(ctypes.byref(sp), nb)
            win = ctypes.c_void_p()
            lib.gp_nccl_comm_window_register(comm.nccl_comm.handle, sp.value, nb, ctypes.byref(win), 1)
            lib.gp_memset_async(sp.value, 0, n * 4, 0)
            us_sym = timeit(lambda: comm.nccl_comm.allReduce(sp.value, sp.value, n, 7, nccl.NCCL_SUM, 0))
            lib.gp_nccl_comm_window_deregister(comm.nccl_comm.handle, win.value)
            lib.gp_nccl_mem_free(sp.value)
        except Exception as e:
            if rank == 0:
                print('   symmetric NCCL experiment failed: %s' % e)
        rows = []
        for mode in (1,):
            for threads in (512,):
                for ctas in (148, 296):
                    lib.gp_p2p_set_tuning(ctas, threads, mode)
                    us = timeit(lambda: comm._p2p.allreduce(np.float32, 0, n, None))
                    rows.append((us, mode, threads, ctas))
        mc_rows = []
        if comm._p2p.multicast_supported and os.environ.get('P2P_MULTICAST', '1') == '1':
            alloc = comm._p2p.mc_allocate(n * 4)
            if alloc is None:
                if rank == 0:
                    print('   multicast unavailable: %s' % comm._p2p.multicast_error)
            else:
                for dt, dname in ((np.float32, 'f32'), (np.float16, 'f16')):
                    ne = n if dt is np.float32 else 2 * n
                    for threads in (256, 512):
                        for ctas in (32, 74, 148, 296):
                            for unroll in (4, 8):
                                if dname == 'f16' and (threads, unroll) != (512, 4):
                                    continue
                                lib.gp_mc_set_tuning(ctas, threads, unroll)
                                us = timeit(lambda: comm._p2p.mc_allreduce(dt, 0, ne, None))
                                mc_rows.append((dname, us, threads, ctas, unroll))
                lib.gp_mc_set_tuning(0, 512, 4)
                comm._p2p.mc_release()
        if rank == 0:
            S = n * 4
            f = 2.0 * (world - 1) / world
            for dname in ('f32', 'f16'):
                sel = sorted(r for r in mc_rows if r[0] == dname)
                for _, us, threads, ctas, unroll in sel[:4] + sel[-1:]:
                    print('   multicast %s t%d c%d u%d: %.1f us  busBW %.0f GB/s  (x%.2f vs NCCL f32)' % (
                        dname, threads, ctas, unroll, us, S / us / 1e3 * f, us_nccl / us))
            print('n=%d (%.1f MB)  NCCL %.1f us  busBW %.0f GB/s' % (n, S / 1e6, us_nccl, S / us_nccl / 1e3 * f))
            if us_sym is not None:
                print('   NCCL symmetric window: %.1f us  busBW %.0f GB/s' % (us_sym, S / us_sym / 1e3 * f))
            for us, mode, threads, ctas in sorted(rows)[:6]:
                print('   p2p mode%d t%d c%d: %.1f us  busBW %.0f GB/s  (x%.2f vs NCCL)' % (
                    mode, threads, ctas, us, S / us / 1e3 * f, us_nccl / us))
            print('   worst: %s' % (sorted(rows)[-1],), flush=True)
    lib.gp_p2p_set_tuning(0, 512, 0)
    comm.finalize()


if __name__ == '__main__':
    main()
