#!/usr/bin/env python
"""Benchmark of the ChainerMN gradient path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

A "step" is one pass of the hot path over one synthetic gradient set of the
workload: pack+cast -> in-place NCCL allreduce (N > 1) -> fused unpack + 1/N
scale + optimizer update, through the public API
(`create_multi_node_optimizer(...).update()`), with parameters, optimizer state
and gradients resident in HBM.  Default workload: BASELINE configs[1], ResNet-50
(162 tensors, 25,557,096 elements) fp32 gradients, MomentumSGD.

Rank 0 prints ONE JSON line (see DESIGN.md "Measurement" for every key).
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = 'mean_grad+update GB/s (ResNet-50 gradient path; algorithmic HBM bytes / time)'


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=400)
    ap.add_argument('--warmup', type=int, default=20)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference', 'reference-worker'])
    ap.add_argument('--workload', default='resnet50', choices=['resnet50', 'seq2seq', 'mnist_mlp'])
    ap.add_argument('--optimizer', default=None, choices=[None, 'momentum_sgd', 'adam'])
    ap.add_argument('--allreduce-dtype', default='float32',
                    choices=['float32', 'float16', 'bfloat16'])
    ap.add_argument('--no-write-grad', action='store_true',
                    help='do not keep param.grad observable after the fused update')
    ap.add_argument('--bucket-mb', type=float, default=None)
    ap.add_argument('--no-p2p', action='store_true', help='NCCL allreduce instead of the peer-memory kernel')
    ap.add_argument('--zero-embedding-rows', type=float, default=0.0,
                    help='seq2seq (config 4 variant): fraction of embedding-gradient rows set to zero '
                         '(token sparsity; values only, layout unchanged)')
    ap.add_argument('--mc-chunk-mb', type=float, default=None, help='multicast path: pipeline chunk size')
    ap.add_argument('--multicast', choices=['auto', 'on', 'off'], default='auto',
                    help='NVSwitch multicast allreduce kernel (auto: 4 and 8 GPUs)')
    ap.add_argument('--p2p-chunk-mb', type=float, default=None)
    ap.add_argument('--p2p-ctas', type=int, default=None)
    ap.add_argument('--cpu-seconds', type=float, default=10.0,
                    help='budget of the CPU baseline sample')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    return ap.parse_args()


def default_optimizer(workload):
    # BASELINE configs: ResNet-50 -> MomentumSGD(lr .01, momentum .9)
    # (train_imagenet.py:193-194); seq2seq and the MNIST MLP -> Adam defaults
    return 'momentum_sgd' if workload == 'resnet50' else 'adam'


def bytes_per_elem(optimizer, buf_itemsize, write_grad):
    """Algorithmic HBM bytes per element."""
    pack = 4 + buf_itemsize
    if optimizer == 'momentum_sgd':
        upd = buf_itemsize + 8 + 8
    else:
        upd = buf_itemsize + 8 + 8 + 8
    if write_grad:
        upd += 4
    return pack, upd


# ------------------------------------------------------------------- clocks --
class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons of one GPU through NVML while the
    benchmark runs (the recipe's nvidia-smi clocks line, in-process so that
    short timed regions are still covered)."""

    def __init__(self, index, period=0.02):
        super(ClockSampler, self).__init__(daemon=True)
        self.index = index
        self.period = period
        self.samples = []
        self.stop_flag = False
        self.window = None
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.max_sm = None

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        while not self.stop_flag:
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                try:
                    reasons = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    reasons = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.samples.append((time.perf_counter(), sm, int(reasons)))
            except Exception:
                pass
            time.sleep(self.period)

    def summary(self, t0, t1):
        names = {0x1: 'gpu_idle', 0x2: 'applications_clocks_setting', 0x4: 'sw_power_cap',
                 0x8: 'hw_slowdown', 0x10: 'sync_boost', 0x20: 'sw_thermal_slowdown',
                 0x40: 'hw_thermal_slowdown', 0x80: 'hw_power_brake_slowdown',
                 0x100: 'display_clock_setting'}
        inside = [s for s in self.samples if t0 <= s[0] <= t1]
        where = 'timed_region'
        if len(inside) < 2:
            inside = self.samples
            where = 'warmup+timed_region'
        if not inside:
            return {'sm_mhz': None, 'sm_max_mhz': self.max_sm, 'reasons': [], 'samples': 0}
        mhz = sorted(s[1] for s in inside)
        bits = 0
        for s in inside:
            bits |= s[2]
        reasons = [n for b, n in names.items() if bits & b and n != 'gpu_idle']
        return {'sm_mhz': mhz[len(mhz) // 2], 'sm_max_mhz': self.max_sm, 'reasons': reasons,
                'samples': len(inside), 'window': where}


# ----------------------------------------------------------------- workload --
def workload_sizes(name):
    from chainer_b200 import workloads
    plist = workloads.WORKLOADS[name]()
    return plist, [int(np.prod(s)) for _, s in plist]


# ------------------------------------------------------------ reference arm --
def run_cpu_reference(args, sizes, optimizer_name, budget_s, steps=None, warmup=1):
    """The reference's CPU implementation of the path (oracle/naive.py: NumPy port
    of NaiveCommunicator.multi_node_mean_grad + update_core_cpu), one process,
    one rank, timed with perf_counter.  Returns (elements/s, ms/step, steps)."""
    from oracle import naive
    rng = np.random.default_rng(7)
    params = [(rng.standard_normal(k) * 0.05).astype(np.float32) for k in sizes]
    grads0 = [(rng.standard_normal(k) * 1e-2).astype(np.float32) for k in sizes]
    opt = naive.MomentumSGD(params, 0.01, 0.9) if optimizer_name == 'momentum_sgd' \
        else naive.Adam(params)
    n = sum(sizes)
    times = []
    t_start = time.perf_counter()
    i = 0
    while True:
        grads = [g.copy() for g in grads0]           # fresh gradients every step (untimed)
        t0 = time.perf_counter()
        naive.step(params, grads, opt, size=1, allreduce=None)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
        i += 1
        if steps is not None:
            if len(times) >= steps:
                break
        elif time.perf_counter() - t_start > budget_s and len(times) >= 3:
            break
    ms = 1e3 * float(np.median(times))
    return n / (ms * 1e-3), ms, len(times)


def reference_worker_main(args):
    """One CPU rank of the reference arm at N > 1: the `naive` communicator's step
    (per-parameter in-place Allreduce over the host cores, here gloo in place of MPI,
    then `*= 1/size` and update_core_cpu).  Launched by reference_main."""
    import torch
    import torch.distributed as dist
    from oracle import naive
    torch.set_num_threads(1)
    rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
    dist.init_process_group('gloo', rank=rank, world_size=world)
    plist, sizes = workload_sizes(args.workload)
    optimizer_name = args.optimizer or default_optimizer(args.workload)
    rng = np.random.default_rng(7)
    params = [(rng.standard_normal(k) * 0.05).astype(np.float32) for k in sizes]
    grng = np.random.default_rng(1000 + rank)
    grads0 = [(grng.standard_normal(k) * 1e-2).astype(np.float32) for k in sizes]
    opt = naive.MomentumSGD(params, 0.01, 0.9) if optimizer_name == 'momentum_sgd' \
        else naive.Adam(params)

    def allreduce(a):
        if a.size:
            dist.all_reduce(torch.from_numpy(a))          # in place, like MPI.IN_PLACE

    times = []
    for i in range(args.warmup + args.steps):
        grads = [g.copy() for g in grads0]
        dist.barrier()
        t0 = time.perf_counter()
        naive.step(params, grads, opt, size=world, allreduce=allreduce)
        dist.barrier()
        if i >= args.warmup:
            times.append(time.perf_counter() - t0)
    if rank == 0:
        print('CPU_REFERENCE_RESULT ' + json.dumps({'ms': 1e3 * float(np.median(times)),
                                                    'steps': len(times)}), flush=True)
    dist.destroy_process_group()


def run_cpu_reference_ranks(args, n_ranks, steps, warmup):
    """N CPU ranks of the reference's naive path on this box's host cores (what
    `mpiexec -n N` with the `naive` communicator runs); returns (ms/step, steps)."""
    import socket
    import subprocess
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    procs = []
    for r in range(n_ranks):
        env = {k: v for k, v in os.environ.items() if not k.startswith('TORCHELASTIC')}
        env.update(RANK=str(r), WORLD_SIZE=str(n_ranks), LOCAL_RANK=str(r),
                   MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), OMP_NUM_THREADS='1',
                   CUDA_VISIBLE_DEVICES='')
        cmd = [sys.executable, os.path.abspath(__file__), '--impl', 'reference-worker',
               '--gpus', str(n_ranks), '--steps', str(steps), '--warmup', str(warmup),
               '--workload', args.workload]
        if args.optimizer:
            cmd += ['--optimizer', args.optimizer]
        procs.append(subprocess.Popen(cmd, env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=1500)[0] for p in procs]
    if any(p.returncode != 0 for p in procs):
        raise RuntimeError('CPU reference ranks failed:\n' + '\n'.join(o[-2000:] for o in outs))
    for line in outs[0].splitlines():
        if line.startswith('CPU_REFERENCE_RESULT '):
            res = json.loads(line[len('CPU_REFERENCE_RESULT '):])
            return res['ms'], res['steps']
    raise RuntimeError('CPU reference rank 0 printed no result:\n' + outs[0][-2000:])


def reference_main(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    plist, sizes = workload_sizes(args.workload)
    optimizer_name = args.optimizer or default_optimizer(args.workload)
    pack_b, upd_b = bytes_per_elem(optimizer_name, 4, True)
    n = sum(sizes)
    n_ranks = max(1, args.gpus)
    if n_ranks == 1:
        steps = max(1, min(args.steps, 50))
        eps, ms, done = run_cpu_reference(args, sizes, optimizer_name, args.cpu_seconds,
                                          steps=steps, warmup=max(1, min(args.warmup, 3)))
    else:
        # the reference's CPU job of the same shape: N ranks of the naive communicator
        ms, done = run_cpu_reference_ranks(args, n_ranks, steps=max(1, min(args.steps, 20)),
                                           warmup=max(1, min(args.warmup, 2)))
        eps = n_ranks * n / (ms * 1e-3)
    gbs = eps * (pack_b + upd_b) / 1e9
    sample = '{} steps of the full {} workload ({} tensors, {} elements), {} rank{}'.format(
        done, args.workload, len(sizes), n, n_ranks,
        '' if n_ranks == 1 else 's (one process each, gloo allreduce per parameter)')
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': gbs, 'unit': 'GB/s', 'n_gpus': args.gpus,
        'steps': done, 'warmup': max(1, min(args.warmup, 3)), 'ms_per_step': ms,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
        'data': 'synthetic',
        'config': {'workload': workload_name(args, optimizer_name, n_ranks), 'n_tensors': len(sizes),
                   'n_elems': n, 'bytes_per_elem': pack_b + upd_b},
        'cpu_baseline': {'value': gbs, 'unit': 'GB/s', 'cores': n_ranks, 'kind': 'port',
                         'sample': sample, 'host_cores': os.cpu_count()},
        'e2e': {'value': gbs, 'unit': 'GB/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'note': 'NumPy port (oracle/naive.py) of the reference naive communicator + '
                'update_core_cpu; the reference NumPy path is single-threaded per process, so '
                'the job uses one host core per rank (N ranks at --gpus N, gloo standing in '
                'for MPI)',
    }
    print(json.dumps(line), flush=True)


def workload_name(args, optimizer_name, n_gpus):
    return '{} {} grads, {}, {} (BASELINE configs[{}]) x {} GPU'.format(
        args.workload, args.allreduce_dtype, optimizer_name, 'pure_nccl',
        {'resnet50': 1, 'seq2seq': 3, 'mnist_mlp': 0}[args.workload], n_gpus)


# ------------------------------------------------------------------ B200 arm --
def b200_main(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit('--gpus {} needs a torchrun launch with {} ranks'.format(
                args.gpus, args.gpus))
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group(backend='gloo', rank=rank, world_size=world)

    import chainer_b200
    from chainer_b200 import _lib
    from chainer_b200 import device as dev
    from chainer_b200.core.link import link_from_named_arrays
    lib = _lib.get()

    sampler = ClockSampler(local_rank)
    sampler.start()

    plist, sizes = workload_sizes(args.workload)
    n = sum(sizes)
    optimizer_name = args.optimizer or default_optimizer(args.workload)
    write_grad = not args.no_write_grad
    adt = {'float32': np.float32, 'float16': np.float16, 'bfloat16': 'bfloat16'}[args.allreduce_dtype]
    bsz = 4 if args.allreduce_dtype == 'float32' else 2
    pack_b, upd_b = bytes_per_elem(optimizer_name, bsz, write_grad)

    comm = chainer_b200.create_communicator('pure_nccl', allreduce_grad_dtype=adt)
    comm.write_grad = write_grad
    if args.no_p2p:
        comm.use_p2p = False
    if args.multicast != 'auto':
        comm.use_multicast = args.multicast == 'on'
    if args.mc_chunk_mb is not None:
        comm.mc_chunk_bytes = int(args.mc_chunk_mb * (1 << 20))
    if args.p2p_chunk_mb is not None:
        comm.p2p_chunk_bytes = int(args.p2p_chunk_mb * (1 << 20))
    if args.p2p_ctas is not None:
        lib.gp_p2p_set_tuning(args.p2p_ctas, 512, 1)
    if args.bucket_mb is not None:
        comm.bucket_bytes = int(args.bucket_mb * (1 << 20))

    # Arenas in layout order; every parameter / gradient is a view (any device
    # array works -- an arena makes the e2e host copies single transfers).
    gen = torch.Generator(device='cuda')
    gen.manual_seed(7)                                   # params identical on all ranks
    p_arena = torch.randn(n, device='cuda', generator=gen) * 0.05
    gen.manual_seed(1000 + rank)                         # grads differ per rank
    n_sets = 2                                           # gradients move every step, as after cleargrads()
    g_arenas = [torch.randn(n, device='cuda', generator=gen) * 1e-2 for _ in range(n_sets)]
    offs = np.concatenate([[0], np.cumsum(sizes)])
    views = lambda arena: [arena[offs[i]:offs[i + 1]] for i in range(len(sizes))]  # noqa: E731
    p_views = views(p_arena)
    g_views = [views(a) for a in g_arenas]
    if args.zero_embedding_rows > 0:
        # EmbedID's dense W-gradient has non-zero rows only for the tokens of the batch
        for gv in g_views:
            for (nm, shape), g in zip(plist, gv):
                if 'embed' in nm and len(shape) == 2:
                    rows = torch.rand(shape[0], device='cuda', generator=gen) < args.zero_embedding_rows
                    g.view(shape)[rows] = 0
    model = link_from_named_arrays([(nm, v) for (nm, _), v in zip(plist, p_views)])
    params_sorted = [p for _, p in sorted(model.namedparams())]
    actual = chainer_b200.MomentumSGD(lr=0.01, momentum=0.9) if optimizer_name == 'momentum_sgd' \
        else chainer_b200.Adam()
    opt = chainer_b200.create_multi_node_optimizer(actual, comm)
    opt.setup(model)

    def set_grads(k):
        gv = g_views[k % n_sets]
        for p, g in zip(params_sorted, gv):
            p.grad = g

    def step(k):
        set_grads(k)
        opt.update()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    set_grads(0)
    opt.update()                                         # first call: bcast_data only
    W = max(args.warmup, 3)
    K = max(args.steps, 1)
    for k in range(W):
        step(k)
    barrier()

    # ---- timed region: K steps, device-resident inputs --------------------------
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    lib_calls_before = lib.launches
    t0 = time.perf_counter()
    ev0.record()
    for k in range(K):
        step(W + k)
    ev1.record()
    t_enq = time.perf_counter()
    barrier()
    t1 = time.perf_counter()
    ms_total = ev0.elapsed_time(ev1)
    launches = lib.launches - lib_calls_before
    ms_step = ms_total / K
    if world > 1:
        t = torch.tensor([ms_step], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_step = float(t.item())
    clocks = sampler.summary(t0, t1)

    # ---- per-kernel timing of the dominant kernel (fused update) ----------------
    kern = time_kernels(torch, lib, comm, model, actual, opt, set_grads, optimizer_name, n,
                        bsz, write_grad, reps=min(max(K // 4, 10), 50))

    # ---- allreduce bus bandwidth (N > 1) ---------------------------------------
    bus = None
    if world > 1:
        bus = time_allreduce(torch, dist, comm, n, bsz, world)

    # ---- e2e: host buffers in, host buffers out ---------------------------------
    e2e = None
    if not args.no_e2e:
        try:
            e2e = time_e2e(torch, dist, world, opt, params_sorted, p_arena, g_arenas, set_grads,
                           n, pack_b + upd_b, steps=min(K, 30), warmup=3)
        except Exception as e:      # noqa: BLE001 -- the device-resident numbers above stand
            e2e = {'value': None, 'unit': 'GB/s', 'h2d_bytes_per_step': n * 4,
                   'd2h_bytes_per_step': n * 4, 'error': '%s: %s' % (type(e).__name__, e)}

    sampler.stop_flag = True
    if rank != 0:
        comm.finalize()
        return

    value = world * n * (pack_b + upd_b) / (ms_step * 1e-3) / 1e9
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    peak = peaks.get('hbm_gbs', 6650.0)
    peak_src = 'MEASURED_PEAKS.json hbm_gbs (of measured)' if 'hbm_gbs' in peaks \
        else 'B200_PROFILING.md fallback 6.65 TB/s (of fallback)'
    achieved = n * upd_b / (kern['update_us'] * 1e-6) / 1e9
    traffic = None
    try:
        prof = json.load(open(os.path.join(ROOT, 'profiles', 'ncu_traffic.json')))
        if args.workload == 'resnet50':       # the ncu capture is of the ResNet-50 list
            traffic = prof.get('{}_{}_{}'.format(optimizer_name, args.allreduce_dtype,
                                                  'wg' if write_grad else 'nowg'))
    except Exception:
        pass

    line = {
        'metric': METRIC, 'value': value, 'unit': 'GB/s', 'n_gpus': world, 'steps': K,
        'warmup': W, 'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32' if bsz == 4 else args.allreduce_dtype,
        'data': 'synthetic',
        'config': {
            'workload': workload_name(args, optimizer_name, world),
            'n_tensors': len(sizes), 'n_elems': n, 'packed_bytes': n * bsz,
            'bytes_per_elem': pack_b + upd_b, 'write_grad': write_grad,
            'l2': 'working set {} MB per step > 126 MB L2; gradient arrays rotate between {} '
                  'sets'.format((n * (4 + bsz + 4 + 4 + (4 if optimizer_name == "adam" else 0))) >> 20,
                                n_sets),
            'api': 'create_multi_node_optimizer(MomentumSGD|Adam, pure_nccl).update()',
            'bucket_bytes': comm.bucket_bytes if world > 1 else None,
            'allreduce_impl': (None if world == 1 else _allreduce_impl(comm)),
            'zero_embedding_rows': args.zero_embedding_rows or None,
        },
        'roofline': {
            'bound': 'hbm', 'kernel': kern['update_kernel'], 'achieved': achieved, 'peak': peak,
            'unit': 'GB/s', 'frac': achieved / peak, 'traffic': traffic, 'peak_source': peak_src,
            'bytes_per_launch': n * upd_b, 'us_per_launch': kern['update_us'],
            'pack_us': kern['pack_us'], 'pack_gbs': n * pack_b / kern['pack_us'] / 1e3,
            'pack_frac': n * pack_b / kern['pack_us'] / 1e3 / peak,
            'frac_of_nominal_8TBs': achieved / 8000.0,
        },
        'e2e': e2e,
        'gpu_launches': launches,
        'clocks': clocks,
        'img_per_s_gradpath_bound': 32.0 * world / (ms_step * 1e-3),
        'host_us_per_step': 1e6 * (t1 - t0) / K,
        'host_enqueue_us_per_step': 1e6 * (t_enq - t0) / K,
    }
    if bus is not None:
        line['allreduce'] = bus
    if not args.no_cpu_baseline:
        eps, cms, done = run_cpu_reference(args, sizes, optimizer_name, args.cpu_seconds)
        line['cpu_baseline'] = {
            'value': eps * (pack_b + upd_b) / 1e9, 'unit': 'GB/s', 'cores': 1, 'kind': 'port',
            'ms_per_step': cms, 'host_cores': os.cpu_count(),
            'sample': '{} steps of the full {} workload on 1 host core (NumPy port of the '
                      'reference naive communicator + update_core_cpu)'.format(done, args.workload)}
    print(json.dumps(line), flush=True)
    comm.finalize()


def time_kernels(torch, lib, comm, model, actual, opt, set_grads, optimizer_name, n, bsz,
                 write_grad, reps):
    """Average duration of the pack and the fused-update kernels, CUDA events on
    the launching (null) stream, measured live by wrapping the library calls."""
    from chainer_b200 import device as dev
    rec = {'gp_pack': [], 'upd': []}
    upd_name = 'gp_unpack_momentum_sgd' if optimizer_name == 'momentum_sgd' else 'gp_unpack_adam'
    orig_pack, orig_upd = lib.gp_pack, getattr(lib, upd_name)

    def wrap(fn, key):
        def call(*a):
            stream = a[-1] or 0
            e0, e1 = dev.Event(timing=True), dev.Event(timing=True)
            e0.record(stream)
            r = fn(*a)
            e1.record(stream)
            rec[key].append((e0, e1))
            return r
        return call
    lib.gp_pack = wrap(orig_pack, 'gp_pack')
    setattr(lib, upd_name, wrap(orig_upd, 'upd'))
    try:
        for k in range(reps):
            set_grads(k)
            opt.update()
        torch.cuda.synchronize()
    finally:
        lib.gp_pack = orig_pack
        setattr(lib, upd_name, orig_upd)
    per_step = max(len(rec['upd']) // reps, 1)

    def total_us(pairs):
        return sum(a.elapsed_ms(b) for a, b in pairs) * 1e3 / reps
    return {'pack_us': total_us(rec['gp_pack']), 'update_us': total_us(rec['upd']),
            'update_kernel': upd_name + (' x%d buckets' % per_step if per_step > 1 else ''),
            'launches_per_step': per_step}


def _allreduce_impl(comm):
    if comm._p2p is None:
        return 'nccl'
    if comm._mc_active(comm.gpu_buffer_a):
        return 'nvswitch multicast kernel (gp_mc)'
    return 'peer-memory kernel (gp_p2p)'


def time_allreduce(torch, dist, comm, n, bsz, world):
    """NCCL-tests convention: algBW = S / t, busBW = algBW * 2(N-1)/N."""
    from chainer_b200 import nccl
    from chainer_b200.communicators import _communication_utility as cu
    buf = comm.gpu_buffer_a
    type_id = 7 if bsz == 4 else 6
    dt = np.float32 if bsz == 4 else np.float16
    if comm._p2p is not None and comm._mc_active(buf):
        def one():
            comm._p2p.mc_allreduce(dt, 0, n, None)
    elif comm._p2p is not None:
        def one():
            comm._p2p.allreduce(dt, 0, n, None)
    else:
        def one():
            comm.nccl_comm.allReduce(buf.ptr(), buf.ptr(), n, type_id, nccl.NCCL_SUM, 0)
    for _ in range(3):
        one()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20
    e0.record()
    for _ in range(reps):
        one()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / reps
    t = torch.tensor([us], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    us = float(t.item())
    S = n * bsz
    alg = S / us / 1e3
    busbw = alg * 2 * (world - 1) / world
    return {'impl': _allreduce_impl(comm),
            'bytes': S, 'us': us, 'alg_gbs': alg, 'bus_gbs': busbw,
            'frac_of_900': busbw / 900.0, 'frac_of_measured_725': busbw / 725.0}


def time_e2e(torch, dist, world, opt, params_sorted, p_arena, g_arenas, set_grads, n,
             bytes_per_elem_total, steps, warmup):
    """Same step through the public API with HOST buffers: every step copies the
    step's gradients from pinned host memory to the device, runs
    optimizer.update(), and reads the updated parameters back to pinned host
    memory; all inside the timed region.  The copies run on their own streams so
    that the H2D of step k+1 and the D2H of step k-1 overlap the kernels of step k
    (PCIe is full duplex); dependencies are expressed with events:
        H2D(k) -> update(k) -> snapshot(k) -> D2H(k);  H2D(k+2) waits update(k)
    (two gradient arenas, two parameter snapshots)."""
    h_grads = [torch.empty(n, dtype=torch.float32).pin_memory() for _ in range(2)]
    for h, g in zip(h_grads, g_arenas):
        h.copy_(g.cpu())
    h_params = [torch.empty(n, dtype=torch.float32).pin_memory() for _ in range(2)]
    snap = [torch.empty_like(p_arena) for _ in range(2)]
    s_h2d, s_d2h = torch.cuda.Stream(), torch.cuda.Stream()
    main = torch.cuda.current_stream()
    ev_h2d = [torch.cuda.Event() for _ in range(2)]
    ev_upd = [torch.cuda.Event() for _ in range(2)]
    ev_snap = [torch.cuda.Event() for _ in range(2)]
    ev_d2h = [torch.cuda.Event() for _ in range(2)]

    def one(k):
        i = k % 2
        with torch.cuda.stream(s_h2d):
            s_h2d.wait_event(ev_upd[i])                 # arena i was consumed by update(k-2)
            g_arenas[i].copy_(h_grads[i], non_blocking=True)       # H2D, pinned
            ev_h2d[i].record(s_h2d)
        main.wait_event(ev_h2d[i])
        set_grads(k)
        opt.update()
        ev_upd[i].record(main)
        main.wait_event(ev_d2h[i])                      # snapshot i was read back by D2H(k-2)
        snap[i].copy_(p_arena)                          # device snapshot of the result
        ev_snap[i].record(main)
        with torch.cuda.stream(s_d2h):
            s_d2h.wait_event(ev_snap[i])
            h_params[i].copy_(snap[i], non_blocking=True)          # D2H, pinned
            ev_d2h[i].record(s_d2h)
    # the pinned allocations above take rank-dependent time: line the ranks up before
    # the first collective kernel spins on its peers
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    for k in range(warmup):
        one(k)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(main)
    for k in range(steps):
        one(warmup + k)
    main.wait_event(ev_d2h[0])                          # the end event covers all three streams
    main.wait_event(ev_d2h[1])
    e1.record(main)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    assert bool(torch.isfinite(h_params[0][:1024]).all())
    return {'value': world * n * bytes_per_elem_total / (ms * 1e-3) / 1e9, 'unit': 'GB/s',
            'h2d_bytes_per_step': n * 4, 'd2h_bytes_per_step': n * 4, 'ms_per_step': ms,
            'steps': steps,
            'note': 'H2D(k+1) and D2H(k-1) overlap step k on separate streams; CUDA events, the '
                    'end event waits for the last D2H copies'}


def main():
    args = parse_args()
    if args.impl == 'reference':
        reference_main(args)
    elif args.impl == 'reference-worker':
        reference_worker_main(args)
    else:
        b200_main(args)


if __name__ == '__main__':
    main()
