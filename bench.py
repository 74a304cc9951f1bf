#!/usr/bin/env python
"""Benchmark of the ChainerMN gradient path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

A "step" is one pass of the hot path over one synthetic gradient set of the
workload: pack+cast -> sum over the N ranks -> unpack + 1/N scale + optimizer update,
through the public API (`create_multi_node_optimizer(...).update()`), with parameters,
optimizer state and gradients resident in HBM.  Default workload: BASELINE configs[1],
ResNet-50 (162 tensors, 25,557,096 elements) fp32 gradients, MomentumSGD.

Rank 0 prints ONE JSON line (DESIGN.md "Measurement" explains every key):
  value / ms_per_step   device-timed whole-job throughput of the step (CUDA events, max over ranks)
  roofline              the dominant kernel against the measured HBM peak (+ whole-step fractions)
  parity                J more steps through the same call, replayed by the NumPy oracle on a sample
                        of tensors with every rank's gradients, + bit-identical replicas across ranks
  allreduce             (N > 1) bus / wire bandwidth of the stand-alone reduction kernel
  e2e                   the same step with HOST gradients in and HOST parameters out
  train                 the "ResNet-50 img/s" half of the metric: torchvision resnet50 fwd/bwd
                        (batch 32/GPU) handing its .grad arrays to the path by pointer
  cpu_baseline          the reference's CPU path on this box's host cores (N = 1 only)

`--impl reference` times the reference's own CPU implementation of the same step: the
UNMODIFIED chainermn `naive` communicator + chainer `update_core_cpu` from
`baseline/_ref` (kind "reference"), or -- when that install is absent -- the NumPy port
under `oracle/` (kind "port"); N ranks on N host cores at --gpus N.
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
NOMINAL_HBM_GBS = 8000.0       # north_star: "~8 TB/s HBM peak"
NVLINK_GBS = 900.0             # NVLink 5, per direction and GPU


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=200)
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
    ap.add_argument('--no-step', action='store_true',
                    help='separate pack / allreduce / update launches instead of the one-launch step')
    ap.add_argument('--zero-embedding-rows', type=float, default=0.0,
                    help='seq2seq (config 4 variant): fraction of embedding-gradient rows set to zero '
                         '(token sparsity; values only, layout unchanged)')
    ap.add_argument('--mc-chunk-mb', type=float, default=None, help='multicast path: pipeline chunk size')
    ap.add_argument('--multicast', choices=['auto', 'on', 'off'], default='auto',
                    help='NVSwitch multicast transport (auto: 4 and 8 GPUs)')
    ap.add_argument('--p2p-chunk-mb', type=float, default=None)
    ap.add_argument('--p2p-ctas', type=int, default=None)
    ap.add_argument('--step-tuning', default='',
                    help='comma-separated key=value pairs for gp_step_set_tuning')
    ap.add_argument('--no-config3', action='store_true',
                    help='skip the BASELINE configs[2] legs of the default workload (float16 packed '
                         'buffer; MultiNodeBatchNormalization statistics)')
    ap.add_argument('--cpu-seconds', type=float, default=10.0,
                    help='budget of the CPU baseline sample')
    ap.add_argument('--parity-steps', type=int, default=3)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-train', action='store_true')
    ap.add_argument('--legs-deadline-s', type=float, default=180.0,
                    help='time allowed to the legs that follow the headline (configs[2], img/s); '
                         'past it the line is printed without them')
    ap.add_argument('--no-parity', action='store_true')
    ap.add_argument('--reference-kind', default='auto', choices=['auto', 'reference', 'port'])
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


def workload_name(args, optimizer_name, n_gpus):
    return '{} {} grads, {}, {} (BASELINE configs[{}]) x {} GPU'.format(
        args.workload, args.allreduce_dtype, optimizer_name, 'pure_nccl',
        {'resnet50': 1, 'seq2seq': 3, 'mnist_mlp': 0}[args.workload], n_gpus)


def _l2_note(working_set):
    mb = working_set >> 20
    if mb > 126:
        return ('inputs larger than L2: working set {} MB per step > 126 MB; gradient arrays rotate '
                'between 2 sets'.format(mb))
    return ('working set {} MB per step is below the 126 MB L2 (a latency-bound size): gradient arrays '
            'rotate between 2 sets, L2 not flushed'.format(mb))


def make_config(args, optimizer_name, n_gpus, sizes):
    """The `config` object: identical for both arms (b200 / reference)."""
    n = sum(sizes)
    bsz = 4 if args.allreduce_dtype == 'float32' else 2
    write_grad = not args.no_write_grad
    pack_b, upd_b = bytes_per_elem(optimizer_name, bsz, write_grad)
    return {
        'workload': workload_name(args, optimizer_name, n_gpus),
        'n_tensors': len(sizes), 'n_elems': n, 'packed_bytes': n * bsz,
        'bytes_per_elem': pack_b + upd_b, 'write_grad': write_grad,
        'optimizer': optimizer_name, 'allreduce_dtype': args.allreduce_dtype,
        'l2': _l2_note(n * (4 + bsz + 4 + 4 + (4 if optimizer_name == 'adam' else 0))),
        'zero_embedding_rows': args.zero_embedding_rows or None,
    }


# ------------------------------------------------------------------- clocks --
class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons of one GPU through NVML while the
    benchmark runs (the recipe's nvidia-smi clocks line, in-process so that
    short timed regions are still covered)."""

    def __init__(self, index, period=0.005):
        super(ClockSampler, self).__init__(daemon=True)
        self.index = index
        self.period = period
        self.samples = []
        self.stop_flag = False
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
def _ref_kind(args):
    """'reference' when the unmodified chainer/chainermn install is present."""
    if args.reference_kind != 'auto':
        return args.reference_kind
    from baseline import ref_shims
    return 'reference' if ref_shims.available() else 'port'


def _host_arrays(sizes, rank):
    rng = np.random.default_rng(7)                       # parameters identical on all ranks
    params = [(rng.standard_normal(k) * 0.05).astype(np.float32) for k in sizes]
    grng = np.random.default_rng(1000 + rank)            # gradients differ per rank
    grads0 = [(grng.standard_normal(k) * 1e-2).astype(np.float32) for k in sizes]
    return params, grads0


class _RefJob(object):
    """One rank of the reference's CPU job through the reference's own public API:
    chainermn.create_communicator('naive') + create_multi_node_optimizer(chainer.optimizers.*)
    on a chainer.Link holding the workload's parameters
    (chainermn/optimizers.py:17-33 -> naive_communicator.py:10-17 ->
    mpi_communicator_base.py:735-778 -> chainer/optimizer.py:857-894 -> update_core_cpu)."""

    def __init__(self, plist, sizes, optimizer_name, rank, world_comm):
        from baseline import ref_shims
        chainer, chainermn = ref_shims.import_reference(world_comm)
        params, self.grads0 = _host_arrays(sizes, rank)
        link = chainer.Link()
        with link.init_scope():
            for i, ((_, shape), a) in enumerate(zip(plist, params)):
                setattr(link, 'p%05d' % i, chainer.Parameter(a.reshape(shape)))
        self.params = [p for _, p in sorted(link.namedparams())]
        comm = chainermn.create_communicator('naive')
        actual = chainer.optimizers.MomentumSGD(lr=0.01, momentum=0.9) \
            if optimizer_name == 'momentum_sgd' else chainer.optimizers.Adam()
        self.opt = chainermn.create_multi_node_optimizer(actual, comm)
        self.opt.setup(link)
        self.set_grads()
        self.opt.update()                                # first call: bcast_data only

    def set_grads(self):
        for p, g in zip(self.params, self.grads0):
            p.grad = g.reshape(p.shape).copy()           # fresh gradients every step (untimed)

    def step(self):
        self.opt.update()


class _PortJob(object):
    """The same job with the NumPy port (oracle/naive.py); used only when baseline/_ref
    is absent."""

    def __init__(self, plist, sizes, optimizer_name, rank, world, allreduce):
        from oracle import naive
        self.naive = naive
        self.params, self.grads0 = _host_arrays(sizes, rank)
        self.opt = naive.MomentumSGD(self.params, 0.01, 0.9) if optimizer_name == 'momentum_sgd' \
            else naive.Adam(self.params)
        self.world, self.allreduce = world, allreduce
        self.set_grads()

    def set_grads(self):
        self.grads = [g.copy() for g in self.grads0]

    def step(self):
        self.naive.step(self.params, self.grads, self.opt, size=self.world,
                        allreduce=self.allreduce if self.world > 1 else None)


def _time_cpu_job(job, steps, warmup, barrier=None, budget_s=None):
    times = []
    t_start = time.perf_counter()
    i = 0
    while True:
        job.set_grads()
        if barrier is not None:
            barrier()
        t0 = time.perf_counter()
        job.step()
        if barrier is not None:
            barrier()
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
        i += 1
        if steps is not None:
            if len(times) >= steps:
                break
        elif time.perf_counter() - t_start > budget_s and len(times) >= 3:
            break
    return 1e3 * float(np.median(times)), len(times)


def run_cpu_reference(args, plist, sizes, optimizer_name, budget_s=None, steps=None, warmup=1,
                      kind=None):
    """The reference's CPU path, one process, one rank.  Returns (ms/step, steps, kind)."""
    kind = kind or _ref_kind(args)
    if kind == 'reference':
        job = _RefJob(plist, sizes, optimizer_name, 0, None)
    else:
        job = _PortJob(plist, sizes, optimizer_name, 0, 1, None)
    ms, done = _time_cpu_job(job, steps, warmup, budget_s=budget_s)
    return ms, done, kind


def reference_worker_main(args):
    """One CPU rank of the reference arm at N > 1 (launched by reference_main)."""
    import torch
    import torch.distributed as dist
    torch.set_num_threads(1)
    rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
    dist.init_process_group('gloo', rank=rank, world_size=world)
    plist, sizes = workload_sizes(args.workload)
    optimizer_name = args.optimizer or default_optimizer(args.workload)
    kind = _ref_kind(args)
    if kind == 'reference':
        from baseline import ref_shims
        job = _RefJob(plist, sizes, optimizer_name, rank, ref_shims.GlooComm(None))
    else:
        def allreduce(a):
            if a.size:
                dist.all_reduce(torch.from_numpy(a))      # in place, like MPI.IN_PLACE
        job = _PortJob(plist, sizes, optimizer_name, rank, world, allreduce)
    ms, done = _time_cpu_job(job, args.steps, args.warmup, barrier=dist.barrier)
    if rank == 0:
        print('CPU_REFERENCE_RESULT ' + json.dumps({'ms': ms, 'steps': done, 'kind': kind}),
              flush=True)
    dist.destroy_process_group()


def run_cpu_reference_ranks(args, n_ranks, steps, warmup):
    """N CPU ranks of the reference's naive path on this box's host cores (what
    `mpiexec -n N` with the `naive` communicator runs); returns (ms/step, steps, kind)."""
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
               '--workload', args.workload, '--reference-kind', args.reference_kind]
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
            return res['ms'], res['steps'], res['kind']
    raise RuntimeError('CPU reference rank 0 printed no result:\n' + outs[0][-2000:])


def reference_main(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    plist, sizes = workload_sizes(args.workload)
    optimizer_name = args.optimizer or default_optimizer(args.workload)
    cfg = make_config(args, optimizer_name, max(1, args.gpus), sizes)
    n = sum(sizes)
    n_ranks = max(1, args.gpus)
    W, K = max(args.warmup, 0), max(args.steps, 1)
    if n_ranks == 1:
        ms, done, kind = run_cpu_reference(args, plist, sizes, optimizer_name, steps=K, warmup=W)
    else:
        ms, done, kind = run_cpu_reference_ranks(args, n_ranks, steps=K, warmup=W)
    gbs = n_ranks * n * cfg['bytes_per_elem'] / (ms * 1e-3) / 1e9
    what = ('unmodified chainermn naive communicator + chainer update_core_cpu from baseline/_ref '
            '(create_multi_node_optimizer(...).update())') if kind == 'reference' else \
        'NumPy port (oracle/naive.py) of the naive communicator + update_core_cpu'
    sample = '{} steps of the full {} workload ({} tensors, {} elements), {} rank{}; {}'.format(
        done, args.workload, len(sizes), n, n_ranks,
        '' if n_ranks == 1 else 's (one process and one host core each, gloo in place of MPI)', what)
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': gbs, 'unit': 'GB/s', 'n_gpus': args.gpus,
        'steps': done, 'warmup': W, 'ms_per_step': ms,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
        'data': 'synthetic', 'config': cfg,
        'cpu_baseline': {'value': gbs, 'unit': 'GB/s', 'cores': n_ranks, 'kind': kind,
                         'sample': sample, 'host_cores': os.cpu_count()},
        'e2e': {'value': gbs, 'unit': 'GB/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
        'note': 'the reference NumPy path is single-threaded per process: the job uses one host '
                'core per rank (N ranks at --gpus N); per-parameter in-place Allreduce, *= 1/size, '
                'update_core_cpu',
    }
    print(json.dumps(line), flush=True)


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
    # a dead peer must fail the run, not hang the box (in-kernel waits: gp_p2p.cuh)
    os.environ.setdefault('CHAINER_B200_PEER_TIMEOUT_S', '120')
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group(backend='gloo', rank=rank, world_size=world)

    import chainer_b200
    from chainer_b200 import _lib
    from chainer_b200.core.link import link_from_named_arrays
    lib = _lib.get()

    sampler = ClockSampler(local_rank)
    sampler.start()

    plist, sizes = workload_sizes(args.workload)
    n = sum(sizes)
    optimizer_name = args.optimizer or default_optimizer(args.workload)
    cfg = make_config(args, optimizer_name, world, sizes)
    write_grad = not args.no_write_grad
    adt = {'float32': np.float32, 'float16': np.float16, 'bfloat16': 'bfloat16'}[args.allreduce_dtype]
    bsz = 4 if args.allreduce_dtype == 'float32' else 2
    pack_b, upd_b = bytes_per_elem(optimizer_name, bsz, write_grad)

    comm = chainer_b200.create_communicator('pure_nccl', allreduce_grad_dtype=adt)
    comm.write_grad = write_grad
    if args.no_p2p:
        comm.use_p2p = False
    if args.no_step:
        comm.use_step = False
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
    for kv in [x for x in args.step_tuning.split(',') if x]:
        k, v = kv.split('=')
        lib.gp_step_set_tuning(k.encode(), int(v))

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
    sampler.stop_flag = True

    # ---- distribution of single steps (median, p10 / p90) --------------
    dist_us = None
    try:
        n_d = min(max(K, 20), 100)
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(n_d + 1)]
        barrier()
        evs[0].record()
        for k in range(n_d):
            step(W + K + k)
            evs[k + 1].record()
        torch.cuda.synchronize()
        per = np.array([evs[k].elapsed_time(evs[k + 1]) * 1e3 for k in range(n_d)])
        q = np.percentile(per, [10, 50, 90])
        if world > 1:
            t = torch.tensor(q, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            q = t.numpy()
        dist_us = {'p10': float(q[0]), 'median': float(q[1]), 'p90': float(q[2]), 'steps': n_d,
                   'note': 'one CUDA event pair per step, outside the timed region; max over ranks'}
    except Exception as e:      # noqa: BLE001 -- additional information only
        dist_us = {'error': '%s: %s' % (type(e).__name__, e)}

    # ---- per-kernel timing (CUDA events around each library launch) --------------
    kern = time_kernels(torch, dist, world, lib, opt, set_grads, reps=min(max(K // 4, 10), 50))

    # ---- parity: J more steps, replayed by the oracle ----------------------------
    parity = None
    if not args.no_parity:
        parity = check_parity(torch, dist, rank, world, comm, opt, actual, optimizer_name, adt,
                              plist, sizes, offs, p_arena, g_arenas, set_grads, n_sets,
                              write_grad, steps=args.parity_steps)

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

    value = world * n * (pack_b + upd_b) / (ms_step * 1e-3) / 1e9
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    peak = peaks.get('hbm_gbs', 6650.0)
    peak_src = 'MEASURED_PEAKS.json hbm_gbs (of measured)' if 'hbm_gbs' in peaks \
        else 'B200_PROFILING.md fallback 6.65 TB/s (of fallback)'
    # the dominant kernel: the launch that takes the most device time per step
    dom = max(kern['kernels'], key=lambda k: kern['kernels'][k]['us'])
    dom_us = kern['kernels'][dom]['us']
    kbytes = {'gp_pack': n * pack_b, 'gp_unpack_momentum_sgd': n * upd_b, 'gp_unpack_adam': n * upd_b,
              'gp_step_momentum_sgd': n * (pack_b + upd_b), 'gp_step_adam': n * (pack_b + upd_b)}
    achieved = kbytes[dom] / (dom_us * 1e-6) / 1e9
    traffic, traffic_src = None, None
    try:
        prof = json.load(open(os.path.join(ROOT, 'profiles', 'ncu_traffic.json')))
        if args.workload == 'resnet50' and world == 1:   # the ncu captures are of the ResNet-50 list
            key = '{}:{}_{}_{}'.format(dom, optimizer_name, args.allreduce_dtype,
                                       'wg' if write_grad else 'nowg')
            traffic = prof.get(key)
            if traffic is not None:
                traffic_src = 'profiles/ncu_traffic.json (ncu --set full capture of this kernel, ' \
                              'dram__bytes_read.sum + dram__bytes_write.sum; not measured in this run)'
    except Exception:
        pass
    step_bytes = n * (pack_b + upd_b)
    roofline = {
        'bound': 'hbm', 'kernel': dom, 'achieved': achieved, 'peak': peak,
        'unit': 'GB/s', 'frac': achieved / peak, 'traffic': traffic, 'traffic_source': traffic_src,
        'peak_source': peak_src, 'bytes_per_launch': kbytes[dom], 'us_per_launch': dom_us,
        'frac_of_nominal_8TBs': achieved / NOMINAL_HBM_GBS,
        'kernels': {k: dict(v, gbs=kbytes[k] / v['us'] / 1e3, frac=kbytes[k] / v['us'] / 1e3 / peak)
                    for k, v in kern['kernels'].items()},
        # the whole step (all launches, gaps included), per GPU
        'step_gbs_per_gpu': step_bytes / ms_step / 1e6,
        'step_frac_of_measured': step_bytes / ms_step / 1e6 / peak,
        'step_frac_of_nominal_8TBs': step_bytes / ms_step / 1e6 / NOMINAL_HBM_GBS,
    }
    if dom.startswith('gp_step'):
        if world == 1:
            roofline['note'] = (
                'one launch = the whole step; algorithmic bytes follow SURVEY 8(d) (pack {} + update '
                '{} B/elem) although pack and update are fused at the register level: the packed '
                'buffer is written once and never read back, so the real DRAM traffic is {} B/elem '
                '(ncu: `traffic`) and `achieved` can exceed the DRAM copy peak; in DRAM bytes the '
                'launch runs at {:.3f} of the measured peak'.format(
                    pack_b, upd_b, pack_b + upd_b - bsz,
                    n * (pack_b + upd_b - bsz) / (dom_us * 1e-6) / 1e9 / peak))
        else:
            roofline['note'] = (
                'one launch = the whole step incl. the NVLink-bound exchange over {} ranks, which '
                'bounds its duration (see `allreduce`: the stand-alone reduction of the same buffer); '
                'algorithmic HBM bytes follow SURVEY 8(d) (pack {} + update {} B/elem)'.format(
                    world, pack_b, upd_b))
    line = {
        'metric': METRIC, 'value': value, 'unit': 'GB/s', 'n_gpus': world, 'steps': K,
        'warmup': W, 'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32' if bsz == 4 else args.allreduce_dtype,
        'data': 'synthetic', 'config': cfg,
        'impl_detail': {
            'api': 'create_multi_node_optimizer(MomentumSGD|Adam, pure_nccl).update()',
            'launches_per_step': kern['launches_per_step'],
            'one_launch_step': bool(kern['launches_per_step'] == 1),
            'transport': (None if world == 1 else _allreduce_impl(comm)),
        },
        'roofline': roofline,
        'parity': parity,
        'e2e': e2e,
        'gpu_launches': launches,
        'clocks': clocks,
        'step_us': dist_us,
        'host_us_per_step': 1e6 * (t1 - t0) / K,
        'host_enqueue_us_per_step': 1e6 * (t_enq - t0) / K,
    }
    if bus is not None:
        # the step moves S(N+1)/N (multicast) or 2S(N-1)/N (peer memory) bytes per NVLink
        # direction; implied wire rate if the whole step were the exchange
        line['allreduce'] = bus
    if not args.no_cpu_baseline and world == 1 and rank == 0:
        cms, done, kind = run_cpu_reference(args, plist, sizes, optimizer_name,
                                            budget_s=args.cpu_seconds)
        line['cpu_baseline'] = {
            'value': n * (pack_b + upd_b) / (cms * 1e-3) / 1e9, 'unit': 'GB/s', 'cores': 1,
            'kind': kind, 'ms_per_step': cms, 'host_cores': os.cpu_count(),
            'sample': '{} steps of the full {} workload on 1 host core ({})'.format(
                done, args.workload,
                'unmodified chainermn naive communicator + chainer update_core_cpu, baseline/_ref'
                if kind == 'reference' else 'NumPy port of the naive communicator + update_core_cpu')}

    # ---- the legs beside the headline (configs[2], img/s) -------------------------------
    # Everything the contract needs is in `line` now.  The legs below add to it; they run
    # under a deadline so that a leg that stops making progress (a peer that died inside an
    # exchange, say) costs its own numbers, not the line: when the deadline passes, rank 0
    # prints the line as it stands, with the reason, and every rank leaves.
    guard = _LegsGuard(line, rank, args.legs_deadline_s)
    guard.start()
    if _WATCHDOG is not None:
        _WATCHDOG.cancel()      # the line is safe from here on: the legs' own deadline takes over

    # ---- BASELINE configs[2]: float16 packed buffer + MNBN statistics -------------
    mnbn = fp16 = None
    if not args.no_config3 and args.workload == 'resnet50' and args.allreduce_dtype == 'float32':
        try:
            fp16 = time_fp16_buffer(torch, dist, world, comm, step, n, optimizer_name, write_grad,
                                    steps=min(K, 50))
        except Exception as e:      # noqa: BLE001
            fp16 = {'error': '%s: %s' % (type(e).__name__, e)}
        try:
            mnbn = time_mnbn(torch, dist, world, comm, lib)
        except Exception as e:      # noqa: BLE001
            mnbn = {'error': '%s: %s' % (type(e).__name__, e)}

    # ---- the img/s half of the metric -------------------------------------------
    train = None
    if not args.no_train and args.workload == 'resnet50':
        try:
            train = time_train(torch, dist, rank, world, comm, args)
        except Exception as e:      # noqa: BLE001
            train = {'error': '%s: %s' % (type(e).__name__, e)}

    if train is not None:
        line['train'] = train
        if 'img_per_s' in train:
            line['img_per_s'] = train['img_per_s']
    if fp16 is not None or mnbn is not None:
        line['config3'] = {'what': 'BASELINE configs[2]: the same ResNet-50 step with '
                                   'allreduce_grad_dtype=float16 (fused cast) and the '
                                   'MultiNodeBatchNormalization statistics of one step',
                           'float16_buffer': fp16, 'mnbn': mnbn}
    guard.finish()
    try:
        comm.finalize()
    except Exception as e:      # noqa: BLE001 -- the line is out; a failed leg may have left the context unusable
        sys.stderr.write('finalize: %s: %s\n' % (type(e).__name__, e))


class _LegsGuard(object):
    """Owner of the bench line while the legs beside the headline run.  `finish()` prints the
    line (rank 0, once).  If it has not been called `deadline_s` seconds after `start()`,
    every thread's stack goes to stderr, rank 0 prints the line as it stands with
    `legs_error`, and the process exits 0 -- a leg that stops making progress costs its own
    numbers, not the line."""

    def __init__(self, line, rank, deadline_s):
        self.line, self.rank, self.deadline_s = line, rank, deadline_s
        self._printed = threading.Event()
        self._timer = threading.Timer(deadline_s, self._expired)
        self._timer.daemon = True

    def start(self):
        self._timer.start()

    def _emit(self):
        if self.rank == 0 and not self._printed.is_set():
            self._printed.set()
            print(json.dumps(self.line), flush=True)

    def _expired(self):
        import faulthandler
        faulthandler.dump_traceback(file=sys.stderr, all_threads=True)
        self.line['legs_error'] = ('the legs after the headline did not finish within %g s; '
                                   'line printed without them' % self.deadline_s)
        self._emit()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)

    def finish(self):
        self._timer.cancel()
        self._emit()


STEP_FUNCS = ('gp_step_momentum_sgd', 'gp_step_adam')
TIMED_FUNCS = ('gp_pack', 'gp_unpack_momentum_sgd', 'gp_unpack_adam') + STEP_FUNCS


def time_kernels(torch, dist, world, lib, opt, set_grads, reps):
    """Average duration of each library launch of a step, CUDA events on the launching
    (null) stream, measured live by wrapping the library calls."""
    from chainer_b200 import device as dev
    rec = {}
    orig = {name: getattr(lib, name) for name in TIMED_FUNCS}

    def wrap(name, fn):
        def call(*a):
            stream = a[-1] or 0
            e0, e1 = dev.Event(timing=True), dev.Event(timing=True)
            e0.record(stream)
            r = fn(*a)
            e1.record(stream)
            rec.setdefault(name, []).append((e0, e1))
            return r
        return call
    for name, fn in orig.items():
        setattr(lib, name, wrap(name, fn))
    try:
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        for k in range(reps):
            set_grads(k)
            opt.update()
        torch.cuda.synchronize()
    finally:
        for name, fn in orig.items():
            setattr(lib, name, fn)
    out = {}
    total = 0
    for name, pairs in rec.items():
        us = sum(a.elapsed_ms(b) for a, b in pairs) * 1e3 / reps
        out[name] = {'us': us, 'launches_per_step': len(pairs) // reps}
        total += len(pairs) // reps
    return {'kernels': out, 'launches_per_step': total}


def _allreduce_impl(comm):
    if comm._p2p is None:
        return 'nccl'
    if comm._mc_active(comm.gpu_buffer_a):
        return 'nvswitch multicast (multimem.ld_reduce / multimem.st)'
    return 'nvlink peer memory (rank-order sums)'


def _sample_tensors(sizes, budget=1200000):
    """Indices of the tensors the parity replay follows: the first small ones, the last
    one of the layout, and mid-size ones up to `budget` elements in total (whole tensors;
    several tiles and tile boundaries of the step kernels are covered)."""
    idx = [i for i in range(min(len(sizes), 24)) if sizes[i] <= 4096][:8]
    idx.append(len(sizes) - 1)
    total = sum(sizes[i] for i in set(idx))
    for i in sorted(range(len(sizes)), key=lambda i: -sizes[i]):
        if i not in idx and 30000 <= sizes[i] <= 600000 and total + sizes[i] <= budget:
            idx.append(i)
            total += sizes[i]
    return sorted(set(i for i in idx if sizes[i] > 0))


def check_parity(torch, dist, rank, world, comm, opt, actual, optimizer_name, adt, plist, sizes,
                 offs, p_arena, g_arenas, set_grads, n_sets, write_grad, steps):
    """Run `steps` more steps through the SAME public call the timed region used and replay
    them with the NumPy oracle (oracle/gradpath.py) for a sample of tensors, from the
    device state before those steps and every rank's gradients of each step.  Bit-exact
    where the summation order is the oracle's (1 or 2 ranks, peer-memory transport),
    within the rounding bound of a re-ordered float sum otherwise (NVSwitch / NCCL add in
    their own order).  Also checks that all ranks hold bit-identical parameters."""
    from oracle import gradpath as og
    odt = og.BF16 if adt == 'bfloat16' else np.dtype(adt)
    idx = _sample_tensors(sizes)
    sl = [slice(int(offs[i]), int(offs[i + 1])) for i in idx]
    rules = [p.update_rule for _, p in sorted(opt.target.namedparams())]

    def host(t):
        return t.detach().cpu().numpy().copy()

    def state_arrays(i, name):
        st = rules[i].state
        a = st[name]
        return a if isinstance(a, torch.Tensor) else torch.as_tensor(a, device='cuda')
    names = ('v',) if optimizer_name == 'momentum_sgd' else ('m', 'v')
    torch.cuda.synchronize()
    hp = [host(p_arena[s]) for s in sl]
    hs = [{nm: host(state_arrays(i, nm)).reshape(-1) for nm in names} for i in idx]
    t_now = int(actual.t)
    exact_order = world <= 2 or (comm._p2p is not None and not comm._mc_active(comm.gpu_buffer_a))
    exact = exact_order and adt is np.float32
    max_gerr = 0.0
    max_perr = 0.0
    ok = True
    detail = []
    perr_acc = [dict() for _ in idx]
    for j in range(steps):
        k = 1000 + j
        mine = [host(g_arenas[k % n_sets][s]) for s in sl]
        if world > 1:
            gathered = [None] * world
            dist.all_gather_object(gathered, mine)
        else:
            gathered = [mine]
        set_grads(k)
        opt.update()
        torch.cuda.synchronize()
        t_now += 1
        got_p = [host(p_arena[s]) for s in sl]
        got_g = [host(g_arenas[k % n_sets][s]) for s in sl] if write_grad else None
        mean = og.multi_node_mean_grad(gathered, odt)
        for q, (i, s) in enumerate(zip(idx, sl)):
            g = mean[q]
            if optimizer_name == 'momentum_sgd':
                og.momentum_sgd_update(hp[q], g, hs[q]['v'], 0.01, 0.9)
            else:
                og.adam_update_gpu(hp[q], g, hs[q]['m'], hs[q]['v'], t_now)
            if exact:
                same = np.array_equal(got_p[q].view(np.uint32), hp[q].view(np.uint32))
                if write_grad:
                    same = same and np.array_equal(got_g[q].view(np.uint32), g.view(np.uint32))
                if not same:
                    ok = False
                    detail.append('step %d tensor %s: bits differ' % (j, plist[i][0]))
                max_perr = max(max_perr, float(np.abs(got_p[q] - hp[q]).max()))
                continue
            # re-ordered sum: |mean - oracle mean| <= N * eps * sum_r |g_r| / N (+ the rounding of a
            # 16-bit buffer), accumulated into the parameters as MomentumSGD / Adam propagate it
            gmag = np.sum([np.abs(gathered[r][q]) for r in range(world)], axis=0) / world
            eps = 1.2e-7 if adt is np.float32 else (8e-3 if adt == 'bfloat16' else 1e-3)
            gerr = world * eps * gmag + (1e-12 if adt is np.float32 else world * 6e-8)
            if write_grad:
                d = np.abs(got_g[q] - g)
                max_gerr = max(max_gerr, float(d.max()))
                if (d > gerr).any():
                    ok = False
                    detail.append('step %d tensor %s: mean gradient off by %.3g' % (
                        j, plist[i][0], float(d.max())))
            acc = perr_acc[q]
            if optimizer_name == 'momentum_sgd':
                acc['v'] = 0.9 * acc.get('v', 0.0) + 0.01 * gerr
                acc['p'] = acc.get('p', 0.0) + acc['v']
                bound = 1.5 * acc['p'] + 4e-7 * np.abs(hp[q]) + 1e-9
            else:
                # Adam divides by sqrt(v) + eps: elements with |g| ~ eps amplify the difference
                bound = 2e-4 + 1e-6 * np.abs(hp[q])
            d = np.abs(got_p[q] - hp[q])
            max_perr = max(max_perr, float(d.max()))
            if (d > bound).any():
                ok = False
                detail.append('step %d tensor %s: parameter off by %.3g' % (j, plist[i][0],
                                                                           float(d.max())))
            # follow the device values so that the bound stays a per-step bound
            hp[q] = got_p[q].copy()
            for nm in names:
                hs[q][nm] = host(state_arrays(i, nm)).reshape(-1)
            perr_acc[q] = {}
    # replicas: every rank holds the same parameter bits
    replicas = True
    if world > 1:
        v = p_arena.view(torch.int32)
        sig = [int(v.sum(dtype=torch.int64).item()),
               int((v[::7].to(torch.int64) * 2654435761 % 1000003).sum().item())]
        sigs = [None] * world
        dist.all_gather_object(sigs, sig)
        replicas = all(s == sigs[0] for s in sigs)
    all_ok = [None] * world
    if world > 1:
        dist.all_gather_object(all_ok, bool(ok))
    else:
        all_ok = [ok]
    return {
        'ok': bool(all(all_ok) and replicas), 'checker': 'oracle/gradpath.py (NumPy restatement of the '
        'reference, pinned to reference outputs in tests/golden)', 'steps': steps,
        'tensors': len(idx), 'elements': int(sum(sizes[i] for i in idx)),
        'mode': 'bit-exact' if exact else 'rounding bound of a re-ordered sum (N*eps*sum|g_r|/N per '
                                          'step, propagated through the update)',
        'max_abs_err_param': max_perr, 'max_abs_err_mean_grad': max_gerr,
        'replicas_identical': bool(replicas), 'ranks_ok': all_ok, 'detail': detail[:4],
    }


def time_allreduce(torch, dist, comm, n, bsz, world):
    """The stand-alone reduction kernel of the transport in use.  NCCL-tests convention:
    algBW = S / t, busBW = algBW * 2(N-1)/N; `wire_gbs` = bytes that actually cross one
    NVLink direction of one GPU / t (S(N+1)/N for the in-switch reduction, 2S(N-1)/N for
    peer memory and rings)."""
    from chainer_b200 import nccl
    buf = comm.gpu_buffer_a
    type_id = 7 if bsz == 4 else 6
    dt = np.float32 if bsz == 4 else np.float16
    mc = comm._p2p is not None and comm._mc_active(buf)
    if mc:
        def one():
            comm._p2p.mc_allreduce(dt, 0, n, None)
    elif comm._p2p is not None:
        comm._p2p.ensure(buf)

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
    wire = S * ((world + 1.0) / world if mc else 2.0 * (world - 1) / world)
    return {'impl': _allreduce_impl(comm), 'bytes': S, 'us': us, 'alg_gbs': alg, 'bus_gbs': busbw,
            'bus_frac_of_900': busbw / NVLINK_GBS, 'wire_bytes_per_direction': wire,
            'wire_gbs': wire / us / 1e3, 'wire_frac_of_900': wire / us / 1e3 / NVLINK_GBS}


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


def time_train(torch, dist, rank, world, comm, args, batch=32, steps=12, warmup=5):
    """The "ResNet-50 img/s" half of the BASELINE metric (configs[1]): a torchvision
    resnet50 stand-in (random init, synthetic 224x224 batch 32 per GPU) runs forward and
    backward -- the part the north_star leaves on the framework's own path -- and hands its
    `.grad` arrays BY POINTER to create_multi_node_optimizer(MomentumSGD, pure_nccl).update()
    (train_imagenet.py:151-200).  img/s = 32 * N / step time; three device-timed numbers:
    forward+backward alone, the synchronous step (fwd/bwd then the gradient path), and the
    reference's double-buffering mode (the exchange of step k overlaps fwd/bwd of step k+1,
    chainermn/optimizers.py:59-146)."""
    import torchvision
    import chainer_b200
    from chainer_b200.core.link import link_from_named_arrays
    torch.manual_seed(7)
    net = torchvision.models.resnet50(weights=None).cuda()      # contiguous (NCHW) parameters
    net.train()
    x = torch.randn(batch, 3, 224, 224, device='cuda').to(memory_format=torch.channels_last)
    y = torch.randint(0, 1000, (batch,), device='cuda')
    named = [(nm.replace('.', '/'), p) for nm, p in net.named_parameters()]
    n_params = sum(p.numel() for _, p in named)

    def fwd_bwd():
        with torch.autocast('cuda', dtype=torch.bfloat16):
            loss = torch.nn.functional.cross_entropy(net(x), y)
        loss.backward()
        return loss

    out = {'model': 'torchvision resnet50 stand-in ({} tensors, {} parameters), batch {}/GPU, '
                    'synthetic 224x224 (channels_last input), bf16 autocast forward/backward, fp32 '
                    'parameters, gradients and MomentumSGD state'.format(len(named), n_params, batch),
           'batch_per_gpu': batch, 'steps': steps}

    # forward + backward captured ONCE into a CUDA graph (static input, gradients written into
    # the same arrays at every replay): the stand-in then costs device time only, instead of
    # ~10 ms of framework host overhead per step that would hide the path being measured
    graph = None
    try:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                for _, p in named:
                    p.grad = None
                fwd_bwd()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        for _, p in named:
            p.grad = None
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            static_loss = fwd_bwd()
        graph.replay()
        torch.cuda.synchronize()
        assert bool(torch.isfinite(static_loss).item())
    except Exception as e:      # noqa: BLE001 -- eager numbers below still stand
        graph = None
        out['graph_error'] = '%s: %s' % (type(e).__name__, e)

    def run(mode, graphed):
        model = link_from_named_arrays([('/' + nm, p.data) for nm, p in named])
        plink = [p for _, p in sorted(model.namedparams())]
        tparam = [p for _, p in sorted((('/' + nm), p) for nm, p in named)]
        actual = chainer_b200.MomentumSGD(lr=0.01, momentum=0.9)
        opt = chainer_b200.create_multi_node_optimizer(actual, comm,
                                                       double_buffering=(mode == 'double_buffering'))
        opt.setup(model)

        def one(update=True):
            if graphed:
                graph.replay()
            else:
                for tp in tparam:
                    tp.grad = None                        # cleargrads(): new gradient arrays
                fwd_bwd()
            if update:
                for lp, tp in zip(plink, tparam):
                    lp.grad = tp.grad
                opt.update()
        for _ in range(warmup):
            one(mode != 'fwd_bwd')
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            one(mode != 'fwd_bwd')
        e1.record()
        torch.cuda.synchronize()
        if hasattr(opt, 'wait'):
            opt.wait()
        ms = e0.elapsed_time(e1) / steps
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        assert bool(torch.isfinite(next(iter(net.parameters()))).all().item())
        return ms

    def leg(graphed, suffix):
        fb = run('fwd_bwd', graphed)
        sync = run('sync', graphed)
        res = {'fwd_bwd_ms' + suffix: fb, 'step_ms' + suffix: sync, 'gradpath_ms' + suffix: sync - fb,
               'gradpath_share' + suffix: (sync - fb) / sync,
               'img_per_s' + suffix: batch * world / (sync * 1e-3),
               'img_per_s_fwd_bwd_only' + suffix: batch * world / (fb * 1e-3)}
        if not graphed:
            try:
                db = run('double_buffering', graphed)
                res.update({'step_ms_double_buffering' + suffix: db,
                            'img_per_s_double_buffering' + suffix: batch * world / (db * 1e-3)})
            except Exception as e:      # noqa: BLE001
                res['double_buffering_error'] = '%s: %s' % (type(e).__name__, e)
        return res
    if graph is not None:
        out.update(leg(True, ''))
        out['fwd_bwd'] = 'CUDA graph replay (device-bound)'
        out.update(leg(False, '_eager'))
    else:
        out.update(leg(False, ''))
        out['fwd_bwd'] = 'eager (host-bound)'
    return out


def time_fp16_buffer(torch, dist, world, comm, step, n, optimizer_name, write_grad, steps):
    """The default workload with `allreduce_grad_dtype=float16` (the packed buffer and the
    exchange in half precision, cast fused into pack / update) on the same communicator."""
    comm.set_config('allreduce_grad_dtype', np.float16)
    try:
        for k in range(5):
            step(k)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for k in range(steps):
            step(k)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
    finally:
        comm.set_config('allreduce_grad_dtype', np.float32)
        step(0)                       # back on the float32 buffer before the later legs
        torch.cuda.synchronize()
    pack_b, upd_b = bytes_per_elem(optimizer_name, 2, write_grad)
    return {'ms_per_step': ms, 'steps': steps, 'bytes_per_elem': pack_b + upd_b,
            'value': world * n * (pack_b + upd_b) / (ms * 1e-3) / 1e9, 'unit': 'GB/s',
            'packed_bytes': n * 2}


def time_mnbn(torch, dist, world, comm, lib, batch=32):
    """BASELINE configs[2]: the MultiNodeBatchNormalization statistics of one ResNet-50
    step -- 53 layers, forward [mean | E x^2] and backward [sum gy | sum gy x_hat], each
    followed by the exchange of its 2C floats over the N ranks -- through the functions the
    link calls (chainer_b200/functions/batch_normalization.py)."""
    from chainer_b200 import workloads
    from chainer_b200.functions.batch_normalization import _NcclImpl
    impl = _NcclImpl(comm)
    layers = workloads.resnet50_bn_layers(batch)
    uniq = {}
    for _, s in layers:
        if s not in uniq:
            x = torch.randn(*s, device='cuda')
            gy = torch.randn(*s, device='cuda') * 1e-3
            gamma = torch.ones(s[1], device='cuda')
            mean, var = impl.get_mean_and_var(None, gamma, x)
            uniq[s] = (x, gy, gamma, mean, torch.rsqrt(var + 2e-5))

    def one_step():
        for _, s in layers:
            x, gy, gamma, mean, inv_std = uniq[s]
            impl.get_mean_and_var(None, gamma, x)
        for _, s in reversed(layers):
            x, gy, gamma, mean, inv_std = uniq[s]
            impl.get_ggamma_and_gbeta_from_x(None, gamma, gy, x, mean, inv_std)

    def timed(fn, reps):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        t_enq = time.perf_counter() - t0
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / reps
        if world > 1:
            t = torch.tensor([us], dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            us = float(t.item())
        return us, 1e6 * t_enq / reps
    for _ in range(3):
        one_step()
    c0 = lib.launches
    reps = 10
    us, enq = timed(one_step, reps)
    n_launch = (lib.launches - c0) // reps
    nbytes = sum(int(np.prod(s)) * 4 for _, s in layers)
    out = {'layers': len(layers), 'us_per_step': us, 'launches_per_step': n_launch,
           'host_enqueue_us_per_step': enq, 'activation_bytes_read': 3 * nbytes,
           'note': 'statistics of 53 BN layers forward + backward at batch %d incl. the exchange of '
                   '2C floats per layer and direction over %d rank(s) (one kernel per layer and '
                   'direction); eager calls are host-enqueue-bound (us_per_step ~ '
                   'host_enqueue_us_per_step); the kernels count their own epochs, so the same 106 '
                   'launches replay from ONE CUDA graph: us_per_step_graph' % (batch, world)}
    # the same 106 launches captured once into a CUDA graph (the exchange kernels count their
    # epochs in device memory, so a replay is a valid collective step)
    try:
        side = torch.cuda.Stream()
        impl.stream = side.cuda_stream
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            one_step()
        side.synchronize()
        if world > 1:
            dist.barrier()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            one_step()
        g.replay()
        us_g, _ = timed(g.replay, reps)
        out['us_per_step_graph'] = us_g
    except Exception as e:      # noqa: BLE001
        out['graph_error'] = '%s: %s' % (type(e).__name__, e)
    finally:
        impl.stream = 0
    return out


_WATCHDOG = None


def _watchdog_fired():
    import faulthandler
    sys.stderr.write('bench.py: no result after BENCH_WATCHDOG_S seconds; stacks follow\n')
    faulthandler.dump_traceback(file=sys.stderr, all_threads=True)
    sys.stderr.flush()
    os._exit(1)


def main():
    args = parse_args()
    # a bench that stops making progress must say where and end, not sit on the GPU box:
    # after BENCH_WATCHDOG_S seconds (default 600) every thread's stack goes to stderr and the
    # process exits (b200_main replaces this by the legs' own deadline once its line is safe)
    global _WATCHDOG
    if args.impl == 'b200':
        _WATCHDOG = threading.Timer(float(os.environ.get('BENCH_WATCHDOG_S', '600')),
                                    _watchdog_fired)
        _WATCHDOG.daemon = True
        _WATCHDOG.start()
    if args.impl == 'reference':
        reference_main(args)
    elif args.impl == 'reference-worker':
        reference_worker_main(args)
    else:
        b200_main(args)


if __name__ == '__main__':
    main()
