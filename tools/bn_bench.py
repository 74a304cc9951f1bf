"""MultiNodeBatchNormalization statistics kernels at the ResNet-50 layer shapes
(BASELINE config 3): time, GB/s against the measured copy peak, and the same
statistics computed the reference's way with torch ops (x.mean, square(x).mean).

    python tools/bn_bench.py [--out gpurun_out/bn_bench.json]
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import chainer_b200
    from chainer_b200 import workloads
    from chainer_b200.functions.batch_normalization import _NcclImpl
    ap = argparse.ArgumentParser()
    ap.add_argument('--out', default='gpurun_out/bn_bench.json')
    ap.add_argument('--batch', type=int, default=32)
    args = ap.parse_args()
    peak = 6462.1
    try:
        peak = json.load(open('MEASURED_PEAKS.json'))['hbm_gbs']
    except Exception:
        pass
    comm = chainer_b200.create_communicator('pure_nccl')
    from chainer_b200 import _lib
    lib = _lib.get()
    if os.environ.get('BN_CTAS'):
        _lib.get().gp_set_tuning(b'bn_ctas_per_sm', int(os.environ['BN_CTAS']))
    if os.environ.get('BN_THREADS'):
        _lib.get().gp_set_tuning(b'bn_threads', int(os.environ['BN_THREADS']))
    impl = _NcclImpl(comm)
    shapes = sorted(set(s for _, s in workloads.resnet50_bn_layers(args.batch)), key=lambda s: -s[1] * s[2] * s[3])
    counts = {}
    for _, s in workloads.resnet50_bn_layers(args.batch):
        counts[s] = counts.get(s, 0) + 1
    side = torch.cuda.Stream()

    def timeit(fn, reps=12):
        """Device time per call: `reps` calls (rotating input sets: cold DRAM reads)
        are captured into a CUDA graph on a side stream and the replay is timed, so
        that host launch overhead (~20 us of Python per call, more than the small
        kernels take) is not part of the number."""
        torch.cuda.synchronize()
        with torch.cuda.stream(side):
            for i in range(3):
                fn(i, side.cuda_stream)
        side.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            for i in range(reps):
                fn(i, side.cuda_stream)
        g.replay()
        torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            g.replay()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3 / reps)
        ts.sort()
        return ts[len(ts) // 2]

    rows = []
    tot_ours = tot_ref = 0.0
    for s in shapes:
        nbytes = int(np.prod(s)) * 4
        n_sets = max(2, int(400e6 // nbytes) + 1)
        xs = [torch.randn(*s, device='cuda') for _ in range(n_sets)]
        gys = [torch.randn(*s, device='cuda') * 1e-3 for _ in range(n_sets)]
        gamma = torch.ones(s[1], device='cuda')
        mean, var = impl.get_mean_and_var(None, gamma, xs[0])
        inv_std = torch.rsqrt(var + 2e-5)
        C, N, HW = s[1], s[0], s[2] * s[3]
        outs = [torch.empty(2 * C, device='cuda') for _ in range(n_sets)]
        ws = torch.zeros(lib.gp_bn_workspace_bytes(C), dtype=torch.uint8, device='cuda')

        def our_fwd(i, st):
            k = i % n_sets
            lib.gp_bn_fwd_mean_var(xs[k].data_ptr(), 7, N, C, HW, outs[k].data_ptr(), 7,
                                   ws.data_ptr(), st)

        def our_bwd(i, st):
            k = i % n_sets
            lib.gp_bn_bwd_stats(gys[k].data_ptr(), 7, xs[k].data_ptr(), 7, mean.data_ptr(),
                                inv_std.data_ptr(), 7, N, C, HW, outs[k].data_ptr(), 7,
                                ws.data_ptr(), st)
        fwd, bwd = timeit(our_fwd), timeit(our_bwd)

        def ref_fwd(i, st):
            x = xs[i % n_sets]
            m = x.mean(dim=(0, 2, 3))
            q = torch.square(x).mean(dim=(0, 2, 3))
            return m, q - m * m

        def ref_bwd(i, st):
            x, gy = xs[i % n_sets], gys[i % n_sets]
            xh = (x - mean.view(1, -1, 1, 1)) * inv_std.view(1, -1, 1, 1)
            return gy.sum(dim=(0, 2, 3)), (gy * xh).sum(dim=(0, 2, 3))
        rf, rb = timeit(ref_fwd), timeit(ref_bwd)
        row = dict(shape=list(s), layers=counts[s], mbytes=nbytes / 1e6, fwd_us=fwd, bwd_us=bwd,
                   fwd_gbs=nbytes / fwd / 1e3, bwd_gbs=2 * nbytes / bwd / 1e3,
                   fwd_frac=nbytes / fwd / 1e3 / peak, bwd_frac=2 * nbytes / bwd / 1e3 / peak,
                   torch_fwd_us=rf, torch_bwd_us=rb)
        rows.append(row)
        tot_ours += counts[s] * (fwd + bwd)
        tot_ref += counts[s] * (rf + rb)
        print('%-20s x%2d %6.1f MB | fwd %6.1f us %5.0f GB/s (%.2f) | bwd %6.1f us %5.0f GB/s (%.2f) | torch-op '
              'restatement fwd %6.1f bwd %6.1f us' % (s, counts[s], nbytes / 1e6, fwd, row['fwd_gbs'],
                                                       row['fwd_frac'], bwd, row['bwd_gbs'], row['bwd_frac'],
                                                       rf, rb), flush=True)
    print('all 53 BN layers, fwd+bwd statistics per step: %.0f us (library) vs %.0f us (torch-op restatement '
          'of the reference sequence)' % (tot_ours, tot_ref))
    os.makedirs(os.path.dirname(args.out) or '.', exist_ok=True)
    json.dump(dict(peak=peak, batch=args.batch, rows=rows, total_us=tot_ours, torch_total_us=tot_ref,
                   note='device time per call from CUDA-graph replays of the C-ABI calls (one rank: one '
                        'kernel per call); inputs rotate over enough sets to exceed L2'), open(args.out, 'w'), indent=1)


if __name__ == '__main__':
    main()
