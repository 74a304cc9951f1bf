"""CPU oracle for the ChainerMN gradient path -- TEST INFRASTRUCTURE ONLY.

This package is a NumPy restatement of the reference's algorithm for the hot
path (chainer/chainer v7.8.1): pack/unpack layout, allreduce-mean, MomentumSGD /
Adam (and SGD / CorrectedMomentumSGD / NesterovAG) updates, the WeightDecay /
GradientClipping hooks and loss scaling, MultiNodeBatchNormalization statistics,
and the CPU `naive` communicator + NumPy update that serves as the reported CPU
baseline.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline``
/ ``--impl reference`` legs may import it; the product package ``chainer_b200``
never does (it fails loudly when ``libgradpath.so`` is missing).

PARITY PINNING
--------------
The reference's GPU path (CuPy NVRTC kernels + cupy.cuda.nccl) cannot run in
this environment (no CuPy, no mpi4py, CuPy 7 predates sm_100).  The oracle is
pinned instead against

  * the UNMODIFIED reference CPU implementation, imported from
    /root/reference with NumPy-2 shims by ``tests/golden/make_golden.py``
    (MomentumSGDRule / AdamRule ``update_core_cpu``, NaiveCommunicator
    ``multi_node_mean_grad`` through an mpi4py stand-in,
    GeneralBatchNormalizationImpl + ``_MpiImpl`` statistics,
    ``sorted(model.namedparams())`` layouts of the example models, the
    ``WeightDecay`` / ``GradientClipping`` hooks with static loss scaling, SGD /
    CorrectedMomentumSGD / NesterovAG, ``use_fp32_update`` master weights and the
    dynamic loss-scale schedule -- all reproduced bit for bit); the
    generated vectors are committed under ``tests/golden/`` and checked by
    ``tests/test_oracle_golden.py``;
  * the known-answer tests of the reference's own test-suite
    (``tests/chainer_tests/optimizers_tests/test_optimizers.py:278-366`` AdamW
    0.9495 and the AMSGrad vectors; ``tests/chainermn_tests/communicator_tests/
    test_communicator.py:252-319`` rank-filled gradient means).

Unpinned (no executable reference exists): bfloat16 as allreduce dtype (the
reference maps only float16/32/64, ``_communication_utility.py:177-186``) --
defined here as round-to-nearest-even of the float32 restatement; the summation
order inside NCCL / CuPy reductions and FMA contraction inside the NVRTC-built
kernels (third-party, tolerance-level parity only, as the reference's own tests
use: atol 1e-5 / rtol 1e-4, ``chainer/testing/array.py:10``).
"""
