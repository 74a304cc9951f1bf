"""Host-logic tests on CPU: the product's Python layer driven against the
oracle-backed double of libgradpath (tests/fake_lib.py).  They mirror the
reference's own tests of this path:

* tests/chainermn_tests/communicator_tests/test_communicator.py:191-202, 242-319,
  346-425, 453-461, 494-512, 1004-1022
* tests/chainermn_tests/optimizer_tests/test_multi_node_optimizer.py:20-235
* tests/chainermn_tests/optimizer_tests/test_double_buffering_optimizer.py:23-98
"""
import numpy as np
import pytest

import chainer_b200
from chainer_b200 import config
from chainer_b200 import device as _dev
from chainer_b200.communicators import _control_plane
from chainer_b200.communicators.pure_nccl_communicator import PureNcclCommunicator
from chainer_b200.core import link as L
from oracle import gradpath as og
from tests import fake_lib
from tests.helpers import assert_bits_equal


@pytest.fixture
def fake(monkeypatch):
    """The oracle-backed double of the library.  The tests below assert on the names of
    the SEPARATE launches (gp_pack / gp_unpack_*), so the one-launch step is switched off
    here; test_one_launch_step_* cover it."""
    monkeypatch.setenv('CHAINER_B200_STEP', '0')
    f, prev = fake_lib.install()
    _control_plane.reset_world()
    yield f
    fake_lib.uninstall(prev)
    config.set_debug(False)
    config.set_dtype(None)


class Linear(L.Link):
    """Stand-in of chainer.links.Linear(in, out): W (out, in) and b (out,);
    in_size None leaves W uninitialised like the reference's lazy Linear."""

    def __init__(self, in_size, out_size, dtype=np.float32):
        super(Linear, self).__init__()
        with self.init_scope():
            self.W = L.Parameter(None if in_size is None else
                                 np.zeros((out_size, in_size), dtype=dtype))
            self.b = L.Parameter(np.zeros((out_size,), dtype=dtype))


class ExampleModel(L.Chain):
    """tests/chainermn_tests/communicator_tests/test_communicator.py:30-41"""

    def __init__(self, dtype=np.float32):
        super(ExampleModel, self).__init__()
        with self.init_scope():
            self.a = Linear(2, 3, dtype)
            self.b = Linear(3, 4, dtype)
            self.c = Linear(None, 5, dtype)


class ExampleMixedModel(L.Chain):
    """test_communicator.py:44-57: float16 / float32 layers alternate."""

    def __init__(self):
        super(ExampleMixedModel, self).__init__()
        with self.init_scope():
            self.a = Linear(2, 3, np.float16)
            self.b = Linear(3, 4, np.float32)
            self.c = Linear(4, 5, np.float16)
            self.d = Linear(5, 6, np.float32)


def _fill_grads(model, rank):
    model.a.W.grad = np.full_like(model.a.W.data, rank)
    model.a.b.grad = np.full_like(model.a.b.data, rank)
    model.b.W.grad = np.full_like(model.b.W.data, rank + 1)
    model.b.b.grad = np.full_like(model.b.b.data, rank + 1)
    model.c.b.grad = np.full_like(model.c.b.data, rank + 2)


# --------------------------------------------------------------- config API --
def test_create_communicator_and_config(fake):
    comm = chainer_b200.create_communicator('pure_nccl')
    assert isinstance(comm, PureNcclCommunicator)
    assert (comm.rank, comm.size, comm.intra_rank, comm.intra_size, comm.inter_rank,
            comm.inter_size) == (0, 1, 0, 1, 0, 1)
    assert comm.get_config('batched_copy') is True
    assert comm.get_config('allreduce_grad_dtype') is None
    comm.set_config('batched_copy', False)
    assert comm.get_config('batched_copy') is False
    comm.set_config('allreduce_grad_dtype', np.float16)
    assert comm.get_config('allreduce_grad_dtype') == np.float16
    comm.set_config('allreduce_grad_dtype', None)
    assert comm.get_config('allreduce_grad_dtype') is None
    with pytest.raises(ValueError):
        comm.set_config('allreduce_grad_dtype', np.int32)       # test_communicator.py:494-512
    with pytest.raises(ValueError):
        comm.set_config('no_such_config')
    with pytest.raises(KeyError):
        comm.get_config('no_such_config')
    comm._configs['foobar'] = 1                                  # test_communicator.py:1015-1022
    assert comm.get_config('foobar') == 1
    del comm._configs['foobar']
    with pytest.raises(ValueError):
        chainer_b200.create_communicator('no_such_communicator')
    with pytest.raises(ValueError):
        chainer_b200.create_communicator('naive', allreduce_grad_dtype=np.float16)
    comm._init_comms()
    comm.finalize()
    assert comm.nccl_comm is None


def test_library_missing_fails_loudly(monkeypatch, tmp_path):
    from chainer_b200 import _lib
    prev = _lib.set_backend_for_testing(None)
    monkeypatch.setenv('CHAINER_B200_LIBGRADPATH', str(tmp_path / 'nope.so'))
    try:
        with pytest.raises(ImportError):
            _lib.load()
        with pytest.raises(ImportError):
            chainer_b200.create_communicator('pure_nccl')
    finally:
        _lib.set_backend_for_testing(prev)


def test_host_arrays_rejected_by_real_backend():
    """Without the test double, NumPy arrays are an unsupported array module
    (reference: _memory_utility.py:50-52)."""
    from chainer_b200 import device
    with pytest.raises(ValueError):
        device.device_ptr(np.zeros(3, np.float32))


# ------------------------------------------------------ multi_node_mean_grad --
def test_multi_node_mean_grad_single_rank(fake):
    """check_multi_node_mean_grad (test_communicator.py:252-269) at size 1, twice."""
    comm = chainer_b200.create_communicator('pure_nccl')
    model = ExampleModel()
    for _ in range(2):
        _fill_grads(model, comm.rank)
        comm.multi_node_mean_grad(model)
        base = (comm.size - 1.0) / 2
        np.testing.assert_allclose(model.a.W.grad, (base + 0) * np.ones((3, 2)))
        np.testing.assert_allclose(model.a.b.grad, (base + 0) * np.ones((3,)))
        np.testing.assert_allclose(model.b.W.grad, (base + 1) * np.ones((4, 3)))
        np.testing.assert_allclose(model.b.b.grad, (base + 1) * np.ones((4,)))
        np.testing.assert_allclose(model.c.b.grad, (base + 2) * np.ones((5,)))
    assert model.c.W.grad is None                      # uninitialised W is skipped


def test_multi_node_mean_grad_empty_and_zero_fill(fake):
    """check_multi_node_mean_grad_empty / _empty_half (test_communicator.py:272-319)."""
    comm = chainer_b200.create_communicator('pure_nccl')
    model = ExampleModel()
    _fill_grads(model, 0)
    model.c.b.grad = None
    comm.multi_node_mean_grad(model, zero_fill=False)
    assert model.c.b.grad is None                      # skipped, not created
    comm.multi_node_mean_grad(model, zero_fill=True)
    assert model.c.b.grad is not None                  # created as zeros
    np.testing.assert_array_equal(model.c.b.grad, np.zeros(5, np.float32))


def test_layout_is_sorted_namedparams(fake):
    comm = chainer_b200.create_communicator('pure_nccl')
    model = ExampleModel()
    rng = np.random.default_rng(0)
    for _, p in model.namedparams():
        if p.data is not None:
            p.grad = rng.standard_normal(p.data.shape).astype(np.float32)
    names = [n for n, p in sorted(model.namedparams()) if p.data is not None]
    assert names == ['/a/W', '/a/b', '/b/W', '/b/b', '/c/b']
    want = og.pack([p.grad for _, p in sorted(model.namedparams()) if p.data is not None],
                   np.float32)
    comm.multi_node_mean_grad(model)
    from tests.fake_lib import _view
    got = _view(comm.gpu_buffer_a.ptr(), want.size, np.float32)
    assert_bits_equal(np.array(got), want, 'packed layout')


@pytest.mark.parametrize('global_dtype,allreduce_dtype,expected', [
    # the table of create_communicator's docstring (communicators/__init__.py:43-53)
    ('float32', None, og.NCCL_FLOAT32), ('float32', np.float16, og.NCCL_FLOAT16),
    ('float32', np.float32, og.NCCL_FLOAT32), ('float16', None, og.NCCL_FLOAT16),
    ('float16', np.float32, og.NCCL_FLOAT32), ('mixed16', None, og.NCCL_FLOAT16),
    ('mixed16', np.float32, og.NCCL_FLOAT32), (None, np.float64, og.NCCL_FLOAT64),
    (None, 'bfloat16', og.NCCL_BFLOAT16),
])
@pytest.mark.parametrize('batched_copy', [True, False])
def test_mixed_dtype_model_and_nccl_dtype(fake, global_dtype, allreduce_dtype, expected,
                                          batched_copy):
    """test_communicator.py:346-425: the dtype id handed to nccl_comm.allReduce."""
    config.set_dtype(global_dtype)
    comm = chainer_b200.create_communicator('pure_nccl', allreduce_grad_dtype=allreduce_dtype,
                                            batched_copy=batched_copy)
    comm._init_comms()
    seen = []
    real = comm.nccl_comm.allReduce

    def spy(sendbuf, recvbuf, count, datatype, op, stream):
        seen.append(datatype)
        return real(sendbuf, recvbuf, count, datatype, op, stream)
    comm.nccl_comm.allReduce = spy
    model = ExampleMixedModel()
    for k, lk in enumerate([model.a, model.b, model.c, model.d]):
        lk.W.grad = np.full_like(lk.W.data, k + 1)
        lk.b.grad = np.full_like(lk.b.data, k + 1)
    comm.multi_node_mean_grad(model)
    assert seen == [expected]
    for k, lk in enumerate([model.a, model.b, model.c, model.d]):
        assert lk.W.grad.dtype == lk.W.data.dtype
        np.testing.assert_allclose(lk.W.grad.astype(np.float64), k + 1)
        np.testing.assert_allclose(lk.b.grad.astype(np.float64), k + 1)


def test_non_float_gradient_rejected(fake):
    comm = chainer_b200.create_communicator('pure_nccl')
    model = ExampleModel()
    _fill_grads(model, 0)
    model.a.b.grad = np.zeros(3, np.int32)
    with pytest.raises(ValueError):
        comm.multi_node_mean_grad(model)


def test_debug_mode_detects_divergence(fake):
    """test_communicator.py:453-461: NaN gradients + debug -> 'diverged'."""
    comm = chainer_b200.create_communicator('pure_nccl')
    model = ExampleModel()
    _fill_grads(model, 0)
    model.b.W.grad[1, 2] = np.nan
    config.set_debug(True)
    with pytest.raises(ValueError, match='.* diverged .*'):
        comm.multi_node_mean_grad(model)
    config.set_debug(False)
    comm.multi_node_mean_grad(model)                   # no check without debug


def test_bcast_data_single_rank(fake):
    comm = chainer_b200.create_communicator('pure_nccl')
    model = ExampleModel()
    model.a.W.data[...] = 3
    comm.bcast_data(model)
    np.testing.assert_array_equal(model.a.W.data, np.full((3, 2), 3, np.float32))
    names = [c[0] for c in fake.calls]
    assert 'gp_pack' in names and 'gp_unpack_scale' in names


def test_multi_node_mean_nccl_on_device_memory(fake):
    """_multi_node_mean_nccl(sendbuf, recvbuf, n_elems, dtype) as MNBN calls it
    (chainermn/functions/batch_normalization.py:57-60), stream=None."""
    from chainer_b200.communicators._memory_utility import DeviceMemory
    comm = chainer_b200.create_communicator('pure_nccl')
    a, b = DeviceMemory(), DeviceMemory()
    a.assign(32)
    b.assign(32)
    from tests.fake_lib import _view
    _view(a.ptr(), 8, np.float32)[...] = np.arange(8)
    comm._multi_node_mean_nccl(a, b, 8, np.dtype(np.float32))
    np.testing.assert_array_equal(_view(b.ptr(), 8, np.float32), np.arange(8, dtype=np.float32))


# ------------------------------------------------------- multi-node optimizer --
def _model_with_values(seed=0, dtype=np.float32):
    model = ExampleModel(dtype)
    rng = np.random.default_rng(seed)
    for _, p in sorted(model.namedparams()):
        if p.data is not None:
            p.data[...] = rng.standard_normal(p.data.shape).astype(dtype)
    return model


def _set_grads(model, seed):
    rng = np.random.default_rng(seed)
    for _, p in sorted(model.namedparams()):
        if p.data is not None:
            p.grad = (rng.standard_normal(p.data.shape) * 1e-2).astype(p.data.dtype)


@pytest.mark.parametrize('opt_name', ['momentum_sgd', 'adam'])
@pytest.mark.parametrize('fused', [True, False])
def test_multi_node_optimizer_protocol(fake, opt_name, fused):
    """test_multi_node_optimizer.py:51-110: first update() only broadcasts
    (t == 0), the second updates (t == 1, every rule.t == 1) and param.grad holds
    the mean; results equal the oracle whether or not the fused path is taken."""
    comm = chainer_b200.create_communicator('pure_nccl')
    model = _model_with_values()
    ref = _model_with_values()
    make = (lambda: chainer_b200.MomentumSGD(lr=0.1, momentum=0.9)) if opt_name == 'momentum_sgd' \
        else (lambda: chainer_b200.Adam(alpha=0.01))
    actual = make()
    opt = chainer_b200.create_multi_node_optimizer(actual, comm)
    opt.setup(model)
    if not fused:
        actual.add_hook(lambda o: None, name='noop')          # any hook disables fusion
    _set_grads(model, 1)
    opt.update()
    assert actual.t == 0
    assert all(p.update_rule.t == 0 for p in model.params())
    np.testing.assert_array_equal(model.a.W.data, ref.a.W.data)   # bcast only

    states = {}
    for step in range(1, 4):
        _set_grads(model, 10 + step)
        _set_grads(ref, 10 + step)
        fake.calls[:] = []
        opt.update()
        names = [c[0] for c in fake.calls]
        if fused:
            assert ('gp_unpack_momentum_sgd' if opt_name == 'momentum_sgd' else 'gp_unpack_adam') in names
            assert 'gp_unpack_scale' not in names
            assert names.count('gp_pack') == 1
        else:
            assert 'gp_unpack_scale' in names
        assert actual.t == step
        for name, p in sorted(model.namedparams()):
            assert p.update_rule.t == step
        for (name, p), (_, q) in zip(sorted(model.namedparams()), sorted(ref.namedparams())):
            if p.data is None:
                continue
            st = states.setdefault(name, {k: np.zeros_like(q.data) for k in ('v', 'm')})
            g = q.grad
            if opt_name == 'momentum_sgd':
                og.momentum_sgd_update(q.data, g, st['v'], 0.1, 0.9)
            else:
                og.adam_update_gpu(q.data, g, st['m'], st['v'], step, alpha=0.01)
            assert_bits_equal(p.data, q.data, name)
            assert_bits_equal(p.grad, g, name + ' grad')       # grad observable: the mean


def test_fused_and_unfused_paths_agree_bitwise(fake):
    results = []
    for fused in (True, False):
        comm = chainer_b200.create_communicator('pure_nccl', allreduce_grad_dtype=np.float16)
        model = _model_with_values(3)
        actual = chainer_b200.MomentumSGD(lr=0.05, momentum=0.8)
        opt = chainer_b200.create_multi_node_optimizer(actual, comm)
        opt.setup(model)
        if not fused:
            actual.add_hook(lambda o: None, name='noop')
        opt.update()
        for step in range(3):
            _set_grads(model, 40 + step)
            opt.update()
        results.append([p.data.copy() for _, p in sorted(model.namedparams()) if p.data is not None])
    for a, b in zip(*results):
        assert_bits_equal(a, b, 'fused vs unfused')


def test_per_parameter_hyperparameters_are_honoured(fake):
    """UpdateRule hyperparameters override the optimizer's (optimizer.py:92-145)."""
    comm = chainer_b200.create_communicator('pure_nccl')
    model = _model_with_values(5)
    ref = _model_with_values(5)
    actual = chainer_b200.MomentumSGD(lr=0.1, momentum=0.9)
    opt = chainer_b200.create_multi_node_optimizer(actual, comm)
    opt.setup(model)
    model.a.W.update_rule.hyperparam.lr = 0.5
    model.b.b.update_rule.enabled = True
    opt.update()
    _set_grads(model, 77)
    _set_grads(ref, 77)
    opt.update()
    for (name, p), (_, q) in zip(sorted(model.namedparams()), sorted(ref.namedparams())):
        if p.data is None:
            continue
        v = np.zeros_like(q.data)
        og.momentum_sgd_update(q.data, q.grad, v, 0.5 if name == '/a/W' else 0.1, 0.9)
        assert_bits_equal(p.data, q.data, name)


def test_disabled_rule_falls_back_and_skips_update(fake):
    comm = chainer_b200.create_communicator('pure_nccl')
    model = _model_with_values(6)
    actual = chainer_b200.MomentumSGD(lr=0.1)
    opt = chainer_b200.create_multi_node_optimizer(actual, comm)
    opt.setup(model)
    model.a.W.update_rule.enabled = False
    before = model.a.W.data.copy()
    opt.update()
    _set_grads(model, 5)
    opt.update()
    np.testing.assert_array_equal(model.a.W.data, before)
    assert model.a.W.update_rule.t == 0
    assert model.a.b.update_rule.t == 1


def test_dynamic_model_rebroadcasts(fake):
    """test_multi_node_optimizer.py:122-235: a new link -> setup again -> the
    next update() broadcasts (t back to 0)."""
    comm = chainer_b200.create_communicator('pure_nccl')
    model = _model_with_values(8)
    actual = chainer_b200.MomentumSGD(lr=0.1)
    opt = chainer_b200.create_multi_node_optimizer(actual, comm)
    opt.setup(model)
    _set_grads(model, 1)
    opt.update()
    assert actual.t == 0
    _set_grads(model, 2)
    opt.update()
    assert actual.t == 1
    with model.init_scope():
        model.d = Linear(4, 4)
    opt.setup(model)
    _set_grads(model, 3)
    opt.update()
    assert actual.t == 0                               # broadcast only
    _set_grads(model, 4)
    opt.update()
    assert actual.t == 1
    assert model.d.W.update_rule.t == 1


def test_double_buffering_optimizer(fake):
    """test_double_buffering_optimizer.py:43-90: means land in
    communicated_target one call late; t increments one call late."""
    comm = chainer_b200.create_communicator('pure_nccl')
    model = _model_with_values(9)
    actual = chainer_b200.MomentumSGD(lr=0.1)
    opt = chainer_b200.create_multi_node_optimizer(actual, comm, double_buffering=True)
    opt.setup(model)
    _fill_grads(model, 0)
    opt.update()                                       # bcast + deep copy
    assert actual.t == 0 and opt.communicated_target is not None
    _fill_grads(model, 0)
    opt.update()                                       # swap, async mean, no update yet
    assert actual.t == 0
    opt.wait()
    ct = opt.communicated_target
    np.testing.assert_allclose(ct.a.W.grad, 0 * np.ones((3, 2)))
    np.testing.assert_allclose(ct.b.W.grad, 1 * np.ones((4, 3)))
    np.testing.assert_allclose(ct.c.b.grad, 2 * np.ones((5,)))
    _fill_grads(model, 0)
    opt.update()
    assert actual.t == 1
    with pytest.raises(ValueError):
        class NotPure(object):
            pass
        chainer_b200.create_multi_node_optimizer(actual, NotPure(), double_buffering=True)


def test_adam_eps_underflow_raises(fake):
    """adam.py:176-187 via the fused path."""
    comm = chainer_b200.create_communicator('pure_nccl')
    model = _model_with_values(2)
    actual = chainer_b200.Adam(eps=1e-60)
    opt = chainer_b200.create_multi_node_optimizer(actual, comm)
    opt.setup(model)
    opt.update()
    _set_grads(model, 1)
    with pytest.raises(ValueError, match='eps of Adam optimizer is too small'):
        opt.update()


def test_update_rule_standalone(fake):
    """GradientMethod.update without a communicator: per-parameter update_core_gpu."""
    model = _model_with_values(4)
    ref = _model_with_values(4)
    opt = chainer_b200.Adam(eta=0.5, weight_decay_rate=0.1)
    opt.setup(model)
    _set_grads(model, 3)
    _set_grads(ref, 3)
    model.c.b.grad = None                              # reallocate_cleared_grads -> zeros
    ref.c.b.grad = np.zeros_like(ref.c.b.data)
    opt.update()
    assert opt.t == 1
    for (name, p), (_, q) in zip(sorted(model.namedparams()), sorted(ref.namedparams())):
        if p.data is None:
            assert p.update_rule.t == 1
            continue
        m, v = np.zeros_like(q.data), np.zeros_like(q.data)
        og.adam_update_gpu(q.data, q.grad, m, v, 1, eta=0.5, weight_decay_rate=0.1)
        assert_bits_equal(p.data, q.data, name)
        assert_bits_equal(p.update_rule.state['m'], m, name)


def test_bucket_bounds():
    comm = PureNcclCommunicator.__new__(PureNcclCommunicator)
    comm.bucket_bytes = 32 << 20

    class M(object):
        size = 8
    comm.mpi_comm = M()
    b = comm._bucket_bounds(25557096, 4)
    assert b[0] == 0 and b[-1] == 25557096
    assert all(x % 1024 == 0 for x in b[:-1])
    assert all(b[i] < b[i + 1] for i in range(len(b) - 1))
    assert len(b) - 1 == 3                              # 102 MB / 32 MB -> 3 buckets (+ folded tail)
    comm.bucket_bytes = 256 << 20
    assert comm._bucket_bounds(25557096, 4) == [0, 25557096]   # default: one allreduce
    M.size = 1
    assert comm._bucket_bounds(1000, 4) == [0, 1000]


def test_standalone_update_is_one_launch_per_group(fake):
    """GradientMethod.update() without a communicator: one multi-tensor launch per
    (dtype, hyperparameter) group instead of one kernel per parameter
    (chainer/optimizer.py:886-889), same results as the per-parameter rules."""
    model = ExampleMixedModel()
    ref = ExampleMixedModel()
    rng = np.random.default_rng(12)
    for (_, p), (_, q) in zip(sorted(model.namedparams()), sorted(ref.namedparams())):
        p.data[...] = rng.standard_normal(p.data.shape).astype(p.data.dtype)
        q.data[...] = p.data
    opt = chainer_b200.MomentumSGD(lr=0.1, momentum=0.9)
    opt.setup(model)
    model.d.b.update_rule.hyperparam.lr = 0.01           # one extra group
    vs = {}
    for step in range(1, 3):
        for (name, p), (_, q) in zip(sorted(model.namedparams()), sorted(ref.namedparams())):
            g = (rng.standard_normal(p.data.shape) * 0.1).astype(p.data.dtype)
            p.grad = g.copy()
            q.grad = g.copy()
        fake.calls[:] = []
        opt.update()
        launches = [c for c in fake.calls if c[0] == 'gp_unpack_momentum_sgd']
        assert len(launches) == 3                          # float16, float32, float32 with lr override
        assert opt.t == step
        for (name, p), (_, q) in zip(sorted(model.namedparams()), sorted(ref.namedparams())):
            assert p.update_rule.t == step
            v = vs.setdefault(name, np.zeros_like(q.data))
            og.momentum_sgd_update(q.data, q.grad, v, 0.01 if name == '/d/b' else 0.1, 0.9)
            assert_bits_equal(p.data, q.data, name)
            assert_bits_equal(p.update_rule.state['v'], v, name)


def test_pipeline_chunk_bounds():
    from chainer_b200.communicators.pure_nccl_communicator import PureNcclCommunicator as C
    assert C._chunk_bounds(1000, 4, 0) == [0, 1000]
    # ResNet-50 fp32: the automatic multicast chunking is two 4096-aligned halves
    n = 25557096
    cb = C._chunk_bounds(n, 4, C._auto_mc_chunk_bytes(n, 4))
    assert len(cb) == 3 and cb[0] == 0 and cb[-1] == n and cb[1] % 4096 == 0
    assert abs(cb[1] - n / 2) < 4096
    # small buffers: one kernel
    assert C._auto_mc_chunk_bytes(600000, 4) == 0
    # explicit chunk size: aligned interior bounds, short tail folded into the last chunk
    cb = C._chunk_bounds(600000, 4, 256 << 10)
    assert all(b % 4096 == 0 for b in cb[1:-1]) and cb[-1] == 600000
    sizes = np.diff(cb)
    assert sizes.min() >= (65536 // 2) and sizes[:-1].max() == 65536
    for isz in (2, 4, 8):
        for n in (1, 4095, 4096, 4097, 100003, 25557096):
            for chunk in (0, 1, 4096, 1 << 20):
                cb = C._chunk_bounds(n, isz, chunk)
                assert cb[0] == 0 and cb[-1] == n and all(np.diff(cb) > 0)
                assert all(b % 4096 == 0 for b in cb[1:-1])


# ------------------------------------------- optimizer hooks and loss scaling --
from tests.hooks_scenario import HOOK_CASES, hook_objects, hooks_golden, run_hooks_scenario  # noqa: E402


@pytest.mark.parametrize('dtype', ['float32', 'float16'])
@pytest.mark.parametrize('variant', sorted(HOOK_CASES))
def test_hooks_and_loss_scale_through_the_multi_node_optimizer(fake, variant, dtype):
    """WeightDecay / GradientClipping hooks and static loss scaling behind
    create_multi_node_optimizer reproduce the unmodified reference (hooks.npz):
    fused into the update kernel for [clip], [wd], [clip, wd]; the reference
    sequence of hook kernels for [wd, clip]."""
    opt_name, spec, ls = HOOK_CASES[variant]
    if opt_name == 'adam' and dtype == 'float16':
        pytest.skip('no reference vector: float16 CPU Adam underflows (make_golden.py)')
    fusable = [k for k, _ in spec] != ['wd', 'clip']
    clip = any(k == 'clip' for k, _ in spec)
    kernel = 'gp_unpack_momentum_sgd' if opt_name == 'sgd' else 'gp_unpack_adam'

    def before_step():
        fake.calls[:] = []

    def after_step(comm):
        called = [c[0] for c in fake.calls]
        if fusable:
            assert (kernel + '_hooked') in called and kernel not in called
            assert ('gp_sqnorm' in called) == clip
            assert 'gp_weight_decay' not in called and 'gp_scale_by_device' not in called
            assert 'gp_divide' not in called
        else:
            assert 'gp_weight_decay' in called and 'gp_scale_by_device' in called

    run_hooks_scenario(variant, dtype, lambda a: a.copy(), np.asarray, before_step, after_step)


def test_standalone_optimizer_hooks(fake):
    """optimizer.update() without a communicator: call_for_each_param hooks and
    optimizer-level hooks run as in chainer/optimizer.py:706-711 (unfused kernels)."""
    z = hooks_golden()
    pre = 'clip_wd|float32|'
    names = sorted(k[len(pre) + 4:] for k in z.files if k.startswith(pre + 'init'))
    model = L.link_from_named_arrays([(n, z[pre + 'init' + n].copy()) for n in names])
    opt = chainer_b200.MomentumSGD(lr=0.01, momentum=0.9)
    opt.setup(model)
    for h in hook_objects(HOOK_CASES['clip_wd'][1]):
        opt.add_hook(h)
    params = dict(sorted(model.namedparams()))
    for step in range(3):
        for n in names:
            params[n].grad = z[pre + 'grad%d%s' % (step, n)].copy()
        opt.update()
        for n in names:
            np.testing.assert_allclose(params[n].data, z[pre + 'param%d%s' % (step, n)],
                                       rtol=2e-6, atol=2e-8)
            np.testing.assert_allclose(params[n].grad, z[pre + 'gradafter%d%s' % (step, n)],
                                       rtol=2e-6, atol=2e-8)


# ------------------------------------------ SGD, CorrectedMomentumSGD, NesterovAG --
from tests.hooks_scenario import run_family_scenario  # noqa: E402


@pytest.mark.parametrize('multi_node', [True, False])
@pytest.mark.parametrize('dtype,hooks', [('float32', False), ('float16', False), ('float64', False),
                                         ('float32', True)])
@pytest.mark.parametrize('rule', ['sgd', 'corrected', 'nesterov'])
def test_sgd_family_rules(fake, rule, dtype, hooks, multi_node):
    """SGD / CorrectedMomentumSGD / NesterovAG reproduce the reference bit-for-bit,
    fused behind the multi-node optimizer (one gp_unpack_sgd_family launch) and as a
    stand-alone multi-tensor optimizer.update()."""
    if multi_node and dtype == 'float64':
        pytest.skip('bcast_data and the allreduce buffer are float32 (chainer.get_dtype()): '
                    'float64 parameters do not stay bit-identical behind the communicator')

    def after_step(comm):
        called = [c[0] for c in fake.calls]
        if 'gp_pack' in called:                # drop the initial broadcast's pack / unpack
            called = called[len(called) - 1 - called[::-1].index('gp_pack'):]
        assert 'gp_unpack_sgd_family' in called
        if multi_node:
            assert 'gp_unpack_scale' not in called           # fused, not the reference sequence
            assert ('gp_sqnorm' in called) == hooks
        n_launch = called.count('gp_unpack_sgd_family')
        assert n_launch == 1, called
        fake.calls[:] = []

    fake.calls[:] = []
    run_family_scenario(rule, dtype, hooks, multi_node, lambda a: a.copy(), np.asarray, after_step)


@pytest.mark.parametrize('multi_node', [True, False])
@pytest.mark.parametrize('dtype', ['float32', 'float16'])
def test_dynamic_loss_scaling(fake, dtype, multi_node):
    """check_nan_in_grads / is_safe_to_update / update_loss_scale
    (chainer/optimizer.py:763-791, 881-894) reproduce the reference's loss-scale
    trajectory, skipped update and parameters bit-for-bit."""
    from tests.hooks_scenario import run_dynamic_loss_scale
    run_dynamic_loss_scale(dtype, multi_node, lambda a: a.copy(), np.asarray)


@pytest.mark.parametrize('multi_node', [True, False])
@pytest.mark.parametrize('case', ['sgd', 'sgd_wd_ls128', 'adam'])
def test_fp32_master_weights(fake, case, multi_node):
    """use_fp32_update: float16 parameters, float32 master + states, bit-for-bit the
    reference for MomentumSGD (also with WeightDecay and loss scaling)."""
    from tests.hooks_scenario import run_fp32_update
    run_fp32_update(case, multi_node, lambda a: a.copy(), np.asarray)


def test_hook_changes_and_mixed_loss_scales_switch_paths(fake):
    """Adding / removing hooks rebuilds the fused plan (hooked <-> plain kernel);
    parameters that carry DIFFERENT loss scales run the reference sequence; all-zero
    gradients clip with rate 1 (no NaN from 0/0)."""
    from chainer_b200 import optimizer_hooks as H
    comm = chainer_b200.create_communicator('pure_nccl')
    model = _model_with_values()
    ref = _model_with_values()
    actual = chainer_b200.MomentumSGD(lr=0.1, momentum=0.9)
    opt = chainer_b200.create_multi_node_optimizer(actual, comm)
    opt.setup(model)
    opt.update()                                           # broadcast
    vs = {n: np.zeros_like(q.data) for n, q in sorted(ref.namedparams()) if q.data is not None}

    def step(seed, expect, decay=None, scales=None, zero=False):
        _set_grads(model, seed)
        _set_grads(ref, seed)
        if zero:
            for m in (model, ref):
                for p in m.params():
                    if p.grad is not None:
                        p.grad[...] = 0
        for i, p in enumerate(p for p in model.params() if p.data is not None):
            p._loss_scale = None if scales is None else scales[i % len(scales)]
        fake.calls[:] = []
        opt.update()
        called = [c[0] for c in fake.calls]
        assert expect in called, called
        for i, ((name, p), (_, q)) in enumerate(zip(sorted(model.namedparams()),
                                                   sorted(ref.namedparams()))):
            if p.data is None:
                continue
            g = q.grad
            if decay is not None:
                og.weight_decay_hook(q.data, g, decay, None if scales is None else
                                     scales[[n for n, r in sorted(ref.namedparams())
                                             if r.data is not None].index(name) % len(scales)])
            if scales is not None:
                og.loss_scale_divide(g, scales[[n for n, r in sorted(ref.namedparams())
                                                if r.data is not None].index(name) % len(scales)])
            og.momentum_sgd_update(q.data, g, vs[name], 0.1, 0.9)
            assert_bits_equal(p.data, q.data, (name, expect))
        return called

    step(1, 'gp_unpack_momentum_sgd')
    opt.add_hook(H.WeightDecay(0.01))
    step(2, 'gp_unpack_momentum_sgd_hooked', decay=0.01)
    # different loss scales on different parameters: per-parameter reference sequence
    called = step(3, 'gp_weight_decay', decay=0.01, scales=[2.0, 4.0])
    assert 'gp_unpack_momentum_sgd_hooked' not in called and 'gp_divide' in called
    step(4, 'gp_unpack_momentum_sgd_hooked', decay=0.01, scales=[8.0])
    opt.remove_hook('WeightDecay')
    step(5, 'gp_unpack_momentum_sgd')
    opt.add_hook(H.GradientClipping(1.0))
    called = step(6, 'gp_sqnorm', zero=True)
    assert 'gp_unpack_momentum_sgd_hooked' in called
    for p in model.params():
        if p.data is not None:
            assert np.isfinite(p.data).all() and np.isfinite(p.grad).all()


# ----------------------------------------- BatchNormalization / create_mnbn_model --
def test_batch_normalization_link_and_create_mnbn_model(fake):
    """BatchNormalization (single worker) matches torch's batch_norm forward / backward
    and running statistics; create_mnbn_model returns a structural copy whose BN links
    are MultiNodeBatchNormalization with the same parameters and persistents
    (chainermn/links/create_mnbn_model.py:7-66), leaving the original untouched."""
    import torch
    from chainer_b200.links import (BatchNormalization, MultiNodeBatchNormalization,
                                    create_mnbn_model)

    class Block(L.Chain):
        def __init__(self):
            super(Block, self).__init__()
            with self.init_scope():
                self.bn = BatchNormalization(6, device='cpu')
                self.fc = Linear(3, 2)

    class Net(L.Chain):
        def __init__(self):
            super(Net, self).__init__()
            with self.init_scope():
                self.block = Block()
                self.bn0 = BatchNormalization(4, decay=0.8, eps=1e-3, device='cpu')
                self.tail = L.ChainList(BatchNormalization(5, use_beta=False, device='cpu'),
                                        Linear(2, 2))

    torch.manual_seed(0)
    net = Net()
    # forward / backward of the single-worker link against torch
    x = torch.randn(8, 4, 3, 3, requires_grad=True)
    net.bn0.gamma.data.copy_(torch.rand(4) + 0.5)
    net.bn0.beta.data.copy_(torch.randn(4))
    net.bn0.gamma.data.requires_grad_(True)
    net.bn0.beta.data.requires_grad_(True)
    y = net.bn0(x)
    gy = torch.randn_like(y)
    y.backward(gy)
    x2 = x.detach().clone().requires_grad_(True)
    g2 = net.bn0.gamma.data.detach().clone().requires_grad_(True)
    b2 = net.bn0.beta.data.detach().clone().requires_grad_(True)
    rm, rv = torch.zeros(4), torch.zeros(4)
    y2 = torch.nn.functional.batch_norm(x2, rm, rv, g2, b2, training=True, momentum=0.2, eps=1e-3)
    y2.backward(gy)
    np.testing.assert_allclose(y.detach().numpy(), y2.detach().numpy(), rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(x.grad.numpy(), x2.grad.numpy(), rtol=1e-3, atol=1e-5)
    np.testing.assert_allclose(net.bn0.gamma.data.grad.numpy(), g2.grad.numpy(), rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(net.bn0.beta.data.grad.numpy(), b2.grad.numpy(), rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(net.bn0.avg_mean.numpy(), rm.numpy(), rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(net.bn0.avg_var.numpy(), rv.numpy(), rtol=1e-4, atol=1e-6)
    net.bn0.N = 5

    comm = chainer_b200.create_communicator('pure_nccl')
    twin = create_mnbn_model(net, comm)
    assert [p for p, _ in twin.namedlinks()] == [p for p, _ in net.namedlinks()]
    assert [n for n, _ in sorted(twin.namedparams())] == [n for n, _ in sorted(net.namedparams())]
    for (path, a), (_, b) in zip(net.namedlinks(), twin.namedlinks()):
        assert a is not b
        if isinstance(a, BatchNormalization):
            assert type(b) is MultiNodeBatchNormalization and b.comm is comm
            assert b.decay == a.decay and b.eps == a.eps and b.N == a.N
            assert hasattr(b, 'beta') == hasattr(a, 'beta')
            for attr in ('avg_mean', 'avg_var'):
                assert getattr(b, attr) is not getattr(a, attr)
                assert torch.equal(getattr(b, attr), getattr(a, attr))
        else:
            assert type(b) is type(a)
    for (n, p), (_, q) in zip(sorted(net.namedparams()), sorted(twin.namedparams())):
        assert p is not q and p.data is not q.data
        np.testing.assert_array_equal(np.asarray(p.data.detach() if hasattr(p.data, 'detach') else p.data),
                                      np.asarray(q.data.detach() if hasattr(q.data, 'detach') else q.data))
    # the copy is independent of the original
    twin.bn0.avg_mean.add_(1.0)
    assert not torch.equal(twin.bn0.avg_mean, net.bn0.avg_mean)
    # and computes the same thing with one worker
    y3 = twin.bn0(x.detach(), train=False)
    y4 = net.bn0(x.detach(), train=False)
    assert y3.shape == y4.shape
    comm.finalize()


@pytest.mark.parametrize('opt_name', ['momentum_sgd', 'adam'])
@pytest.mark.parametrize('adt', [None, np.float16])
def test_one_launch_step_equals_separate_launches(fake, monkeypatch, opt_name, adt):
    """With the one-launch step enabled (the default) update() issues ONE library launch
    (gp_step_*) and leaves the same bits as pack + fused update; configurations it does
    not cover (hooks, mixed dtypes) keep the separate launches."""
    results = []
    for use_step in ('1', '0'):
        monkeypatch.setenv('CHAINER_B200_STEP', use_step)
        comm = chainer_b200.create_communicator('pure_nccl', allreduce_grad_dtype=adt)
        model = _model_with_values()
        actual = chainer_b200.MomentumSGD(lr=0.1, momentum=0.9) if opt_name == 'momentum_sgd' \
            else chainer_b200.Adam(alpha=0.01)
        opt = chainer_b200.create_multi_node_optimizer(actual, comm)
        opt.setup(model)
        _set_grads(model, 1)
        opt.update()
        for step in range(1, 4):
            _set_grads(model, 10 + step)
            fake.calls[:] = []
            opt.update()
            names = [c[0] for c in fake.calls]
            if use_step == '1':
                assert names == ['gp_step_' + opt_name], names
            else:
                assert 'gp_pack' in names and not any(n.startswith('gp_step') for n in names)
            assert actual.t == step
        results.append([(n, p.data.copy(), p.grad.copy()) for n, p in sorted(model.namedparams())
                        if p.data is not None])
    for (n, d1, g1), (_, d0, g0) in zip(*results):
        assert_bits_equal(d1, d0, n)
        assert_bits_equal(g1, g0, n)


def test_one_launch_step_not_taken_with_hooks(fake, monkeypatch):
    from chainer_b200 import optimizer_hooks as H
    monkeypatch.setenv('CHAINER_B200_STEP', '1')
    comm = chainer_b200.create_communicator('pure_nccl')
    model = _model_with_values()
    actual = chainer_b200.MomentumSGD(lr=0.1, momentum=0.9)
    actual_opt = chainer_b200.create_multi_node_optimizer(actual, comm)
    actual_opt.setup(model)
    actual.add_hook(H.WeightDecay(0.01))
    _set_grads(model, 1)
    actual_opt.update()
    _set_grads(model, 2)
    fake.calls[:] = []
    actual_opt.update()
    names = [c[0] for c in fake.calls]
    assert 'gp_unpack_momentum_sgd_hooked' in names and 'gp_pack' in names


@pytest.mark.parametrize('case', ['sgd_wd', 'adam'])
@pytest.mark.parametrize('multi_node', [False, True])
def test_fp32_master_with_dynamic_loss_scaling(fake, case, multi_node):
    """float16 parameters + float32 masters + dynamic loss scaling against vectors of the
    unmodified reference (fp32_dynamic.npz); behind the multi-node optimizer through the
    fused master kernels with the device-side skip word."""
    from tests.hooks_scenario import run_fp32_dynamic
    run_fp32_dynamic(case, multi_node, lambda a: a.copy(), np.asarray, fake=fake)


@pytest.mark.parametrize('case', ['sgd', 'sgd_wd_ls128', 'adam'])
def test_fp32_master_multi_node_is_fused(fake, case):
    """use_fp32_update behind create_multi_node_optimizer: ONE master launch per step, none of
    the per-parameter cast / divide / update launches."""
    from tests.hooks_scenario import run_fp32_update
    run_fp32_update(case, True, lambda a: a.copy(), np.asarray)
    called = [c[0] for c in fake.calls]
    kernel = 'gp_unpack_adam_master' if case == 'adam' else 'gp_unpack_momentum_sgd_master'
    assert called.count(kernel) == 3
    assert 'gp_unpack_momentum_sgd' not in called and 'gp_unpack_adam' not in called
    assert 'gp_divide' not in called


def test_device_array_deepcopy_owns_new_memory(fake):
    """copy.deepcopy of a model after an optimizer step (what the double-buffering optimizer
    does) gives arrays with their OWN device memory: different pointers, same contents, and
    the raw handles refuse to be copied."""
    import copy
    a = _dev.DeviceArray.from_numpy(np.arange(12, dtype=np.float32).reshape(3, 4))
    b = copy.deepcopy(a)
    c = copy.copy(a)
    assert b.data.ptr != a.data.ptr and c.data.ptr != a.data.ptr
    np.testing.assert_array_equal(b.get(), a.get())
    np.testing.assert_array_equal(c.get(), a.get())
    b.fill(7.0)
    assert a.get()[0, 1] == 1.0
    with pytest.raises(TypeError):
        copy.deepcopy(a.data)
    model = L.link_from_named_arrays([('/w', _dev.DeviceArray.from_numpy(np.ones(8, np.float32)))])
    opt = chainer_b200.MomentumSGD(lr=0.1)
    opt.setup(model)
    for _, p in model.namedparams():
        p.grad = _dev.DeviceArray.from_numpy(np.full(8, 0.5, np.float32))
    opt.update()
    twin = copy.deepcopy(model)
    for (_, p), (_, q) in zip(sorted(model.namedparams()), sorted(twin.namedparams())):
        assert q.data.data.ptr != p.data.data.ptr
        np.testing.assert_array_equal(q.data.get(), p.data.get())
        sv, tv = p.update_rule.state['v'], q.update_rule.state['v']
        assert tv.data.ptr != sv.data.ptr
        np.testing.assert_array_equal(tv.get(), sv.get())


class _DictSerializer(object):
    def __init__(self, store, prefix=''):
        self.store, self.prefix = store, prefix

    def __getitem__(self, key):
        return type(self)(self.store, self.prefix + key.strip('/') + '/')

    def __call__(self, key, value):
        self.store[self.prefix + key] = np.array(_dev.to_numpy(value)) if hasattr(value, 'shape') \
            or isinstance(value, _dev.DeviceArray) else value
        return value


class _DictDeserializer(_DictSerializer):
    """Non-strict, like chainer.serializers.NpzDeserializer(strict=False): a key the
    snapshot lacks (the state of a never-updated parameter) comes back as the passed value."""

    def __call__(self, key, value):
        if self.prefix + key not in self.store:
            return value
        got = self.store[self.prefix + key]
        if isinstance(got, np.ndarray) and got.shape != ():
            return got.copy()
        return got


@pytest.mark.parametrize('opt_name', ['momentum_sgd', 'adam'])
def test_resume_restores_optimizer_state(fake, opt_name):
    """Save after N steps, load into a FRESH optimizer (setup() then load -- the rules have
    no state yet) and compare the next step: the moments come back (the reference,
    optimizer.py:433-471); without them Adam would pair a large t with zero moments."""
    def mk():
        return chainer_b200.MomentumSGD(lr=0.1, momentum=0.9) if opt_name == 'momentum_sgd' \
            else chainer_b200.Adam(alpha=0.01)
    model = _model_with_values()
    opt = mk()
    opt.setup(model)
    for step in range(2):
        _set_grads(model, 20 + step)
        opt.update()
    store = {}
    opt.serialize(_DictSerializer(store))
    snap = {n: p.data.copy() for n, p in sorted(model.namedparams()) if p.data is not None}

    resumed = _model_with_values()
    for n, p in sorted(resumed.namedparams()):
        if p.data is not None:
            p.data[...] = snap[n]
    opt2 = mk()
    opt2.setup(resumed)
    opt2.serialize(_DictDeserializer(store))
    assert opt2.t == 2
    names = ('v',) if opt_name == 'momentum_sgd' else ('m', 'v')
    for (n, p), (_, q) in zip(sorted(model.namedparams()), sorted(resumed.namedparams())):
        if p.data is None:
            continue
        assert q.update_rule.t == 2 and q.update_rule.state is not None
        for k in names:
            assert_bits_equal(np.asarray(q.update_rule.state[k]), np.asarray(p.update_rule.state[k]), (n, k))
    _set_grads(model, 30)
    _set_grads(resumed, 30)
    opt.update()
    opt2.update()
    for (n, p), (_, q) in zip(sorted(model.namedparams()), sorted(resumed.namedparams())):
        if p.data is not None:
            assert_bits_equal(q.data, p.data, n)
