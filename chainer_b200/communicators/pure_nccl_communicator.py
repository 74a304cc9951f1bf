"""PureNcclCommunicator: the B200 implementation of the reference class of the
same name (``chainermn/communicators/pure_nccl_communicator.py:13-193``) on top
of ``MpiCommunicatorBase`` (``mpi_communicator_base.py:99-815``).

Same public surface and error behaviour: ``__init__(mpi_comm)``, lazy
``_init_comms`` (the NCCL communicator binds to the CUDA device that is current
when it is first needed, ``:30-34``), ``finalize``, ``set_config`` /
``get_config`` (``allreduce_grad_dtype``, ``batched_copy``), ``bcast_data``,
``multi_node_mean_grad``, ``_multi_node_mean_grad_async(model, zero_fill,
stream)``, ``_multi_node_mean_nccl(sendbuf, recvbuf, n_elems, dtype, stream)``,
``nccl_comm.allReduce(...)`` positional arguments.

What changes underneath (none of it observable through that surface):

* pack is ONE multi-tensor launch with the table staged in shared memory
  (gp_pack) and a single asynchronous table upload;
* the allreduce runs IN PLACE on the packed buffer, split into buckets that
  are issued on a side stream as soon as each bucket is packed, so that the
  allreduce of bucket i overlaps the pack of bucket i+1 and the unpack/update of
  bucket i-1;
* ``div_by_size`` and unpack are one kernel (gp_unpack_scale), with the
  reference's rounding sequence (scale in double, round to the buffer dtype,
  cast to the gradient dtype);
* ``multi_node_mean_grad_and_update`` additionally fuses the optimizer update
  (MomentumSGD / Adam) into that kernel.
"""
import collections
import ctypes
import operator
import warnings

import numpy as np

from chainer_b200 import _lib
from chainer_b200 import config
from chainer_b200 import device as _dev
from chainer_b200 import nccl
from chainer_b200.communicators import _communication_utility
from chainer_b200.communicators import _memory_utility
from chainer_b200.communicators import mpi_communicator_base


class _SelfNcclComm(object):
    """``nccl_comm`` of a single-process world: SUM over one rank is a copy.
    Keeps the ``allReduce`` / ``bcast`` call surface (tests spy on it)."""

    size = 1
    rank = 0

    def allReduce(self, sendbuf, recvbuf, count, datatype, op, stream):
        if sendbuf != recvbuf and count > 0:
            itemsize = 8 if datatype == nccl.NCCL_FLOAT64 else (4 if datatype == nccl.NCCL_FLOAT32 else 2)
            _lib.get().gp_memcpy_async(recvbuf, sendbuf, count * itemsize, 2, stream)

    def bcast(self, buff, count, datatype, root, stream):
        pass

    def reduce(self, sendbuf, recvbuf, count, datatype, op, root, stream):
        self.allReduce(sendbuf, recvbuf, count, datatype, op, stream)

    def destroy(self):
        pass


class PureNcclCommunicator(mpi_communicator_base.MpiCommunicatorBase):

    def __init__(self, mpi_comm):
        super(PureNcclCommunicator, self).__init__(mpi_comm)
        _lib.get()  # fail loudly now if libgradpath.so is missing
        if self.size > 1:
            try:
                version = nccl.get_version()
            except Exception as e:
                raise RuntimeError(
                    'PureNcclCommunicator requires NCCL 2.0+, '
                    'but NCCL is not available. ({})'.format(e))
            if version < 2000:
                raise RuntimeError(
                    'PureNcclCommunicator requires NCCL 2.0+, '
                    'but found {}.'.format(version))
            if version < 2302:
                warnings.warn('NCCL 2.2 and older versions are deprecated.',
                              DeprecationWarning)

        # NCCL communicators bind to the current CUDA device: delay the
        # initialisation until the user has selected it
        # (pure_nccl_communicator.py:30-34).
        self.nccl_comm = None

        self.gpu_tmp_buffer = _memory_utility.DeviceMemory()
        self.gpu_buffer_a = _memory_utility.DeviceMemory()
        self.gpu_buffer_b = _memory_utility.DeviceMemory()

        with self.config_scope():
            self.allreduce_grad_dtype = None
        self.grad_dtype_to_allreduce_dtype_kernel = None
        self.allreduce_dtype_to_grad_dtype_kernel = None
        self.params_data = None

        # B200 additions (not configs of the reference)
        self.bucket_bytes = 256 << 20    # NCCL path: allreduce bucket size when size > 1
        self.write_grad = True           # fused update keeps param.grad observable
        self._comm_stream = None
        self._events = []
        self._table_grad = None
        self._table_data = None
        self._table_fused = None
        self._finite_flag = None
        self._fused_plan = None
        self.use_p2p = None              # None: decide at first use; True/False: forced
        self.p2p_chunk_bytes = 0         # peer-memory path: pipeline chunk (0: one kernel; measured best)
        self.p2p_max_bytes = 256 << 20   # beyond this NCCL (NVLS) is faster on 4/8 GPUs (measured)
        self._p2p = None
        # NVSwitch-multicast allreduce (csrc/gp_mc.cu).  None: 4 and 8 ranks when the
        # devices support it (or as CHAINER_B200_MULTICAST says); True/False: forced
        self.use_multicast = None
        self.mc_max_bytes = 0            # 0: no limit
        # multicast path: pipeline chunk (0: one kernel; None: two chunks from 32 MB
        # up -- measured best at N = 4 and 8, profiles/r01_bench_n{4,8}_mc*.json)
        self.mc_chunk_bytes = None
        # one-launch step (csrc/gp_step.cu): pack -> sum over ranks -> fused update as ONE
        # kernel per rank, tile by tile.  None: whenever it covers the configuration
        # (or as CHAINER_B200_STEP says); False: the separate launches
        self.use_step = None

    # ------------------------------------------------------------ lifecycle --
    def finalize(self):
        super(PureNcclCommunicator, self).finalize()
        if self.nccl_comm is not None:
            _dev.Stream.null.synchronize()
            if self._comm_stream is not None:
                self._comm_stream.synchronize()
            self.mpi_comm.barrier()
            if self._p2p is not None:
                self._p2p.destroy()
                self._p2p = None
            self.nccl_comm.destroy()
            self.nccl_comm = None

    def _init_comms(self):
        if self.nccl_comm is not None:
            return
        if self.size == 1:
            self.nccl_comm = _SelfNcclComm()
        else:
            self.nccl_comm = _communication_utility.init_nccl_comm(self.mpi_comm)
            self._init_p2p()

    def _init_p2p(self):
        """Peer-memory allreduce (csrc/gp_p2p.cu) when all ranks share one box."""
        from chainer_b200.communicators import _p2p
        want = self.use_p2p
        if want is None:
            want = _p2p.enabled_by_env() and not _lib.get().accepts_host_pointers
        ok = bool(want) and self.size in (2, 4, 8) and self.intra_size == self.size
        # the decision must be collective
        ok = all(self.mpi_comm.allgather(ok))
        if not ok:
            self._p2p = None
            return
        try:
            p2p = _p2p.PeerAllreduce(self.mpi_comm)
            good = True
        except Exception as e:       # no peer access between these devices
            warnings.warn('peer-memory allreduce unavailable, using NCCL: {}'.format(e))
            p2p, good = None, False
        if all(self.mpi_comm.allgather(good)):
            self._p2p = p2p
            want_mc = self.use_multicast
            if want_mc is None:
                env = _p2p.multicast_enabled_by_env()
                want_mc = (self.size >= 4) if env is None else env not in ('0', '', 'false', 'False')
            if all(self.mpi_comm.allgather(bool(want_mc) and p2p.multicast_supported)):
                # the packed buffer becomes a multicast-bound allocation (collective)
                self.gpu_buffer_a.allocator = self._mc_allocate
        else:
            if p2p is not None:
                p2p._close(p2p._flag_maps)
            self._p2p = None

    def _mc_allocate(self, nbytes):
        if self._p2p is None or not self._p2p.multicast_supported:
            return None
        alloc = self._p2p.mc_allocate(nbytes)
        if alloc is None:
            warnings.warn('NVSwitch multicast unavailable, using the peer-memory kernel: {}'
                          .format(self._p2p.multicast_error))
        return alloc

    def _mc_active(self, buf):
        from chainer_b200.communicators import _p2p
        return self._p2p is not None and isinstance(buf._alloc, _p2p._McAllocation)

    def set_config(self, name, value=True, **kwargs):
        if name == 'allreduce_grad_dtype':
            if value is not None:
                if isinstance(value, str) and value == 'bfloat16':
                    allreduce_grad_dtype = 'bfloat16'       # extension: no NumPy dtype exists
                else:
                    allreduce_grad_dtype = np.dtype(value)
                    if allreduce_grad_dtype.kind != 'f':
                        raise ValueError(
                            'allreduce_grad_dtype must be'
                            'numpy.float16, numpy.float32,'
                            'numpy.float64, or None.')
            else:
                allreduce_grad_dtype = None

            with self.config_scope():
                self.allreduce_grad_dtype = allreduce_grad_dtype
        else:
            super(PureNcclCommunicator, self).set_config(name, value, **kwargs)

    def get_config(self, name=None):
        if name == 'allreduce_grad_dtype':
            return self.allreduce_grad_dtype
        else:
            return super(PureNcclCommunicator, self).get_config(name)

    # ---------------------------------------------------------- bcast_data --
    def bcast_data(self, model):
        """``pure_nccl_communicator.py:82-99`` -- same layout and transfer dtype
        (``chainer.get_dtype()``); batched pack / unpack instead of one cast +
        memcpy per parameter."""
        self._init_comms()
        params = _memory_utility.extract_params_set_data(model)
        data_dtype = config.get_dtype()
        n_elems = sum(_dev.array_size(param.data) for param in params)
        data_grad_n_bytes = data_dtype.itemsize * n_elems
        if self.gpu_tmp_buffer.size != data_grad_n_bytes:
            self.gpu_tmp_buffer.assign(data_grad_n_bytes)
        stream = _dev.Stream.null
        if n_elems == 0:
            return
        if self._table_data is None:
            self._table_data = _memory_utility.DeviceTable()
        pd = _memory_utility.ParamsData(params, 'data', False, stream=stream,
                                        table=self._table_data)
        _memory_utility._batched_pack_params(pd, self.gpu_tmp_buffer, data_dtype, stream)
        self.nccl_comm.bcast(self.gpu_tmp_buffer.ptr(), n_elems,
                             _communication_utility._get_nccl_type_id(data_dtype),
                             0, stream.ptr)
        _memory_utility._batched_unpack_params(pd, self.gpu_tmp_buffer, data_dtype, stream)

    # ------------------------------------------------- multi_node_mean_grad --
    def multi_node_mean_grad(self, model, zero_fill=False):
        stream = _dev.Stream.null
        self._multi_node_mean_grad_async(model, zero_fill, stream)

    def _allreduce_dtype(self):
        # NOTE: explicit `is None` as in the reference (:111-114)
        if self.allreduce_grad_dtype is not None:
            return self.allreduce_grad_dtype
        return config.get_dtype()

    def _multi_node_mean_grad_async(self, model, zero_fill, stream):
        """``pure_nccl_communicator.py:105-139``."""
        self._init_comms()
        params = _memory_utility.extract_params_set_grad(model, zero_fill)
        allreduce_grad_dtype = self._allreduce_dtype()
        assert allreduce_grad_dtype is not None

        if self._table_grad is None:
            self._table_grad = _memory_utility.DeviceTable()
        # zero_fill creates missing gradients here (ParamsData, :40-42)
        pd = _memory_utility.ParamsData(params, 'grad', zero_fill, stream=stream,
                                        table=self._table_grad)
        n_elems = pd.n_elems
        needs_sync = self._prepare_allreduce_pack_buffer(allreduce_grad_dtype, n_elems)
        if stream != _dev.Stream.null and needs_sync:
            _dev.Stream.null.synchronize()
        if n_elems == 0:
            return
        self.params_data = None

        def unpack(begin, end):
            _memory_utility._batched_unpack_params(
                pd, self.gpu_buffer_a, allreduce_grad_dtype, stream,
                scale=1.0 / self.size, elem_begin=begin, elem_end=end)

        self._pipeline(pd, allreduce_grad_dtype, stream, unpack)

    def _prepare_allreduce_pack_buffer(self, allreduce_grad_dtype, n_elems):
        """``:141-152``.  Only buffer A is needed: the allreduce is in place."""
        allreduce_grad_n_bytes = _dev.dtype_itemsize(allreduce_grad_dtype) * n_elems
        needs_sync = False
        if self.gpu_buffer_a.size != allreduce_grad_n_bytes:
            self.gpu_buffer_a.assign(allreduce_grad_n_bytes)
            needs_sync = True
        return needs_sync

    def _bucket_bounds(self, n_elems, itemsize):
        if self.size == 1 or self.bucket_bytes <= 0:
            return [0, n_elems]
        per = max(self.bucket_bytes // itemsize, 1024) // 1024 * 1024
        bounds = list(range(0, n_elems, per)) + [n_elems]
        if len(bounds) > 2 and bounds[-1] - bounds[-2] < per // 4:
            del bounds[-2]                       # fold a small tail into the last bucket
        return bounds

    @staticmethod
    def _auto_mc_chunk_bytes(n_elems, itemsize):
        """Two chunks from 32 MB up, one kernel below (measured: profiles/r01_bench_n*_mc*)."""
        if n_elems * itemsize < (32 << 20):
            return 0
        half = ((n_elems + 1) // 2 + 4095) // 4096 * 4096
        return half * itemsize

    @staticmethod
    def _chunk_bounds(n_elems, itemsize, chunk_bytes):
        """Element bounds of the pipeline chunks: multiples of 4096 elements (16-byte
        vectors for every dtype, whole walker tiles), a short tail folded in."""
        if chunk_bytes <= 0:
            return [0, n_elems]
        per = max(chunk_bytes // itemsize, 4096) // 4096 * 4096
        cb = list(range(0, n_elems, per)) + [n_elems]
        if len(cb) > 2 and cb[-1] - cb[-2] < per // 2:
            del cb[-2]
        return cb

    def _event(self, i):
        while len(self._events) <= i:
            self._events.append(_dev.Event())
        return self._events[i]

    def _pipeline(self, pd, dtype, stream, consume):
        """pack -> in-place allreduce -> consume(begin, end), bucketed.

        Stream `stream` runs pack and consume; the allreduce of every bucket is
        issued on a side stream as soon as that bucket is packed (events), so it
        overlaps the packing of later buckets and the consumption of earlier
        ones.  With one rank there is nothing to reduce and no side stream."""
        buf = self.gpu_buffer_a
        type_id = _communication_utility._get_nccl_type_id(dtype)
        itemsize = _dev.dtype_itemsize(dtype)
        n = pd.n_elems
        debug = config.is_debug()
        if debug:
            self._check_ready_to_allreduce_meta(n, dtype)
        bounds = self._bucket_bounds(n, itemsize)
        nb = len(bounds) - 1
        nv = _lib.nvtx_range
        if self.size == 1:
            with nv('gradpath.pack'):
                _memory_utility._batched_pack_params(pd, buf, dtype, stream)
            # a one-rank SUM is the identity: nothing to launch (the reference
            # would copy A to B here)
            self.nccl_comm.allReduce(buf.ptr(), buf.ptr(), n, type_id, nccl.NCCL_SUM,
                                     stream.ptr)
            if debug:
                self._ensure_all_finite_device(buf.ptr(), dtype, n, stream)
            with nv('gradpath.update'):
                consume(0, n)
            return
        reduce_range = None
        if self._mc_active(buf):
            if itemsize <= 4 and (self.mc_max_bytes <= 0 or n * itemsize <= self.mc_max_bytes):
                # ONE kernel per rank: the switch reduces (multimem.ld_reduce) and
                # replicates (multimem.st) this rank's 1/N shard; barriers inside
                reduce_range = self._p2p.mc_allreduce
                chunk_bytes = self.mc_chunk_bytes
                if chunk_bytes is None:
                    chunk_bytes = self._auto_mc_chunk_bytes(n, itemsize)
            # float64 or over the limit: NCCL below (works on any device pointer)
        elif self._p2p is not None and (self.size == 2 or n * itemsize <= self.p2p_max_bytes):
            # ONE kernel per rank reduces over NVLink peer memory (the cross-GPU
            # barriers are inside it)
            self._p2p.ensure(buf, stream)
            reduce_range = self._p2p.allreduce
            chunk_bytes = self.p2p_chunk_bytes
        if reduce_range is not None:
            # Optional chunking: the NVLink-bound reduction of chunk i runs on a side
            # stream under the HBM-bound pack of chunk i+1 and update of chunk i-1.
            cb = self._chunk_bounds(n, itemsize, chunk_bytes)
            if len(cb) == 2:
                with nv('gradpath.pack'):
                    _memory_utility._batched_pack_params(pd, buf, dtype, stream)
                with nv('gradpath.allreduce'):
                    reduce_range(dtype, 0, n, stream)
                if debug:
                    self._ensure_all_finite_device(buf.ptr(), dtype, n, stream)
                with nv('gradpath.update'):
                    consume(0, n)
                return
            if self._comm_stream is None:
                self._comm_stream = _dev.Stream(non_blocking=True)
            cs = self._comm_stream
            for b in range(len(cb) - 1):
                lo, hi = cb[b], cb[b + 1]
                _memory_utility._batched_pack_params(pd, buf, dtype, stream, elem_begin=lo,
                                                     elem_end=hi)
                ev = self._event(2 * b)
                ev.record(stream)
                cs.wait_event(ev)
                reduce_range(dtype, lo, hi - lo, cs)
                self._event(2 * b + 1).record(cs)
            for b in range(len(cb) - 1):
                stream.wait_event(self._event(2 * b + 1))
                if debug:
                    self._ensure_all_finite_device(buf.ptr() + cb[b] * itemsize, dtype,
                                                   cb[b + 1] - cb[b], stream)
                consume(cb[b], cb[b + 1])
            return
        if self._comm_stream is None:
            self._comm_stream = _dev.Stream(non_blocking=True)
        cs = self._comm_stream
        for b in range(nb):
            lo, hi = bounds[b], bounds[b + 1]
            _memory_utility._batched_pack_params(pd, buf, dtype, stream, elem_begin=lo,
                                                 elem_end=hi)
            ev = self._event(2 * b)
            ev.record(stream)
            cs.wait_event(ev)
            self.nccl_comm.allReduce(buf.ptr() + lo * itemsize, buf.ptr() + lo * itemsize,
                                     hi - lo, type_id, nccl.NCCL_SUM, cs.ptr)
            self._event(2 * b + 1).record(cs)
        for b in range(nb):
            stream.wait_event(self._event(2 * b + 1))
            if debug:
                self._ensure_all_finite_device(buf.ptr() + bounds[b] * itemsize, dtype,
                                               bounds[b + 1] - bounds[b], stream)
            consume(bounds[b], bounds[b + 1])

    # -------------------------------------------------- debug-mode checks --
    def _check_ready_to_allreduce_meta(self, n_elems, dtype):
        """Shape agreement across ranks (``mpi_communicator_base.py:717-728``)."""
        my_shapes = (((n_elems,), str(dtype)), (n_elems,), str(dtype))
        all_shapes = self.gather_obj((self.rank, my_shapes))
        if self.rank == 0:
            for rank, shapes in all_shapes:
                if my_shapes != shapes:
                    raise ValueError('Shape does not match: {}'
                                     ' at rank 0 while {} at rank {}'
                                     .format(my_shapes, shapes, rank))

    def _ensure_all_finite_device(self, ptr, dtype, n_elems, stream):
        """``_ensure_all_finite`` (``mpi_communicator_base.py:730-733``)."""
        lib = _lib.get()
        if self._finite_flag is None:
            self._finite_flag = _dev.DeviceArray.zeros((1,), np.int32)
        flag = self._finite_flag
        lib.gp_memset_async(flag.data.ptr, 0, 4, stream.ptr)
        lib.gp_check_finite(ptr, _dev.dtype_id(dtype), n_elems, flag.data.ptr, stream.ptr)
        if int(flag.get(stream)[0]) != 0:
            raise ValueError('Parameters diverged after allreduce.')

    # ----------------------------------------------- _multi_node_mean_nccl --
    def _multi_node_mean_nccl(self, sendbuf, recvbuf, n_elems, dtype, stream=None):
        """``pure_nccl_communicator.py:154-193``: ``recvbuf = mean over ranks of
        sendbuf`` for ``DeviceMemory``-like buffers (used by MNBN)."""
        if stream is None:
            stream = _dev.Stream.null
        if config.is_debug():
            stream.synchronize()
            self._check_ready_to_allreduce_meta(n_elems, dtype)
        self._init_comms()
        type_id = _communication_utility._get_nccl_type_id(dtype)
        self.nccl_comm.allReduce(sendbuf.ptr(), recvbuf.ptr(), n_elems,
                                 type_id, nccl.NCCL_SUM, stream.ptr)
        _lib.get().gp_scale(recvbuf.ptr(), _dev.dtype_id(dtype), n_elems, 1.0 / self.size,
                            stream.ptr)
        if config.is_debug():
            self._ensure_all_finite_device(recvbuf.ptr(), dtype, n_elems, stream)

    # ----------------------------------------------------- fused update ----
    def multi_node_mean_grad_and_update(self, model, optimizer, zero_fill=False, stream=None):
        """``multi_node_mean_grad(model, zero_fill)`` followed by
        ``optimizer.update(None)`` as ONE pipeline: pack, in-place allreduce, and
        a fused unpack + descale + MomentumSGD / Adam kernel that reads every
        mean gradient once.

        Returns False -- having done nothing -- when the optimizer cannot be
        fused (hooks, loss scaling, custom or disabled update rules, fp32 master
        weights, mixed gradient/parameter dtypes); the caller then runs the two
        reference steps.  Observable state afterwards equals the reference's:
        ``optimizer.t`` and every ``rule.t`` advanced by one, states created,
        and (``self.write_grad``) ``param.grad`` holding the mean.

        The walk over the model, the fusability checks and the device tables are
        cached in a plan that stays valid while the model structure and the
        update rules are unchanged (version counters of ``chainer_b200.core``);
        gradients are re-read every step because Chainer reallocates them
        (``cleargrads``), and optimizer-level hyperparameters are re-read every
        step because training extensions change them.
        """
        plan = self._fused_plan
        if plan is None or not plan.matches(model, optimizer, zero_fill):
            plan = _FusedPlan.build(self, model, optimizer, zero_fill)
            self._fused_plan = plan if (plan is not None and plan.cacheable) else None
            if plan is None:
                return False
        return plan.run(stream if stream is not None else _dev.Stream.null)

    def invalidate_plans(self):
        """Drop cached plans (call after replacing parameter or state arrays
        behind the back of the Link / UpdateRule classes)."""
        self._fused_plan = None


def _version_cells():
    from chainer_b200.core import link as _link
    from chainer_b200.core import optimizer as _opt
    return _link._structure_version, _opt._rules_version


_cells = []


def _versions():
    """(structure version, rules version) of chainer_b200.core -- what a cached plan
    is valid for."""
    if not _cells:
        _cells.extend(_version_cells())
    return (_cells[0][0], _cells[1][0])


class _FusedPlan(object):
    """Everything about one (model, optimizer) pair that does not change from
    step to step on the fused path."""

    @classmethod
    def build(cls, comm, model, optimizer, zero_fill):
        fp = _fusion_plan(model, optimizer, zero_fill)
        if fp is None:
            return None
        self = cls()
        self.comm = comm
        self.model = model
        self.optimizer = optimizer
        self.zero_fill = zero_fill
        self.params, self.others = fp
        self.clip_hook, self.wd_hook = _fusable_hooks(optimizer)
        self.hook_struct = _lib.GpHooks()
        self.rules = [p.update_rule for p in self.params]
        self.other_rules = [p.update_rule for p in self.others
                            if p.update_rule is not None and p.update_rule.enabled]
        self.cacheable = bool(getattr(model, '_b200_versioned', False) and
                              getattr(optimizer, '_b200_versioned', False) and
                              all(getattr(r, '_b200_versioned', False) for r in self.rules))
        self.master = bool(_is_master_plan(self.params))
        self.skip_flag = None                # device word of the dynamic-loss-scaling check
        # states are created on first use (chainer/optimizer.py:473-484); with float32
        # master weights on the float32 copy of the parameter (:262-282)
        for p in self.params:
            rule = p.update_rule
            rule._init_states(rule.fp32_param_for(p) if self.master else p)
        self.versions = _versions()          # after state creation
        for p in self.params:                # ParamsData zero_fill (_memory_utility.py:40-42)
            if p.grad is None:
                p.grad = _dev.zeros_like(p.data)
        self.extra = []
        for p in self.params:
            st = p.update_rule.state
            states = [st[k] for k in p.update_rule.state_names]
            if self.master:
                self.extra.append((p.update_rule._fp32_param.data, states, p.data))
            else:
                self.extra.append((p.data, states))
        self.table = _memory_utility.DeviceTable()
        self.pd = _memory_utility.ParamsData(self.params, 'grad', False, extra_ptrs=self.extra,
                                             table=self.table)
        self.cache = collections.OrderedDict()
        self.sizes = [_dev.array_size(p.data) for p in self.params]
        self.dtypes = [_dev.array_dtype(p.data) for p in self.params]
        # launch groups: rules with equal hyperparameters and step count.  The
        # grouping only depends on per-rule overrides, which are version-tracked.
        groups = {}
        order = []
        for i, rule in enumerate(self.rules):
            key = rule.fused_signature()
            if key not in groups:
                groups[key] = []
                order.append(key)
            groups[key].append(i)
        self.groups = []
        if len(order) == 1:
            self.groups.append((self.rules[0], None, self.pd))
        else:
            for key in order:
                idx = np.asarray(groups[key])
                sub = _memory_utility.ParamsData(
                    [self.params[i] for i in idx], 'grad', False,
                    extra_ptrs=[self.extra[i] for i in idx],
                    buf_offsets=self.pd.host_csum[idx])
                self.groups.append((self.rules[idx[0]], idx, sub))
        return self

    def has_param_loss_scale(self):
        # backward(loss_scale=...) marks every parameter; looking at one is enough to
        # notice loss scaling driven from outside the optimizer object
        return bool(self.params) and getattr(self.params[0], '_loss_scale', None) is not None

    def matches(self, model, optimizer, zero_fill):
        return (model is self.model and optimizer is self.optimizer and
                zero_fill == self.zero_fill and self.versions == _versions())

    def _tables(self, stream):
        """Device tables for the CURRENT gradient arrays.

        Chainer reallocates gradients every step (cleargrads), frameworks with
        persistent gradient buffers do not, and benchmarks rotate a few sets.  The
        tables are cached per tuple of gradient ADDRESSES -- what the kernels consume,
        so a key can never go stale and the cache holds no reference to any array (no
        old gradient set is kept alive) -- up to four sets, each with its own device
        copy.  A hit costs one tuple of addresses; a miss validates the new arrays like
        ParamsData does and uploads one table.
        """
        params = self.params
        grads = list(map(_GET_GRAD, params))
        try:
            key = _dev.ptr_key(grads)
        except ValueError:
            if not any(g is None for g in grads):
                raise
            for p in params:
                if p.grad is None:
                    p.grad = _dev.zeros_like(p.data)         # zero_fill
            grads = list(map(_GET_GRAD, params))
            key = _dev.ptr_key(grads)
        # (the dtype of a gradient equals its parameter's -- checked on a miss; an array
        # of another dtype at a cached address is caught by this probe of both ends)
        probe = (grads[0].dtype, grads[-1].dtype) if grads else None
        ent = self.cache.get(key)
        if ent is not None and ent.probe == probe:
            if len(self.cache) > 1:
                self.cache.move_to_end(key)
            return ent
        for i, g in enumerate(grads):
            dt = _dev.array_dtype(g)
            if isinstance(dt, str) or dt != self.dtypes[i]:
                if isinstance(dt, str) or dt not in (np.float16, np.float32, np.float64):
                    raise ValueError('dtype must be float16, float32 or float64.')
                self.comm._fused_plan = None
                raise _PlanStale()
            if _dev.array_size(g) != self.sizes[i]:
                raise ValueError('gradient of size {} for a parameter of size {}'.format(
                    _dev.array_size(g), self.sizes[i]))
            _dev.device_ptr(g)                               # contiguity / device checks
        ptrs = np.asarray(key, dtype=np.uint64)
        if ent is None:
            if len(self.cache) >= 4:
                _, ent = self.cache.popitem(last=False)      # recycle the oldest set's tables
            else:
                ent = _TableSet(self)
        ent.update(ptrs, probe, stream)
        self.cache[key] = ent
        return ent

    def run(self, stream):
        """One fused step; False (nothing done) when this step must run unfused."""
        comm = self.comm
        comm._init_comms()
        dtype = comm._allreduce_dtype()
        # optimizer hooks and loss scaling of this step (chainer/optimizer.py:881-883,
        # 286-291): the loss scale is what backward() left on the parameters
        clip, wdh = self.clip_hook, self.wd_hook
        loss_scale = None
        if self.optimizer._loss_scale is not None or self.has_param_loss_scale():
            scales = set(getattr(p, '_loss_scale', None) for p in self.params)
            if len(scales) != 1:
                return False              # per-parameter loss scales: the unfused sequence
            loss_scale = scales.pop()
        try:
            tables = self._tables(stream)
        except _PlanStale:
            # a gradient changed dtype: rebuild through the slow path
            if not comm.multi_node_mean_grad_and_update(self.model, self.optimizer,
                                                        self.zero_fill, stream):
                raise ValueError('gradient dtype does not match its parameter')
            return True
        # t bookkeeping of GradientMethod.update / UpdateRule.update
        # (chainer/optimizer.py:857-894, 236-250)
        self.optimizer.t += 1
        for rule in self.other_rules:
            rule.t += 1
        for rule in self.rules:
            rule.t += 1
        pd = tables.pd
        n_elems = pd.n_elems
        needs_sync = comm._prepare_allreduce_pack_buffer(dtype, n_elems)
        if stream != _dev.Stream.null and needs_sync:
            _dev.Stream.null.synchronize()
        if n_elems == 0:
            return True
        lib = _lib.get()
        buf_id = _dev.dtype_id(dtype)
        scale = 1.0 / comm.size
        wg = 1 if comm.write_grad else 0
        buf_ptr = comm.gpu_buffer_a.ptr()
        sp = stream.ptr
        hooked = clip is not None or wdh is not None or loss_scale is not None
        hk_addr = None
        if hooked:
            hk = self.hook_struct
            decay = 0.0
            if wdh is not None:
                decay = float(wdh.rate) * (loss_scale if loss_scale is not None else 1.0)
            hk.weight_decay = decay
            hk.loss_scale = float(loss_scale) if loss_scale is not None else 0.0
            hk.clip_rate = clip.scratch().rate_ptr if clip is not None else None
            hk_addr = ctypes.addressof(hk)
        launches = []
        for rep, idx, t in tables.groups:
            key = rep.fused_key()                 # re-read hyperparameters (and alpha_t)
            if key[0] == 'adam':
                ddt = self.dtypes[0 if idx is None else idx[0]]
                rep._check_eps(np.float32 if ddt == np.float16 else ddt.type)
            launches.append((key, t))

        master = self.master
        dynamic = bool(getattr(self.optimizer, '_loss_scaling_is_dynamic', False))
        skip_ptr = None
        if dynamic:
            # is_safe_to_update() on the device: gp_check_finite over the reduced buffer sets
            # this word, the master kernels skip on it -- no host decision between the
            # allreduce and the update (chainer/optimizer.py:763-779)
            if self.skip_flag is None:
                self.skip_flag = _dev.DeviceArray.zeros((1,), np.int32)
            skip_ptr = self.skip_flag.data.ptr

        def launch(key, t, begin, end):
            hint = t.layout_hint(dtype)
            if master:
                if key[0] == 'momentum_sgd':
                    lib.gp_unpack_momentum_sgd_master(buf_ptr, buf_id, t.d_csum, t.d_segs,
                                                      t.n_params, begin, end, scale, key[1], key[2],
                                                      wg, hk_addr, skip_ptr, sp)
                else:
                    lib.gp_unpack_adam_master(buf_ptr, buf_id, t.d_csum, t.d_segs, t.n_params,
                                              begin, end, scale, key[1], key[2], key[3], key[4],
                                              key[5], key[6], key[7], key[8], key[9], wg, hk_addr,
                                              skip_ptr, sp)
            elif key[0] == 'momentum_sgd':
                if hooked:
                    lib.gp_unpack_momentum_sgd_hooked(buf_ptr, buf_id, t.d_csum, t.d_segs,
                                                      t.n_params, begin, end, scale, key[1],
                                                      key[2], wg, hint, hk_addr, sp)
                else:
                    lib.gp_unpack_momentum_sgd(buf_ptr, buf_id, t.d_csum, t.d_segs, t.n_params,
                                               begin, end, scale, key[1], key[2], wg, hint, sp)
            elif key[0] == 'sgd_family':
                lib.gp_unpack_sgd_family(buf_ptr, buf_id, t.d_csum, t.d_segs, t.n_params, begin,
                                         end, scale, key[1], key[2], key[3], wg, hint, hk_addr, sp)
            elif hooked:
                lib.gp_unpack_adam_hooked(buf_ptr, buf_id, t.d_csum, t.d_segs, t.n_params, begin,
                                          end, scale, key[1], key[2], key[3], key[4], key[5],
                                          key[6], key[7], key[8], key[9], wg, hint, hk_addr, sp)
            else:
                lib.gp_unpack_adam(buf_ptr, buf_id, t.d_csum, t.d_segs, t.n_params, begin, end,
                                   scale, key[1], key[2], key[3], key[4], key[5], key[6], key[7],
                                   key[8], key[9], wg, hint, sp)

        if clip is not None or dynamic:
            # the global norm (and the finiteness verdict) need every bucket reduced: one
            # reduction over the whole allreduced buffer forms the rate (sets the skip word)
            # on the device, then the update(s) run
            threshold = float(clip.threshold) if clip is not None else 0.0
            sc = clip.scratch() if clip is not None else None

            def consume(begin, end):
                if end == n_elems:
                    if dynamic:
                        lib.gp_memset_async(skip_ptr, 0, 4, sp)
                        lib.gp_check_finite(buf_ptr, buf_id, n_elems, skip_ptr, sp)
                    if clip is not None:
                        lib.gp_sqnorm(buf_ptr, buf_id, n_elems, scale, 0, threshold, sc.ws, sc.out,
                                      sp)
                    for key, t in launches:
                        launch(key, t, 0, t.n_elems)
        elif len(launches) == 1:
            key0, t0 = launches[0]

            def consume(begin, end):
                launch(key0, t0, begin, end)
        else:
            def consume(begin, end):
                # grouped launches walk their own segment lists: run them once
                # the last bucket has been reduced
                if end == n_elems:
                    for key, t in launches:
                        launch(key, t, 0, t.n_elems)
        if (not hooked and not master and len(launches) == 1 and not config.is_debug() and
                self._run_step(lib, launches[0][0], launches[0][1], dtype, buf_id, scale, wg,
                               n_elems, stream)):
            return True
        comm._pipeline(pd, dtype, stream, consume)
        if dynamic:
            self._finish_dynamic_loss_scale(stream)
        return True

    def _finish_dynamic_loss_scale(self, stream):
        """The host half of dynamic loss scaling, AFTER the whole step is enqueued: read the
        4-byte verdict (the only synchronisation of the step, at its end -- the reference
        synchronises per parameter before it may launch the updates), undo the step counters
        of a skipped update, warn like ``check_nan_in_grads`` and move the scale
        (``chainer/optimizer.py:763-791``)."""
        import warnings
        opt = self.optimizer
        bad = int(self.skip_flag.get(stream)[0]) != 0
        opt._loss_scaling_isnan = bad
        if bad:
            opt._loss_scaling_isnan_ever = True
            for rule in self.rules:          # param.update() was not called
                rule.t -= 1
            for rule in self.other_rules:
                rule.t -= 1
            names = self._nonfinite_names()
            for name in names or ['(the reduced gradient buffer)']:
                warnings.warn(
                    'Non finite number found in param.grad of {}'
                    ' (iteration: {}, loss_scale: {})'
                    .format(name, opt.t - 1, opt._loss_scale))
        opt.update_loss_scale()

    def _nonfinite_names(self):
        """Names of the parameters whose (mean) gradient is not finite -- only run after a
        skipped step, to word the warning like the reference."""
        lib = _lib.get()
        named = [(n, p) for n, p in sorted(self.model.namedparams()) if p.grad is not None]
        if not named or not self.comm.write_grad:
            return []
        flags = _dev.DeviceArray.zeros((len(named),), np.int32)
        for i, (_, p) in enumerate(named):
            g = p.grad
            lib.gp_check_finite(_dev.device_ptr(g), _dev.dtype_id(_dev.array_dtype(g)),
                                _dev.array_size(g), flags.data.ptr + 4 * i, 0)
        host = flags.get()
        return [n for (n, _), b in zip(named, host) if b]

    def _run_step(self, lib, key, t, dtype, buf_id, scale, wg, n_elems, stream):
        """The whole step as ONE launch (csrc/gp_step.cu) when it covers this
        configuration: float32 arrays, float32 / float16 / bfloat16 buffer, MomentumSGD or
        Adam without hooks, 1 rank or 2 / 4 / 8 ranks of one NVSwitch box.  False: not
        covered, nothing launched.  The decision depends only on replicated state, so all
        ranks take it together."""
        comm = self.comm
        want = comm.use_step
        if want is None:
            want = _step_enabled_by_env()
        if not want or key[0] not in ('momentum_sgd', 'adam'):
            return False
        hint = _lib.GP_F32 if t.all_float32 else 0     # no alignment requirement here
        adam_flags = key[9] if key[0] == 'adam' else 0
        if not lib.gp_step_supported(comm.size, buf_id, hint, scale, adam_flags):
            return False
        buf = comm.gpu_buffer_a
        handle = mc_ptr = None
        if comm.size > 1:
            p2p = comm._p2p
            if p2p is None:
                return False
            if comm._mc_active(buf):
                mc_ptr = buf._alloc.mc_ptr
            else:
                p2p.ensure(buf, stream)
            p2p.step_prepare(n_elems)
            handle = p2p.handle
        sp = stream.ptr
        with _lib.nvtx_range('gradpath.step (pack + allreduce + update, one launch)'):
            if key[0] == 'momentum_sgd':
                lib.gp_step_momentum_sgd(handle, mc_ptr, buf.ptr(), buf_id, t.d_csum, t.d_segs,
                                         t.n_params, n_elems, scale, key[1], key[2], wg, hint, sp)
            else:
                lib.gp_step_adam(handle, mc_ptr, buf.ptr(), buf_id, t.d_csum, t.d_segs,
                                 t.n_params, n_elems, scale, key[1], key[2], key[3], key[4],
                                 key[5], key[6], key[7], key[8], key[9], wg, hint, sp)
        return True


class _PlanStale(Exception):
    pass


def _step_enabled_by_env():
    import os
    return os.environ.get('CHAINER_B200_STEP', '1') not in ('0', '', 'false', 'False')


_GET_GRAD = operator.attrgetter('grad')


class _TableSet(object):
    """Device tables of a fused plan for ONE set of gradient arrays: a copy of the
    plan's segment tables with its own gradient-pointer column and device memory."""

    def __init__(self, plan):
        self.pd = plan.pd.clone()
        self.groups = []
        for rep, idx, sub in plan.groups:
            self.groups.append((rep, idx, self.pd if idx is None else sub.clone()))
        self.probe = None

    def update(self, ptrs, probe, stream):
        self.probe = probe
        self.pd.set_ptr0(ptrs, stream)
        for _, idx, sub in self.groups:
            if idx is not None:
                sub.set_ptr0(ptrs[idx], stream)


# hook classes the fused kernels implement (chainer_b200.integration adds the reference's)
WEIGHT_DECAY_HOOKS = ()
GRADIENT_CLIPPING_HOOKS = ()


def _hook_classes():
    global WEIGHT_DECAY_HOOKS, GRADIENT_CLIPPING_HOOKS
    if not WEIGHT_DECAY_HOOKS:
        from chainer_b200 import optimizer_hooks as H
        WEIGHT_DECAY_HOOKS = (H.WeightDecay,) + tuple(WEIGHT_DECAY_HOOKS)
        GRADIENT_CLIPPING_HOOKS = (H.GradientClipping,) + tuple(GRADIENT_CLIPPING_HOOKS)
    return WEIGHT_DECAY_HOOKS, GRADIENT_CLIPPING_HOOKS


def _fusable_hooks(optimizer):
    """The optimizer-level hooks as (GradientClipping or None, WeightDecay or None) when
    the fused kernels can apply them in registration order -- no hooks,
    [WeightDecay], [GradientClipping], [GradientClipping, WeightDecay] -- else None
    (other orders, other hooks, 'post' hooks: the reference sequence runs unfused)."""
    wd_classes, clip_classes = _hook_classes()
    hookable = getattr(optimizer, '_hookable', None)
    if hookable is None or hookable._post:
        return None
    pre = list(hookable._pre.values())
    kinds = ['wd' if type(h) in wd_classes else ('clip' if type(h) in clip_classes else '?')
             for h in pre]
    if not kinds:
        return (None, None)
    if kinds == ['wd']:
        return (None, pre[0])
    if kinds == ['clip']:
        return (pre[0], None)
    if kinds == ['clip', 'wd']:
        return (pre[0], pre[1])
    return None


def _is_master_plan(params):
    """True: every parameter is float16 with use_fp32_update and a rule the master kernels
    cover; False: none uses fp32 update; None: a mixture or an uncovered rule (not fusable)."""
    flags = []
    for p in params:
        rule = getattr(p, 'update_rule', None)
        flags.append(bool(rule is not None and rule._use_fp32_update and p.data is not None and
                          _dev.array_dtype(p.data) == np.float16))
    if not any(flags):
        return False
    if not all(flags):
        return None
    for p in params:
        rule = p.update_rule
        if getattr(rule, 'fused_kind', None) not in ('momentum_sgd', 'adam'):
            return None
        if rule.fused_kind == 'adam' and rule.hyperparam.amsgrad:
            return None
        if not hasattr(rule, 'fp32_param_for'):
            return None
    return True


def _fusion_plan(model, optimizer, zero_fill):
    """Decide whether ``optimizer.update(None)`` can be fused; returns
    (params_in_layout_order, other_params) or None."""
    if _fusable_hooks(optimizer) is None:
        return None
    if getattr(optimizer, 'target', None) is not model:
        return None
    params = _memory_utility.extract_params_set_grad(model, zero_fill)
    chosen = set(id(p) for p in params)
    others = [p for p in model.params() if id(p) not in chosen]
    # float32 master weights behind float16 parameters (use_fp32_update): fused when EVERY
    # parameter of the step is such a pair and the rule is MomentumSGD or Adam (gp_master.cu)
    master = _is_master_plan(params)
    if master is None:
        return None
    if getattr(optimizer, '_loss_scaling_is_dynamic', False) and not master:
        # dynamic loss scaling is fused with the master-weight kernels only (float16
        # training); elsewhere the reference sequence runs
        return None
    for p in params:
        rule = getattr(p, 'update_rule', None)
        if rule is None or getattr(rule, 'fused_kind', None) is None:
            return None
        if not rule.enabled or rule._hookable.has_hooks():
            return None
        ddt = _dev.array_dtype(p.data)
        if isinstance(ddt, str) or ddt not in (np.float16, np.float32, np.float64):
            return None
        if p.grad is not None and _dev.array_dtype(p.grad) != ddt:
            return None
        if rule.fused_kind == 'adam':
            interm = np.float32 if ddt == np.float16 else ddt.type
            rule._check_eps(interm)
    for p in others:
        # An initialised parameter outside the mean-grad set (grad None with
        # zero_fill=False) is still updated by the reference's optimizer.update():
        # reallocate_cleared_grads() gives it a zero gradient first
        # (chainer/optimizer.py:834-851, 881).  Leave that case to the unfused path.
        if p.data is not None:
            return None
        rule = getattr(p, 'update_rule', None)
        if rule is not None and rule._hookable.has_hooks():
            return None
    return params, others
