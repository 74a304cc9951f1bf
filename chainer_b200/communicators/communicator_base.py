"""Interface of all communicators: mirror of
``chainermn/communicators/communicator_base.py:8-441`` (same method names,
same configuration mechanism, including the class-level ``_configs`` dict that
the reference shares between instances, ``communicator_base.py:38, 425-441``).
"""
from abc import ABCMeta
from abc import abstractmethod
import contextlib
import warnings


class CommunicatorBase(metaclass=ABCMeta):

    _configs = {}

    def __init__(self):
        self._within_config_scope = False

    @property
    def rank(self):
        raise NotImplementedError()

    @property
    def size(self):
        raise NotImplementedError()

    @property
    def intra_rank(self):
        raise NotImplementedError()

    @property
    def intra_size(self):
        raise NotImplementedError()

    @property
    def inter_rank(self):
        raise NotImplementedError()

    @property
    def inter_size(self):
        raise NotImplementedError()

    def set_config(self, name, **kwargs):
        raise ValueError('Unknown config: {}'.format(name))

    def get_config(self, name=None):
        if name is not None:
            return self._configs[name]
        return self._configs

    @abstractmethod
    def split(self, color, key):
        raise NotImplementedError()

    @abstractmethod
    def alltoall(self, xs):
        raise NotImplementedError()

    @abstractmethod
    def send(self, data, dest, tag):
        raise NotImplementedError()

    @abstractmethod
    def recv(self, source, tag):
        raise NotImplementedError()

    @abstractmethod
    def bcast(self, data, max_buf_len=None, root=0):
        raise NotImplementedError()

    @abstractmethod
    def gather(self, data, root=0):
        raise NotImplementedError()

    @abstractmethod
    def allgather(self, x):
        raise NotImplementedError()

    @abstractmethod
    def allreduce(self, data):
        raise NotImplementedError()

    @abstractmethod
    def scatter(self, xs, root=0):
        raise NotImplementedError()

    def finalize(self):
        pass

    @abstractmethod
    def send_obj(self, obj, dest, tag):
        raise NotImplementedError()

    @abstractmethod
    def recv_obj(self, source, tag):
        raise NotImplementedError()

    @abstractmethod
    def bcast_obj(self, obj, max_buf_len=None, root=0):
        raise NotImplementedError()

    @abstractmethod
    def gather_obj(self, obj, root=0):
        raise NotImplementedError()

    @abstractmethod
    def allreduce_obj(self, obj):
        raise NotImplementedError()

    @abstractmethod
    def bcast_data(self, model):
        raise NotImplementedError()

    def broadcast_data(self, model):
        warnings.warn('broadcast_data() is deprecated.', DeprecationWarning)
        self.bcast_data(model)

    @abstractmethod
    def multi_node_mean_grad(self, model, zero_fill=False):
        raise NotImplementedError()

    def allreduce_grad(self, model, zero_fill=False):
        warnings.warn('allreduce_grad() is deprecated.', DeprecationWarning)
        self.multi_node_mean_grad(model, zero_fill)

    @property
    def within_config_scope(self):
        return getattr(self, '_within_config_scope', False)

    @contextlib.contextmanager
    def config_scope(self):
        old_flag = self.within_config_scope
        self._within_config_scope = True
        try:
            yield
        finally:
            self._within_config_scope = old_flag

    def __setattr__(self, name, value):
        if self.within_config_scope:
            self._configs[name] = value
        super(CommunicatorBase, self).__setattr__(name, value)
