"""The communicator interface of this package.

API contract (names, arguments, errors) of ``chainermn.CommunicatorBase``
(``chainermn/communicators/communicator_base.py:8-441``): the six rank / size
properties, the array and object collectives, ``bcast_data`` /
``multi_node_mean_grad`` with their deprecated aliases, and the configuration
mechanism.  The implementation is this package's own:

* the abstract surface is generated from one table (``_COLLECTIVES``) instead of being
  spelled out method by method;
* configuration values are ordinary instance attributes that are ALSO recorded in a
  per-class registry while a ``config_scope()`` is open -- the reference's behaviour that
  ``get_config()`` reports what was assigned inside such a scope, and that the registry is
  shared by the instances of a class (``communicator_base.py:38, 425-441``).

When the real ``chainermn`` is imported, ``chainer_b200.integration.install`` registers
the concrete communicator as a ``chainermn.CommunicatorBase`` too.
"""
import abc
import contextlib
import warnings

# name -> signature text; every entry becomes an abstract method that concrete
# communicators (MpiCommunicatorBase) implement
_COLLECTIVES = (
    ('split', '(self, color, key)'),
    ('alltoall', '(self, xs)'),
    ('send', '(self, data, dest, tag)'),
    ('recv', '(self, source, tag)'),
    ('bcast', '(self, data, max_buf_len=None, root=0)'),
    ('gather', '(self, data, root=0)'),
    ('allgather', '(self, x)'),
    ('allreduce', '(self, data)'),
    ('scatter', '(self, xs, root=0)'),
    ('send_obj', '(self, obj, dest, tag)'),
    ('recv_obj', '(self, source, tag)'),
    ('bcast_obj', '(self, obj, max_buf_len=None, root=0)'),
    ('gather_obj', '(self, obj, root=0)'),
    ('allreduce_obj', '(self, obj)'),
    ('bcast_data', '(self, model)'),
    ('multi_node_mean_grad', '(self, model, zero_fill=False)'),
)
_TOPOLOGY = ('rank', 'size', 'intra_rank', 'intra_size', 'inter_rank', 'inter_size')


def _abstract(name, signature):
    scope = {}
    exec('def {}{}:\n    raise NotImplementedError()'.format(name, signature), scope)
    fn = scope[name]
    fn.__doc__ = 'Abstract: see chainermn.CommunicatorBase.{}.'.format(name)
    return abc.abstractmethod(fn)


def _unimplemented_property(name):
    def getter(self):
        raise NotImplementedError()
    getter.__name__ = name
    return property(getter)


class _Meta(abc.ABCMeta):
    """Builds the abstract surface from the tables above."""

    def __new__(mcs, cls_name, bases, ns):
        if ns.get('_is_interface_root', False):
            for name, signature in _COLLECTIVES:
                ns.setdefault(name, _abstract(name, signature))
            for name in _TOPOLOGY:
                ns.setdefault(name, _unimplemented_property(name))
        return super(_Meta, mcs).__new__(mcs, cls_name, bases, ns)


class CommunicatorBase(metaclass=_Meta):

    _is_interface_root = True
    #: registry of configuration values assigned inside ``config_scope()``
    _configs = {}

    def __init__(self):
        object.__setattr__(self, '_config_depth', 0)

    # -- configuration ---------------------------------------------------------
    def set_config(self, name, **kwargs):
        """Subclasses handle the names they know and delegate here for the rest."""
        raise ValueError('Unknown config: {}'.format(name))

    def get_config(self, name=None):
        registry = type(self)._configs
        return registry if name is None else registry[name]

    @property
    def within_config_scope(self):
        return self.__dict__.get('_config_depth', 0) > 0

    @contextlib.contextmanager
    def config_scope(self):
        """Attributes assigned inside the scope are configuration values."""
        depth = self.__dict__.get('_config_depth', 0)
        object.__setattr__(self, '_config_depth', depth + 1)
        try:
            yield
        finally:
            object.__setattr__(self, '_config_depth', depth)

    def __setattr__(self, name, value):
        if self.__dict__.get('_config_depth', 0) > 0:
            type(self)._configs[name] = value
        object.__setattr__(self, name, value)

    # -- lifecycle and deprecated aliases --------------------------------------
    def finalize(self):
        pass

    def broadcast_data(self, model):
        warnings.warn('broadcast_data() is deprecated.', DeprecationWarning)
        self.bcast_data(model)

    def allreduce_grad(self, model, zero_fill=False):
        warnings.warn('allreduce_grad() is deprecated.', DeprecationWarning)
        self.multi_node_mean_grad(model, zero_fill)
