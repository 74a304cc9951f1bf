"""Minimal ``Parameter`` / ``Link`` / ``Chain`` / ``ChainList``.

Only what the gradient path consumes is kept:
``Link.namedparams`` / ``params`` (ordering as ``chainer/link.py:480, 967``),
``cleargrads`` (``link.py:574``), ``zerograds``, ``add_link``; ``Parameter``
with ``data`` / ``array``, ``grad``, ``update_rule``, ``update()``
(``chainer/variable.py:1941``).  Arrays are opaque device buffers.
"""
import collections
import contextlib
import copy

from chainer_b200 import device as _dev

# Bumped whenever the set of parameters, their initialised-ness or their update
# rules change.  Lets the per-step host code (is_changed, the fused-update plan)
# skip re-walking an unchanged model; gradients are NOT covered (they move every
# step and are re-read every step).
_structure_version = [0]


def structure_version():
    return _structure_version[0]


class Parameter(object):

    _b200_versioned = True

    def __init__(self, data=None, name=None):
        self._data = data
        self.grad = None
        self.name = name
        self._update_rule = None
        self._loss_scale = None
        _structure_version[0] += 1

    @property
    def data(self):
        return self._data

    @data.setter
    def data(self, value):
        if (value is None) != (self._data is None) or value is not self._data:
            _structure_version[0] += 1
        self._data = value

    @property
    def update_rule(self):
        return self._update_rule

    @update_rule.setter
    def update_rule(self, rule):
        _structure_version[0] += 1
        self._update_rule = rule

    @property
    def array(self):
        return self.data

    @array.setter
    def array(self, value):
        self.data = value

    @property
    def dtype(self):
        if self.data is None:
            raise RuntimeError('uninitialized parameter has no dtype')
        return _dev.array_dtype(self.data)

    @property
    def shape(self):
        return None if self.data is None else tuple(self.data.shape)

    @property
    def size(self):
        return _dev.array_size(self.data)

    def cleargrad(self):
        self.grad = None

    def zerograd(self):
        if self.data is None:
            return
        self.grad = _dev.zeros_like(self.data)

    def update(self):
        """``Parameter.update`` (``variable.py:1941``): run the update rule."""
        if self.update_rule is not None:
            self.update_rule.update(self)

    def __deepcopy__(self, memo):
        new = Parameter(name=self.name)
        new.data = _copy_array(self.data)
        new.grad = _copy_array(self.grad)
        new.update_rule = copy.deepcopy(self.update_rule, memo)
        return new


def _copy_array(a):
    if a is None:
        return None
    if _dev.is_torch(a):
        return a.clone()
    return a.copy()


def _assign_array(dst, src):
    """dst[...] = src for arrays of one module (torch tensors, NumPy arrays)."""
    if _dev.is_torch(dst):
        dst.copy_(src)
    else:
        dst[...] = src


class Link(object):

    _b200_versioned = True

    def __init__(self, **params):
        self._params = []
        self._persistent = []
        self._within_init_scope = False
        self.name = None
        for name, value in params.items():
            self.add_param(name, value)

    @contextlib.contextmanager
    def init_scope(self):
        old = self._within_init_scope
        self._within_init_scope = True
        try:
            yield
        finally:
            self._within_init_scope = old

    def __setattr__(self, name, value):
        if getattr(self, '_within_init_scope', False) and isinstance(value, Parameter):
            value.name = name
            if name not in self._params:
                self._params.append(name)
            _structure_version[0] += 1
        super(Link, self).__setattr__(name, value)

    def add_param(self, name, data=None):
        p = data if isinstance(data, Parameter) else Parameter(data)
        with self.init_scope():
            setattr(self, name, p)
        return p

    def add_persistent(self, name, value):
        """``link.py:463-481``: a value saved with the link that is not a parameter
        (running statistics of batch normalisation, counters)."""
        if name in self.__dict__ and name not in self._persistent:
            raise AttributeError('cannot register a new persistent value %s: attribute exists'
                                 % name)
        if name not in self._persistent:
            self._persistent.append(name)
        super(Link, self).__setattr__(name, value)

    def register_persistent(self, name):
        """``link.py:483-497``: mark an existing attribute as persistent."""
        if not hasattr(self, name):
            raise AttributeError('cannot register non-existent attribute %s as a persistent '
                                 'value' % name)
        if name not in self._persistent:
            self._persistent.append(name)

    def params(self, include_uninit=True):
        for name in sorted(self._params):
            p = self.__dict__[name]
            if include_uninit or p.data is not None:
                yield p

    def namedparams(self, include_uninit=True):
        for name in sorted(self._params):
            p = self.__dict__[name]
            if include_uninit or p.data is not None:
                yield '/' + name, p

    def links(self, skipself=False):
        if not skipself:
            yield self

    def namedlinks(self, skipself=False):
        """``link.py:704-715``: (path, link) pairs, the link itself at '/'."""
        if not skipself:
            yield '/', self

    def children(self):
        return iter(())

    def copyparams(self, link, copy_persistent=True):
        """``link.py:1008-1040``: copy the parameter values (and persistents) of a link
        of the same structure into this one, array by array."""
        src = dict(link.namedparams())
        for path, p in self.namedparams():
            q = src[path]
            if q.data is None:
                continue
            if p.data is None:
                p.data = _copy_array(q.data)
            else:
                _assign_array(p.data, q.data)
        if copy_persistent:
            mine = dict(self.namedlinks())
            for path, other in link.namedlinks():
                dst = mine[path]
                for name in other._persistent:
                    value = other.__dict__[name]
                    cur = dst.__dict__.get(name)
                    if hasattr(value, 'dtype') and cur is not None and hasattr(cur, 'dtype'):
                        _assign_array(cur, value)
                    else:
                        dst.add_persistent(name, copy.deepcopy(value))

    def cleargrads(self):
        for p in self.params():
            p.cleargrad()

    def zerograds(self):
        for p in self.params():
            p.zerograd()


class Chain(Link):

    def __init__(self, **links):
        super(Chain, self).__init__()
        self._children = []
        for name, link in links.items():
            self.add_link(name, link)

    def __setattr__(self, name, value):
        if getattr(self, '_within_init_scope', False) and isinstance(value, Link):
            value.name = name
            if name not in self._children:
                self._children.append(name)
            _structure_version[0] += 1
        super(Chain, self).__setattr__(name, value)

    def add_link(self, name, link):
        with self.init_scope():
            setattr(self, name, link)

    def params(self, include_uninit=True):
        for p in super(Chain, self).params(include_uninit):
            yield p
        for name in sorted(self._children):
            for p in self.__dict__[name].params(include_uninit):
                yield p

    def namedparams(self, include_uninit=True):
        for ret in super(Chain, self).namedparams(include_uninit):
            yield ret
        for name in sorted(self._children):
            prefix = '/' + name
            for path, p in self.__dict__[name].namedparams(include_uninit):
                yield prefix + path, p

    def links(self, skipself=False):
        if not skipself:
            yield self
        for name in sorted(self._children):
            for link in self.__dict__[name].links():
                yield link

    def namedlinks(self, skipself=False):
        if not skipself:
            yield '/', self
        for name in sorted(self._children):
            child = self.__dict__[name]
            prefix = '/' + name
            yield prefix, child
            for path, link in child.namedlinks(True):
                yield prefix + path, link

    def children(self):
        for name in sorted(self._children):
            yield self.__dict__[name]


class ChainList(Link):

    def __init__(self, *links):
        super(ChainList, self).__init__()
        self._children = []
        for link in links:
            self.add_link(link)

    def add_link(self, link):
        link.name = str(len(self._children))
        self._children.append(link)
        _structure_version[0] += 1

    append = add_link

    def __getitem__(self, i):
        return self._children[i]

    def __len__(self):
        return len(self._children)

    def params(self, include_uninit=True):
        for p in super(ChainList, self).params(include_uninit):
            yield p
        for link in self._children:
            for p in link.params(include_uninit):
                yield p

    def namedparams(self, include_uninit=True):
        for ret in super(ChainList, self).namedparams(include_uninit):
            yield ret
        for idx, link in enumerate(self._children):
            prefix = '/{}'.format(idx)
            for path, p in link.namedparams(include_uninit):
                yield prefix + path, p

    def links(self, skipself=False):
        if not skipself:
            yield self
        for child in self._children:
            for link in child.links():
                yield link

    def namedlinks(self, skipself=False):
        if not skipself:
            yield '/', self
        for idx, child in enumerate(self._children):
            prefix = '/{}'.format(idx)
            yield prefix, child
            for path, link in child.namedlinks(True):
                yield prefix + path, link

    def children(self):
        return iter(self._children)


def link_from_named_arrays(named_arrays):
    """Build a (possibly nested) Chain whose ``namedparams()`` yields exactly the
    given ``(path, array)`` pairs -- used to instantiate the benchmark models
    (``chainer_b200.workloads``) without their forward code."""
    root = Chain()
    for path, arr in named_arrays:
        parts = path.strip('/').split('/')
        node = root
        for part in parts[:-1]:
            child = node.__dict__.get(part)
            if child is None:
                child = Chain()
                node.add_link(part, child)
            node = child
        node.add_param(parts[-1], arr)
    return root
