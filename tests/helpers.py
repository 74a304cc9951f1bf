"""Shared helpers of the test-suite."""
import numpy as np


class P(object):
    """Protocol-conformant parameter: .data, .grad, .update_rule, .name."""

    def __init__(self, data=None, grad=None, name=None):
        self.data = data
        self.grad = grad
        self.name = name
        self.update_rule = None

    @property
    def array(self):
        return self.data

    @property
    def dtype(self):
        return self.data.dtype


def torch_dtype(np_dtype):
    import torch
    return {np.dtype(np.float16): torch.float16, np.dtype(np.float32): torch.float32,
            np.dtype(np.float64): torch.float64}[np.dtype(np_dtype)]


def to_dev(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def to_host(t):
    import torch
    if t.dtype == torch.bfloat16:
        return t.float().cpu().numpy()
    return t.cpu().numpy()


def bits(a):
    a = np.ascontiguousarray(a)
    return a.view({2: np.uint16, 4: np.uint32, 8: np.uint64}[a.dtype.itemsize])


def assert_bits_equal(a, b, what=''):
    a = np.ascontiguousarray(a)
    b = np.ascontiguousarray(b)
    assert a.dtype == b.dtype and a.shape == b.shape, (what, a.dtype, b.dtype, a.shape, b.shape)
    if a.size == 0:
        return
    neq = bits(a) != bits(b)
    # NaNs with different payloads are still equal for our purposes
    both_nan = np.isnan(a) & np.isnan(b)
    bad = neq & ~both_nan
    if bad.any():
        i = np.argwhere(bad)[0]
        raise AssertionError('{}: {} of {} elements differ; first at {}: {!r} vs {!r}'.format(
            what, int(bad.sum()), a.size, tuple(i), a[tuple(i)], b[tuple(i)]))
