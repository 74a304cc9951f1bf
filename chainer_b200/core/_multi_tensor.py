"""Multi-tensor ``GradientMethod.update``: every ``UpdateRule.update(param)`` of
one optimizer step (``chainer/optimizer.py:886-889``: one ElementwiseKernel launch
per parameter in the reference) as one fused launch per (dtype, hyperparameter,
step-count) group, reading each gradient array directly (``buffer == NULL`` mode
of gp_unpack_momentum_sgd / gp_unpack_adam)."""
import numpy as np

from chainer_b200 import _lib
from chainer_b200 import device as _dev
from chainer_b200.communicators import _memory_utility


def _eligible(opt):
    # (optimizer-level hooks run before/after this in GradientMethod.update and do
    # not interfere; per-rule hooks and loss scaling need the per-parameter path)
    params = []
    for p in opt.target.params():
        rule = p.update_rule
        if rule is None:
            continue
        if getattr(rule, 'fused_kind', None) is None or rule._hookable.has_hooks() or \
                rule._use_fp32_update or getattr(p, '_loss_scale', None) is not None:
            return None
        params.append(p)
    return params


def update(opt):
    if getattr(opt, '_loss_scale', None) is not None:
        return False
    params = _eligible(opt)
    if params is None:
        return False
    lib = _lib.get()
    groups = {}
    for p in params:
        rule = p.update_rule
        if not rule.enabled:
            continue
        # UpdateRule.update (optimizer.py:236-250): t += 1 even when nothing is updated
        rule.t += 1
        if p.data is None or p.grad is None:
            continue
        ddt = _dev.array_dtype(p.data)
        if isinstance(ddt, str) or ddt not in (np.float16, np.float32, np.float64) or \
                _dev.array_dtype(p.grad) != ddt:
            # cannot happen for well-formed models; fall back for this parameter
            rule.t -= 1
            p.update()
            continue
        rule._init_states(p)
        if rule.fused_kind == 'adam':
            rule._check_eps(np.float32 if ddt == np.float16 else ddt.type)
        key = (ddt.str,) + rule.fused_key()
        groups.setdefault(key, []).append(p)
    # one DeviceTable (a ring of pinned staging slots + device copies) per launch group,
    # reused every step: no per-step cudaHostAlloc / cudaMalloc / cudaFree
    tables = getattr(opt, '_mt_tables', None)
    if tables is None:
        tables = opt._mt_tables = []
    for gi, (key, plist) in enumerate(groups.items()):
        extra = [(p.data, [p.update_rule.state[k] for k in p.update_rule.state_names])
                 for p in plist]
        while len(tables) <= gi:
            tables.append(_memory_utility.DeviceTable())
        pd = _memory_utility.ParamsData(plist, 'grad', False, extra_ptrs=extra, table=tables[gi])
        dt_id = _dev.dtype_id(np.dtype(key[0]))
        k = key[1:]
        if k[0] == 'momentum_sgd':
            lib.gp_unpack_momentum_sgd(None, dt_id, pd.d_csum, pd.d_segs, pd.n_params, 0,
                                       pd.n_elems, 1.0, k[1], k[2], 0,
                                       pd.layout_hint(np.dtype(key[0])), 0)
        elif k[0] == 'sgd_family':
            lib.gp_unpack_sgd_family(None, dt_id, pd.d_csum, pd.d_segs, pd.n_params, 0, pd.n_elems,
                                     1.0, k[1], k[2], k[3], 0, pd.layout_hint(np.dtype(key[0])),
                                     None, 0)
        else:
            lib.gp_unpack_adam(None, dt_id, pd.d_csum, pd.d_segs, pd.n_params, 0, pd.n_elems, 1.0,
                               k[1], k[2], k[3], k[4], k[5], k[6], k[7], k[8], k[9], 0,
                               pd.layout_hint(np.dtype(key[0])), 0)
    return True
