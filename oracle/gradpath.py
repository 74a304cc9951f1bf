"""NumPy restatement of the reference `pure_nccl` gradient path and optimizers.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Every function cites the
reference lines it restates (paths relative to chainer/chainer v7.8.1).
"""
import math

import numpy as np

NCCL_FLOAT16, NCCL_FLOAT32, NCCL_FLOAT64, NCCL_BFLOAT16 = 6, 7, 8, 9

BF16 = 'bfloat16'  # marker: NumPy has no bfloat16; values are carried as float32


# ----------------------------------------------------------------- bfloat16 --
def bf16_round(x):
    """Round float32/float64 values to the nearest bfloat16 (ties to even) and
    return them as float32.  No reference counterpart (extension dtype)."""
    x = np.asarray(x)
    if x.dtype == np.float64:
        # float64 -> float32 -> bf16 could double-round: round directly
        return _bf16_from_f64(x)
    x = x.astype(np.float32)
    u = x.view(np.uint32).astype(np.uint64)
    rounding_bias = ((u >> 16) & 1) + 0x7FFF
    r = ((u + rounding_bias) >> 16) << 16
    out = r.astype(np.uint32).view(np.float32)
    nan = np.isnan(x)
    if np.any(nan):
        out = np.where(nan, np.float32(np.nan), out)
    return out


def _bf16_from_f64(x):
    """float64 -> bfloat16 with ONE rounding (ties to even), returned as float32."""
    x = np.asarray(x, dtype=np.float64)
    u = x.view(np.uint64)
    # normal range: keep sign(1) + exponent(11) + 7 mantissa bits, RNE on the other 45
    bias = ((u >> np.uint64(45)) & np.uint64(1)) + np.uint64((1 << 44) - 1)
    r = ((u + bias) >> np.uint64(45)) << np.uint64(45)
    normal = r.view(np.float64)
    # bfloat16 subnormals (|x| < 2^-126) are multiples of 2^-133: the classic
    # add-and-subtract trick rounds to that grid with ties to even
    c = np.copysign(np.float64(2.0) ** -81, x)
    sub = (x + c) - c
    out = np.where(np.abs(x) < np.float64(2.0) ** -126, sub, normal)
    out = np.where(np.isfinite(x), out, x)
    with np.errstate(over='ignore'):
        return out.astype(np.float32)   # exact (or +-inf on overflow)


def cast(x, dtype):
    """x.astype(dtype) with bfloat16 support (values carried as float32)."""
    if _is_bf16(dtype):
        return bf16_round(x)
    return np.asarray(x).astype(dtype)


def _is_bf16(dtype):
    return isinstance(dtype, str) and dtype == BF16


def carrier_dtype(dtype):
    return np.dtype(np.float32) if _is_bf16(dtype) else np.dtype(dtype)


def itemsize(dtype):
    return 2 if _is_bf16(dtype) else np.dtype(dtype).itemsize


def nccl_type_id(dtype):
    """_communication_utility.py:177-186 (+ bfloat16 = 9 as an extension)."""
    if _is_bf16(dtype):
        return NCCL_BFLOAT16
    dtype = np.dtype(dtype)
    if dtype == np.float16:
        return NCCL_FLOAT16
    elif dtype == np.float32:
        return NCCL_FLOAT32
    elif dtype == np.float64:
        return NCCL_FLOAT64
    raise ValueError('dtype must be float16, float32, or float64.')


# ------------------------------------------------------------------- layout --
def extract_params_set_data(named_params):
    """_memory_utility.py:154-156.  `named_params`: iterable of (name, param);
    a param is anything with .data / .grad."""
    return [p for _, p in sorted(named_params, key=lambda kv: kv[0]) if p.data is not None]


def extract_params_set_grad(named_params, zero_fill):
    """_memory_utility.py:159-165."""
    if zero_fill:
        return [p for _, p in sorted(named_params, key=lambda kv: kv[0]) if p.data is not None]
    return [p for _, p in sorted(named_params, key=lambda kv: kv[0])
            if p.data is not None and p.grad is not None]


def size_csum(arrays):
    """ParamsData.size_csum (_memory_utility.py:36-57), as int64."""
    cs = np.zeros(len(arrays) + 1, dtype=np.int64)
    for i, a in enumerate(arrays):
        cs[i + 1] = cs[i] + a.size
    return cs


def pack(arrays, buf_dtype, scale=1.0):
    """cupy_batched_pack_params (_memory_utility.py:289-358): element k of array
    j lands at flat index csum[j] + k (C order), cast to the buffer dtype.
    `scale` != 1 is the pre-scaled variant (product evaluated in double, one
    rounding)."""
    cs = size_csum(arrays)
    out = np.zeros(int(cs[-1]), dtype=carrier_dtype(buf_dtype))
    for j, a in enumerate(arrays):
        v = np.asarray(a).reshape(-1)
        if scale != 1.0:
            v = v.astype(np.float64) * np.float64(scale)
        out[cs[j]:cs[j + 1]] = cast(v, buf_dtype)
    return out


def scale_buffer(buf, buf_dtype, scale):
    """div_by_size (pure_nccl_communicator.py:183-189): `x *= (1.0/size)` with a
    double literal: the product is formed in double and rounded once to the
    buffer dtype."""
    if scale == 1.0:
        return buf
    return cast(buf.astype(np.float64) * np.float64(scale), buf_dtype)


def unpack(buf, arrays_like, buf_dtype):
    """cupy_batched_unpack_params (_memory_utility.py:361-429): returns the list
    of arrays cast to each destination's dtype."""
    cs = size_csum(arrays_like)
    out = []
    for j, a in enumerate(arrays_like):
        out.append(buf[cs[j]:cs[j + 1]].astype(a.dtype).reshape(a.shape))
    return out


def allreduce_sum(buffers, buf_dtype):
    """nccl allReduce(SUM) of the packed buffers of all ranks
    (pure_nccl_communicator.py:180-182).  The summation order inside NCCL is
    not specified; here ranks are added in order, every partial sum rounded to
    the buffer dtype (what a ring does for float16)."""
    acc = buffers[0].copy()
    for b in buffers[1:]:
        if _is_bf16(buf_dtype):
            acc = bf16_round(acc + b)
        else:
            acc = (acc + b).astype(carrier_dtype(buf_dtype))
    return acc


def multi_node_mean_grad(rank_grads, buf_dtype):
    """PureNcclCommunicator._multi_node_mean_grad_async
    (pure_nccl_communicator.py:105-139) for all ranks at once.

    rank_grads[r] = list of gradient arrays of rank r (already in layout order).
    Returns the list of mean gradients (identical on all ranks), each in the
    dtype of the corresponding gradient array of rank 0.
    """
    size = len(rank_grads)
    packed = [pack(g, buf_dtype) for g in rank_grads]
    summed = allreduce_sum(packed, buf_dtype)
    mean = scale_buffer(summed, buf_dtype, 1.0 / size)
    return unpack(mean, rank_grads[0], buf_dtype)


# --------------------------------------------------------------- MomentumSGD --
def momentum_sgd_update(param, grad, v, lr=0.01, momentum=0.9):
    """MomentumSGDRule.update_core_cpu (momentum_sgd.py:61-73), in place:
        v *= momentum;  v -= lr * grad;  param += v
    (the GPU kernel, momentum_sgd.py:80-85, computes the same expression; its
    only possible difference is FMA contraction by NVRTC)."""
    v *= momentum
    v -= lr * grad
    param += v


# ------------------------------------------ SGD, CorrectedMomentumSGD, NesterovAG --
def sgd_update(param, grad, lr=0.01):
    """SGDRule.update_core_cpu (chainer/optimizers/sgd.py:45-52), in place."""
    param -= lr * grad


def corrected_momentum_sgd_update(param, grad, v, lr=0.01, momentum=0.9):
    """CorrectedMomentumSGDRule.update_core_cpu
    (chainer/optimizers/corrected_momentum_sgd.py:61-75), in place."""
    v *= momentum
    v -= grad
    param += lr * v


def nesterov_ag_update(param, grad, v, lr=0.01, momentum=0.9):
    """NesterovAGRule.update_core_cpu (chainer/optimizers/nesterov_ag.py:60-71), in
    place.  (The reference's GPU kernel, :73-84, adds `m*m*v - (1+m)*lr*grad` to the
    parameter in one expression: one rounding less, tolerance-level difference.)"""
    v *= momentum
    v -= lr * grad
    param += momentum * momentum * v
    param -= (1 + momentum) * lr * grad


# ------------------------------------------------- optimizer hooks, loss scale --
def weight_decay_hook(param, grad, rate, loss_scale=None):
    """WeightDecay.__call__, CPU branch (chainer/optimizer_hooks/weight_decay.py:44-57),
    in place: ``rate *= param._loss_scale`` (when set); ``g += rate * p``."""
    if loss_scale is not None:
        rate = rate * loss_scale
    grad += rate * param


def gradient_clipping_hook(grads, threshold):
    """GradientClipping.__call__, CPU branch (chainer/optimizer_hooks/
    gradient_clipping.py:9-52, 84-106), in place on the list of gradient arrays:
    sqnorm = sum_i g_i.ravel().dot(g_i.ravel()); rate = threshold / sqrt(sqnorm);
    if rate < 1: g_i *= rate.  Returns the rate (>= 1: nothing was scaled)."""
    dots = []
    for g in grads:
        r = g.ravel()
        dots.append(r.dot(r))
    sqnorm = sum(dots)
    norm = np.sqrt(sqnorm)
    with np.errstate(divide='ignore'):
        rate = threshold / norm
    if rate >= 1:
        return rate
    for g in grads:
        g *= rate
    return rate


def loss_scale_divide(grad, loss_scale):
    """UpdateRule.update (chainer/optimizer.py:286-291), in place: ``grad /= loss_scale``."""
    grad /= loss_scale


# ---------------------------------------------------------------------- Adam --
def adam_alpha_t(alpha, beta1, beta2, t):
    """_learning_rate (adam.py:47-54)."""
    if t == 0:
        raise RuntimeError('Can\'t determine the learning rate of Adam optimizer '
                           'because the update steps have not been started.')
    fix1 = 1. - math.pow(beta1, t)
    fix2 = 1. - math.pow(beta2, t)
    return alpha * math.sqrt(fix2) / fix1


def adam_bounds(final_lr, alpha, initial_alpha, gamma, t):
    """AdamRule.bounds (adam.py:346-358)."""
    final_lr = final_lr * alpha / initial_alpha
    lower = final_lr * (1.0 - 1.0 / (gamma * t + 1))
    upper = final_lr * (1.0 + 1.0 / (gamma * t))
    return lower, upper


def _intermediate_dtype(dtype):
    """_get_intermediate_dtype (adam.py:57-63)."""
    return np.dtype(np.float32) if np.dtype(dtype) == np.float16 else np.dtype(dtype)


def adam_update_cpu(param, grad, m, v, t, alpha=0.001, beta1=0.9, beta2=0.999, eps=1e-8,
                    eta=1.0, weight_decay_rate=0.0, amsgrad=False, vhat=None,
                    adabound=False, final_lr=0.1, gamma=1e-3, initial_alpha=None):
    """AdamRule.update_core_cpu (adam.py:189-222), in place."""
    dtype = _intermediate_dtype(param.dtype).type
    grad = grad.astype(dtype, copy=False)
    m += (1.0 - beta1) * (grad - m)
    v += (1.0 - beta2) * (grad * grad - v)
    if amsgrad:
        np.maximum(vhat, v, out=vhat)
        vh = vhat
    else:
        vh = v
    vh = vh.astype(dtype, copy=False)
    step = adam_alpha_t(alpha, beta1, beta2, t) / (np.sqrt(vh) + eps)
    if adabound:
        lower, upper = adam_bounds(final_lr, alpha, initial_alpha or alpha, gamma, t)
        step = np.clip(step, lower, upper)
    a = 1.0 - eta * weight_decay_rate
    if a == 1:
        param += -eta * (step * m)
    else:
        param[...] = a * param + -eta * (step * m)


def adam_update_gpu(param, grad, m, v, t, alpha=0.001, beta1=0.9, beta2=0.999, eps=1e-8,
                    eta=1.0, weight_decay_rate=0.0, amsgrad=False, vhat=None,
                    adabound=False, final_lr=0.1, gamma=1e-3, initial_alpha=None):
    """The `adam` / `amsgrad` / `adabound` / `amsbound` ElementwiseKernels
    (adam.py:237-332) restated operation by operation, every operation rounded
    to the intermediate type T (no FMA contraction), in place:

        T grad_ = grad; T m_ = m; T v_ = v; [T vhat_ = vhat;]
        m_ += one_minus_beta1 * (grad_ - m_);
        v_ += one_minus_beta2 * (grad_ * grad_ - v_);
        [vhat_ = max(vhat_, v_); vhat = vhat_;]
        m = m_; v = v_;
        param -= eta * (alpha_t * m_ / (sqrt(v_) + eps) + weight_decay_rate * param);
    """
    P = param.dtype
    T = _intermediate_dtype(P).type
    alpha_t = T(adam_alpha_t(alpha, beta1, beta2, t))
    omb1, omb2 = T(1 - beta1), T(1 - beta2)
    eps_, eta_, wd_ = T(eps), T(eta), T(weight_decay_rate)
    g_ = grad.astype(T)
    m_ = m.astype(T)
    v_ = v.astype(T)
    m_ = m_ + omb1 * (g_ - m_)
    v_ = v_ + omb2 * (g_ * g_ - v_)
    if amsgrad:
        vh_ = np.maximum(vhat.astype(T), v_)
        vhat[...] = vh_.astype(P)
        d_ = vh_
    else:
        d_ = v_
    m[...] = m_.astype(P)
    v[...] = v_.astype(P)
    denom = np.sqrt(d_) + eps_
    if adabound:
        lower, upper = adam_bounds(final_lr, alpha, initial_alpha or alpha, gamma, t)
        step = np.maximum(np.minimum(alpha_t / denom, T(upper)), T(lower)) * m_
    else:
        step = alpha_t * m_ / denom
    p_ = param.astype(T)
    param[...] = (p_ - eta_ * (step + wd_ * p_)).astype(P)


# ---------------------------------------------------- fused path restatement --
def mean_grad_value(summed_buf, buf_dtype, size, grad_dtype):
    """What the fused kernels feed the update with: buffer -> x*(1.0/size) in
    double, rounded to the buffer dtype -> cast to the gradient dtype."""
    return scale_buffer(summed_buf, buf_dtype, 1.0 / size).astype(grad_dtype)


# ------------------------------------------------- batch-norm statistics ------
def bn_fwd_stats(x, out_dtype):
    """_NcclImpl.get_mean_and_var, local part
    (chainermn/functions/batch_normalization.py:53-56): per-channel mean and
    mean of squares over every axis but 1, accumulated in gamma.dtype."""
    axis = (0,) + tuple(range(2, x.ndim))
    mean = x.mean(axis=axis, dtype=out_dtype)
    sqmean = np.square(x).mean(axis=axis, dtype=out_dtype)
    return np.concatenate([mean, sqmean]).astype(out_dtype)


def bn_mean_var_from_stats(rank_stats, dtype):
    """allreduce + div_by_size + `var = sqmean - square(mean)`
    (functions/batch_normalization.py:57-67)."""
    size = len(rank_stats)
    s = allreduce_sum([np.asarray(r, dtype=dtype) for r in rank_stats], dtype)
    s = scale_buffer(s, dtype, 1.0 / size)
    C = s.size // 2
    mean, sqmean = s[:C], s[C:]
    var = sqmean - np.square(mean)
    return mean, var


def bn_bwd_stats(gy, x_hat, out_dtype):
    """_NcclImpl.get_ggamma_and_gbeta, local part (:79-82): [sum gy, sum gy*x_hat]."""
    axis = (0,) + tuple(range(2, gy.ndim))
    gbeta = gy.sum(axis=axis, dtype=out_dtype)
    ggamma = (gy * x_hat).sum(axis=axis, dtype=out_dtype)
    return np.concatenate([gbeta, ggamma]).astype(out_dtype)


def x_hat(x, mean, inv_std):
    """_x_hat (chainer/functions/normalization/batch_normalization.py):
    x_mu = x - mean; x_mu *= inv_std."""
    shape = (1, -1) + (1,) * (x.ndim - 2)
    x_mu = x - mean.reshape(shape)
    x_mu *= inv_std.reshape(shape)
    return x_mu


# --------------------------------------- batch-norm elementwise halves ------
def _interm(x_dtype, stat_dtype):
    """The type the reference's `bn_fwd` / `bn_bwd` ElementwiseKernels compute in: float16
    operands are promoted to float, so float unless a float64 is involved."""
    return np.float64 if np.float64 in (np.dtype(x_dtype), np.dtype(stat_dtype)) else np.float32


def bn_inv_std(var, eps):
    """`inv_std = rsqrt(var + eps)` (chainer/functions/normalization/
    batch_normalization.py:40-45), correctly rounded in the statistics' type."""
    T = np.float64 if var.dtype == np.float64 else np.float32
    v = var.astype(T) + T(eps)
    return (1.0 / np.sqrt(v.astype(np.float64))).astype(T).astype(var.dtype)


def bn_fwd_apply(x, mean, var, gamma, beta, eps):
    """The `bn_fwd` kernel (:864-867) `y = gamma * (x - mean) * inv_std + beta` with
    `inv_std = rsqrt(var + eps)`, every operation rounded to the intermediate type in the
    kernel's order (no FMA contraction), result cast to x.dtype."""
    T = _interm(x.dtype, gamma.dtype)
    shape = (1, -1) + (1,) * (x.ndim - 2)
    inv_std = bn_inv_std(var, eps).astype(T).reshape(shape)
    g, b, m = (a.astype(T).reshape(shape) for a in (gamma, beta, mean))
    y = g * (x.astype(T) - m)
    y = y * inv_std
    y = y + b
    return y.astype(x.dtype)


def bn_running_update(running_mean, running_var, mean, var, decay, adjust):
    """`update_mean_var` (:69-77), in place:
    r_mean = r_mean * decay + mean * (1 - decay); r_var = r_var * decay + var * (1 - decay) * adjust."""
    T = _interm(mean.dtype, running_mean.dtype)
    d, a = T(decay), T(adjust)
    omd = T(1) - d
    rm = running_mean.astype(T) * d + mean.astype(T) * omd
    rv = running_var.astype(T) * d + (var.astype(T) * omd) * a
    running_mean[...] = rm.astype(running_mean.dtype)
    running_var[...] = rv.astype(running_var.dtype)


def bn_bwd_apply(gy, x, mean, inv_std, gamma, ggamma, gbeta, inv_m):
    """The `bn_bwd` kernel (:121-133) with x_hat formed on the fly (`_x_hat`, :826-829):
    gx = (gamma * inv_std) * (gy - (x_hat * ggamma + gbeta) * inv_m)."""
    T = _interm(x.dtype, gamma.dtype)
    shape = (1, -1) + (1,) * (x.ndim - 2)
    m, s, g, gg, gb = (a.astype(T).reshape(shape) for a in (mean, inv_std, gamma, ggamma, gbeta))
    xh = (x.astype(T) - m) * s
    t = xh * gg
    t = t + gb
    t = t * T(inv_m)
    gx = (g * s) * (gy.astype(T) - t)
    return gx.astype(x.dtype)
