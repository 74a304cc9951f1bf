"""INTEGRATION.md applied to an imported, UNMODIFIED chainer / chainermn (v7.8.1).

``install(chainer, chainermn)`` is the patch a maintainer would make, done at run time so
that it can be exercised against the reference's own classes without editing its files
(``tests/test_dropin_reference.py`` does, with the C-ABI double on CPU; with CuPy arrays
and the real library the same objects run on the GPU):

1. ``chainermn.create_communicator('pure_nccl')`` returns this package's
   :class:`PureNcclCommunicator`, which is also what
   ``chainermn.communicators.pure_nccl_communicator.PureNcclCommunicator`` names afterwards
   (the ``isinstance`` checks of ``chainermn/optimizers.py:174-178`` and
   ``chainermn/functions/batch_normalization.py:104-114``) and a registered
   ``CommunicatorBase``.
2. ``chainermn.optimizers._MultiNodeOptimizer.update`` offers the step to
   ``communicator.multi_node_mean_grad_and_update`` first and
   runs the reference's two calls when that declines.
3. The reference's ``MomentumSGDRule`` / ``AdamRule`` and hook containers get the few
   read-only accessors the fused plan asks an update rule for (``fused_kind``,
   ``state_names``, ``fused_key`` ...): their hyperparameters, ``t``, ``state`` and
   ``alpha_t`` are used as they are.
4. ``chainermn.functions.batch_normalization._NcclImpl`` is this package's (section 5).

Nothing here touches arithmetic: it only tells the fused path where the reference keeps
what it needs.
"""
from chainer_b200 import _lib


def _register_hookable(chainer):
    H = chainer.optimizer._Hookable
    if getattr(H, '_b200_patched', False):
        return
    H._pre = property(lambda self: self._pre_update_hooks)
    H._post = property(lambda self: self._post_update_hooks)
    H.has_hooks = lambda self: bool(self._pre_update_hooks) or bool(self._post_update_hooks)
    H._b200_patched = True


def _register_rules(chainer):
    from chainer.optimizers import adam as ref_adam
    from chainer.optimizers import momentum_sgd as ref_sgd
    S = ref_sgd.MomentumSGDRule
    S.fused_kind = 'momentum_sgd'
    S.state_names = ('v',)
    S.fused_key = lambda self: ('momentum_sgd', float(self.hyperparam.lr),
                                float(self.hyperparam.momentum))
    S.fused_signature = lambda self: self.fused_key() + (self.t,)

    A = ref_adam.AdamRule
    A.fused_kind = 'adam'
    A.state_names = property(
        lambda self: ('m', 'v', 'vhat') if self.hyperparam.amsgrad else ('m', 'v'))

    def kernel_args(self):
        hp = self.hyperparam
        lower, upper = self.bounds if hp.adabound else (0.0, 0.0)
        flags = (_lib.GP_ADAM_AMSGRAD if hp.amsgrad else 0) | \
            (_lib.GP_ADAM_ADABOUND if hp.adabound else 0)
        return (float(self.alpha_t), float(1 - hp.beta1), float(1 - hp.beta2), float(hp.eps),
                float(hp.eta), float(hp.weight_decay_rate), float(lower), float(upper), flags)
    A.kernel_args = kernel_args
    A.fused_key = lambda self: ('adam',) + self.kernel_args()
    A.fused_signature = lambda self: ('adam', self.t, getattr(self, 'initial_alpha', None)) + \
        tuple(sorted((k, float(v)) for k, v in self.hyperparam.get_dict().items()))


def _register_hooks(chainer):
    """The reference's WeightDecay / GradientClipping hook objects are recognised by the
    fused plan (``rate`` / ``threshold`` are read from them)."""
    from chainer_b200 import optimizer_hooks as H
    from chainer_b200.communicators import pure_nccl_communicator as pnc
    ref = chainer.optimizer_hooks
    wd, clip = pnc._hook_classes()
    if ref.WeightDecay not in wd:
        pnc.WEIGHT_DECAY_HOOKS = wd + (ref.WeightDecay,)
    if ref.GradientClipping not in clip:
        pnc.GRADIENT_CLIPPING_HOOKS = clip + (ref.GradientClipping,)

    def scratch(self):
        sc = getattr(self, '_b200_scratch', None)
        if sc is None:
            sc = self._b200_scratch = H._NormScratch()
        return sc
    ref.GradientClipping.scratch = scratch


def _patch_multi_node_optimizer(chainermn):
    M = chainermn.optimizers._MultiNodeOptimizer
    if getattr(M, '_b200_patched', False):
        return

    def update(self, lossfun=None, *args, **kwds):
        # chainermn/optimizers.py:17-33 with INTEGRATION.md section 4 applied
        target = self.target
        if lossfun is not None:
            use_cleargrads = getattr(self, '_use_cleargrads', True)
            loss = lossfun(*args, **kwds)
            if use_cleargrads:
                target.cleargrads()
            else:
                target.zerograds()
            loss.backward(loss_scale=self.actual_optimizer._loss_scale)
            del loss
        if self.is_changed(target):
            self.communicator.bcast_data(target)
        else:
            fused = getattr(self.communicator, 'multi_node_mean_grad_and_update', None)
            if fused is None or not fused(target, self.actual_optimizer, self.zero_fill):
                self.communicator.multi_node_mean_grad(target, self.zero_fill)
                self.actual_optimizer.update(None, *args, **kwds)
    M.update = update
    M._b200_patched = True


def _patch_communicators(chainermn):
    from chainer_b200.communicators import create_communicator as ours_create
    from chainer_b200.communicators.pure_nccl_communicator import PureNcclCommunicator
    import chainermn.communicators as cc
    from chainermn.communicators import communicator_base, pure_nccl_communicator
    if getattr(cc, '_b200_patched', False):
        return
    communicator_base.CommunicatorBase.register(PureNcclCommunicator)
    pure_nccl_communicator.PureNcclCommunicator = PureNcclCommunicator
    ref_create = cc.create_communicator

    def create_communicator(communicator_name='pure_nccl', mpi_comm=None, **kwargs):
        if communicator_name == 'pure_nccl':
            return ours_create('pure_nccl', mpi_comm=mpi_comm, **kwargs)
        return ref_create(communicator_name, mpi_comm, **kwargs)
    cc.create_communicator = create_communicator
    chainermn.create_communicator = create_communicator
    cc._b200_patched = True


def _patch_mnbn(chainermn):
    # the two method bodies of chainermn/functions/batch_normalization.py:44-93; the class
    # (a GeneralBatchNormalizationImpl: y, gx, running statistics) stays the reference's
    from chainer_b200.functions import batch_normalization as ours
    import chainermn.functions.batch_normalization as ref
    ref._NcclImpl.get_mean_and_var = ours._NcclImpl.get_mean_and_var
    ref._NcclImpl.get_ggamma_and_gbeta = ours._NcclImpl.get_ggamma_and_gbeta
    ref._NcclImpl._mean_over_ranks = ours._NcclImpl._mean_over_ranks


def install(chainer, chainermn):
    """Apply the integration to the imported reference modules (idempotent)."""
    _register_hookable(chainer)
    _register_rules(chainer)
    _register_hooks(chainer)
    _patch_multi_node_optimizer(chainermn)
    _patch_communicators(chainermn)
    _patch_mnbn(chainermn)
