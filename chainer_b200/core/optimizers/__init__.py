from chainer_b200.core.optimizers.adam import Adam, AdamRule  # NOQA
from chainer_b200.core.optimizers.momentum_sgd import MomentumSGD, MomentumSGDRule  # NOQA
from chainer_b200.core.optimizers.sgd_family import (  # NOQA
    SGD, SGDRule, CorrectedMomentumSGD, CorrectedMomentumSGDRule, NesterovAG, NesterovAGRule)
