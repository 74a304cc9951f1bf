"""The slice of Chainer's ``Link`` / ``Optimizer`` surface that the gradient
path sits behind, for machines where Chainer itself cannot hold GPU arrays
(Chainer accepts only CuPy arrays on GPUs, ``chainer/variable.py:47-110``, and
CuPy 7 does not exist for sm_100).  The classes keep the reference's names,
attributes and semantics, so code written against ``chainer.Link`` /
``chainer.optimizers.MomentumSGD`` / ``Adam`` ports by changing the import."""
from chainer_b200.core.link import Chain, ChainList, Link, Parameter  # NOQA
from chainer_b200.core.optimizer import (GradientMethod, Hyperparameter,  # NOQA
                                          HyperparameterProxy, Optimizer, UpdateRule)
