from chainer_b200.links.batch_normalization import MultiNodeBatchNormalization  # NOQA
from chainer_b200.links.batch_normalization import BatchNormalization  # NOQA
from chainer_b200.links.create_mnbn_model import create_mnbn_model  # NOQA
