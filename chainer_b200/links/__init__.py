from chainer_b200.links.batch_normalization import MultiNodeBatchNormalization  # NOQA
