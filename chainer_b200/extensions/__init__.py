from chainer_b200.extensions.allreduce_persistent import AllreducePersistent  # NOQA
