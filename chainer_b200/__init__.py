"""chainer_b200 -- B200-native (sm_100a) implementation of ChainerMN's
data-parallel gradient path."""
__version__ = '0.1.0'
