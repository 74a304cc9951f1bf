"""One forward + backward of the BN link at two ResNet-50 layer shapes (for ncu)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402


def main():
    from chainer_b200.links import BatchNormalization
    for shape in ((32, 64, 112, 112), (32, 512, 28, 28), (32, 2048, 7, 7)):
        bn = BatchNormalization(shape[1])
        x = torch.randn(*shape, device='cuda', requires_grad=True)
        for _ in range(2):
            y = bn(x)
            y.backward(torch.randn_like(y))
        torch.cuda.synchronize()


if __name__ == '__main__':
    main()
