"""Parameter lists (name, shape) of the benchmark configurations, generated from
the structure of the reference's example models -- not hard-coded size tables.

* ResNet-50: ``examples/chainermn/imagenet/models/resnet50.py:10-108``
* seq2seq:   ``examples/chainermn/seq2seq/seq2seq.py:63-72`` (3 layers, 1024
  units, vocabulary 40,000 as in ``europal.py:61``)
* MNIST MLP: ``examples/chainermn/mnist/train_mnist.py:16-29`` (784-1000-1000-10)

The lists are returned in ``sorted(model.namedparams())`` order, i.e. the
packed-buffer layout (``_memory_utility.py:154-165``).
``tests/golden/layouts.json`` holds the same lists dumped from the real
reference models; ``tests/test_workloads.py`` checks that they agree.
"""


def _bn(prefix, ch):
    return [(prefix + '/gamma', (ch,)), (prefix + '/beta', (ch,))]


def _bottleneck_a(prefix, in_size, ch, out_size):
    p = []
    p.append((prefix + '/conv1/W', (ch, in_size, 1, 1)))
    p += _bn(prefix + '/bn1', ch)
    p.append((prefix + '/conv2/W', (ch, ch, 3, 3)))
    p += _bn(prefix + '/bn2', ch)
    p.append((prefix + '/conv3/W', (out_size, ch, 1, 1)))
    p += _bn(prefix + '/bn3', out_size)
    p.append((prefix + '/conv4/W', (out_size, in_size, 1, 1)))
    p += _bn(prefix + '/bn4', out_size)
    return p


def _bottleneck_b(prefix, in_size, ch):
    p = []
    p.append((prefix + '/conv1/W', (ch, in_size, 1, 1)))
    p += _bn(prefix + '/bn1', ch)
    p.append((prefix + '/conv2/W', (ch, ch, 3, 3)))
    p += _bn(prefix + '/bn2', ch)
    p.append((prefix + '/conv3/W', (in_size, ch, 1, 1)))
    p += _bn(prefix + '/bn3', in_size)
    return p


def _block(prefix, layer, in_size, ch, out_size):
    p = _bottleneck_a(prefix + '/0', in_size, ch, out_size)
    for i in range(1, layer):
        p += _bottleneck_b(prefix + '/{}'.format(i), out_size, ch)
    return p


def resnet50():
    p = [('/conv1/W', (64, 3, 7, 7)), ('/conv1/b', (64,))]
    p += _bn('/bn1', 64)
    p += _block('/res2', 3, 64, 64, 256)
    p += _block('/res3', 4, 256, 128, 512)
    p += _block('/res4', 6, 512, 256, 1024)
    p += _block('/res5', 3, 1024, 512, 2048)
    p += [('/fc/W', (1000, 2048)), ('/fc/b', (1000,))]
    return sorted(p)


def resnet50_bn_layers(batch=32):
    """(name, (N, C, H, W)) of the input of every BatchNormalization layer of
    ResNet-50 for 224x224 images (shapes follow resnet50.py:80-104)."""
    layers = [('/bn1', (batch, 64, 112, 112))]
    hw = 56
    cfg = [('/res2', 3, 64, 256, 1), ('/res3', 4, 128, 512, 2),
           ('/res4', 6, 256, 1024, 2), ('/res5', 3, 512, 2048, 2)]
    for name, n, ch, out, stride in cfg:
        hw_out = hw // stride
        for i in range(n):
            pre = '{}/{}'.format(name, i)
            layers.append((pre + '/bn1', (batch, ch, hw_out, hw_out)))
            layers.append((pre + '/bn2', (batch, ch, hw_out, hw_out)))
            layers.append((pre + '/bn3', (batch, out, hw_out, hw_out)))
            if i == 0:
                layers.append((pre + '/bn4', (batch, out, hw_out, hw_out)))
        hw = hw_out
    return layers


def seq2seq(n_layers=3, n_source_vocab=40000, n_target_vocab=40000, n_units=1024):
    p = [('/embed_x/W', (n_source_vocab, n_units)), ('/embed_y/W', (n_target_vocab, n_units))]
    for rnn in ('/encoder', '/decoder'):
        for layer in range(n_layers):
            for k in range(8):
                p.append(('{}/{}/w{}'.format(rnn, layer, k), (n_units, n_units)))
                p.append(('{}/{}/b{}'.format(rnn, layer, k), (n_units,)))
    p += [('/W/W', (n_target_vocab, n_units)), ('/W/b', (n_target_vocab,))]
    return sorted(p)


def mnist_mlp(n_units=1000, n_in=784, n_out=10):
    p = [('/l1/W', (n_units, n_in)), ('/l1/b', (n_units,)),
         ('/l2/W', (n_units, n_units)), ('/l2/b', (n_units,)),
         ('/l3/W', (n_out, n_units)), ('/l3/b', (n_out,))]
    return sorted(p)


def n_elements(plist):
    total = 0
    for _, shape in plist:
        n = 1
        for s in shape:
            n *= s
        total += n
    return total


def scaled_histogram(target_elems, base=None, min_elems=64):
    """BASELINE config 5 (ii): the ResNet-50 size histogram scaled to about
    `target_elems` elements in total (every size a multiple of 4, >= min_elems)."""
    base = base or resnet50()
    total = n_elements(base)
    out = []
    for name, shape in base:
        n = 1
        for s in shape:
            n *= s
        m = max(min_elems, int(round(n * float(target_elems) / total)) // 4 * 4)
        out.append((name, (m,)))
    return out


WORKLOADS = {
    'resnet50': resnet50,
    'seq2seq': seq2seq,
    'mnist_mlp': mnist_mlp,
}
