import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import chainer_b200
from chainer_b200.core.link import link_from_named_arrays
from oracle import gradpath as og
PLIST = sorted([('/a/W', (3, 2)), ('/a/b', (3,)), ('/d/W', (257, 129))])
rng = np.random.default_rng(1)
host = [np.asarray(rng.standard_normal(s) * 0.05).astype(np.float64).reshape(s) for _, s in PLIST]
model = link_from_named_arrays([(n, torch.from_numpy(a).cuda()) for (n, _), a in zip(PLIST, host)])
comm = chainer_b200.create_communicator('pure_nccl')
actual = chainer_b200.MomentumSGD(lr=0.01, momentum=0.9)
opt = chainer_b200.create_multi_node_optimizer(actual, comm); opt.setup(model); opt.update()
vs=[np.zeros_like(a) for a in host]
for step in range(1,3):
    grads = [np.asarray(rng.standard_normal(a.shape) * 1e-2).astype(np.float64).reshape(a.shape) for a in host]
    for (_, p), g in zip(sorted(model.namedparams()), grads): p.grad = torch.from_numpy(g).cuda()
    opt.update(); torch.cuda.synchronize()
    from tests.fake_lib import _view
    buf = torch.empty(0)
    for (name, p), q, g, v in zip(sorted(model.namedparams()), host, grads, vs):
        g32 = g.astype(np.float32).astype(np.float64)
        og.momentum_sgd_update(q, g32, v, 0.01, 0.9)
        got = p.data.cpu().numpy(); gv = p.update_rule.state['v'].cpu().numpy(); gg = p.grad.cpu().numpy()
        print(step, name, 'param maxdiff', np.abs(got-q).max(), 'v maxdiff', np.abs(gv-v).max(), 'grad maxdiff', np.abs(gg-g32).max(),
              'v vs -(0.01*g32)', np.abs(gv+0.01*g32).max() if step==1 else '', 'v vs f32 lr', np.abs(gv+float(np.float32(0.01))*g32).max() if step==1 else '')
