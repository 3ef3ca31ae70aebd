"""One-shot diagnostic dump (not a test): per-stage max errors against the goldens."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from golden_util import CASES, FORWARD_ONLY_CASES, Golden, subsample
from gpu_util import model_from_golden, dense_grads, table_grads
from oracle import satrans_oracle as O

for name in CASES + FORWARD_ONLY_CASES:
    g = Golden(name)
    print('=====', name, g.meta['flag'], g.meta['meta_mode'], 'H', g.meta['att_head_num'], 'L', g.meta['domain_att_layer_num'], flush=True)
    try:
        model = model_from_golden(g)
        model.engine.precision = os.environ.get('SAT_PRECISION', 'fp32')
        X, y = g.X().cuda().contiguous(), g.y().cuda().contiguous()
        E = model.engine.gather(X)
        print('gather exact:', torch.equal(E.cpu(), g.arr('E0')))
        yp = model.eval()(X)
        torch.cuda.synchronize()
        print('y_pred max err:', (yp.cpu() - g.arr('y_pred')).abs().max().item())
        if name in FORWARD_ONLY_CASES:
            continue
        l2 = g.cfg.l2_reg_embedding
        yp2, w = model.engine.loss_and_grads(X, y, training=False, l2_emb=l2)
        torch.cuda.synchronize()
        print('y_pred(train-path) max err:', (yp2.cpu() - g.arr('y_pred').squeeze(1)).abs().max().item())
        print('bce', float(w['scalars'][0]), float(g.z['bce']), 'reg', float(w['scalars'][1]), float(g.z['reg']), 'nuniq', float(w['scalars'][2]), 'err', float(w['scalars'][3]))
        grads = dense_grads(model, w)
        tg, _ = table_grads(model, w, l2)
        grads.update(tg)
        for k in O.trainable_keys(g.cfg):
            ref = g.arr('grad/' + k)
            got = grads[k][subsample(k)]
            err = (got - ref).abs().max().item()
            sc = ref.abs().max().item()
            flag = '' if err <= 1e-4 * sc + 1e-9 else '   <<<<<< FAIL'
            print('  grad %-55s err %.3e scale %.3e rel %.2e%s' % (k, err, sc, err / (sc + 1e-30), flag))
    except Exception as e:
        import traceback
        traceback.print_exc()
