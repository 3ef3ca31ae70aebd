"""Diagnostic: tensor-core forward vs goldens (prints errors; not a test)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from golden_util import CASES, FORWARD_ONLY_CASES, Golden
from gpu_util import model_from_golden
for name in CASES + FORWARD_ONLY_CASES:
    g = Golden(name)
    model = model_from_golden(g).eval()
    X = g.X().cuda().contiguous()
    for prec in ('fp32', 'tf32x3', 'tf32'):
        model.engine.precision = prec
        try:
            yp = model(X)
            torch.cuda.synchronize()
            e = (yp.cpu() - g.arr('y_pred')).abs()
            print(name, prec, 'max err %.3e mean %.3e nan %d' % (e.max().item(), e.mean().item(), int(torch.isnan(yp).sum())), flush=True)
        except Exception as ex:
            print(name, prec, 'EXC', str(ex)[:300], flush=True)
            raise
