"""Per-element error budget for one step of a NORMALISING optimiser (torch.optim.Adagrad / RMSprop as the reference's compile()
string table builds them, meta_basemodel.py:612-625), derived from the stated GRADIENT tolerance by exact interval arithmetic.

Why: u = lr * g / (sqrt(S) + eps) with S accumulated from g^2 is +-lr/sqrt(c) whatever |g| is, so an element whose gradient is
within round-off of zero moves by the full step in a direction decided by the last bit of the gradient.  No two fp32
implementations (the reference on two devices included) agree on those elements, while every well-conditioned element must still
agree tightly.  The bound below says exactly that, element by element:

    g_hat in [g - delta, g + delta]   (delta = the gradient tolerance of the path under test)
    S_hat in [S_lo, S_hi]             (interval carried over from the earlier steps, same rule)
    u(g, S_prev) = lr * g / (sqrt(alpha * S_prev + c * g^2) + eps)  is increasing in g and decreasing in S_prev,

so the largest |u_hat - u| is attained at one of the four corners.  Nothing is linearised and no statistical slack is used."""
import torch

HPARAMS = {  # torch defaults = what the reference's string table instantiates
    'adagrad': dict(alpha=1.0, c=1.0, eps=1e-10, state='sum'),
    'rmsprop': dict(alpha=0.99, c=0.01, eps=1e-8, state='square_avg'),
}


def _u(lr, g, s_prev, alpha, c, eps):
    return lr * g / (torch.sqrt(alpha * s_prev + c * g * g) + eps)


def step_bound(opt, lr, g, s_prev, s_lo, s_hi, delta):
    """g, s_prev: the exact (fp64) gradient and accumulator BEFORE this step; [s_lo, s_hi]: interval of the implementation's
    accumulator before this step; delta: per-element gradient tolerance.  Returns (bound on |u_hat - u|, new s_lo, new s_hi)."""
    h = HPARAMS[opt]
    a, c, eps = h['alpha'], h['c'], h['eps']
    u = _u(lr, g, s_prev, a, c, eps)
    corners = [_u(lr, g + sg * delta, s, a, c, eps) for sg in (-1.0, 1.0) for s in (s_lo, s_hi)]
    bound = torch.stack([(cu - u).abs() for cu in corners]).max(dim=0).values
    g_lo = (g.abs() - delta).clamp_min(0.0)
    g_hi = g.abs() + delta
    return bound, a * s_lo + c * g_lo * g_lo, a * s_hi + c * g_hi * g_hi


def grad_tolerance(g, gtol, exact_mask=None):
    """Per-element gradient tolerance: gtol x the tensor's max-norm (how the gradient bar of the parity tests is stated);
    elements whose gradient is a single product (table rows the batch never reads: g = 2 * l2 * w) carry relative round-off only."""
    delta = torch.full_like(g, gtol * g.abs().max().item())
    if exact_mask is not None:
        delta = torch.where(exact_mask, 8 * 1.2e-7 * g.abs(), delta)
    return delta


def check_normalising_steps(O, g, opt, loss, lr, impl_step, impl_state, gtol, steps=2, subsample=None, report=None):
    """Drive `steps` training steps of the implementation under test (impl_step(X, y) -> loss; impl_state() -> {key: fp32 tensor})
    beside the fp64 oracle and require EVERY parameter element to stay inside `step_bound` (+ fp32 round-off of the update
    arithmetic).  After each step the oracle is re-based on the implementation's own parameters, so a near-zero element decided the
    other way at step t (inside its budget) cannot leak into the gradients of step t+1 and be charged to other elements."""
    h = HPARAMS[opt]
    tr = O.Trainer({k: t.double() for k, t in g.state().items()}, g.cfg, lr=lr, optimizer=opt, loss=loss)
    keys = O.trainable_keys(g.cfg)
    s_lo = {k: torch.zeros_like(tr.params[k]) for k in keys}
    s_hi = {k: torch.zeros_like(tr.params[k]) for k in keys}
    losses = []
    for step in range(steps):
        X, y = g.X(step), g.y(step)
        s_prev = {k: (tr.optim.state[tr.params[k]][h['state']].clone() if tr.params[k] in tr.optim.state else torch.zeros_like(tr.params[k]))
                  for k in keys}
        losses.append(impl_step(X, y))
        tr.step(X.double(), y.double(), train_mode=False)
        got = impl_state()
        for k in keys:
            p = tr.params[k]
            exact = None
            if k.startswith('embedding_dict.'):
                col = g.cfg.sparse_names.index(k.split('.')[1])
                touched = torch.zeros(p.shape[0], dtype=torch.bool)
                touched[X[:, col].long()] = True
                exact = (~touched)[:, None].expand_as(p)
            delta = grad_tolerance(p.grad, gtol, exact)
            bound, s_lo[k], s_hi[k] = step_bound(opt, lr, p.grad, s_prev[k], s_lo[k], s_hi[k], delta)
            err = (got[k].double() - p.detach()).abs()
            scale = max(p.detach().abs().max().item(), 1e-3)
            slack = 2e-6 * scale + 1e-7 + 4 * 1.2e-7 * lr / h['c'] ** 0.5       # fp32 storage of p and of the step
            excess = err - bound - slack
            if report is not None:
                report.append((step, k, err.max().item(), bound.max().item(), (bound > slack).float().mean().item()))
            assert excess.max().item() <= 0.0, (step, k, err.flatten()[excess.argmax()].item(), bound.flatten()[excess.argmax()].item(), slack)
        with torch.no_grad():                     # re-base the oracle on the implementation's parameters
            for k in keys:
                tr.params[k].copy_(got[k].double())
    return losses
