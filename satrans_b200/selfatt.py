"""Field self-attention layers of the baselines that share SATrans' attention front end ("next" rows of SURVEY.md section 8f, the
AutoInt / STAR-Trans scoring of BASELINE config 5), run through the same per-sample field-attention kernel code as SATrans
(`sat_field_attention_fwd` / `sat_field_attention_bwd`).  Same constructors and state-dict keys as the reference classes.
All three layers are differentiable (a torch.autograd.Function around the two kernels: gradients w.r.t. the input and every
parameter, dropout masks of the backward = the forward's through the shared counter-based seed); MDR_InteractingLayer's backward
runs the backward kernel once per scenario on that scenario's sample list.

    SelfAttention_Layer    reference submodules.py:178-238 (key `W_Key`) and star_trans.py:335-396 (key `W_key`)
    MDR_InteractingLayer   reference star_trans.py:146-242: one W_Query and one LayerNorm per scenario
    InteractingLayer       deepctr-torch 0.2.9 (AutoInt); the package is not vendored in the reference tree, so this follows its
                           published algorithm -- PARITY UNPINNED (checked against oracle.interacting_layer only)
"""
import torch
import torch.nn as nn

from . import _lib


def _ptr(t):
    return t.detach().contiguous().data_ptr()


def _check_shape(embedding_size, head_num):
    if head_num <= 0:
        raise ValueError('head_num must be a int > 0')
    if embedding_size % head_num != 0:
        raise ValueError('embedding_size is not an integer multiple of head_num!')
    if embedding_size != _lib.SAT_D or head_num not in (1, 2, 4, 8):
        raise NotImplementedError("kernels are built for embedding_size=32 and head_num in {1,2,4,8}")


class _FieldAttention(torch.autograd.Function):
    """y = layer(x) for the single-W_Query layers; backward = sat_field_attention_bwd (recomputes the forward from x)."""

    @staticmethod
    def forward(ctx, x, wq, wk, wv, wres, ln_g, ln_b, heads, scaling, training, p_drop, seed):
        x = x.contiguous().float()
        B, F, _ = x.shape
        out = torch.empty_like(x)
        c = lambda t: None if t is None else t.detach().contiguous()
        wq_, wk_, wv_, wres_, g_, b_ = c(wq), c(wk), c(wv), c(wres), c(ln_g), c(ln_b)
        p = lambda t: None if t is None else t.data_ptr()
        _lib.check(_lib.lib().sat_field_attention_fwd(
            x.data_ptr(), out.data_ptr(), p(wq_), 1, p(wk_), p(wv_), p(wres_), p(g_), p(b_), None, B, F, heads,
            int(wres is not None), int(scaling), int(training), float(p_drop), seed, torch.cuda.current_stream().cuda_stream),
            'sat_field_attention_fwd')
        ctx.save_for_backward(x, wq_, wk_, wv_, wres_, g_)
        ctx.cfg = (heads, scaling, training, p_drop, seed)
        return out

    @staticmethod
    def backward(ctx, dy):
        x, wq, wk, wv, wres, g = ctx.saved_tensors
        heads, scaling, training, p_drop, seed = ctx.cfg
        B, F, _ = x.shape
        dy = dy.contiguous().float()
        dx = torch.empty_like(x)
        dwq, dwk, dwv = torch.empty_like(wq), torch.empty_like(wk), torch.empty_like(wv)
        dwres = torch.empty_like(wres) if wres is not None else None
        dg = torch.empty_like(g) if g is not None else None
        db = torch.empty_like(g) if g is not None else None
        l = _lib.lib()
        scratch = torch.empty(max(1, l.sat_field_attention_bwd_scratch(B, F) // 4), dtype=torch.float32, device=x.device)
        p = lambda t: None if t is None else t.data_ptr()
        _lib.check(l.sat_field_attention_bwd(
            x.data_ptr(), dy.data_ptr(), dx.data_ptr(), p(wq), p(wk), p(wv), p(wres), p(g), None, B, F, heads,
            int(wres is not None), int(scaling), int(training), float(p_drop), seed, p(dwq), p(dwk), p(dwv), p(dwres), p(dg), p(db),
            scratch.data_ptr(), torch.cuda.current_stream().cuda_stream), 'sat_field_attention_bwd')
        return dx, dwq, dwk, dwv, dwres, dg, db, None, None, None, None, None


class _ScenarioAttention(torch.autograd.Function):
    """y = MDR_InteractingLayer(x, dom) with the effective per-scenario W_Query stack wq (n, 32, 32) and LayerNorm stacks (n, 32).
    Forward: ONE launch of sat_field_attention_fwd (the kernel picks every sample's set).  Backward: the samples are grouped by
    scenario (stable sort of the ids) and sat_field_attention_bwd runs once per scenario on that scenario's sample list with its
    own W_Query / LayerNorm -- the reference's masked per-scenario loop (star_trans.py:218-239) under autograd; gradients of the
    shared W_key / W_Value / W_Res are added in scenario order (fixed order: bit-reproducible).  Samples whose id is outside 1..n
    have a zero output and receive a zero input gradient."""

    @staticmethod
    def forward(ctx, x, dom, wq, wk, wv, wres, ln_g, ln_b, heads, scaling, training, p_drop, seed):
        x = x.contiguous().float()
        B, F, _ = x.shape
        n = wq.shape[0]
        out = torch.empty_like(x)
        c = lambda t: None if t is None else t.detach().contiguous().float()
        wq_, wk_, wv_, wres_, g_, b_ = c(wq), c(wk), c(wv), c(wres), c(ln_g), c(ln_b)
        p = lambda t: None if t is None else t.data_ptr()
        _lib.check(_lib.lib().sat_field_attention_fwd(
            x.data_ptr(), out.data_ptr(), p(wq_), n, p(wk_), p(wv_), p(wres_), p(g_), p(b_), dom.data_ptr(), B, F, heads,
            int(wres is not None), int(scaling), int(training), float(p_drop), seed, torch.cuda.current_stream().cuda_stream),
            'sat_field_attention_fwd')
        ctx.save_for_backward(x, dom, wq_, wk_, wv_, wres_, g_)
        ctx.cfg = (heads, scaling, training, p_drop, seed)
        return out

    @staticmethod
    def backward(ctx, dy):
        x, dom, wq, wk, wv, wres, g = ctx.saved_tensors
        heads, scaling, training, p_drop, seed = ctx.cfg
        B, F, _ = x.shape
        n = wq.shape[0]
        dy = dy.contiguous().float()
        dx = torch.zeros_like(x)
        dwq, dg, db = torch.zeros_like(wq), torch.zeros_like(g), torch.zeros_like(g)
        shared = torch.zeros((n, 3 if wres is not None else 2, 32, 32), dtype=torch.float32, device=x.device)
        key = torch.where((dom >= 1) & (dom <= n), dom, torch.full_like(dom, n + 1))
        order = torch.argsort(key, stable=True).to(torch.int32)
        counts = torch.bincount(key.long(), minlength=n + 2)[1:n + 1].tolist()      # (one host sync per layer and backward)
        l = _lib.lib()
        scratch = torch.empty(max(1, l.sat_field_attention_bwd_scratch(max(counts + [1]), F) // 4), dtype=torch.float32,
                              device=x.device)
        s = torch.cuda.current_stream().cuda_stream
        start = 0
        for d, cnt in enumerate(counts):
            if cnt:
                idx = order[start:start + cnt].contiguous()
                _lib.check(l.sat_field_attention_bwd(
                    x.data_ptr(), dy.data_ptr(), dx.data_ptr(), wq[d].data_ptr(), wk.data_ptr(), wv.data_ptr(),
                    wres.data_ptr() if wres is not None else None, g[d].data_ptr(), idx.data_ptr(), cnt, F, heads,
                    int(wres is not None), int(scaling), int(training), float(p_drop), seed, dwq[d].data_ptr(),
                    shared[d, 0].data_ptr(), shared[d, 1].data_ptr(), shared[d, 2].data_ptr() if wres is not None else None,
                    dg[d].data_ptr(), db[d].data_ptr(), scratch.data_ptr(), s), 'sat_field_attention_bwd')
            start += cnt
        tot = shared[0].clone()
        for d in range(1, n):                                                        # scenario order
            tot += shared[d]
        return (dx, None, dwq, tot[0], tot[1], tot[2] if wres is not None else None, dg, db, None, None, None, None, None)


def _run(layer, inputs, wq, n_wq, wk, ln_g, ln_b, dom, scaling, p_drop):
    if len(inputs.shape) != 3:
        raise ValueError("Unexpected inputs dimensions %d, expect to be 3 dimensions" % (len(inputs.shape)))
    if not inputs.is_cuda:
        raise _lib.SatError('%s needs CUDA tensors; there is no CPU fallback' % type(layer).__name__)
    x = inputs.detach().contiguous().float()
    B, F, _ = x.shape
    out = torch.empty_like(x)
    layer._calls += 1
    _lib.check(_lib.lib().sat_field_attention_fwd(
        x.data_ptr(), out.data_ptr(), wq, n_wq, wk, _ptr(layer.W_Value), _ptr(layer.W_Res) if layer.use_res else None,
        ln_g, ln_b, dom, B, F, layer.head_num, int(layer.use_res), int(scaling), int(layer.training), float(p_drop),
        (int(torch.initial_seed()) + layer._calls) & (2 ** 64 - 1), torch.cuda.current_stream().cuda_stream),
        'sat_field_attention_fwd')
    return out


def _differentiable(layer, inputs, wk, ln_g, ln_b, scaling, p_drop):
    if len(inputs.shape) != 3:
        raise ValueError("Unexpected inputs dimensions %d, expect to be 3 dimensions" % (len(inputs.shape)))
    if not inputs.is_cuda:
        raise _lib.SatError('%s needs CUDA tensors; there is no CPU fallback' % type(layer).__name__)
    layer._calls += 1
    seed = (int(torch.initial_seed()) + layer._calls) & (2 ** 64 - 1)
    return _FieldAttention.apply(inputs, layer.W_Query, wk, layer.W_Value, layer.W_Res if layer.use_res else None, ln_g, ln_b,
                                 layer.head_num, bool(scaling), bool(layer.training), float(p_drop), seed)


class SelfAttention_Layer(nn.Module):
    def __init__(self, embedding_size, head_num=2, use_res=True, scaling=True, seed=1024, device='cpu', key_name='W_Key'):
        """key_name: 'W_Key' (models/submodules.py) or 'W_key' (the copy of the class inside models/star_trans.py)."""
        super().__init__()
        _check_shape(embedding_size, head_num)
        if key_name not in ('W_Key', 'W_key'):
            raise ValueError("key_name must be 'W_Key' or 'W_key'")
        self.att_embedding_size = embedding_size // head_num
        self.head_num, self.use_res, self.scaling, self.seed, self._key_name = head_num, use_res, scaling, seed, key_name
        d = embedding_size
        self.W_Query = nn.Parameter(torch.empty(d, d))
        setattr(self, key_name, nn.Parameter(torch.empty(d, d)))
        self.W_Value = nn.Parameter(torch.empty(d, d))
        self.W_Out = nn.Parameter(torch.empty(d, d))          # declared, never used (reference submodules.py:195)
        self.layer_norm = nn.LayerNorm(d, eps=1e-6)
        if use_res:
            self.W_Res = nn.Parameter(torch.empty(d, d))
        for t in self.parameters():                            # reference submodules.py:203-204
            nn.init.normal_(t, mean=0.0, std=0.05)
        self.dropout_rate = 0.1
        self._calls = 0
        self.to(device)

    def forward(self, inputs):
        if torch.is_grad_enabled() and (inputs.requires_grad or self.W_Query.requires_grad):
            return _differentiable(self, inputs, getattr(self, self._key_name), self.layer_norm.weight, self.layer_norm.bias,
                                   self.scaling, self.dropout_rate)
        return _run(self, inputs, _ptr(self.W_Query), 1, _ptr(getattr(self, self._key_name)), _ptr(self.layer_norm.weight),
                    _ptr(self.layer_norm.bias), None, self.scaling, self.dropout_rate)


class InteractingLayer(nn.Module):
    """deepctr-torch 0.2.9 InteractingLayer (AutoInt): relu(softmax(Q K^T [/ sqrt(dh) if scaling]) V + X W_Res)."""

    def __init__(self, embedding_size, head_num=2, use_res=True, scaling=False, seed=1024, device='cpu'):
        super().__init__()
        _check_shape(embedding_size, head_num)
        self.att_embedding_size = embedding_size // head_num
        self.head_num, self.use_res, self.scaling, self.seed = head_num, use_res, scaling, seed
        d = embedding_size
        self.W_Query = nn.Parameter(torch.empty(d, d))
        self.W_key = nn.Parameter(torch.empty(d, d))
        self.W_Value = nn.Parameter(torch.empty(d, d))
        if use_res:
            self.W_Res = nn.Parameter(torch.empty(d, d))
        for t in self.parameters():
            nn.init.normal_(t, mean=0.0, std=0.05)
        self._calls = 0
        self.to(device)

    def forward(self, inputs):
        if torch.is_grad_enabled() and (inputs.requires_grad or self.W_Query.requires_grad):
            return _differentiable(self, inputs, self.W_key, None, None, self.scaling, 0.0)
        return _run(self, inputs, _ptr(self.W_Query), 1, _ptr(self.W_key), None, None, None, self.scaling, 0.0)


class MDR_InteractingLayer(nn.Module):
    MAX_DOMAINS = 8

    def __init__(self, embedding_size, num_domains, head_num=2, merge='no', use_res=True, scaling=True, seed=1024, device='cpu'):
        super().__init__()
        _check_shape(embedding_size, head_num)
        if merge not in ('no', 'sum', 'dot'):
            raise NotImplementedError('not implemented parameters merge')
        if not 1 <= num_domains <= self.MAX_DOMAINS:
            raise NotImplementedError("1 <= num_domains <= %d" % self.MAX_DOMAINS)
        self.att_embedding_size = embedding_size // head_num
        self.head_num, self.use_res, self.scaling, self.seed = head_num, use_res, scaling, seed
        self.merge, self.embedding_size, self.num_domains = merge, embedding_size, num_domains
        d = embedding_size
        self.W_Query = nn.Parameter(torch.empty(d, d))
        self.domain_W_Queries = nn.ParameterList([nn.Parameter(torch.empty(d, d)) for _ in range(num_domains)])
        self.W_key = nn.Parameter(torch.empty(d, d))
        self.W_Value = nn.Parameter(torch.empty(d, d))
        self.domain_layer_norms = nn.ModuleList([nn.LayerNorm(d, eps=1e-6) for _ in range(num_domains)])
        if use_res:
            self.W_Res = nn.Parameter(torch.empty(d, d))
        for t in self.parameters():                            # reference star_trans.py:192-193
            nn.init.normal_(t, mean=0.0, std=0.05)
        self.dropout_rate = 0.1
        self._calls = 0
        self.to(device)

    def forward(self, inputs, domain_ids):
        if torch.is_grad_enabled() and (inputs.requires_grad or self.W_key.requires_grad):
            return self._forward_differentiable(inputs, domain_ids)
        with torch.no_grad():
            wq = torch.stack([w.detach() for w in self.domain_W_Queries])              # (n, 32, 32)
            if self.merge == 'sum':                                                     # reference star_trans.py:209-217
                wq = wq + self.W_Query.detach()
            elif self.merge == 'dot':
                wq = wq * self.W_Query.detach()
            wq = wq.contiguous().float()
            ln_g = torch.stack([ln.weight.detach() for ln in self.domain_layer_norms]).contiguous().float()
            ln_b = torch.stack([ln.bias.detach() for ln in self.domain_layer_norms]).contiguous().float()
            dom = domain_ids.detach().to(device=inputs.device, dtype=torch.int32).contiguous()
        if dom.numel() != inputs.shape[0]:
            raise ValueError('domain_ids must hold one id per sample')
        return _run(self, inputs, wq.data_ptr(), self.num_domains, _ptr(self.W_key), ln_g.data_ptr(), ln_b.data_ptr(),
                    dom.data_ptr(), self.scaling, self.dropout_rate)

    def _forward_differentiable(self, inputs, domain_ids):
        if len(inputs.shape) != 3:
            raise ValueError("Unexpected inputs dimensions %d, expect to be 3 dimensions" % (len(inputs.shape)))
        if not inputs.is_cuda:
            raise _lib.SatError('MDR_InteractingLayer needs CUDA tensors; there is no CPU fallback')
        dom = domain_ids.detach().to(device=inputs.device, dtype=torch.int32).contiguous()
        if dom.numel() != inputs.shape[0]:
            raise ValueError('domain_ids must hold one id per sample')
        wq = torch.stack(list(self.domain_W_Queries))              # torch autograd carries the merge and the stacking
        if self.merge == 'sum':                                    # reference star_trans.py:209-217
            wq = wq + self.W_Query
        elif self.merge == 'dot':
            wq = wq * self.W_Query
        ln_g = torch.stack([ln.weight for ln in self.domain_layer_norms])
        ln_b = torch.stack([ln.bias for ln in self.domain_layer_norms])
        self._calls += 1
        seed = (int(torch.initial_seed()) + self._calls) & (2 ** 64 - 1)
        return _ScenarioAttention.apply(inputs, dom, wq, self.W_key, self.W_Value, self.W_Res if self.use_res else None, ln_g, ln_b,
                                        self.head_num, bool(self.scaling), bool(self.training), float(self.dropout_rate), seed)
