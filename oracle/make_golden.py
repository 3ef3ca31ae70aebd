"""Generate tests/golden/*.npz from the UNMODIFIED reference — TEST INFRASTRUCTURE.

Run in the build container only (needs /root/reference):

    python oracle/make_golden.py            # writes tests/golden/{en_small,enp_small,q_small,relu_small_fwd,
                                            #   selfatt_layer,compile_variants}.npz

The reference modules `models/{satrans,meta_basemodel,submodules}.py` are imported as they lie
under /root/reference, with `oracle/ref_shims` supplying the third-party names they import
(SURVEY.md §8c / Appendix C).  For each case we dump, in fp32 exactly as the reference produced them:

  state/<key>        the reference `state_dict()` at construction (canonical keys; aliased keys dropped)
  X, y               the (B, n_cols) float32 input matrix (ids stored as floats, as `fit` feeds it) and labels
  E0                 gathered field embeddings (B, F, d)          — bit-exact target
  y_pred             eval-mode forward (B, 1)
  bce, reg           BCE(sum) and L2 regulariser of that batch
  grad/<key>         every parameter gradient of (bce + reg), eval-mode forward (dropout off)
  stepK/<key>, stepK_m/<key>, stepK_v/<key>   K in {1,3}: parameters and Adam moments after K dense
                     torch.optim.Adam steps on batches X[k] (eval-mode forward, dropout off)
  predict            `model.predict(x_dict, batch_size)` through the reference's own API (N,1) float64

While dumping, the script also asserts that `oracle/satrans_oracle.py` reproduces the reference
(fp32 vs fp32) — that is the pin.
"""
import json
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get('SATRANS_REFERENCE', '/root/reference')


def import_reference():
    sys.path.insert(0, os.path.join(HERE, 'ref_shims'))
    sys.path.insert(0, REF)
    from models.satrans import SATrans  # noqa
    from deepctr_torch.inputs import SparseFeat, DenseFeat  # noqa
    return SATrans, SparseFeat, DenseFeat


CASES = {
    # name: (schema, vocab_scale, B, ctor overrides)
    'en_small': ('aliccp', 1e-4, 96, dict(domain_att_layer_num=3, att_head_num=4, meta_mode='QK', flag='sota')),
    'enp_small': ('alimama', 1e-4, 80, dict(domain_att_layer_num=3, att_head_num=4, meta_mode='QK',
                                            flag='sota-pos')),
    'q_small': ('aliccp', 5e-5, 40, dict(domain_att_layer_num=2, att_head_num=2, meta_mode='Q', flag='sota')),
    # 'relu' in flag: the reference's own eval-mode backward raises (in-place `result += inputs` on the ReLU
    # output, satrans.py:91-97, when dropout is the identity), so this case pins the forward only.
    'relu_small_fwd': ('aliccp', 5e-5, 40, dict(domain_att_layer_num=2, att_head_num=4, meta_mode='QK',
                                                flag='sota-relu')),
    # the ablation variants of the scenario-specific projection (satrans.py:59-81,158-176): a generated gate / per-head bilinear
    # map instead of the MetaNet, a wide scenario embedding instead of the scenario encoder; with and without 'pos'
    'gate_small': ('aliccp', 5e-5, 40, dict(domain_att_layer_num=2, att_head_num=4, meta_mode='QK', flag='sota-gate')),
    'gatepos_small': ('alimama', 5e-5, 40, dict(domain_att_layer_num=2, att_head_num=2, meta_mode='K', flag='sota-gate-pos')),
    'bilinear_small': ('aliccp', 5e-5, 40, dict(domain_att_layer_num=2, att_head_num=4, meta_mode='QK', flag='sota-bilinear')),
    'bilinearpos_small': ('aliccp', 5e-5, 40, dict(domain_att_layer_num=2, att_head_num=2, meta_mode='Q',
                                                   flag='sota-bilinear-pos')),
    # (B = 48: at B = 40 one element of the shared Q/K MetaNet LayerNorm bias has |g| = 2e-8 ~ Adam's eps, 5e-5 of the tensor's
    #  max-norm -- its step direction is decided by round-off in any fp32 implementation, a useless fixture for Adam parity)
    'onlyemb_small': ('aliccp', 5e-5, 48, dict(domain_att_layer_num=2, att_head_num=4, meta_mode='QK', flag='sota-onlyemb')),
    'onlyembpos_small': ('alimama', 5e-5, 40, dict(domain_att_layer_num=2, att_head_num=4, meta_mode='QK',
                                                   flag='sota-onlyemb-pos')),
}
FORWARD_ONLY = {'relu_small_fwd'}
VARIANT_CASES = ('gate_small', 'gatepos_small', 'bilinear_small', 'bilinearpos_small', 'onlyemb_small', 'onlyembpos_small')


def subsample(key):
    """The scenario-encoder weight (4096 x 32|64) dominates the fixture size: post-step snapshots keep
    and gradients keep every 8th row of it (the initial state is stored in full)."""
    return slice(None, None, 8) if key == 'domain_map_dnn_Q.linears.0.weight' else slice(None)


def zero_grad_key(key):
    """K-side MetaNet LayerNorm bias shifts every key by a constant => softmax-invariant => its true
    gradient is 0; only present as its own tensor with 'pos'."""
    return key.endswith('K_meta_mlp.ffn_layer_norm.bias')


def build_case(name):
    sys.path.insert(0, ROOT)
    from satrans_b200 import synth
    from oracle import satrans_oracle as O
    SATrans, SparseFeat, DenseFeat = import_reference()
    schema_name, scale, B, kw = CASES[name]
    schema = (synth.aliccp_schema if schema_name == 'aliccp' else synth.alimama_schema)(scale)
    cols = ([SparseFeat(f, vocabulary_size=schema.data_max[f] + 2, embedding_dim=32) for f in schema.fields]
            + [DenseFeat(f, 1) for f in schema.dense])
    torch.set_num_threads(1)
    model = SATrans(linear_feature_columns=cols, dnn_feature_columns=cols,
                    domain_column_list=[schema.domain_col], num_domains_list=[schema.num_domains],
                    att_layer_num=0, use_linear=False, use_dnn=False, seed='1021', device='cpu', **kw)
    # give the MetaNet LayerNorms and the encoder non-trivial values so that every gradient path is exercised
    g = torch.Generator().manual_seed(7)
    with torch.no_grad():
        for k, v in model.state_dict().items():
            if 'ffn_layer_norm' in k:
                v.add_(torch.randn(v.shape, generator=g) * 0.1)
            if k == 'domain_map_dnn_Q.linears.0.weight':
                v.copy_(torch.randn(v.shape, generator=g) * 0.05)
            if k.startswith('embedding_dict.'):
                v.copy_(torch.randn(v.shape, generator=g) * 0.3)
    state0 = {k: v.detach().clone() for k, v in model.state_dict().items()}

    names = schema.fields + schema.dense
    xs, ys = synth.make_inputs(schema, 3 * B, seed=1021, ctr=0.3)
    Xall = synth.stack_columns(xs, names)
    yall = ys.astype(np.float32)
    X = torch.from_numpy(Xall[:B])
    y = torch.from_numpy(yall[:B])

    cfg = O.OracleConfig(sparse_names=schema.fields, vocab_sizes=schema.vocab_sizes, dense_names=schema.dense,
                         domain_col=schema.domain_col, num_domains=schema.num_domains, embedding_dim=32,
                         layers=kw['domain_att_layer_num'], heads=kw['att_head_num'], meta_mode=kw['meta_mode'],
                         flag=kw['flag'])
    out = {'config': np.frombuffer(json.dumps(dict(
        schema=schema_name, vocab_scale=scale, B=B, sparse_names=schema.fields, vocab_sizes=schema.vocab_sizes,
        dense_names=schema.dense, domain_col=schema.domain_col, num_domains=schema.num_domains, **kw)).encode(),
        dtype=np.uint8)}
    for k, v in state0.items():
        if O.canonical_key(cfg, k) == k:      # aliases (K/V MetaNet LN, encoder _K/_V) are re-created by the tests
            out['state/' + k] = v.numpy()
    out['X'] = Xall
    out['y'] = yall

    # ---- eval-mode forward / gather --------------------------------------------------------------
    model.eval()
    emb_list, _ = model.input_from_feature_columns(X, model.dnn_feature_columns, model.embedding_dict)
    E0 = torch.cat(emb_list, dim=1).detach()
    out['E0'] = E0.numpy()
    yp = model(X)
    out['y_pred'] = yp.detach().numpy()
    if name in FORWARD_ONLY:
        st = {k: v.clone() for k, v in state0.items()}
        with torch.no_grad():
            assert torch.allclose(O.forward(st, X, cfg), yp.detach(), atol=1e-6, rtol=0)
        import contextlib
        import io
        with contextlib.redirect_stderr(io.StringIO()):
            out['predict'] = model.predict({k: v[:B] for k, v in xs.items()}, batch_size=37)
        return out
    # ---- loss + grads (eval mode ⇒ dropout off), exactly the ops of meta_basemodel.py:314-327 -----
    model.zero_grad()
    bce = torch.nn.functional.binary_cross_entropy(yp.squeeze(), y, reduction='sum')
    reg = model.get_regularization_loss()
    (bce + reg).backward()
    out['bce'] = np.float32(bce.item())
    out['reg'] = np.float32(reg.item())
    named = dict(model.named_parameters())
    for k in O.trainable_keys(cfg):
        gr = named[k].grad
        out['grad/' + k] = (gr if gr is not None else torch.zeros_like(named[k])).detach().numpy()[subsample(k)].copy()

    # ---- pin the oracle against the reference ---------------------------------------------------
    st = {k: v.clone() for k, v in state0.items()}
    o_ids, o_E0 = O.gather_fields(st, X, cfg)
    assert torch.equal(o_E0, E0), 'oracle gather != reference'
    o_yp, o_bce, o_reg, o_g = O.loss_and_grads(st, X, y, cfg)
    assert torch.allclose(o_yp, yp.detach(), atol=1e-6, rtol=0), (o_yp - yp).abs().max()
    assert abs(o_bce.item() - bce.item()) <= 1e-5 * abs(bce.item())
    assert abs(o_reg.item() - reg.item()) <= 1e-5 * abs(reg.item())
    for k in O.trainable_keys(cfg):
        a, b = o_g[k][subsample(k)], torch.from_numpy(out['grad/' + k])
        scale_ = b.abs().max().item()
        assert (a - b).abs().max().item() <= 1e-4 * scale_ + 1e-9, (k, (a - b).abs().max().item(), scale_)

    # ---- dense Adam steps (eval-mode forward) ---------------------------------------------------
    lr = 0.005 if schema_name == 'aliccp' else 0.001
    optim = torch.optim.Adam(model.parameters(), lr=lr)
    o_opt = O.adam_init(st, cfg)
    for step in range(1, 4):
        Xk = torch.from_numpy(Xall[(step - 1) * B: step * B])
        yk = torch.from_numpy(yall[(step - 1) * B: step * B])
        ypk = model(Xk).squeeze()
        optim.zero_grad()
        total = torch.nn.functional.binary_cross_entropy(ypk, yk, reduction='sum') + model.get_regularization_loss()
        total.backward()
        optim.step()
        O.train_step(st, Xk, yk, cfg, o_opt, lr)
        if step in (1, 3):
            sd = model.state_dict()
            for k in O.trainable_keys(cfg):
                sub = subsample(k)
                out['step%d/%s' % (step, k)] = sd[k].detach().numpy()[sub].copy()
                s = optim.state[named[k]]
                if step == 3:
                    out['step%d_m/%s' % (step, k)] = s['exp_avg'].numpy()[sub].copy()
                    out['step%d_v/%s' % (step, k)] = s['exp_avg_sq'].numpy()[sub].copy()
                # pin: the oracle's Adam restatement tracks the reference within fp32 round-off
                ref_p = sd[k].detach()
                err = (st[k] - ref_p).abs().max().item()
                tol = 2e-5 * max(ref_p.abs().max().item(), 1e-3) + 1e-7
                if zero_grad_key(k):
                    tol = 3 * lr * step     # analytically-zero gradient: Adam amplifies round-off noise
                assert err <= tol, (step, k, err)
    out['lr'] = np.float32(lr)

    # ---- predict() through the reference API (restore the step-0 weights first) ------------------
    model.load_state_dict(state0)
    import contextlib
    import io
    with contextlib.redirect_stderr(io.StringIO()):
        pred = model.predict({k: v[:B] for k, v in xs.items()}, batch_size=37)
    out['predict'] = pred
    return out


def build_selfatt():
    """SelfAttention_Layer of the unmodified reference (submodules.py:178-238): eval-mode forward goldens and the gradients of
    sum(y * cot) w.r.t. the input and every parameter (torch autograd through the reference module).  Seeded: the module's
    constructor draws from torch's global stream."""
    sys.path.insert(0, ROOT)
    from oracle import satrans_oracle as O
    import_reference()
    from models.submodules import SelfAttention_Layer
    out = {}
    g = torch.Generator().manual_seed(3)
    for idx, (heads, Fn, res) in enumerate(((4, 19, True), (2, 15, True), (4, 7, False))):
        torch.manual_seed(1021 + idx)
        layer = SelfAttention_Layer(32, heads, use_res=res).eval()
        with torch.no_grad():
            layer.layer_norm.weight.add_(1.0)          # N(0, .05) gamma would squash the output to ~0
        x = torch.randn((24, Fn, 32), generator=g) * 0.5
        with torch.no_grad():
            y = layer(x)
        tag = 'h%d_f%d_r%d' % (heads, Fn, int(res))
        out[tag + '/x'] = x.numpy()
        out[tag + '/y'] = y.numpy()
        sd = {k: v.detach().clone() for k, v in layer.state_dict().items()}
        for k, v in sd.items():
            out[tag + '/state/' + k] = v.numpy()
        with torch.no_grad():
            o = O.self_attention_layer(x, sd, heads, res)
        assert torch.allclose(o, y, atol=1e-6, rtol=0), (o - y).abs().max()
        # backward goldens
        cot = torch.randn((24, Fn, 32), generator=g)
        xg = x.clone().requires_grad_(True)
        layer.zero_grad()
        (layer(xg) * cot).sum().backward()
        out[tag + '/cot'] = cot.numpy()
        out[tag + '/grad/x'] = xg.grad.numpy()
        for k, p_ in layer.named_parameters():
            if p_.grad is not None:
                out[tag + '/grad/' + k] = p_.grad.detach().numpy()
    return out


def build_star_trans():
    """MDR_InteractingLayer (star_trans.py:146-242) and the Star_Trans model (star_trans.py:674-775) of the unmodified reference:
    eval-mode forward goldens (+ gradients of sum(y * cot) for the layer), the oracle restatement asserted against them."""
    sys.path.insert(0, ROOT)
    from oracle import satrans_oracle as O
    import_reference()
    from models.star_trans import MDR_InteractingLayer, Star_Trans
    from deepctr_torch.inputs import SparseFeat
    from satrans_b200 import synth
    out = {}
    g = torch.Generator().manual_seed(11)
    for idx, (heads, Fn, nd, merge, res) in enumerate(((4, 18, 3, 'no', True), (2, 14, 4, 'sum', True), (4, 6, 3, 'dot', False))):
        torch.manual_seed(2021 + idx)
        layer = MDR_InteractingLayer(32, nd, heads, merge, res).eval()
        with torch.no_grad():
            for ln in layer.domain_layer_norms:
                ln.weight.add_(1.0)                     # N(0, .05) gamma would squash the output to ~0
        B = 40
        x = torch.randn((B, Fn, 32), generator=g) * 0.5
        dom = torch.randint(0, nd + 1, (B,), generator=g)          # 0 = a scenario outside 1..nd: the layer leaves zeros
        with torch.no_grad():
            y = layer(x, dom)
        tag = 'mdr_h%d_f%d_d%d_%s_r%d' % (heads, Fn, nd, merge, int(res))
        out[tag + '/x'], out[tag + '/dom'], out[tag + '/y'] = x.numpy(), dom.numpy(), y.numpy()
        sd = {k: v.detach().clone() for k, v in layer.state_dict().items()}
        for k, v in sd.items():
            out[tag + '/state/' + k] = v.numpy()
        with torch.no_grad():
            o = O.mdr_interacting_layer(x, dom, sd, heads, nd, merge, res)
        assert torch.allclose(o, y, atol=1e-6, rtol=0), (o - y).abs().max()
        cot = torch.randn((B, Fn, 32), generator=g)
        xg = x.clone().requires_grad_(True)
        layer.zero_grad()
        (layer(xg, dom) * cot).sum().backward()
        out[tag + '/cot'], out[tag + '/grad/x'] = cot.numpy(), xg.grad.numpy()
        for k, p_ in layer.named_parameters():
            if p_.grad is not None:
                out[tag + '/grad/' + k] = p_.grad.detach().numpy()
    # the whole Star_Trans model on an Ali-CCP-shaped schema (main.py's configuration: scenario column filtered out of the fields)
    schema = synth.aliccp_schema(1e-4)
    cols = [SparseFeat(f, vocabulary_size=schema.data_max[f] + 2, embedding_dim=32) for f in schema.fields]
    torch.manual_seed(2030)
    model = Star_Trans(cols, cols, schema.domain_col, schema.num_domains, att_layer_num=2, domain_att_layer_num=1, att_head_num=4,
                       merge='no', use_linear=False, seed=1021, device='cpu').eval()
    with torch.no_grad():
        for k, v in model.state_dict().items():
            if k.startswith('embedding_dict.'):
                v.copy_(torch.randn(v.shape, generator=g) * 0.3)
            if 'layer_norm' in k and k.endswith('weight'):
                v.add_(1.0)
    xs, _ = synth.make_inputs(schema, 64, seed=1021)
    X = torch.from_numpy(synth.stack_columns(xs, schema.fields))
    with torch.no_grad():
        yp = model(X)
    st = {k: v.detach().clone() for k, v in model.state_dict().items()}
    o = O.star_trans_forward(st, X, schema.fields, schema.domain_col, schema.num_domains, 4, 2, 1)
    assert torch.allclose(o, yp, atol=1e-6, rtol=0), (o - yp).abs().max()
    out['model/X'], out['model/y_pred'] = X.numpy(), yp.numpy()
    out['model/fields'] = np.frombuffer(json.dumps(dict(fields=schema.fields, vocab=schema.vocab_sizes, domain_col=schema.domain_col,
                                                        num_domains=schema.num_domains)).encode(), dtype=np.uint8)
    for k, v in st.items():
        if not k.startswith('linear_model.'):
            out['model/state/' + k] = v.numpy()
    # training goldens of the same model (reference basemodel.py:309-328: BCE(sum) + L2(embeddings), dense Adam), dropout off
    # (eval mode with gradients) so that they do not depend on torch's mask stream; drawn AFTER everything above, so the forward
    # goldens keep their bits
    yl = (torch.rand(64, generator=g) < 0.3).float()
    out['model/y'] = yl.numpy()
    state0 = {k: v.detach().clone() for k, v in model.state_dict().items()}

    def ref_step(first):
        model.optim.zero_grad()
        loss = model.loss_func(model(X).squeeze(), yl, reduction='sum')
        reg = model.get_regularization_loss()
        (loss + reg).sum().backward()
        if first:
            out['model/loss'], out['model/reg'] = loss.detach().numpy(), reg.detach().numpy()
            for k, p_ in model.named_parameters():
                if p_.grad is not None and not k.startswith('linear_model.'):
                    out['model/grad/' + k] = p_.grad.detach().clone().numpy()
            # the restatement under torch autograd gives the same gradients
            so = {k: v.detach().clone().requires_grad_(True) for k, v in st.items() if v.dtype.is_floating_point}
            lo = F.binary_cross_entropy(O.star_trans_forward(so, X, schema.fields, schema.domain_col, schema.num_domains, 4, 2, 1).squeeze(),
                                        yl, reduction='sum')
            lo = lo + sum(1e-5 * (v * v).sum() for k, v in so.items() if k.startswith('embedding_dict.'))
            lo.backward()
            for k, p_ in model.named_parameters():
                if p_.grad is not None and not k.startswith('linear_model.'):
                    gmax = float(p_.grad.abs().max())
                    assert float((so[k].grad - p_.grad).abs().max()) <= 1e-5 * gmax + 1e-12, k
        model.optim.step()
        return float(loss)

    # (a) two plain-SGD steps (the reference's 'sgd' string: lr 0.01): the update is linear in the gradient, so the end state
    #     can be held to the gradient tolerance element by element
    model.compile('sgd', 'binary_crossentropy', metrics=[])
    ref_step(True)
    ref_step(False)
    for k, v in model.state_dict().items():
        if not k.startswith('linear_model.'):
            out['model/sgd_after2/' + k] = v.detach().clone().numpy()
    # (b) three Adam steps (main.py:343; lr 0.01) from the same start: the loss trajectory
    model.load_state_dict(state0)
    model.compile(torch.optim.Adam(model.parameters(), lr=0.01), 'binary_crossentropy', metrics=[])
    out['model/adam_losses'] = np.array([ref_step(False) for _ in range(3)], dtype=np.float64)
    # a second configuration with the optional logit terms: the linear model (use_linear, over EVERY column), one head per
    # scenario (use_domain_dnn_linear), the scenario column kept as a field, merge='dot': forward, loss and every gradient
    torch.manual_seed(2031)
    model2 = Star_Trans(cols, cols, schema.domain_col, schema.num_domains, domain_id_as_feature=True, att_layer_num=1,
                        domain_att_layer_num=2, att_head_num=2, merge='dot', use_linear=True, use_domain_dnn_linear=True,
                        seed=1021, device='cpu').eval()
    with torch.no_grad():
        for k, v in model2.state_dict().items():
            if k.startswith('embedding_dict.'):
                v.copy_(torch.randn(v.shape, generator=g) * 0.3)
            if k.startswith('linear_model.'):
                v.copy_(torch.randn(v.shape, generator=g) * 0.2)
            if 'layer_norm' in k and k.endswith('weight'):
                v.add_(1.0)
    st2 = {k: v.detach().clone() for k, v in model2.state_dict().items()}
    model2.zero_grad()
    yp2 = model2(X)
    loss2 = F.binary_cross_entropy(yp2.squeeze(), yl, reduction='sum')
    (loss2 + model2.get_regularization_loss()).sum().backward()
    out['model2/y_pred'], out['model2/loss'] = yp2.detach().numpy(), np.float32(loss2.item())
    for k, v in st2.items():
        out['model2/state/' + k] = v.numpy()
    so = {k: v.detach().clone().requires_grad_(True) for k, v in st2.items()}
    o2 = O.star_trans_forward(so, X, schema.fields, schema.domain_col, schema.num_domains, 2, 1, 2, merge='dot', use_linear=True,
                              use_domain_dnn_linear=True, domain_id_as_feature=True)
    assert torch.allclose(o2.detach(), yp2.detach(), atol=1e-6, rtol=0), (o2 - yp2).abs().max()
    lo = F.binary_cross_entropy(o2.squeeze(), yl, reduction='sum') + sum(1e-5 * (v * v).sum() for k, v in so.items()
                                                                          if k.startswith('embedding_dict.'))
    lo.backward()
    for k, p_ in model2.named_parameters():
        if p_.grad is not None:
            out['model2/grad/' + k] = p_.grad.detach().clone().numpy()
            gmax = float(p_.grad.abs().max())
            assert float((so[k].grad - p_.grad).abs().max()) <= 1e-5 * gmax + 1e-12, k
    return out


def build_meta_module():
    """BaseModel.meta_transformation (models/basemodel.py:191-199) + MetaNet with units (32, 64, 32) of the unmodified reference --
    the scenario-specific pre-transformation of the single-task baselines ('usemetatrans' in flag, e.g. autoint.py:86-89), the
    part of that path that IS vendored: output and gradients of sum(out * cot) for 'metanorm' on / off, eval mode."""
    sys.path.insert(0, ROOT)
    from oracle import satrans_oracle as O
    import_reference()
    from models.basemodel import BaseModel
    from deepctr_torch.inputs import SparseFeat
    out = {}
    g = torch.Generator().manual_seed(23)
    cols = [SparseFeat('a', 10, embedding_dim=32), SparseFeat('301', 5, embedding_dim=32)]
    B, Fn = 24, 8
    shared = None
    for flag in ('usemetatrans', 'usemetatrans-metanorm'):
        torch.manual_seed(77)
        m = BaseModel(cols, cols, flag=flag, domain_column='301', num_domains=3).eval()
        with torch.no_grad():
            if shared is None:                                # (both variants share the scenario embedding + encoder: stored once)
                m.domain_map_dnn.linears[0].weight.copy_(torch.randn(m.domain_map_dnn.linears[0].weight.shape, generator=g) * 0.05)
                shared = {k: v.detach().clone() for k, v in m.state_dict().items() if k.startswith(('domain_embeddings.', 'domain_map_dnn.'))}
            else:
                m.load_state_dict(shared, strict=False)
            if 'metanorm' in flag:
                m.meta_net.ffn_layer_norm.weight.add_(torch.randn(32, generator=g) * 0.1)
                m.meta_net.ffn_layer_norm.bias.add_(torch.randn(32, generator=g) * 0.1)
        X = torch.stack([torch.randint(0, 10, (B,), generator=g), torch.randint(0, 4, (B,), generator=g)], dim=1).float()
        x = torch.randn((B, Fn, 32), generator=g) * 0.5
        cot = torch.randn((B, Fn, 32), generator=g)
        xg = x.clone().requires_grad_(True)
        m.zero_grad()
        y = m.meta_transformation(X, xg)
        (y * cot).sum().backward()
        tag = 'metanorm' if 'metanorm' in flag else 'plain'
        out[tag + '/X'], out[tag + '/x'], out[tag + '/cot'], out[tag + '/y'] = X.numpy(), x.numpy(), cot.numpy(), y.detach().numpy()
        out[tag + '/grad/x'] = xg.grad.numpy()
        keys = [k for k in m.state_dict() if k.startswith(('domain_embeddings.', 'domain_map_dnn.', 'meta_net.'))]
        named = dict(m.named_parameters())
        for k in keys:
            out[('shared' if k in shared else tag) + '/state/' + k] = m.state_dict()[k].detach().clone().numpy()
            out[tag + '/grad/' + k] = named[k].grad.detach().clone().numpy()[subsample('domain_map_dnn_Q.' + k.split('.', 1)[1])
                                                                             if k.startswith('domain_map_dnn.') else slice(None)]
        # the oracle's restatement (autoint_forward's 'usemetatrans' branch uses exactly these calls)
        st = {k: torch.from_numpy(out[('shared' if k in shared else tag) + '/state/' + k]) for k in keys}
        e = F.relu(st['domain_embeddings.weight'][X[:, 1].long()])
        theta = e @ st['domain_map_dnn.linears.0.weight'].t() + st['domain_map_dnn.linears.0.bias']
        norm = 'metanorm' in flag
        o = O.meta_net(x, theta, [32, 64, 32], st['meta_net.ffn_layer_norm.weight'] if norm else None,
                       st['meta_net.ffn_layer_norm.bias'] if norm else None)
        assert torch.allclose(o, y.detach(), atol=1e-6, rtol=0), (o - y).abs().max()
    return out


VARIANTS = (('sgd', 'binary_crossentropy'), ('adagrad', 'mse'), ('rmsprop', 'mae'))


def build_compile_variants():
    """compile(<optimizer string>, <loss string>) of the unmodified reference (meta_basemodel.py:598-640) on the q_small
    model: two steps of its training-loop body (:311-328, eval-mode forward so that dropout is off) per variant.
    Initial state, X and y are those of q_small.npz (same construction), so only losses and end states are stored."""
    sys.path.insert(0, ROOT)
    from satrans_b200 import synth
    from oracle import satrans_oracle as O
    SATrans, SparseFeat, DenseFeat = import_reference()
    schema_name, scale, B, kw = CASES['q_small']
    schema = synth.aliccp_schema(scale)
    cols = [SparseFeat(f, vocabulary_size=schema.data_max[f] + 2, embedding_dim=32) for f in schema.fields]
    cfg = O.OracleConfig(sparse_names=schema.fields, vocab_sizes=schema.vocab_sizes, dense_names=schema.dense,
                         domain_col=schema.domain_col, num_domains=schema.num_domains, embedding_dim=32,
                         layers=kw['domain_att_layer_num'], heads=kw['att_head_num'], meta_mode=kw['meta_mode'], flag=kw['flag'])
    xs, ys = synth.make_inputs(schema, 3 * B, seed=1021, ctr=0.3)
    Xall = synth.stack_columns(xs, schema.fields)
    yall = ys.astype(np.float32)
    base = np.load(os.path.join(ROOT, 'tests', 'golden', 'q_small.npz'))
    out = {}
    torch.set_num_threads(1)
    for opt, loss in VARIANTS:
        model = SATrans(linear_feature_columns=cols, dnn_feature_columns=cols, domain_column_list=[schema.domain_col],
                        num_domains_list=[schema.num_domains], att_layer_num=0, use_linear=False, use_dnn=False, seed='1021',
                        device='cpu', **kw)
        g = torch.Generator().manual_seed(7)
        with torch.no_grad():
            for k, v in model.state_dict().items():
                if 'ffn_layer_norm' in k:
                    v.add_(torch.randn(v.shape, generator=g) * 0.1)
                if k == 'domain_map_dnn_Q.linears.0.weight':
                    v.copy_(torch.randn(v.shape, generator=g) * 0.05)
                if k.startswith('embedding_dict.'):
                    v.copy_(torch.randn(v.shape, generator=g) * 0.3)
        state0 = {k: v.detach().clone() for k, v in model.state_dict().items()}
        for k in O.trainable_keys(cfg):
            assert np.array_equal(state0[k].numpy(), base['state/' + k]), k      # same start as q_small.npz
        model.compile(opt, loss)                                                 # the reference's own string tables
        model.eval()
        tr = O.Trainer({k: v.clone() for k, v in state0.items()}, cfg, lr=model.optim.param_groups[0]['lr'], optimizer=opt, loss=loss)
        tag = '%s_%s' % (opt, loss)
        for step in range(2):
            Xk = torch.from_numpy(Xall[step * B:(step + 1) * B])
            yk = torch.from_numpy(yall[step * B:(step + 1) * B])
            y_pred = model(Xk).squeeze()
            model.optim.zero_grad()
            lv = model.loss_func(y_pred, yk.squeeze(), reduction='sum')           # meta_basemodel.py:317
            reg = model.get_regularization_loss()
            total = lv + reg + model.aux_loss
            total.backward()
            model.optim.step()
            out['%s/loss%d' % (tag, step)] = np.float32(lv.item())
            out['%s/reg%d' % (tag, step)] = np.float32(reg.item())
            o_l, o_r = tr.step(Xk, yk, train_mode=False)
            assert abs(o_l - lv.item()) <= 1e-5 * abs(lv.item()), (tag, step, o_l, lv.item())
        sd = model.state_dict()
        for k in O.trainable_keys(cfg):
            sub = subsample(k)
            out['%s/end/%s' % (tag, k)] = sd[k].detach().numpy()[sub].copy()
            ref_p = sd[k].detach()
            err = (tr.params[k].detach() - ref_p).abs().max().item()
            assert err <= 2e-5 * max(ref_p.abs().max().item(), 1e-3) + 1e-7, (tag, k, err)   # pin: oracle == reference
        out['%s/lr' % tag] = np.float32(model.optim.param_groups[0]['lr'])
    return out


def main():
    os.makedirs(os.path.join(ROOT, 'tests', 'golden'), exist_ok=True)
    if os.environ.get('SAT_GOLDEN_ONLY_VARIANTS'):
        path = os.path.join(ROOT, 'tests', 'golden', 'compile_variants.npz')
        np.savez_compressed(path, **build_compile_variants())
        print('compile_variants ->', path, '%.1f KB' % (os.path.getsize(path) / 1024))
        return
    if os.environ.get('SAT_GOLDEN_ONLY_META'):
        path = os.path.join(ROOT, 'tests', 'golden', 'meta_module.npz')
        np.savez_compressed(path, **build_meta_module())
        print('meta_module ->', path, '%.1f KB' % (os.path.getsize(path) / 1024))
        return
    if os.environ.get('SAT_GOLDEN_ONLY_STAR'):
        path = os.path.join(ROOT, 'tests', 'golden', 'star_trans.npz')
        np.savez_compressed(path, **build_star_trans())
        print('star_trans ->', path, '%.1f KB' % (os.path.getsize(path) / 1024))
        return
    if os.environ.get('SAT_GOLDEN_ONLY_CASES'):        # comma-separated case names
        for name in os.environ['SAT_GOLDEN_ONLY_CASES'].split(','):
            path = os.path.join(ROOT, 'tests', 'golden', name + '.npz')
            np.savez_compressed(path, **build_case(name))
            print(name, '->', path, '%.1f KB' % (os.path.getsize(path) / 1024))
        return
    path = os.path.join(ROOT, 'tests', 'golden', 'selfatt_layer.npz')
    np.savez_compressed(path, **build_selfatt())
    print('selfatt_layer ->', path, '%.1f KB' % (os.path.getsize(path) / 1024))
    if os.environ.get('SAT_GOLDEN_ONLY_SELFATT'):
        return
    path = os.path.join(ROOT, 'tests', 'golden', 'star_trans.npz')
    np.savez_compressed(path, **build_star_trans())
    print('star_trans ->', path, '%.1f KB' % (os.path.getsize(path) / 1024))
    path = os.path.join(ROOT, 'tests', 'golden', 'meta_module.npz')
    np.savez_compressed(path, **build_meta_module())
    print('meta_module ->', path, '%.1f KB' % (os.path.getsize(path) / 1024))
    for name in CASES:
        out = build_case(name)
        path = os.path.join(ROOT, 'tests', 'golden', name + '.npz')
        np.savez_compressed(path, **out)
        print(name, '->', path, '%.1f KB' % (os.path.getsize(path) / 1024))
    path = os.path.join(ROOT, 'tests', 'golden', 'compile_variants.npz')
    np.savez_compressed(path, **build_compile_variants())
    print('compile_variants ->', path, '%.1f KB' % (os.path.getsize(path) / 1024))


if __name__ == '__main__':
    main()
