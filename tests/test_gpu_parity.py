"""GPU parity tests: the CUDA path (through the C ABI) against the golden vectors dumped from the unmodified
reference and against the CPU oracle on the same seeded inputs.  Tolerances (SURVEY.md §8c, fp32 noise floor
of the reference itself): gathered rows bit-exact; y_pred atol 2e-6; BCE / reg rtol 1e-5; gradients
max-norm rtol 1e-4 (+1e-9 abs); post-Adam parameters per `assert_adam_close`.

Every training-path test runs on BOTH arithmetic paths: 'tf32x3' (tcgen05 tensor cores, error-compensated 3xTF32: the
package default and what bench.py times) and 'fp32' (CUDA-core FFMA kernels)."""
import numpy as np
import pytest
import torch

from golden_util import CASES, FORWARD_ONLY_CASES, VARIANT_CASES, Golden, subsample
from oracle import satrans_oracle as O
from test_oracle_golden import assert_adam_close

pytestmark = pytest.mark.gpu


def _dev(t):
    return t.cuda().contiguous()


@pytest.mark.parametrize('name', CASES + FORWARD_ONLY_CASES)
def test_gather_bit_exact(name):
    from gpu_util import model_from_golden
    g = Golden(name)
    model = model_from_golden(g)
    E = model.engine.gather(_dev(g.X()))
    assert torch.equal(E.cpu(), g.arr('E0'))
    # integer id paths give the same rows (dense-free schemas only: a mixed matrix has to be float)
    if not g.cfg.dense_names:
        for dt in (torch.int32, torch.int64):
            assert torch.equal(model.engine.gather(_dev(g.X().to(dt))).cpu(), g.arr('E0'))
    # the API-level call returns the reference's list-of-(B,1,d) structure
    embs, dense = model.input_from_feature_columns(_dev(g.X()), model.dnn_feature_columns, model.embedding_dict)
    assert len(embs) == len(g.cfg.sparse_names) and embs[0].shape == (g.B, 1, 32)
    assert torch.equal(torch.cat(embs, dim=1).cpu(), g.arr('E0'))
    assert len(dense) == len(g.cfg.dense_names)


@pytest.mark.parametrize('name', CASES + FORWARD_ONLY_CASES)
def test_forward_eval_matches_reference(name):
    from gpu_util import model_from_golden
    g = Golden(name)
    model = model_from_golden(g).eval()
    yp = model(_dev(g.X()))
    assert yp.shape == (g.B, 1)
    assert torch.allclose(yp.cpu(), g.arr('y_pred'), atol=2e-6, rtol=0), (yp.cpu() - g.arr('y_pred')).abs().max()
    pred = model.predict(g.x_dict(), batch_size=37)
    assert pred.dtype == np.float64 and pred.shape == (g.B, 1)
    assert np.allclose(pred, g.z['predict'], atol=2e-6, rtol=0)


PRECISIONS = ['tf32x3', 'fp32']
GTOL = {'fp32': 1e-5, 'tf32x3': 2e-5}          # gradient tolerance (max-norm relative) fed to tests/optim_bounds.py; the stated bar is 1e-4


@pytest.mark.parametrize('precision', PRECISIONS)
@pytest.mark.parametrize('name', CASES)
def test_loss_and_gradients_match_reference(name, precision):
    """Loss and every gradient against the reference's golden vectors at the stated bar (max-norm rtol 1e-4) on both paths.
    A hidden unit of a MetaNet that the reference evaluated within round-off of zero may be decided the other way by the
    3xTF32 products; if any golden-vector comparison misses the bar, the gradient must then meet it against the fp64 oracle
    differentiating with the kernel's own ReLU decisions injected (same piecewise-linear function) -- never a looser bar."""
    from gpu_util import dense_grads, model_from_golden, relu_masks, table_grads
    g = Golden(name)
    model = model_from_golden(g, precision=precision)
    l2 = g.cfg.l2_reg_embedding
    yp, w = model.engine.loss_and_grads(_dev(g.X()), _dev(g.y()), training=False, l2_emb=l2)
    torch.cuda.synchronize()
    bce, reg = float(w['scalars'][0]), float(w['scalars'][1])
    assert abs(bce - float(g.z['bce'])) <= 1e-5 * abs(float(g.z['bce']))
    assert abs(reg - float(g.z['reg'])) <= 1e-5 * abs(float(g.z['reg']))
    grads = dense_grads(model, w)
    tg, _ = table_grads(model, w, l2)
    grads.update(tg)
    missed = []
    for k in O.trainable_keys(g.cfg):
        ref = g.arr('grad/' + k)
        got = grads[k][subsample(k)]
        tol = 1e-4 * ref.abs().max().item() + 1e-9
        err = (got - ref).abs().max().item()
        if err > tol:
            missed.append((k, err, tol))
    if missed:
        assert precision != 'fp32', missed            # the CUDA-core path takes the reference's own ReLU decisions
        rm = relu_masks(model, w, g.B)
        _, _, _, og = O.loss_and_grads(g.state(torch.float64), g.X().double(), g.y().double(), g.cfg, relu_masks=rm)
        for k in O.trainable_keys(g.cfg):
            ref = og[k].float()
            err = (grads[k] - ref).abs().max().item()
            assert err <= 1e-4 * ref.abs().max().item() + 1e-9, (k, err, 'vs fp64 oracle with the kernel ReLU decisions', missed)


@pytest.mark.parametrize('precision', PRECISIONS)
@pytest.mark.parametrize('name', CASES)
def test_dense_adam_steps_match_reference(name, precision):
    from gpu_util import model_from_golden
    g = Golden(name)
    model = model_from_golden(g, precision=precision)
    lr = float(g.z['lr'])
    eng = model.engine
    for step in range(1, 4):
        eng.train_step(_dev(g.X(step - 1)), _dev(g.y(step - 1)), lr=lr, l2_emb=g.cfg.l2_reg_embedding,
                       dense_mode=True, training=False)
        if step in (1, 3):
            torch.cuda.synchronize()
            sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
            for k in O.trainable_keys(g.cfg):
                assert_adam_close(sd[k][subsample(k)], g.arr('step%d/%s' % (step, k)), lr, step, k)
    # Adam moments after 3 steps
    st = eng.adam_state()
    names = {id(p): n for n, p in model.named_parameters()}
    m_flat, v_flat = st['dense_m'].cpu(), st['dense_v'].cpu()
    for p, off in model._flat_slots():
        k = names[id(p)]
        if k.endswith('K_meta_mlp.ffn_layer_norm.bias'):
            continue
        m_ref, v_ref = g.arr('step3_m/' + k), g.arr('step3_v/' + k)
        m_got = m_flat[off:off + p.numel()].view(p.shape)[subsample(k)]
        v_got = v_flat[off:off + p.numel()].view(p.shape)[subsample(k)]
        assert (m_got - m_ref).abs().max().item() <= 1e-4 * m_ref.abs().max().item() + 1e-9, k
        assert (v_got - v_ref).abs().max().item() <= 1e-4 * v_ref.abs().max().item() + 1e-12, k
    tm, tv = st['tab_m'].cpu(), st['tab_v'].cpu()
    for f, a, b in zip(model._sparse_feats, model._row_off[:-1], model._row_off[1:]):
        k = 'embedding_dict.%s.weight' % f.embedding_name
        m_ref, v_ref = g.arr('step3_m/' + k), g.arr('step3_v/' + k)
        assert (tm[int(a):int(b)] - m_ref).abs().max().item() <= 1e-4 * m_ref.abs().max().item() + 1e-9, k
        assert (tv[int(a):int(b)] - v_ref).abs().max().item() <= 1e-4 * v_ref.abs().max().item() + 1e-12, k


@pytest.mark.parametrize('name', VARIANT_CASES)
def test_flag_variants_match_reference(name):
    """The reference's ablation flags (satrans.py:59-81,158-176) -- 'gate' (Q' = Q * g * 2), 'bilinear' (per-head Q'_h = Q_h B_h),
    'onlyemb' (wide scenario embedding, no encoder), each with and without 'pos' -- against golden vectors of the unmodified
    reference at the same bars as the main configurations: gathered rows bit-exact, y_pred 2e-6, loss / reg 1e-5, EVERY gradient
    max-norm 1e-4, parameters after 1 and 3 dense-Adam steps.  'gate' / 'bilinear' run in the CUDA-core kernels (the engine picks
    precision='fp32' for them and the tensor-core entry points refuse); 'onlyemb' runs on both arithmetic paths."""
    from gpu_util import dense_grads, model_from_golden, relu_masks, table_grads
    g = Golden(name)
    variant_only_fp32 = 'onlyemb' not in g.cfg.flag
    for precision in (['fp32'] if variant_only_fp32 else PRECISIONS):
        model = model_from_golden(g, precision=None if variant_only_fp32 else precision)
        assert model.precision == precision
        if variant_only_fp32:
            with pytest.raises(NotImplementedError):
                model.precision = 'tf32x3'
        assert torch.equal(model.engine.gather(_dev(g.X())).cpu(), g.arr('E0'))
        yp = model.eval()(_dev(g.X()))
        assert torch.allclose(yp.cpu(), g.arr('y_pred'), atol=2e-6, rtol=0), (yp.cpu() - g.arr('y_pred')).abs().max()
        assert np.allclose(model.predict(g.x_dict(), batch_size=37), g.z['predict'], atol=2e-6, rtol=0)
        l2 = g.cfg.l2_reg_embedding
        yp, w = model.engine.loss_and_grads(_dev(g.X()), _dev(g.y()), training=False, l2_emb=l2)
        torch.cuda.synchronize()
        assert abs(float(w['scalars'][0]) - float(g.z['bce'])) <= 1e-5 * abs(float(g.z['bce']))
        assert abs(float(w['scalars'][1]) - float(g.z['reg'])) <= 1e-5 * abs(float(g.z['reg']))
        grads = dense_grads(model, w)
        tg, _ = table_grads(model, w, l2)
        grads.update(tg)
        keys = O.trainable_keys(g.cfg)
        assert ('domain_map_dnn_Q.linears.0.weight' in keys) == ('onlyemb' not in g.cfg.flag)
        missed = []
        for k in keys:
            ref = g.arr('grad/' + k)
            err = (grads[k][subsample(k)] - ref).abs().max().item()
            if err > 1e-4 * ref.abs().max().item() + 1e-9:
                missed.append((k, err, ref.abs().max().item()))
        if missed:      # tensor-core MetaNet only: same rule as test_loss_and_gradients_match_reference (kernel ReLU decisions injected)
            assert precision != 'fp32', missed
            rm = relu_masks(model, w, g.B)
            _, _, _, og = O.loss_and_grads(g.state(torch.float64), g.X().double(), g.y().double(), g.cfg, relu_masks=rm)
            for k in keys:
                ref = og[k].float()
                err = (grads[k] - ref).abs().max().item()
                assert err <= 1e-4 * ref.abs().max().item() + 1e-9, (k, err, 'vs fp64 oracle with the kernel ReLU decisions', missed)
        lr = float(g.z['lr'])
        for step in range(1, 4):
            model.engine.train_step(_dev(g.X(step - 1)), _dev(g.y(step - 1)), lr=lr, l2_emb=l2, dense_mode=True, training=False)
            if step in (1, 3):
                torch.cuda.synchronize()
                sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
                for k in keys:
                    assert_adam_close(sd[k][subsample(k)], g.arr('step%d/%s' % (step, k)), lr, step, k, grad1=g.arr('grad/' + k))


@pytest.mark.parametrize('precision', PRECISIONS)
def test_reg_and_loss_reported_by_train_step(precision):
    from gpu_util import model_from_golden
    g = Golden('en_small')
    model = model_from_golden(g, precision=precision)
    _, w = model.engine.train_step(_dev(g.X()), _dev(g.y()), lr=1e-3, l2_emb=g.cfg.l2_reg_embedding, training=False)
    torch.cuda.synchronize()
    assert abs(float(w['scalars'][0]) - float(g.z['bce'])) <= 1e-5 * abs(float(g.z['bce']))
    assert abs(float(w['scalars'][1]) - float(g.z['reg'])) <= 1e-5 * abs(float(g.z['reg']))


# ---- tensor-core (tcgen05) forward ---------------------------------------------------------------------------
@pytest.mark.parametrize('name', CASES + FORWARD_ONLY_CASES)
@pytest.mark.parametrize('precision,atol', [('tf32x3', 2e-6), ('tf32', 2e-3)])
def test_tensor_core_forward_matches_reference(name, precision, atol):
    """tcgen05 forward: 3xTF32 keeps the fp32 tolerance of the CUDA-core path; plain TF32 is the fast mode
    (10-bit mantissa products: tolerance 2e-3 on probabilities)."""
    from gpu_util import model_from_golden
    g = Golden(name)
    model = model_from_golden(g).eval()
    model.engine.precision = precision
    yp = model(_dev(g.X()))
    err = (yp.cpu() - g.arr('y_pred')).abs().max().item()
    assert err <= atol, (precision, err)


BF16_YPRED_ATOL, BF16_LOSS_RTOL, BF16_GRAD_COS = 2e-2, 2e-2, 0.98


@pytest.mark.parametrize('name', CASES)
def test_bf16_variant_stated_tolerance(name):
    """precision='bf16': the projections as tcgen05 kind::f16 MMAs on bf16 operands with fp32 accumulation (8-bit mantissa
    products) -- reported separately from the fp32-parity paths, never the headline.  Its own tolerance against the reference's
    golden vectors: y_pred atol 2e-2, BCE rtol 2e-2, every gradient with a cosine similarity >= 0.98 to the reference gradient;
    and it must actually differ from the 3xTF32 path (i.e. the bf16 kernels ran)."""
    from gpu_util import dense_grads, model_from_golden, table_grads
    g = Golden(name)
    model = model_from_golden(g, precision='bf16')
    yp = model.eval()(_dev(g.X()))
    err = (yp.cpu() - g.arr('y_pred')).abs().max().item()
    assert 1e-7 < err <= BF16_YPRED_ATOL, err
    l2 = g.cfg.l2_reg_embedding
    _, w = model.engine.loss_and_grads(_dev(g.X()), _dev(g.y()), training=False, l2_emb=l2)
    torch.cuda.synchronize()
    assert abs(float(w['scalars'][0]) - float(g.z['bce'])) <= BF16_LOSS_RTOL * abs(float(g.z['bce']))
    grads = dense_grads(model, w)
    tg, _ = table_grads(model, w, l2)
    grads.update(tg)
    worst = 1.0
    for k in O.trainable_keys(g.cfg):
        ref = g.arr('grad/' + k).flatten().double()
        got = grads[k][subsample(k)].flatten().double()
        if ref.norm().item() < 1e-12 or k.endswith('K_meta_mlp.ffn_layer_norm.bias'):
            continue
        cos = float(torch.dot(ref, got) / (ref.norm() * got.norm() + 1e-300))
        worst = min(worst, cos)
        assert cos >= BF16_GRAD_COS, (k, cos)
    print('%s: bf16 y_pred err %.2e, worst gradient cosine %.5f' % (name, err, worst))


@pytest.mark.parametrize('bwd', ['fp32', None])
@pytest.mark.parametrize('name', CASES)
def test_tensor_core_gradients_with_injected_relu_decisions(name, bwd):
    """3xTF32 tensor-core forward feeding the CUDA-core ('fp32') or the tensor-core (None) backward: with the forward's own
    ReLU decisions injected into the fp64 oracle, 100 % of the gradient entries meet the fp32 bar (max-norm rtol 1e-4) -- the
    only thing the 3xTF32 rounding can change beyond that bar is which side of zero a hidden pre-activation within ~1e-7 of
    it falls (that claim is what this test pins: no entry is excused)."""
    from gpu_util import dense_grads, model_from_golden, relu_masks, table_grads
    g = Golden(name)
    model = model_from_golden(g, precision='tf32x3')
    model.engine.backward_precision = bwd
    l2 = g.cfg.l2_reg_embedding
    yp, w = model.engine.loss_and_grads(_dev(g.X()), _dev(g.y()), training=False, l2_emb=l2)
    torch.cuda.synchronize()
    assert abs(float(w['scalars'][0]) - float(g.z['bce'])) <= 1e-5 * abs(float(g.z['bce']))
    grads = dense_grads(model, w)
    tg, _ = table_grads(model, w, l2)
    grads.update(tg)
    rm = relu_masks(model, w, g.B)
    # the injected decisions differ from the sign of the fp64 pre-activations in at most a handful of units
    yp64, _, _, og = O.loss_and_grads(g.state(torch.float64), g.X().double(), g.y().double(), g.cfg, relu_masks=rm)
    assert (yp.cpu() - yp64[:, 0].float()).abs().max().item() <= 2e-6
    for k in O.trainable_keys(g.cfg):
        ref = og[k].float()
        err = (grads[k] - ref).abs().max().item()
        assert err <= 1e-4 * ref.abs().max().item() + 1e-9, (k, err, ref.abs().max().item())


@pytest.mark.parametrize('precision', PRECISIONS)
def test_two_runs_are_bit_identical(precision):
    """No floating-point atomics: two runs of three train-mode (dropout on) default-eps Adam steps from the same state give
    bit-identical parameters, Adam moments, predictions and losses."""
    from gpu_util import model_from_golden
    g = Golden('enp_small')
    outs = []
    for run in range(2):
        model = model_from_golden(g, precision=precision)
        eng = model.engine
        rec = []
        for step in range(3):
            yp, w = eng.train_step(_dev(g.X(step)), _dev(g.y(step)), lr=float(g.z['lr']), l2_emb=g.cfg.l2_reg_embedding,
                                   dense_mode=True, training=True, dropout_p=0.1, seed=77 + step)
            rec.append(yp.clone())
            rec.append(w['scalars'][0:2].clone())
        torch.cuda.synchronize()
        st = eng.adam_state()
        outs.append([t.cpu() for t in rec] + [model._tables.detach().cpu().clone(), model._dense.detach().cpu().clone(),
                                              st['tab_m'].cpu().clone(), st['tab_v'].cpu().clone(),
                                              st['dense_m'].cpu().clone(), st['dense_v'].cpu().clone()])
    for a, b in zip(*outs):
        assert torch.equal(a, b)


@pytest.mark.parametrize('precision', PRECISIONS)
@pytest.mark.parametrize('opt,loss', [('sgd', 'binary_crossentropy'), ('adagrad', 'mse'), ('rmsprop', 'mae')])
def test_compile_string_variants_match_reference(opt, loss, precision):
    """compile('<optimizer>', '<loss>') (reference meta_basemodel.py:612-640) through the drop-in API: two train_on_batch steps on
    the q_small golden model against the reference's own two steps (tests/golden/compile_variants.npz).

    sgd: every parameter within fp32 round-off of the reference's end state.  Adagrad / RMSprop start from a zero accumulator, so
    their normalised steps lr * g / (sqrt(S) + eps) move an element whose gradient is within round-off of zero by the full step in
    a direction decided by the gradient's last bit: the reference on two devices does not agree with itself there.  For them every
    element must stay inside the budget that the gradient tolerance (`GTOL`, 5-10 x TIGHTER than the stated 1e-4) allows after
    exact interval propagation through the update formula (tests/optim_bounds.py; fp64 oracle as the centre, re-based on the
    kernel's parameters after each step), the losses must equal the reference's, and the bulk of every tensor (as many elements as
    the fp64 oracle's end state, or at least half) must match the reference's end state to round-off."""
    import os
    from golden_util import GOLDEN_DIR
    from gpu_util import model_from_golden
    import optim_bounds
    g = Golden('q_small')
    v = np.load(os.path.join(GOLDEN_DIR, 'compile_variants.npz'))
    tag = '%s_%s' % (opt, loss)
    lr = float(v[tag + '/lr'])
    model = model_from_golden(g, precision=precision)
    model.dropout_rate = 0.0
    model.compile(opt, loss)
    assert abs(model._opt_hparams['lr'] - lr) < 1e-8          # (the fixture stores lr as float32)

    def impl_step(X, y):
        return model.train_on_batch(X.numpy(), y.numpy())[0]

    def impl_state():
        return {k: t.detach().cpu().clone() for k, t in model.state_dict().items()}

    if opt == 'sgd':
        losses = [impl_step(g.X(step), g.y(step)) for step in range(2)]
    else:
        losses = optim_bounds.check_normalising_steps(O, g, opt, loss, lr, impl_step, impl_state, gtol=GTOL[precision])
    for step in range(2):
        ref_l = float(v['%s/loss%d' % (tag, step)])
        assert abs(losses[step] - ref_l) <= 2e-5 * abs(ref_l), (step, losses[step], ref_l)
    sd = impl_state()
    tr64 = O.Trainer({k: t.double() for k, t in g.state().items()}, g.cfg, lr=lr, optimizer=opt, loss=loss)
    for step in range(2):
        tr64.step(g.X(step).double(), g.y(step).double(), train_mode=False)
    for k in O.trainable_keys(g.cfg):
        ref = torch.from_numpy(v['%s/end/%s' % (tag, k)])
        err = (sd[k][subsample(k)] - ref).abs()
        tight = 2e-5 * max(ref.abs().max().item(), 1e-3) + 1e-7
        if opt == 'sgd':
            assert err.max().item() <= tight, k
        else:
            e64 = (tr64.params[k].detach().float()[subsample(k)] - ref).abs()
            # (one near-zero element decided the other way at step 1 -- inside its budget above -- shifts every gradient of step 2
            #  a little, which RMSprop's normalisation turns into ~1e-4 moves: only the bulk can be compared with the golden end state)
            assert (err <= tight).float().mean().item() >= min((e64 <= tight).float().mean().item() - 0.05, 0.5), k


def _run_steps(g, precision, n_steps, defer, opt_kind=0):
    from gpu_util import model_from_golden
    model = model_from_golden(g, precision=precision)
    eng = model.engine
    preds, regs = [], []
    for step in range(n_steps):
        k = step % 3
        yp, w = eng.train_step(_dev(g.X(k)), _dev(g.y(k)), lr=float(g.z['lr']), l2_emb=g.cfg.l2_reg_embedding, dense_mode=True,
                               training=True, dropout_p=0.1, seed=1234 + step, opt_kind=opt_kind, defer=defer)
        preds.append(yp.clone())
        if defer <= 1:
            regs.append(w['scalars'][1].clone())
        elif eng.defer_pending == defer or step == n_steps - 1:          # the epoch is full (or the run ends): close it
            eng.sync_tables()
            regs += list(eng.deferred_reg()[0].clone())
    st = eng.adam_state()
    torch.cuda.synchronize()
    state = [model._tables.detach().cpu().clone(), model._dense.detach().cpu().clone()]
    state += [st[k].cpu().clone() for k in ('tab_m', 'tab_v', 'dense_m', 'dense_v') if st[k] is not None]
    return [p.cpu() for p in preds], torch.stack([r.double().cpu() for r in regs]), state


@pytest.mark.parametrize('opt_kind', [0, 3])
@pytest.mark.parametrize('precision', PRECISIONS)
def test_deferred_table_sweep_is_bit_identical(precision, opt_kind):
    """The every-row table update (reference meta_basemodel.py:577-593 + main.py:343) batched over sweep epochs of K steps
    (sat_defer_*): predictions of every step, parameters and optimiser state must be BIT-identical to sweeping after every step
    (the batches re-read rows touched one, two and three steps earlier and rows never touched before), and every step's
    regulariser term must come out the same (different summation grouping: 1e-9 relative)."""
    g = Golden('en_small')
    n_steps = 11
    p1, r1, s1 = _run_steps(g, precision, n_steps, 1, opt_kind)
    for K in (4, 16):
        pk, rk, sk = _run_steps(g, precision, n_steps, K, opt_kind)
        for a, b in zip(p1, pk):
            assert torch.equal(a, b)
        for a, b in zip(s1, sk):
            assert torch.equal(a, b)
        assert rk.shape == r1.shape
        assert torch.allclose(rk, r1, rtol=1e-6, atol=0), (rk, r1)


def test_deferred_sweep_through_the_api():
    """fit() with table_sweep_interval = 8 (default) against 1: same parameters bit for bit, same epoch losses; train_on_batch(
    defer_reg=True) + sync_tables() against train_on_batch(): same state_dict; reading the state closes the epoch by itself."""
    from gpu_util import model_from_golden
    g = Golden('en_small')
    X = np.concatenate([g.X(k).numpy() for k in range(3)] * 4)
    y = np.concatenate([g.y(k).numpy() for k in range(3)] * 4)
    hist, sds = [], []
    for K in (1, 8):
        torch.manual_seed(7)
        model = model_from_golden(g)
        model.table_sweep_interval = K
        model.compile('adam', 'binary_crossentropy')
        h = model.fit(X, y, batch_size=g.B // 2, epochs=2, verbose=0, shuffle=True)
        hist.append(h.history['loss'])
        sds.append({k: v.detach().cpu().clone() for k, v in model.state_dict().items()})
    for k in sds[0]:
        assert torch.equal(sds[0][k], sds[1][k]), k
    assert np.allclose(hist[0], hist[1], rtol=1e-6, atol=0), hist
    sds = []
    for defer in (False, True):
        torch.manual_seed(7)
        model = model_from_golden(g)
        model.table_sweep_interval = 4
        model.compile('adam', 'binary_crossentropy')
        for step in range(7):
            bce, reg = model.train_on_batch(g.X(step % 3).numpy(), g.y(step % 3).numpy(), defer_reg=defer)
            assert (reg is None) == defer
        sds.append({k: v.detach().cpu().clone() for k, v in model.state_dict().items()})      # (state_dict() syncs the tables)
        assert model.engine.defer_pending == 0
    for k in sds[0]:
        assert torch.equal(sds[0][k], sds[1][k]), k


@pytest.mark.parametrize('shuffle', [False, True])
def test_fit_streaming_host_batches_equals_resident(shuffle):
    """fit() with the training matrix left in pinned host memory and each batch copied one step ahead on a copy stream
    (stream_host_batches, the reference's DataLoader + per-batch .to(device), meta_basemodel.py:311-313) == fit() with the matrix
    resident in HBM: same parameters bit for bit, same epoch losses (ragged last batch included)."""
    from gpu_util import model_from_golden
    g = Golden('en_small')
    X = np.concatenate([g.X(k).numpy() for k in range(3)] * 3)[:-7]
    y = np.concatenate([g.y(k).numpy() for k in range(3)] * 3)[:-7]
    hist, sds = [], []
    for stream in (False, True):
        torch.manual_seed(11)
        model = model_from_golden(g)
        model.stream_host_batches = stream
        model.compile('adam', 'binary_crossentropy')
        h = model.fit(X, y, batch_size=64, epochs=2, verbose=0, shuffle=shuffle)
        hist.append(h.history['loss'])
        sds.append({k: v.detach().cpu().clone() for k, v in model.state_dict().items()})
    for k in sds[0]:
        assert torch.equal(sds[0][k], sds[1][k]), k
    assert hist[0] == hist[1], hist


def test_tensor_core_weight_gradients():
    """SAT_WGRAD_TC=1 selects backward_tc_wg.cu: the MetaNet / QKV WEIGHT gradients as tcgen05 3xTF32 GEMMs with both operands
    MN-major in shared memory (SWIZZLE_128B_BASE32B; the token index is the MMA's K).  Same bar as the default path: the
    golden-vector gradient tests, the config-2-sized check and run-to-run bit identity, in a fresh process with the switch set."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, SAT_WGRAD_TC='1')
    r = subprocess.run([sys.executable, '-m', 'pytest', os.path.join(root, 'tests', 'test_gpu_parity.py'), '-m', 'gpu', '-q', '-x',
                        '--no-header', '-p', 'no:cacheprovider', '-k', 'tf32x3 and (gradients or two_runs or dense_adam)'],
                       env=env, capture_output=True, text=True, timeout=900, cwd=root)
    assert r.returncode == 0 and ' passed' in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]


@pytest.mark.parametrize('precision', PRECISIONS)
def test_custom_loss_callable(precision):
    """compile(loss=<callable>) (the reference passes any callable through, meta_basemodel.py:627-640, and calls it as
    loss(y_pred, y, reduction='sum'), :317): torch evaluates the loss and its derivative on the prediction vector between the fused
    forward and the fused backward.  A callable that IS binary cross-entropy / mse must reproduce the built-in fused losses
    (two optimiser steps: losses and every parameter to fp32 round-off); a focal-style loss must match torch autograd through the
    oracle."""
    import torch.nn.functional as F
    from gpu_util import model_from_golden
    g = Golden('en_small')
    for name, fn in (('binary_crossentropy', F.binary_cross_entropy), ('mse', F.mse_loss)):
        sds, losses = [], []
        for loss in (name, fn):
            model = model_from_golden(g, precision=precision)
            model.dropout_rate = 0.0
            model.compile('sgd', loss)
            losses.append([model.train_on_batch(g.X(k).numpy(), g.y(k).numpy())[0] for k in range(2)])
            sds.append({k: v.detach().cpu().clone() for k, v in model.state_dict().items()})
        assert np.allclose(losses[0], losses[1], rtol=2e-6, atol=0), (name, losses)
        for k in sds[0]:
            scale = max(sds[0][k].abs().max().item(), 1e-3)
            assert (sds[0][k] - sds[1][k]).abs().max().item() <= 2e-6 * scale, (name, k)

    def focal(yp, y, reduction='sum'):
        pt = torch.where(y > 0.5, yp, 1.0 - yp)
        v = -((1.0 - pt) ** 2) * torch.log(pt.clamp_min(1e-12))
        return v.sum() if reduction == 'sum' else v.mean()

    model = model_from_golden(g, precision=precision)
    model.dropout_rate = 0.0
    lr = 0.01
    model.compile(torch.optim.SGD(model.parameters(), lr=lr), focal)
    got_loss = model.train_on_batch(g.X().numpy(), g.y().numpy())[0]
    st = {k: t.clone().requires_grad_(k in O.trainable_keys(g.cfg)) for k, t in g.state().items()}
    yp = O.forward(st, g.X(), g.cfg).squeeze(-1)
    ref_loss = focal(yp, g.y())
    (ref_loss + O.reg_loss(st, g.cfg)).backward()
    assert abs(got_loss - ref_loss.item()) <= 1e-5 * abs(ref_loss.item())
    sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    for k in O.trainable_keys(g.cfg):
        step = (st[k].detach() - sd[k]) / lr                      # = the gradient the fused path applied
        ref = st[k].grad
        tol = 1e-3 * ref.abs().max().item() + 4e-7 * st[k].detach().abs().max().item() / lr + 1e-6   # (+ fp32 storage of p -/- lr g)
        assert (step - ref)[subsample(k)].abs().max().item() <= tol, k


def test_attention_backward_with_two_threads_per_token():
    """SAT_BWD_ATTN_SPLIT=1 selects backward_attn2.cu (every token split over two threads in sibling warps: 256 threads, <= 128
    registers, 16 warps per SM).  Same bar as the default kernel -- golden-vector gradients (2- and 4-head models), Adam steps,
    run-to-run bit identity -- in a fresh process with the switch set."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, SAT_BWD_ATTN_SPLIT='1')
    r = subprocess.run([sys.executable, '-m', 'pytest', os.path.join(root, 'tests', 'test_gpu_parity.py'), '-m', 'gpu', '-q', '-x',
                        '--no-header', '-p', 'no:cacheprovider', '-k', 'gradients_match or two_runs or dense_adam'],
                       env=env, capture_output=True, text=True, timeout=900, cwd=root)
    assert r.returncode == 0 and ' passed' in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]
