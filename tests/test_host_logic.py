"""CPU-only tests: host-side logic of the drop-in API and the C-ABI surface (no compute calls)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from golden_util import Golden
from satrans_b200 import DenseFeat, SATrans, SparseFeat, build_input_features, get_feature_names, synth
from satrans_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _model(g, device='cpu'):
    m = g.meta
    cols = ([SparseFeat(n, v, 32) for n, v in zip(m['sparse_names'], m['vocab_sizes'])] +
            [DenseFeat(n, 1) for n in m['dense_names']])
    return SATrans(cols, cols, [m['domain_col']], [m['num_domains']], att_layer_num=0,
                   domain_att_layer_num=m['domain_att_layer_num'], att_head_num=m['att_head_num'],
                   meta_mode=m['meta_mode'], seed='1021', device=device, flag=m['flag'])


def test_header_symbols_exported_and_bound():
    hdr = open(os.path.join(ROOT, 'include', 'satrans_b200.h')).read()
    declared = set(re.findall(r'^\s*(?:int|int64_t|const char\*)\s+(sat_\w+)\s*\(', hdr, flags=re.M))
    assert declared, 'no declarations parsed'
    assert declared == set(_lib.PROTOTYPES), declared ^ set(_lib.PROTOTYPES)
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    l = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(l, name), name
    assert l.sat_abi_version() == _lib.SAT_ABI_VERSION == 4
    # struct mirrors: sizes implied by the header layout
    assert ctypes.sizeof(_lib.SatModel) == 10 * 4 + 10 * 8 + 5 * 8 + 8 + 8 + 7 * 8 * 8
    assert ctypes.sizeof(_lib.SatWork) == 24 + _lib.SAT_WORK_BUFFERS * 8
    assert ctypes.sizeof(_lib.SatAdam) == 32 + 7 * 8


def test_workspace_sizes_host_only():
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    l = _lib.lib()
    d = _lib.SatModel()
    d.F, d.L, d.H, d.n_dom, d.total_rows, d.n_dense_params, d.flags = 19, 3, 4, 4, 6571961, 148580, 0
    d.n_shards, d.shard_rank, d.local_rows = 1, 0, 6571961
    sizes = (ctypes.c_int64 * _lib.SAT_WORK_BUFFERS)()
    mt = ctypes.c_int32()
    assert l.sat_workspace_sizes(ctypes.byref(d), 8192, 1, sizes, ctypes.byref(mt)) == 0
    s = dict(zip(_lib.WORK_BUFFERS, sizes))
    assert s['acts'] == (3 * _lib.SAT_ACT_PER_LAYER + 1) * 8192 * 19 * 32 * 4 + 3 * 2 * 8192 * 19 * 2 * 4
    assert s['bitmap'] == (6571961 + 31) // 32 * 4
    assert mt.value == (8192 + 9) // 10 + 4
    assert l.sat_workspace_sizes(ctypes.byref(d), 8192, 0, sizes, ctypes.byref(mt)) == 0
    assert dict(zip(_lib.WORK_BUFFERS, sizes))['acts'] == 0
    d.F = 40
    assert l.sat_workspace_sizes(ctypes.byref(d), 8192, 1, sizes, ctypes.byref(mt)) != 0
    assert b'bad argument' in l.sat_last_error()


def test_feature_columns_and_index():
    cols = [SparseFeat('a', 10, 32), SparseFeat('b', 5, 32), DenseFeat('p', 1)]
    fi = build_input_features(cols + cols)
    assert list(fi.items()) == [('a', (0, 1)), ('b', (1, 2)), ('p', (2, 3))]
    assert get_feature_names(cols) == ['a', 'b', 'p']
    assert SparseFeat('x', 100, 'auto').embedding_dim == 6 * int(pow(100, 0.25))
    assert SparseFeat('x', 7).embedding_name == 'x'


@pytest.mark.parametrize('name', ['en_small', 'enp_small'])
def test_state_dict_is_reference_compatible(name):
    g = Golden(name)
    model = _model(g)
    full = g.full_state_dict()
    sd = model.state_dict()
    assert set(sd) == set(full)
    for k in sd:
        assert tuple(sd[k].shape) == tuple(full[k].shape), k
    model.load_state_dict(full)
    # every kernel-visible parameter is a view into the flat buffer / the concatenated table
    for p, off in model._flat_slots():
        assert p.data_ptr() == model._dense.data_ptr() + 4 * off
    o = model._layout['off_layers']
    assert torch.equal(model._dense[o:o + 1024].view(32, 32), full['domain_int_layers.0.W_Query'])
    first = g.cfg.sparse_names[0]
    assert torch.equal(model._tables[:g.cfg.vocab_sizes[0]], full['embedding_dict.%s.weight' % first])
    # aliases stay aliases (reference satrans.py:44-47,179-180)
    assert model.domain_map_dnn_K is model.domain_map_dnn_Q
    lay0 = model.domain_int_layers[0]
    assert (lay0.K_meta_mlp is lay0.Q_meta_mlp) == ('pos' not in g.meta['flag'])
    # .to() keeps the views intact
    model.to('cpu')
    model._ensure_flat()
    for p, off in model._flat_slots():
        assert p.data_ptr() == model._dense.data_ptr() + 4 * off


def test_init_follows_reference_distributions():
    g = Golden('en_small')
    model = _model(g)
    lay = model.domain_int_layers[0]
    assert abs(lay.layer_norm.weight.mean().item()) < 0.05          # N(0, .05), not ones (satrans.py:42-43)
    assert torch.equal(lay.Q_meta_mlp.ffn_layer_norm.weight, torch.ones(32))
    assert model._tables.std().item() < 2e-4
    assert model.domain_map_dnn_Q.linears[0].weight.std().item() < 2e-4


def test_errors_match_reference_contract():
    cols = [SparseFeat('a', 10, 32), SparseFeat('301', 5, 32)]
    with pytest.raises(ValueError):
        SATrans(cols, cols, ['301'], [3], att_head_num=0, flag='sota')
    with pytest.raises(ValueError):
        SATrans(cols, cols, ['301'], [3], att_head_num=3, flag='sota')
    with pytest.raises(TypeError):
        SATrans(cols, cols, ['301'], [3], att_head_num=2, flag=None)
    with pytest.raises(ValueError):
        SATrans([SparseFeat('a', 10, 32), SparseFeat('301', 5, 16)], [SparseFeat('a', 10, 32), SparseFeat('301', 5, 16)],
                ['301'], [3], flag='sota')
    m = SATrans(cols, cols, ['301'], [3], att_head_num=2, flag='sota')
    with pytest.raises(NotImplementedError):
        m.compile('nadam', 'binary_crossentropy')
    for name in ('sgd', 'adam', 'adagrad', 'rmsprop'):
        m.compile(name, 'binary_crossentropy')
    for loss in ('mse', 'mae'):
        m.compile('adam', loss)
    with pytest.raises(NotImplementedError):
        m.compile(torch.optim.SGD(m.parameters(), lr=0.1, momentum=0.9), 'binary_crossentropy')
    with pytest.raises(NotImplementedError):
        m.compile('adam', 'hinge')
    m.compile(torch.optim.Adam(m.parameters(), lr=0.005), 'binary_crossentropy', metrics=['binary_crossentropy', 'auc'])
    assert m._opt_hparams['lr'] == 0.005 and m.metrics_names == ['loss', 'binary_crossentropy', 'auc']
    # no CPU fallback: using the engine off-GPU fails loudly
    with pytest.raises(Exception):
        m(torch.zeros((2, 2)))


def test_stack_inputs_layouts():
    cols = [SparseFeat('a', 10, 32), SparseFeat('301', 5, 32), DenseFeat('p', 1)]
    m = SATrans(cols, cols, ['301'], [3], att_head_num=2, flag='sota')
    x = {'a': np.array([1, 2, 3]), '301': np.array([1, 2, 3]), 'p': np.array([.5, .25, 1.], dtype=np.float32)}
    X = m._stack_inputs(x)
    assert X.dtype == np.float32 and X.shape == (3, 3) and X[1, 2] == 0.25
    cols2 = cols[:2]
    m2 = SATrans(cols2, cols2, ['301'], [3], att_head_num=2, flag='sota')
    X2 = m2._stack_inputs({'a': np.array([1, 2 ** 25]), '301': np.array([1, 2])})
    assert X2.dtype == np.int32 and X2[1, 0] == 2 ** 25            # exact beyond 2**24 on the integer path
    with pytest.raises(ValueError):
        m._stack_inputs({'a': np.array([2 ** 25]), '301': np.array([1]), 'p': np.array([.5], dtype=np.float32)})


def test_synth_schemas():
    s = synth.aliccp_schema()
    assert s.total_rows == 6571961 and len(s.fields) == 19 and s.num_domains == 3
    x, y = synth.make_inputs(synth.aliccp_schema(1e-3), 1000, seed=1)
    assert set(np.unique(x['301'])) <= {1, 2, 3} and y.shape == (1000,)
    a = synth.alimama_schema()
    assert len(a.fields) == 15 and a.dense == ['price'] and a.num_domains == 4
    xz, _ = synth.make_inputs(synth.aliccp_schema(1e-2), 1000, seed=1, dist='zipf')
    for f in s.fields:
        assert xz[f].min() >= 0


def test_precision_knob():
    """model.precision selects the projection arithmetic of the engine ('tf32x3' default); unknown values raise."""
    s = synth.aliccp_schema(1e-4)
    cols = s.feature_columns(32)
    m = SATrans(cols, cols, ['301'], [3], att_layer_num=0, domain_att_layer_num=1, att_head_num=2, meta_mode='QK',
                device='cpu', flag='sota')
    assert m.precision == 'tf32x3'
    m.precision = 'fp32'
    assert m.engine.precision == 'fp32'
    m.precision = 'bf16'            # the kind::f16 variant (reported separately)
    assert m.engine.precision == 'bf16'
    with pytest.raises(ValueError):
        m.precision = 'fp8'


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the unmodified reference vendored into oracle/_ref where present -- kind "reference" --
    else the oracle port of its step -- kind "port" -- on the host cores) prints exactly ONE JSON line on stdout with the
    keys of the measurement contract; a tiny vocabulary / batch keeps this a seconds-long CPU test."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, 'bench.py'), '--impl', 'reference', '--vocab-scale', '0.001',
                        '--batch', '128', '--steps', '1', '--warmup', '1'], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    for k in ('metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling', 'vs_baseline',
              'dtype', 'data', 'config', 'e2e', 'cpu_baseline', 'impl'):
        assert k in d, k
    assert d['impl'] == 'reference' and d['metric'] == 'train_samples_per_sec' and d['unit'] == 'samples/s'
    have_ref = os.path.exists(os.path.join(root, 'oracle', '_ref', 'models', 'satrans.py'))
    assert d['value'] > 0 and d['cpu_baseline']['kind'] == ('reference' if have_ref else 'port') and d['cpu_baseline']['cores'] >= 1
    assert d['e2e']['h2d_bytes_per_step'] == 0 and d['e2e']['d2h_bytes_per_step'] == 0 and 'workload' in d['config']


def test_constructor_error_contract():
    """What the drop-in constructor accepts and what it refuses: same ValueErrors as the reference
    (satrans.py:16-19, meta_basemodel.py:685), NotImplementedError -- never a silent fallback -- for configurations whose
    kernels are not built."""
    import pytest
    from satrans_b200 import SATrans, SparseFeat

    def cols(d=32, n=3):
        return [SparseFeat('f%d' % i, 7 + i, d) for i in range(n - 1)] + [SparseFeat('dom', 5, d)]

    def make(c=None, **kw):
        c = c or cols()
        args = dict(att_layer_num=0, domain_att_layer_num=2, att_head_num=4, meta_mode='QK', seed=3, device='cpu', flag='sota')
        args.update(kw)
        return SATrans(c, c, ['dom'], [3], **args)

    make()                                                     # the supported configuration builds on the CPU (parameters only)
    make(flag='sota-pos', meta_mode='Q', att_head_num=8, att_res=False, seed='1021')
    for bad_heads in (0, -1):
        with pytest.raises(ValueError):
            make(att_head_num=bad_heads)
    with pytest.raises(ValueError):
        make(att_head_num=3)                                   # 32 % 3 != 0
    with pytest.raises(ValueError):                            # unequal embedding dims (meta_basemodel.py:677-686)
        c = cols()
        c[0] = SparseFeat('f0', 7, 16)
        make(c)
    for d in (16, 64):
        with pytest.raises(NotImplementedError, match='embedding_dim=32'):
            make(cols(d))
    with pytest.raises(NotImplementedError, match='meta_dnn_hidden_units'):
        make(meta_dnn_hidden_units=(32, 32))
    with pytest.raises(NotImplementedError, match='att_head_num'):
        make(att_head_num=16)
    # the ablation flags build (state-dict shapes of reference satrans.py:158-176); their invalid combinations raise
    sd = make(flag='sota-gate', meta_mode='QK').state_dict()
    assert sd['domain_map_dnn_Q.linears.0.weight'].shape == (32, 32)
    sd = make(flag='sota-bilinear-pos', att_head_num=4).state_dict()
    assert sd['domain_map_dnn_Q.linears.0.weight'].shape == (256, 64)
    sd = make(flag='sota-onlyemb').state_dict()
    assert sd['domain_embeddings.weight'].shape[1] == 4096 and not any(k.startswith('domain_map_dnn') for k in sd)
    with pytest.raises(ValueError, match="cannot be combined"):
        make(flag='sota-gate-bilinear')
    with pytest.raises(NotImplementedError, match="onlyemb"):
        make(flag='sota-gate-onlyemb')
    m = make(flag='sota-gate', meta_mode='Q')
    with pytest.raises(NotImplementedError, match="CUDA-core"):
        m.precision = 'tf32x3'
    with pytest.raises(NotImplementedError, match='embedding_name'):
        c = cols()
        c[0] = SparseFeat('f0', 8, 32, embedding_name='f1')      # shares f1's table
        make(c)
    with pytest.raises(NotImplementedError, match='multi-column'):
        c = cols()
        SATrans(c, c, ['dom', 'f0'], [3, 7], att_layer_num=0, domain_att_layer_num=2, att_head_num=4, device='cpu', flag='sota')
    with pytest.raises(NotImplementedError):
        make(cols(n=40))                                       # more than SAT_MAXF = 32 sparse fields
    with pytest.raises(TypeError):
        make(flag=None)                                        # the reference substring-tests the flag: it must be a str
    m = make()
    with pytest.raises(NotImplementedError):
        m.compile('adamw')                                     # unknown optimizer string (meta_basemodel.py:623)
    with pytest.raises(NotImplementedError):
        m.compile('adam', loss='hinge')                        # unknown loss string (meta_basemodel.py:637)
    m.compile('adam', loss=lambda p, y, reduction='sum': (p - y).abs().sum())          # a callable is accepted (meta_basemodel.py:638-639)
    assert callable(m._loss_fn)
    with pytest.raises(NotImplementedError):
        m.compile('adam', loss=3.14)                               # neither a known string nor a callable
    import torch
    with pytest.raises(NotImplementedError):
        m.compile(torch.optim.Adam(m.parameters(), lr=0.01, amsgrad=True), 'binary_crossentropy')
    from satrans_b200 import _lib
    m.compile('adam', 'binary_crossentropy')
    with pytest.raises(_lib.SatError, match='no CPU fallback'):
        m.engine.bind()                                        # the kernels need the model on a CUDA device


def test_no_floating_point_atomics_in_the_library():
    """Determinism by construction: the SASS of libsatrans_b200.so contains no floating-point atomic / reduction instruction
    (every gradient sum has a fixed order; the remaining ATOM* / RED* are integer counters, bitmap ORs and mbarrier-style
    shared-memory bit operations).  Also: the tcgen05 / TMA instructions the design rests on are present."""
    import shutil
    import subprocess
    if shutil.which('cuobjdump') is None:
        pytest.skip('cuobjdump not on PATH')
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    sass = subprocess.run(['cuobjdump', '-sass', _lib.LIB_PATH], capture_output=True, text=True, timeout=300).stdout
    atomics = set(re.findall(r'\b((?:ATOMS|ATOMG|ATOM|REDG|RED)\.[A-Z0-9_.]+)', sass))
    assert atomics, 'no atomics parsed: cuobjdump output changed?'
    bad = [a for a in atomics if re.search(r'F32|F16|F64|BF16|FADD|\.F\b', a)]
    assert not bad, bad
    for op in ('UTCHMMA', 'UTMALDG', 'UBLKCP', 'LDTM', 'STTM'):
        assert op in sass, op


def test_header_is_plain_c_and_matches_the_library(tmp_path):
    """include/satrans_b200.h compiles as C99 (gcc -pedantic -Werror) and a C program that dlopens the library sees the same
    ABI version and struct sizes: what a cgo / JNI / plain-C binding of the reference's maintainers would be built on."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / 'abi_check')
    subprocess.run(['gcc', '-std=c99', '-Wall', '-Wextra', '-pedantic', '-Werror', '-I', os.path.join(root, 'include'),
                    os.path.join(root, 'tests', 'abi_check.c'), '-o', exe, '-ldl'], check=True)
    r = subprocess.run([exe, os.path.join(root, 'satrans_b200', 'lib', 'libsatrans_b200.so')], capture_output=True, text=True)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    assert 'abi 4 header 4' in r.stdout


def test_column_store_to_input_matrix(tmp_path):
    """SURVEY 8f rank 4 (host logic): a column store with the reference's wire content (one array per column and group,
    utils.py:266-278) -> the model's (N, n_cols) input matrix; the scenario column is shifted to start from 1 (main.py:108-110);
    the HDF5 container itself needs h5py (absent here: a clear ImportError, never a silent fallback)."""
    import pytest
    from satrans_b200 import SATrans, data, synth
    schema = synth.aliccp_schema(1e-4)
    xs, ys = synth.make_inputs(schema, 500, seed=5)
    stored = {k: v.astype(np.int32) for k, v in xs.items()}
    stored['301'] = stored['301'] - 1                              # ids stored from 0, as in the reference's file
    stored['click'] = ys.astype(np.int64)
    store = data.NpyDirColumnStore(str(tmp_path))
    store.write('ctr_train', stored)
    x, y = data.load_aliccp(store, 'ctr_train')
    assert list(x) == data.ALICCP_SPARSE == schema.fields and np.array_equal(y, ys)
    assert x['301'].min() == 1 and np.array_equal(x['301'], xs['301'])
    cols = schema.feature_columns(32)
    model = SATrans(cols, cols, [schema.domain_col], [schema.num_domains], att_layer_num=0, domain_att_layer_num=1, att_head_num=4,
                    meta_mode='QK', seed=3, device='cpu', flag='sota')
    X = data.to_input_matrix(model, x, pin=False)
    assert X.dtype == torch.int32 and X.shape == (500, 19)
    assert np.array_equal(X.numpy(), np.stack([xs[f] for f in schema.fields], axis=1))
    with pytest.raises(KeyError):
        store.read('ctr_test', ['101'])
    try:
        import h5py  # noqa: F401
    except ImportError:
        with pytest.raises(ImportError, match='h5py'):
            data.H5ColumnStore(str(tmp_path / 'alicpp.h5'))
    row = data.append_result_row(str(tmp_path / 'alicpp_results.csv'), 'SATrans_32_0.005_3_4_QK_1021_301_sota', [0.61, 0.62], 0.1623)
    assert row.endswith('SATrans_32_0.005_3_4_QK_1021_301_sota,0.610000,0.620000,0.162300')
    assert open(str(tmp_path / 'alicpp_results.csv')).read().strip() == row


def test_attention_baselines_host_contract():
    """Star_Trans / AutoInt on the host (parameter containers only): checkpoint keys equal the unmodified reference's
    (tests/golden/star_trans.npz holds its state dicts), compile() builds real torch.optim instances from the reference's
    string table (basemodel.py:554-568), unsupported options raise, and nothing computes without CUDA (no CPU fallback)."""
    import json
    import os
    import pytest
    from satrans_b200 import AutoInt, SparseFeat, Star_Trans
    from satrans_b200._lib import SatError
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'star_trans.npz'))
    meta = json.loads(bytes(z['model/fields']).decode())
    cols = [SparseFeat(f, v, 32) for f, v in zip(meta['fields'], meta['vocab'])]
    m1 = Star_Trans(cols, cols, meta['domain_col'], meta['num_domains'], att_layer_num=2, domain_att_layer_num=1, att_head_num=4,
                    merge='no', use_linear=False, seed=1021, device='cpu')
    ref_keys = {k[len('model/state/'):] for k in z.files if k.startswith('model/state/')}
    assert ref_keys == {k for k in m1.state_dict() if not k.startswith('linear_model.')}
    m2 = Star_Trans(cols, cols, meta['domain_col'], meta['num_domains'], domain_id_as_feature=True, att_layer_num=1,
                    domain_att_layer_num=2, att_head_num=2, merge='dot', use_linear=True, use_domain_dnn_linear=True, device='cpu')
    assert {k[len('model2/state/'):] for k in z.files if k.startswith('model2/state/')} == set(m2.state_dict())
    with pytest.raises(NotImplementedError, match='att_layer_type'):
        Star_Trans(cols, cols, meta['domain_col'], meta['num_domains'], att_layer_type='trans', use_linear=False, device='cpu')
    a = AutoInt(cols, cols, device='cpu', flag='usemetatrans-metanorm', domain_column=meta['domain_col'], num_domains=meta['num_domains'])
    keys = set(a.state_dict())
    assert {'domain_embeddings.weight', 'domain_map_dnn.linears.0.weight', 'domain_map_dnn.linears.0.bias',
            'meta_net.ffn_layer_norm.weight', 'meta_net.ffn_layer_norm.bias', 'dnn_linear.weight', 'out.bias',
            'int_layers.2.W_key', 'dnn.linears.1.bias'} <= keys
    assert 'meta_net.ffn_layer_norm.weight' not in AutoInt(cols, cols, device='cpu', flag='usemetatrans',
                                                           domain_column=meta['domain_col'], num_domains=3).state_dict()
    with pytest.raises(ValueError, match='domain_column'):
        AutoInt(cols, cols, device='cpu', flag='usemetatrans')
    for name, cls in (('sgd', torch.optim.SGD), ('adam', torch.optim.Adam), ('adagrad', torch.optim.Adagrad), ('rmsprop', torch.optim.RMSprop)):
        m1.compile(name, 'binary_crossentropy', metrics=['auc'])
        assert isinstance(m1.optim, cls) and m1.metrics_names == ['loss', 'auc']
    assert m1.optim.param_groups[0]['lr'] == 0.01
    with pytest.raises(NotImplementedError):
        m1.compile('adamw', 'binary_crossentropy')
    with pytest.raises(NotImplementedError):
        m1.compile('adam', 'hinge')
    X = np.zeros((4, len(cols)), dtype=np.int64)
    with pytest.raises(SatError):
        m1.predict({f: X[:, i] for i, f in enumerate(meta['fields'])})
    with pytest.raises(SatError):
        m1.train_on_batch({f: X[:, i] for i, f in enumerate(meta['fields'])}, np.zeros(4))
