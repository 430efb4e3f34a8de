"""Host-side logic that needs no GPU: parameter layouts, config parsing, mask drawing, C-ABI symbols."""
import ctypes
import json
import os
import random
import re

import numpy as np
import pytest

from conftest import ROOT


def test_layout_roundtrip_and_padding():
    from avsi_b200.layout import HP, ParamLayout, init_canonical
    for in_dim, ncls in ((393, 0), (257, 34), (136, 0)):
        L = ParamLayout(in_dim, 250, 3, 257, ncls)
        params = init_canonical(L, seed=2, bias_scale=0.1)
        flat = L.pack(params, np.float64)
        back = L.unpack(flat)
        assert set(back) == set(params)
        for k in params:
            assert np.array_equal(back[k], params[k]), k
        assert L.n_params() == sum(v.size for v in params.values())
        assert int(L.pad_mask().sum()) <= L.n_params()
        # spot-check the gate-interleaved layout: column n = dir*1024 + unit*4 + gate
        wih = L.view(flat, 'wih0')
        k = params['cudnn_lstm/stack_bidirectional_rnn/cell_0/bidirectional_rnn/bw/cudnn_compatible_lstm_cell/kernel']
        assert wih[1024 + 17 * 4 + 2, 5] == k[5, 2 * 250 + 17]
        whh = L.view(flat, 'whh1')
        k1 = params['cudnn_lstm/stack_bidirectional_rnn/cell_1/bidirectional_rnn/fw/cudnn_compatible_lstm_cell/kernel']
        assert whh[9 * 4 + 3, 100] == k1[500 + 100, 3 * 250 + 9]
        wih1 = L.view(flat, 'wih1')
        assert wih1[9 * 4 + 1, HP + 3] == k1[250 + 3, 250 + 9]       # bw half of the layer input lives at 256+
        assert np.all(wih1[:, 250:256] == 0) and np.all(wih1[250 * 4:1024] == 0)
    assert ParamLayout(393, 250, 3, 257, 0).n_params() == 4420757     # AV-SI parameter count
    assert ParamLayout(257, 250, 3, 257, 0).n_params() == 4148757
    with pytest.raises(ValueError):
        L.pack({k: v[..., :1] for k, v in params.items()})


def test_config_parser_matches_reference(golden_dir, tmp_path):
    from avsi_b200 import config_utils
    confs = json.load(open(os.path.join(golden_dir, 'configs.json')))
    for name, c in confs.items():
        p = tmp_path / (name + '.config')
        p.write_text(c['text'])
        parsed = config_utils.load_configfile(str(p))
        assert parsed == c['parsed'], name
        if '__error__' in c['checked']:
            with pytest.raises(ValueError):
                config_utils.check_trainconfiguration(dict(parsed))
            continue
        checked = config_utils.check_trainconfiguration(dict(parsed))
        ref = dict(c['checked'])
        ref.setdefault('ctc_loss', 1)        # documented deviation: the reference never defaults this key
        assert checked == ref, name
    bad = tmp_path / 'bad.config'
    bad.write_text('model = a blstm\n')
    with pytest.raises(ValueError):
        config_utils.load_configfile(str(bad))
    with pytest.raises(ValueError):
        config_utils.load_configfile(str(tmp_path / 'missing.config'))


def test_draw_intrusions_matches_reference_function(golden_dir):
    from avsi_b200.dataset_generator import draw_intrusions
    cases = json.load(open(os.path.join(golden_dir, 'maskgen_cases.json')))
    gold = np.load(os.path.join(golden_dir, 'maskgen.npz'))
    last = None
    for c in cases:
        key = (c['seed'], c['n_max'], c['mean'], c['std'])
        if key != last:
            random.seed(c['seed'])
            last = key
        iv, cov, n_intr = draw_intrusions(250, c['mean'], c['std'], c['n_max'])
        col = np.ones(250, np.uint8)
        for o, l in iv:
            col[o:o + l] = 0
        assert n_intr == c['n_intr'] and cov == c['cov']
        assert np.array_equal(col, np.unpackbits(gold[c['key']])[:250])


def test_library_exports_every_declared_symbol():
    import __graft_entry__
    __graft_entry__.build()
    from avsi_b200 import _lib
    header = open(os.path.join(ROOT, 'include', 'avsi_b200.h')).read()
    declared = set(re.findall(r'\b(avsi_[a-z0-9_]+)\s*\(', header))
    assert declared, 'no declarations found'
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    loaded = _lib.load()
    assert b'sm_100a' in loaded.avsi_version()
    assert loaded.avsi_sizeof_frontend_args() == ctypes.sizeof(_lib.FrontendArgs)
    assert loaded.avsi_sizeof_istft_args() == ctypes.sizeof(_lib.IstftArgs)


def test_ctypes_signatures_match_the_header():
    """Every prototype of include/avsi_b200.h has as many parameters, of the same kind (pointer / integer / float), as
    its ctypes mirror in _lib.SIGNATURES: an ABI drift between the two would corrupt arguments silently."""
    from avsi_b200 import _lib
    header = open(os.path.join(ROOT, 'include', 'avsi_b200.h')).read()
    header = re.sub(r'/\*.*?\*/', ' ', header, flags=re.S)
    protos = re.findall(r'\b[A-Za-z_][A-Za-z0-9_ ]*?[ *]+(avsi_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;', header)
    assert len(protos) == len(_lib.SIGNATURES), (len(protos), len(_lib.SIGNATURES))

    def kind_c(param):
        param = param.strip()
        if '*' in param:
            return 'ptr'
        base = param.rsplit(' ', 1)[0].strip() if ' ' in param else param
        if base in ('float', 'double'):
            return base
        return 'int64' if ('64' in base) else 'int32'

    def kind_py(t):
        name = getattr(t, '__name__', str(t))
        if name in ('c_void_p', 'c_char_p') or name.startswith('LP_'):
            return 'ptr'
        return {'c_float': 'float', 'c_double': 'double', 'c_int': 'int32', 'c_int32': 'int32', 'c_uint32': 'int32', 'c_uint': 'int32',
                'c_int64': 'int64', 'c_uint64': 'int64', 'c_long': 'int64', 'c_ulong': 'int64'}[name]
    for name, params in protos:
        params = params.strip()
        c_kinds = [] if params in ('', 'void') else [kind_c(x) for x in params.split(',')]
        py_kinds = [kind_py(t) for t in _lib.SIGNATURES[name][1]]
        assert c_kinds == py_kinds, (name, c_kinds, py_kinds)


def test_product_path_fails_loudly_without_gpu():
    import torch
    from avsi_b200 import _lib
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from avsi_b200 import audio_processing, blstm
    with pytest.raises(_lib.AvsiError):
        blstm.BLSTMEngine(393)
    with pytest.raises(_lib.AvsiError):
        audio_processing.fused_features(torch.zeros(1, 4800), 384, 192)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, 'audio-visual-speech-inpainting_b200')
    for fn in os.listdir(pkg):
        if fn.endswith('.py'):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r'^\s*(from|import)\s+oracle\b', src, re.M), fn
