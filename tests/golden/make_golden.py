"""Generate the committed golden vectors from the reference checkout.

Run HERE (build container) only:  python tests/golden/make_golden.py
/root/reference does not exist on the GPU box, so nothing else may read it at test time.

What can be taken from the reference itself:
  * docs/files/{800ms,1600ms}/ex{1,2}/{target,masked}.wav -- the only known-answer fixtures the
    reference ships; stored as int16 arrays with the zeroed frame ranges.
  * get_intrusions_mask (dataset_generator.py:11-48), get_motion_vector (face_landmarks.py:30-39):
    pure Python/numpy function bodies, extracted with ``ast`` (their modules import pydub / dlib,
    which are absent) and executed on seeded inputs.
  * load_configfile / check_trainconfiguration (config_utils.py) on the shipped config files.
TensorFlow is absent, so no TF op output can be recorded (parity of BLSTM/CTC/Adam/mel is
unpinned by the reference, see oracle/__init__.py).
"""
import ast
import contextlib
import io
import json
import os
import random
import sys

import numpy as np
from scipy.io import wavfile

REF = '/root/reference'
OUT = os.path.dirname(os.path.abspath(__file__))


def extract_function(path, name, glob):
    src = open(path).read()
    tree = ast.parse(src)
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name == name:
            mod = ast.Module(body=[node], type_ignores=[])
            exec(compile(mod, path, 'exec'), glob)
            return glob[name]
    raise KeyError(name)


def main():
    pkg = os.path.join(REF, 'av_speech_inpainting')
    # 1. docs fixtures
    ranges = {'800ms/ex1': (77, 144), '800ms/ex2': (35, 102), '1600ms/ex1': (24, 157), '1600ms/ex2': (19, 152)}
    fx = {}
    for k, (a, b) in ranges.items():
        _, t = wavfile.read(os.path.join(REF, 'docs/files', k, 'target.wav'))
        _, m = wavfile.read(os.path.join(REF, 'docs/files', k, 'masked.wav'))
        key = k.replace('/', '_')
        fx[key + '_target'] = t.astype(np.int16)
        fx[key + '_masked'] = m.astype(np.int16)
        fx[key + '_range'] = np.array([a, b], np.int32)
    np.savez_compressed(os.path.join(OUT, 'docs_fixtures.npz'), **fx)

    # 2. get_intrusions_mask on seeded draws
    g = {'random': random, 'np': np}
    gim = extract_function(os.path.join(pkg, 'dataset_generator.py'), 'get_intrusions_mask', g)
    cases = []
    masks = {}
    idx = 0
    for seed in (30, 1, 2, 3, 7, 11, 12345):
        for (n_max, mean, std) in ((1, 0.27, 0.1), (3, 0.3, 0.15), (5, 0.5, 0.2), (8, 0.2, 0.3)):
            random.seed(seed)
            for rep in range(3):
                mask, cov, n_intr = gim(257, 250, mean, std, n_max)
                cases.append({'seed': seed, 'rep': rep, 'n_max': n_max, 'mean': mean, 'std': std,
                              'cov': cov, 'n_intr': n_intr, 'key': 'm%d' % idx})
                masks['m%d' % idx] = np.packbits(mask[:, 0].astype(np.uint8))
                idx += 1
    np.savez_compressed(os.path.join(OUT, 'maskgen.npz'), **masks)
    json.dump(cases, open(os.path.join(OUT, 'maskgen_cases.json'), 'w'), indent=0)

    # 3. get_motion_vector
    g2 = {'np': np}
    gmv = extract_function(os.path.join(pkg, 'face_landmarks.py'), 'get_motion_vector', g2)
    rng = np.random.default_rng(5)
    lm = np.round(rng.uniform(50, 300, (40, 136)))
    np.savez_compressed(os.path.join(OUT, 'motion_vector.npz'), landmarks=lm, delta1=gmv(lm, delta=1), delta2=gmv(lm, delta=2))

    # 4. config parsing
    sys.path.insert(0, pkg)
    import config_utils as ref_cu
    confs = {}
    for name in ('blstm', 'blstm_ctc', 'blstm_asr', 'unet'):
        path = os.path.join(REF, 'scripts/config', name + '.config')
        text = open(path).read()
        parsed = ref_cu.load_configfile(path)
        try:
            with contextlib.redirect_stderr(io.StringIO()):
                checked = ref_cu.check_trainconfiguration(dict(parsed))
        except ValueError as e:           # e.g. unet.config lacks audio_feat_mean
            checked = {'__error__': str(e.args[0])}
        confs[name] = {'text': text, 'parsed': parsed, 'checked': checked}
    json.dump(confs, open(os.path.join(OUT, 'configs.json'), 'w'), indent=0)
    print('golden vectors written to', OUT)


if __name__ == '__main__':
    main()
