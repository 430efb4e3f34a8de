"""Checkpoints under the reference's variable names.

tf.train.Saver() in the reference writes every global variable: the model variables in canonical layout, the
unnamed global step `Variable`, and the Adam slots `<var>/Adam`, `<var>/Adam_1`, `beta1_power`, `beta2_power`
(training.py:114,267,335).  Two containers hold the same name -> array mapping: one `.npz` per checkpoint
(`netmodel/ckpt.npz`, `netmodel/sinet.npz`; '/' in names kept as-is via a name table) and TensorFlow's own tensor
bundle (`sinet.index` + `sinet.data-00000-of-00001`, tf_bundle.py) so that checkpoints move between the two
implementations in either direction.  `load` / `restore` take either; `save(..., fmt='tf')` writes the bundle."""
import json
import os

import numpy as np


def _adam_slots(model):
    eng = model.engine
    L = eng.layout
    m = L.unpack(eng.adam_m.detach().cpu().numpy())
    v = L.unpack(eng.adam_v.detach().cpu().numpy())
    out = {}
    for k in m:
        out[model._scoped(k) + '/Adam'] = m[k]
        out[model._scoped(k) + '/Adam_1'] = v[k]
    t = eng.step_count
    out['beta1_power'] = np.asarray(0.9 ** t, np.float32)
    out['beta2_power'] = np.asarray(0.999 ** t, np.float32)
    out['__adam_step__'] = np.asarray(t, np.int64)
    return out


def save(model, path, with_optimizer=True, fmt='npz'):
    """Write `<path>.npz`, or with fmt='tf' the tensor bundle `<path>.index` / `<path>.data-00000-of-00001`
    (path as given to saver.save in the reference, e.g. .../netmodel/sinet)."""
    variables = dict(model.all_vars)
    if with_optimizer and model.optimizer_choice == 'adam':
        variables.update(_adam_slots(model))
    if fmt == 'tf':
        from . import tf_bundle
        variables.pop('__adam_step__', None)                 # not a TF variable: recovered from beta1_power
        return tf_bundle.write_bundle(path, {k: np.asarray(v) for k, v in variables.items()})
    names = sorted(variables)
    arrays = {'v%05d' % i: np.asarray(variables[n]) for i, n in enumerate(names)}
    arrays['__names__'] = np.frombuffer(json.dumps(names).encode(), np.uint8)
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    np.savez(path + '.npz', **arrays)
    return path


def load(path):
    """`<path>.npz` -> {tf variable name: array}.  Raises ValueError like saver.restore on a bad checkpoint."""
    if os.path.exists(path + '.index'):
        from . import tf_bundle
        return tf_bundle.read_bundle(path)
    f = path if path.endswith('.npz') else path + '.npz'
    if not os.path.exists(f):
        raise ValueError('%s is not a valid checkpoint' % path)
    z = np.load(f)
    names = json.loads(bytes(z['__names__']).decode())
    return {n: z['v%05d' % i] for i, n in enumerate(names)}


def restore(model, path, train_vars_only=False):
    """saver.restore(sess, path): model variables (+ global step and Adam slots unless train_vars_only)."""
    import torch
    variables = load(path)
    model.assign_vars(variables)
    if train_vars_only:
        return model
    if '__adam_step__' not in variables:
        if 'beta1_power' not in variables:
            return model
        # a TensorFlow bundle: Adam's step count is beta1_power = 0.9 ** t
        variables['__adam_step__'] = int(round(np.log(max(float(variables['beta1_power']), 1e-300)) / np.log(0.9)))
    eng = model.engine
    want = eng.layout.canonical_shapes()
    m = {k: variables[model._scoped(k) + '/Adam'] for k in want}
    v = {k: variables[model._scoped(k) + '/Adam_1'] for k in want}
    eng.adam_m.copy_(torch.from_numpy(eng.layout.pack(m, np.float32)))
    eng.adam_v.copy_(torch.from_numpy(eng.layout.pack(v, np.float32)))
    eng.step_count = int(variables['__adam_step__'])
    return model
