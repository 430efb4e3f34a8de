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


B1, B2 = 0.9, 0.999


def _opt_names(model, key):
    """TF-1 names of the optimiser's variables (UNVERIFIED against a TF-written file -- TensorFlow cannot run here).
    The reference calls optimizer.minimize inside tf.variable_scope(<model>) + tf.name_scope('optimizer')
    (training_ctc.py:85, models.py:164-178): slot variables are created under variable_scope(None, primary.op.name + '/Adam')
    nested in the active variable scope -> '<model>/<model>/<var>/Adam'; the non-slot accumulators are plain variables
    under the name scope -> '<model>/optimizer/beta1_power'.  `restore` does not depend on these guesses: it matches by suffix."""
    scope = model.var_scope
    if key in ('beta1_power', 'beta2_power'):
        return (scope + '/optimizer/' + key) if scope else key
    return (scope + '/' + key) if scope else key


def _optimizer_slots(model):
    if model.optimizer_choice not in ('adam', 'momentum'):
        return {}
    eng = model.engine
    L = eng.layout
    out = {}
    t = eng.step_count
    if model.optimizer_choice == 'adam':
        m = L.unpack(eng.adam_m.detach().cpu().numpy())
        v = L.unpack(eng.adam_v.detach().cpu().numpy())
        for k in m:
            out[_opt_names(model, model._scoped(k) + '/Adam')] = m[k]
            out[_opt_names(model, model._scoped(k) + '/Adam_1')] = v[k]
        # TF initialises beta1_power to beta1 and multiplies it AFTER each step: after t steps it holds beta1^(t+1)
        # (float32, as TF stores it; it underflows for t >~ 980 / 87000 -- `restore` then falls back on the global step)
        out[_opt_names(model, 'beta1_power')] = np.asarray(B1 ** (t + 1), np.float32)
        out[_opt_names(model, 'beta2_power')] = np.asarray(B2 ** (t + 1), np.float32)
    elif model.optimizer_choice == 'momentum' and hasattr(eng, 'momentum_acc'):
        acc = L.unpack(eng.momentum_acc.detach().cpu().numpy())
        for k in acc:
            out[_opt_names(model, model._scoped(k) + '/Momentum')] = acc[k]
    out['__adam_step__'] = np.asarray(t, np.int64)
    return out


def save(model, path, with_optimizer=True, fmt='npz'):
    """Write `<path>.npz`, or with fmt='tf' the tensor bundle `<path>.index` / `<path>.data-00000-of-00001`
    (path as given to saver.save in the reference, e.g. .../netmodel/sinet)."""
    variables = dict(model.all_vars)
    if with_optimizer:
        variables.update(_optimizer_slots(model))
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


def _find_suffix(variables, suffix):
    """The value whose key is `suffix` or ends with '/' + suffix (shortest such key), else None."""
    if suffix in variables:
        return variables[suffix]
    hits = sorted((k for k in variables if k.endswith('/' + suffix)), key=len)
    return variables[hits[0]] if hits else None


def _adam_step_from(variables, model):
    """Number of Adam updates behind a checkpoint: our own counter, else TF's beta1_power = beta1^(t+1), else -- when that
    float32 has underflowed or is missing -- the global step `Variable` (one minimize() per step, models.py:178)."""
    if '__adam_step__' in variables:
        return int(variables['__adam_step__'])
    b1p = _find_suffix(variables, 'beta1_power')
    if b1p is not None and 1e-30 < float(b1p) <= B1 * (1 + 1e-6):
        return max(0, int(round(np.log(float(b1p)) / np.log(B1))) - 1)
    return int(model.global_step)


def restore(model, path, train_vars_only=False):
    """saver.restore(sess, path): model variables (+ global step and optimiser slots unless train_vars_only).

    Optimiser slots are matched by suffix ('<anything>/<scoped var>/Adam').  A checkpoint without per-variable slots
    -- e.g. one written by the reference's TRAINING graph, whose CudnnLSTM saveable keeps the weights canonical but whose
    Adam slots belong to the opaque cuDNN blob ('cudnn_lstm/opaque_kernel/Adam') -- restores the weights and the global
    step only, with a warning, and the moments restart from zero.  Anything else that is missing raises ValueError
    (what the drivers catch, training.py:154-166)."""
    import torch
    variables = load(path)
    try:
        model.assign_vars(variables)
    except KeyError as e:
        raise ValueError('%s: variable %s is missing from the checkpoint' % (path, e))
    if train_vars_only:
        return model
    eng = model.engine
    want = eng.layout.canonical_shapes()
    eng.step_count = _adam_step_from(variables, model)
    for slot, target in (('Adam', 'adam_m'), ('Adam_1', 'adam_v'), ('Momentum', 'momentum_acc')):
        found = {k: _find_suffix(variables, model._scoped(k) + '/' + slot) for k in want}
        have = [k for k, v in found.items() if v is not None]
        if not have:
            if slot != 'Momentum' and model.optimizer_choice == 'adam' and any(k.endswith('/' + slot) for k in variables):
                print('WARNING: %s holds no per-variable %s slots (cuDNN opaque-kernel slots?): optimiser moments restart from zero'
                      % (path, slot))
            continue
        if len(have) != len(want):
            raise ValueError('%s: %s slots are incomplete (%d of %d variables)' % (path, slot, len(have), len(want)))
        flat = torch.from_numpy(eng.layout.pack(found, np.float32))
        if target == 'momentum_acc' and not hasattr(eng, 'momentum_acc'):
            eng.momentum_acc = torch.zeros_like(eng.theta)
        getattr(eng, target).copy_(flat)
    return model
