"""Oracle (test infrastructure): CTC negative log-likelihood and gradient.

Restates tf.nn.ctc_loss as called at models.py:1950-1953 /
models_asr.py:146-148: inputs are UNNORMALISED logits [T,B,C] (softmax inside
the op), time major, blank = C-1, ctc_merge_repeated=True,
preprocess_collapse_repeated=False, frames t >= sequence_length[b] ignored
(zero gradient).  Labels come dense [B,Lmax] + lengths
(ctc_label_dense_to_sparse, models.py:1760).

Three independent forms, cross-checked in tests/test_oracle_ctc.py (and
against torch.nn.functional.ctc_loss):
  ctc_nll_torch   alpha recursion in torch (autograd gives the gradient)
  ctc_alpha_beta  numpy log-space alpha/beta with the analytic gradient
                  softmax - posterior  (the algorithm the CUDA kernel uses)
  ctc_bruteforce  enumeration of all alignments, tiny cases only
"""
import itertools

import numpy as np
import torch

NEG_INF = -1e30


def _extended(labels, blank):
    ext = [blank]
    for l in labels:
        ext += [int(l), blank]
    return ext


def ctc_nll_torch(logits_tbc, labels, lab_len, seq_len, blank=None):
    """logits_tbc torch [T,B,C]; labels [B,Lmax] int; returns nll [B] (differentiable)."""
    T, B, C = logits_tbc.shape
    blank = C - 1 if blank is None else blank
    logp = torch.log_softmax(logits_tbc, dim=2)
    out = []
    for b in range(B):
        L = int(lab_len[b])
        Tb = int(seq_len[b])
        ext = _extended([int(v) for v in labels[b][:L]], blank)
        S = len(ext)
        ext_t = torch.tensor(ext)
        can_skip = torch.zeros(S, dtype=torch.bool)
        for s in range(2, S):
            can_skip[s] = ext[s] != blank and ext[s] != ext[s - 2]
        neg = logp.new_full((S,), NEG_INF)
        alpha = neg.clone()
        alpha[0] = logp[0, b, blank]
        if S > 1:
            alpha[1] = logp[0, b, ext[1]]
        for t in range(1, Tb):
            a1 = torch.cat([neg[:1], alpha[:-1]])
            a2 = torch.where(can_skip, torch.cat([neg[:2], alpha[:-2]]), neg)
            alpha = torch.logsumexp(torch.stack([alpha, a1, a2]), dim=0) + logp[t, b, ext_t]
        tail = alpha[S - 1:] if S == 1 else alpha[S - 2:]
        out.append(-torch.logsumexp(tail, dim=0))
    return torch.stack(out)


def _lse(a, b):
    m = np.maximum(a, b)
    return m + np.log(np.exp(a - m) + np.exp(b - m))


def ctc_alpha_beta(logits_tbc, labels, lab_len, seq_len, blank=None, dtype=np.float64):
    """numpy alpha/beta.  Returns (nll [B], dlogits [T,B,C]) with d(sum_b nll_b)/dlogits."""
    logits = np.asarray(logits_tbc, dtype=dtype)
    T, B, C = logits.shape
    blank = C - 1 if blank is None else blank
    mx = logits.max(axis=2, keepdims=True)
    logp = logits - mx - np.log(np.exp(logits - mx).sum(axis=2, keepdims=True))
    nll = np.zeros(B, dtype)
    grad = np.zeros_like(logits)
    for b in range(B):
        L = int(lab_len[b])
        Tb = int(seq_len[b])
        ext = np.array(_extended([int(v) for v in np.asarray(labels[b])[:L]], blank))
        S = len(ext)
        skip = np.zeros(S, bool)
        skip[2:] = (ext[2:] != blank) & (ext[2:] != ext[:-2])
        alpha = np.full((Tb, S), NEG_INF, dtype)
        beta = np.full((Tb, S), NEG_INF, dtype)
        alpha[0, 0] = logp[0, b, blank]
        if S > 1:
            alpha[0, 1] = logp[0, b, ext[1]]
        for t in range(1, Tb):
            prev = alpha[t - 1]
            a = prev.copy()
            a[1:] = _lse(a[1:], prev[:-1])
            a[2:] = np.where(skip[2:], _lse(a[2:], prev[:-2]), a[2:])
            alpha[t] = a + logp[t, b, ext]
        beta[Tb - 1, S - 1] = logp[Tb - 1, b, ext[S - 1]]
        if S > 1:
            beta[Tb - 1, S - 2] = logp[Tb - 1, b, ext[S - 2]]
        for t in range(Tb - 2, -1, -1):
            nxt = beta[t + 1]
            a = nxt.copy()
            a[:-1] = _lse(a[:-1], nxt[1:])
            a[:-2] = np.where(skip[2:], _lse(a[:-2], nxt[2:]), a[:-2])
            beta[t] = a + logp[t, b, ext]
        ll = alpha[Tb - 1, S - 1] if S == 1 else _lse(alpha[Tb - 1, S - 1], alpha[Tb - 1, S - 2])
        nll[b] = -ll
        # posterior over classes: sum_{s: ext[s]=k} alpha*beta / y_t(k)
        post = np.full((Tb, C), NEG_INF, dtype)
        ab = alpha + beta
        for s in range(S):
            post[:, ext[s]] = _lse(post[:, ext[s]], ab[:, s])
        grad[:Tb, b, :] = np.exp(logp[:Tb, b, :]) - np.exp(post - logp[:Tb, b, :] - ll)
    return nll, grad


def _collapse(path, blank):
    out = []
    prev = None
    for p in path:
        if p != prev and p != blank:
            out.append(p)
        prev = p
    return out


def ctc_bruteforce(logits_tc, labels, blank=None):
    """Sum over every alignment of length T (tiny T, C only).  Returns nll."""
    logits = np.asarray(logits_tc, dtype=np.float64)
    T, C = logits.shape
    blank = C - 1 if blank is None else blank
    p = np.exp(logits - logits.max(1, keepdims=True))
    p /= p.sum(1, keepdims=True)
    tot = 0.0
    target = [int(v) for v in labels]
    for path in itertools.product(range(C), repeat=T):
        if _collapse(path, blank) == target:
            tot += np.prod([p[t, k] for t, k in enumerate(path)])
    return -np.log(tot)


def greedy_decode(logits_tbc, seq_len, blank=None):
    """Best-path decoding (merge repeats, drop blanks); PER monitoring helper."""
    logits = np.asarray(logits_tbc)
    T, B, C = logits.shape
    blank = C - 1 if blank is None else blank
    best = logits.argmax(2)
    return [_collapse(list(best[:int(seq_len[b]), b]), blank) for b in range(B)]


def ctc_beam_search(logits, beam_width=20, merge_repeated=True):
    """tf.nn.ctc_beam_search_decoder(inputs, sequence_length, beam_width=20, top_paths=1, merge_repeated=True) as the
    reference calls it (models.py:1627, :2027; models_asr.py:139), restated: prefix beam search over the log-softmax
    of logits [T, C], blank = C - 1, every label extended at every frame.  Returns (labels, log_prob).
    merge_repeated is TF-1's post-processing: consecutive equal labels of the emitted path are collapsed."""
    lg = np.asarray(logits, np.float64)
    T, C = lg.shape
    blank = C - 1
    lp = lg - (lg.max(1, keepdims=True) + np.log(np.exp(lg - lg.max(1, keepdims=True)).sum(1, keepdims=True)))
    beam = {(): (0.0, -np.inf)}                       # prefix -> (log p ending in blank, log p ending in its last label)
    for t in range(T):
        nxt = {}

        def add(key, pb, pnb):
            ob, onb = nxt.get(key, (-np.inf, -np.inf))
            nxt[key] = (np.logaddexp(ob, pb), np.logaddexp(onb, pnb))
        for y, (pb, pnb) in beam.items():
            tot = np.logaddexp(pb, pnb)
            add(y, tot + lp[t, blank], pnb + lp[t, y[-1]] if y else -np.inf)
            for c in range(C):
                if c == blank:
                    continue
                frm = pb if (y and y[-1] == c) else tot
                if frm > -np.inf:
                    add(y + (c,), -np.inf, frm + lp[t, c])
        items = sorted(nxt.items(), key=lambda kv: (-np.logaddexp(*kv[1]), kv[0]))
        beam = dict(items[:beam_width])
    best, (pb, pnb) = max(beam.items(), key=lambda kv: np.logaddexp(*kv[1]))
    out = [c for i, c in enumerate(best) if not (merge_repeated and i > 0 and c == best[i - 1])]
    return out, float(np.logaddexp(pb, pnb))


def ctc_best_labeling_bruteforce(logits):
    """argmax over LABELINGS of the total alignment probability, by enumerating all C**T alignments (tiny cases)."""
    import itertools
    lg = np.asarray(logits, np.float64)
    T, C = lg.shape
    blank = C - 1
    lp = lg - np.log(np.exp(lg).sum(1, keepdims=True))
    tot = {}
    for path in itertools.product(range(C), repeat=T):
        lab, prev = [], -1
        for k in path:
            if k != prev and k != blank:
                lab.append(k)
            prev = k
        p = float(sum(lp[t, k] for t, k in enumerate(path)))
        key = tuple(lab)
        tot[key] = np.logaddexp(tot.get(key, -np.inf), p)
    best = max(tot.items(), key=lambda kv: kv[1])
    return list(best[0]), float(best[1]), tot
