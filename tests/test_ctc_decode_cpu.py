"""CTC beam-search decoding (tf.nn.ctc_beam_search_decoder as called at models.py:1627): the host routine of the
library against the Python restatement, and both against exhaustive enumeration on tiny problems.  No GPU involved:
the reference's decoder is a CPU op as well."""
import ctypes

import numpy as np

from oracle import ctc as octc


def _lib_decode(logits_tbc, seq_len, beam_width=20, merge_repeated=True, max_out=64, n_threads=2):
    from avsi_b200 import _lib
    lib = _lib.load()
    T, B, C = logits_tbc.shape
    ldl = C + 3                                             # classes sit at a column offset inside wider rows
    buf = np.zeros((T * B, ldl), np.float32)
    buf[:, 2:2 + C] = logits_tbc.reshape(T * B, C)
    seq = np.ascontiguousarray(seq_len, np.int32)
    out = np.zeros((B, max_out), np.int32)
    out_len = np.zeros(B, np.int32)
    logp = np.zeros(B, np.float32)
    rc = lib.avsi_ctc_beam_search_host(buf.ctypes.data_as(ctypes.c_void_p), T, B, ldl, 2, C, seq.ctypes.data_as(ctypes.c_void_p),
                                       beam_width, int(merge_repeated), max_out, out.ctypes.data_as(ctypes.c_void_p),
                                       out_len.ctypes.data_as(ctypes.c_void_p), logp.ctypes.data_as(ctypes.c_void_p), n_threads)
    assert rc == 0, _lib.last_error()
    return out, out_len, logp


def test_beam_search_finds_the_best_labeling_on_tiny_problems():
    rng = np.random.default_rng(0)
    for trial in range(12):
        T, C = int(rng.integers(2, 6)), int(rng.integers(2, 4))
        lg = rng.standard_normal((T, 1, C)) * 2.0
        best, best_lp, tot = octc.ctc_best_labeling_bruteforce(lg[:, 0])
        ref, ref_lp = octc.ctc_beam_search(lg[:, 0], beam_width=1000, merge_repeated=False)
        assert ref == best and abs(ref_lp - best_lp) < 1e-9
        out, n, lp = _lib_decode(lg.astype(np.float32), [T], beam_width=1000, merge_repeated=False)
        assert out[0, :n[0]].tolist() == best and abs(lp[0] - best_lp) < 1e-4
        assert np.all(out[0, n[0]:] == -1)


def test_beam_search_matches_restatement_grid_shape():
    """GRID-like shapes: C = 34 (blank 33), beam 20, ragged sequence lengths, peaky logits from a label sequence."""
    rng = np.random.default_rng(1)
    T, B, C = 60, 5, 34
    lg = rng.standard_normal((T, B, C)).astype(np.float32)
    for b in range(B):
        labs = rng.integers(0, 33, 12)
        labs[3] = labs[2]                                   # a genuine repeat: merge_repeated collapses it in the output
        for i, c in enumerate(labs):
            lg[4 * i + 1:4 * i + 3, b, c] += 6.0
            lg[4 * i + 3:4 * i + 5, b, 33] += 4.0
    seq = np.array([60, 55, 60, 13, 1], np.int32)
    for merge in (True, False):
        out, n, lp = _lib_decode(lg, seq, beam_width=20, merge_repeated=merge)
        for b in range(B):
            ref, ref_lp = octc.ctc_beam_search(lg[:seq[b], b], beam_width=20, merge_repeated=merge)
            assert out[b, :n[b]].tolist() == ref, (b, merge)
            assert abs(lp[b] - ref_lp) < 1e-3 * max(1.0, abs(ref_lp))
    merged = _lib_decode(lg, seq, merge_repeated=True)[1]
    plain = _lib_decode(lg, seq, merge_repeated=False)[1]
    assert merged[0] < plain[0]                              # the repeated label was collapsed


def test_beam_search_truncates_and_reports_full_length():
    rng = np.random.default_rng(2)
    lg = rng.standard_normal((30, 1, 5)).astype(np.float32) * 3
    full, n, _ = _lib_decode(lg, [30], max_out=64, merge_repeated=False)
    cut, n2, _ = _lib_decode(lg, [30], max_out=3, merge_repeated=False)
    assert n2[0] == n[0] and cut[0].tolist() == full[0, :3].tolist()
