"""ctypes binding of libavsi_b200.so (the C ABI declared in include/avsi_b200.h).

There is NO fallback: if the library is missing or a call fails, an exception is raised.
Pointers are raw device addresses taken from torch tensors (``tensor.data_ptr()``); torch is
only the allocator / stream provider here.
"""
import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_double, c_float, c_int, c_int64, c_uint32, c_uint64, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, 'libavsi_b200.so')


class AvsiError(RuntimeError):
    pass


class FrontendArgs(Structure):
    _fields_ = [
        ('wav', c_void_p), ('B', c_int), ('N', c_int),
        ('frame_len', c_int), ('hop', c_int), ('nfft', c_int), ('T', c_int), ('F', c_int),
        ('window', c_void_p), ('twiddle', c_void_p),
        ('mean', c_void_p), ('stdev', c_void_p), ('mask', c_void_p),
        ('video', c_void_p), ('V', c_int),
        ('power', c_float), ('log_flag', c_int), ('stft_masked', c_int),
        ('stft_out', c_void_p), ('spec_out', c_void_p), ('feat_out', c_void_p),
        ('xh_out', c_void_p), ('ldx', c_int),
        ('logmel_out', c_void_p), ('mel_w', c_void_p), ('n_mel', c_int), ('mel_eps', c_float),
        ('hole_count', c_void_p), ('xh_video_only', c_int), ('mel_masked', c_int), ('xh_skip_pad', c_int),
        ('mel_bands', c_void_p),
    ]


class IstftArgs(Structure):
    _fields_ = [
        ('mag', c_void_p), ('phase_src', c_void_p), ('mask', c_void_p),
        ('mean', c_void_p), ('stdev', c_void_p),
        ('inv_window', c_void_p), ('twiddle', c_void_p),
        ('B', c_int), ('T', c_int), ('F', c_int), ('frame_len', c_int), ('hop', c_int), ('nfft', c_int),
        ('num_samples', c_int),
        ('out', c_void_p),
    ]


# name -> (restype, argtypes); every symbol include/avsi_b200.h declares
SIGNATURES = {
    'avsi_last_error': (c_char_p, []),
    'avsi_version': (c_char_p, []),
    'avsi_launch_count': (c_int64, []),
    'avsi_sizeof_frontend_args': (c_int, []),
    'avsi_sizeof_istft_args': (c_int, []),
    'avsi_frontend_fwd': (c_int, [POINTER(FrontendArgs), c_void_p]),
    'avsi_istft_fwd': (c_int, [POINTER(IstftArgs), c_void_p]),
    'avsi_features_to_x0': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int,
                                    c_void_p]),
    'avsi_cast_pad_f16': (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_int64, c_void_p]),
    'avsi_leaky_relu': (c_int, [c_void_p, c_int64, c_float, c_void_p, c_void_p]),
    'avsi_leaky_relu_bwd': (c_int, [c_void_p, c_void_p, c_int64, c_float, c_void_p, c_void_p]),
    'avsi_masked_time_mean': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    'avsi_masked_time_mean_bwd': (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    'avsi_time_sum': (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    'avsi_tile_embedding': (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_int, c_void_p]),
    'avsi_video_features': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    'avsi_expand_mask': (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    'avsi_ctc_beam_search_host': (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int, c_int,
                                          c_void_p, c_void_p, c_void_p, c_int]),
    'avsi_parse_av_sample_host': (c_int, [c_void_p, c_uint64, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64,
                                          c_void_p, c_int64, c_void_p, c_int, c_void_p]),
    'avsi_crc32c_host': (c_uint32, [c_void_p, c_uint64, c_uint32]),
    'avsi_spectrogram': (c_int, [c_void_p, c_int64, c_float, c_int, c_void_p, c_void_p]),
    'avsi_log_mel': (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int, c_float, c_void_p, c_void_p]),
    'avsi_preemphasis': (c_int, [c_void_p, c_int, c_int, c_float, c_void_p, c_void_p]),
    'avsi_mfcc': (c_int, [c_void_p, c_int64, c_int, c_int, c_void_p, c_void_p]),
    'avsi_resample_workspace_bytes': (c_int64, [c_int, c_int64, c_int64]),
    'avsi_resample_fft': (c_int, [c_void_p, c_int, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_void_p]),
    'avsi_select_c64': (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p]),
    'avsi_delta_features': (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    'avsi_dropout_f16': (c_int, [c_void_p, c_int, c_void_p, c_int, c_int64, c_int, c_float, c_uint64, c_uint64, c_void_p,
                                 c_int, c_void_p]),
    'avsi_feature_stats': (c_int, [c_void_p, c_int, c_void_p, c_int, c_int64, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    'avsi_gemm_f16': (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_void_p,
                              c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    'avsi_lstm_fwd': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    'avsi_lstm_bwd_scratch_bytes': (c_int64, [c_int]),
    'avsi_set_reduce_scratch': (c_int, [c_void_p, c_int64, c_void_p]),
    'avsi_reduce_scratch_min_bytes': (c_int64, []),
    'avsi_gemm_f16_scratch_bytes': (c_int64, [c_int, c_int, c_int, c_int, c_int, c_int]),
    'avsi_lstm_bwd': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    'avsi_masked_l1': (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                               c_float, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    'avsi_head_l1': (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_int,
                             c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float, c_void_p, c_void_p, c_void_p,
                             c_void_p, c_int, c_void_p]),
    'avsi_head_l1_workspace_bytes': (c_int, []),
    'avsi_mtl_scales': (c_int, [c_void_p, c_int, c_float, c_void_p, c_void_p, c_void_p]),
    'avsi_colsum_f16': (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    'avsi_ctc_workspace_bytes': (c_int64, [c_int, c_int, c_int]),
    'avsi_ctc_loss': (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int,
                              c_float, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    'avsi_adam_tf': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_double, c_double, c_double, c_double,
                             c_int, c_float, c_void_p, c_float, c_void_p, c_void_p]),
    'avsi_sgd_momentum': (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_double, c_double, c_float, c_void_p, c_float,
                                  c_void_p, c_void_p]),
    'avsi_grad_guard_init': (c_int, [c_void_p, c_void_p]),
    'avsi_grad_guard_check': (c_int, [c_void_p, c_int64, c_void_p, c_void_p]),
    'avsi_grad_guard_update': (c_int, [c_void_p, c_int, c_void_p]),
    'avsi_reload_env': (None, []),
    'avsi_debug_lstm4_timing': (c_int, [c_void_p]),
    'avsi_debug_lstm4_bwd_timing': (c_int, [c_void_p]),
    'avsi_cast_to_f32': (c_int, [c_void_p, c_int, c_int64, c_void_p, c_void_p]),
    'avsi_cast_weights': (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p]),
    'avsi_gate_bias_prescale': (c_int, [c_void_p, c_int, c_void_p, c_void_p]),
}

_lib = None


def load():
    """Load the shared library (once).  Raises AvsiError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise AvsiError('libavsi_b200.so not found at %s -- run `python -c "import __graft_entry__ as g; g.build()"` '
                        '(there is no CPU fallback)' % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)      # AttributeError if a declared symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if lib.avsi_sizeof_frontend_args() != ctypes.sizeof(FrontendArgs) or \
            lib.avsi_sizeof_istft_args() != ctypes.sizeof(IstftArgs):
        raise AvsiError('ctypes struct mirrors do not match include/avsi_b200.h (rebuild the library)')
    _lib = lib
    return lib


def set_env(**kw):
    """Set / clear (value None) AVSI_* switches of the library inside this process and make it re-read them."""
    for k, v in kw.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = str(v)
    load().avsi_reload_env()


def check(rc, what=''):
    if rc != 0:
        msg = load().avsi_last_error().decode(errors='replace')
        raise AvsiError('%s failed (code %d): %s' % (what, rc, msg))


def ptr(t):
    """Device pointer of a torch tensor (None -> NULL)."""
    if t is None:
        return None
    return c_void_p(t.data_ptr())


def stream_ptr():
    import torch
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def launch_count():
    return int(load().avsi_launch_count())


# ---- optional per-kernel CUDA-event timing (bench.py); zero cost when disabled -----------------
_prof = None


class _Span(object):
    __slots__ = ('name', 'bytes', 'flops', 'start', 'stop')

    def __init__(self, name, nbytes, flops):
        self.name, self.bytes, self.flops = name, nbytes, flops

    def __enter__(self):
        import torch
        self.start = torch.cuda.Event(enable_timing=True)
        self.stop = torch.cuda.Event(enable_timing=True)
        self.start.record()
        return self

    def __exit__(self, *exc):
        self.stop.record()
        _prof.append(self)
        return False


class _NullSpan(object):
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


_NULL = _NullSpan()


class _NvtxSpan(object):
    """AVSI_NVTX=1: every span is an NVTX range (`ncu --nvtx --nvtx-include "lstm_bwd/"` then selects a stage of the step by
    name instead of by kernel regex and launch index)."""
    __slots__ = ('name',)

    def __init__(self, name):
        self.name = name

    def __enter__(self):
        import torch
        torch.cuda.nvtx.range_push(self.name)
        return self

    def __exit__(self, *exc):
        import torch
        torch.cuda.nvtx.range_pop()
        return False


_NVTX = os.environ.get('AVSI_NVTX', '0') == '1'


def span(name, nbytes=0, flops=0):
    """``with span('lstm_fwd'):`` brackets kernel launches with CUDA events on the current stream
    while profiling is enabled (profile_start / profile_stop)."""
    if _prof is None:
        return _NvtxSpan(name) if _NVTX else _NULL
    return _Span(name, nbytes, flops)


def profile_start():
    global _prof
    _prof = []


def profile_stop():
    """Returns {name: dict(ms, launches, bytes, flops)} summed over the recorded spans."""
    global _prof
    import torch
    torch.cuda.synchronize()
    spans, _prof = _prof, None
    out = {}
    for s in spans or []:
        d = out.setdefault(s.name, {'ms': 0.0, 'launches': 0, 'bytes': 0, 'flops': 0})
        d['ms'] += s.start.elapsed_time(s.stop)
        d['launches'] += 1
        d['bytes'] += s.bytes
        d['flops'] += s.flops
    return out
