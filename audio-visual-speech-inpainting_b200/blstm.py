"""Stacked-BLSTM engine: weights, workspaces and the forward / backward / update schedule.

Host-side orchestration of the sm_100a kernels for the network of models.py:89-125
(CudnnLSTM(num_layers, num_units, bidirectional, linear_input) + linear head(s)) and its
gradient.  torch is used for device memory, streams and (in parallel.py) NCCL only; every
arithmetic op is a kernel from libavsi_b200.so.  No autograd: the backward schedule is explicit.

Per layer l (time-major rows r = t*B + b, fp16 activations, fp32 accumulate; G_l and C_l are stored in the
interleaved layout of common.cuh, the forward weight copies / bias carry the 1/2 of the sigmoid gates):
  forward   G_l = X_l . Wih_l^T                  tcgen05 GEMM  [T*B,Kp] x [2048,Kp]^T -> f16 (interleaved)
            (Y_l, C_l, G_l<-gates) = recur(G_l)  cluster-persistent LSTM kernel
  head      logits = Y_last . Whead^T + b        tcgen05 GEMM -> f32
  backward  G_l <- dgates = bptt(G_l, C_l, dY_l) cluster-persistent BPTT kernel
            dWih_l += G_l^T . X_l ; dWhh_l += G_l^T . Y_l(shifted)   tcgen05 GEMM (MN-major, split-K)
            dY_{l-1} = G_l . Wih_l               tcgen05 GEMM -> f16
Gradients are computed on loss-scaled fp16 activations gradients and unscaled inside Adam.
"""
import os

import numpy as np
import torch

from . import _lib
from .layout import GATES, HP, ParamLayout

NG = 2 * GATES * HP     # 2048 gate columns (both directions)
NY = 2 * HP             # 512 output columns (both directions)


def _p(t):
    return _lib.ptr(t)


def ctypes_ptr(addr):
    import ctypes
    return ctypes.c_void_p(int(addr))


A_IL, C_IL, B_IL = 1, 2, 4     # avsi_gemm_f16 layout bits: A operand / f16 output / B operand stored interleaved (include/avsi_b200.h)


# ---- bit-reproducible reductions (include/avsi_b200.h, avsi_set_reduce_scratch) ---------------------------------------
# One scratch buffer per device, registered with the library: the split-K weight-gradient GEMMs, the bias-gradient column
# sums and the loss sums then add their partial results in a fixed order instead of through floating-point atomics, and a
# training step repeated on the same inputs gives the same bits.  AVSI_DETERMINISTIC=0 keeps the atomics (A/B runs).
_REDUCE = {}          # device index -> {'buf': uint8 tensor registered with the library, 'keep': outgrown buffers, 'need': cache}


def deterministic():
    return os.environ.get('AVSI_DETERMINISTIC', '1') != '0'


def ensure_reduce_scratch(nbytes=0):
    """Register (or grow) this device's reduction scratch so that it holds at least nbytes.  Outgrown buffers are kept
    alive: kernels in flight, or a captured CUDA graph, may still address them."""
    if not deterministic():
        return
    lib = _lib.load()
    dev = torch.cuda.current_device()
    st = _REDUCE.setdefault(dev, {'buf': None, 'keep': [], 'need': {}})
    need = max(int(nbytes), int(lib.avsi_reduce_scratch_min_bytes()))
    if st['buf'] is not None and st['buf'].numel() >= need:
        return
    if torch.cuda.is_current_stream_capturing():
        raise RuntimeError('the reduction scratch cannot grow inside a CUDA-graph capture: run one eager step first')
    need = -(-need // (8 << 20)) * (8 << 20)
    buf = torch.empty(need, dtype=torch.uint8, device='cuda:%d' % dev)
    _lib.check(lib.avsi_set_reduce_scratch(_p(buf), need, _lib.stream_ptr()), 'avsi_set_reduce_scratch')
    if st['buf'] is not None:
        st['keep'].append(st['buf'])
    st['buf'] = buf


def release_reduce_scratch():
    """Unregister this device's scratch (the library falls back to atomics); tests of the atomic path use it."""
    lib = _lib.load()
    dev = torch.cuda.current_device()
    _lib.check(lib.avsi_set_reduce_scratch(None, 0, _lib.stream_ptr()), 'avsi_set_reduce_scratch')
    st = _REDUCE.pop(dev, None)
    if st is not None:
        torch.cuda.synchronize()


def gemm(A, lda, B, ldb, C, ldc, bias, M, N, K, trans, out_mode, split_k=1, tag='gemm', layout=0, nbytes=0):
    """nbytes: algorithmic operand bytes, given for the shapes that sit on the HBM roofline rather than the tensor one."""
    lib = _lib.load()
    if out_mode == 2 and deterministic():
        st = _REDUCE.get(torch.cuda.current_device())
        key = (M, N, K, trans, split_k)
        need = st['need'].get(key) if st is not None else None
        if need is None:
            need = int(lib.avsi_gemm_f16_scratch_bytes(M, N, K, trans, out_mode, split_k))
            ensure_reduce_scratch(need)
            _REDUCE[torch.cuda.current_device()]['need'][key] = need
    with _lib.span(tag, nbytes=nbytes, flops=2 * M * N * K):
        _lib.check(lib.avsi_gemm_f16(A, lda, B, ldb, C, ldc, bias, M, N, K, trans, out_mode, split_k, layout,
                                     _lib.stream_ptr()), 'avsi_gemm_f16')


def to_il(x, chunk=8):
    """Row-major [R, C] tensor -> interleaved [ceil32(R)/32, C/chunk, 32, chunk] copy (test / debug helper)."""
    R, C = x.shape
    Rp = -(-R // 32) * 32
    xp = torch.zeros(Rp, C, dtype=x.dtype, device=x.device)
    xp[:R] = x
    return xp.view(Rp // 32, 32, C // chunk, chunk).permute(0, 2, 1, 3).contiguous().view(Rp, C)


def from_il(x, R, chunk=8):
    """Inverse of to_il: interleaved storage (any shape with ceil32(R)*C elements) -> row-major [R, C]."""
    Rp = -(-R // 32) * 32
    C = x.numel() // Rp
    return x.view(Rp // 32, C // chunk, 32, chunk).permute(0, 2, 1, 3).contiguous().view(Rp, C)[:R]


def pick_split_k(m_rows, n_cols, k, target_ctas=296):
    tiles = -(-m_rows // 128) * -(-n_cols // 128)
    kb = -(-k // 64)
    return int(max(1, min(-(-target_ctas // tiles), 32, kb)))


class BLSTMEngine(object):
    def __init__(self, in_dim, hidden=250, n_layers=3, out_dim=257, n_classes=0, device='cuda', dense=()):
        if not torch.cuda.is_available():
            raise _lib.AvsiError('BLSTMEngine needs a CUDA device (there is no CPU fallback)')
        _lib.load()
        self.layout = ParamLayout(in_dim, hidden, n_layers, out_dim, n_classes, dense=dense)
        self.device = torch.device(device)
        L = self.layout
        n = L.n_params_padded
        self.theta = torch.zeros(n, dtype=torch.float32, device=self.device)
        self.grad = torch.zeros(n + 8, dtype=torch.float32, device=self.device)   # + loss scalars tail
        self.adam_m = torch.zeros(n, dtype=torch.float32, device=self.device)
        self.adam_v = torch.zeros(n, dtype=torch.float32, device=self.device)
        self.step_count = 0
        # overflow guard of the fp16 gradient path (include/avsi_b200.h: avsi_grad_guard_*): 8 device words
        self.guard = torch.zeros(8, dtype=torch.int32, device=self.device)
        _lib.check(_lib.load().avsi_grad_guard_init(_p(self.guard), _lib.stream_ptr()), 'avsi_grad_guard_init')
        self.guard_growth_interval = 2000
        # fp16 operand copies
        self.half = {}
        for l in range(L.n_layers):
            kp = L.layer_k(l)
            self.half['wih%d' % l] = torch.zeros(NG, kp, dtype=torch.float16, device=self.device)
            self.half['wihT%d' % l] = torch.zeros(kp, NG, dtype=torch.float16, device=self.device)
            self.half['whh%d' % l] = torch.zeros(NG, HP, dtype=torch.float16, device=self.device)
            self.half['whhT%d' % l] = torch.zeros(HP, NG, dtype=torch.float16, device=self.device)
        self.bias_fwd = [torch.zeros(NG, dtype=torch.float32, device=self.device) for _ in range(L.n_layers)]
        self.half['head'] = torch.zeros(L.nop, NY, dtype=torch.float16, device=self.device)
        self.half['headT'] = torch.zeros(NY, L.nop, dtype=torch.float16, device=self.device)
        for k in range(len(L.dense)):                 # dense layers: [fan_out, Kp] and its transpose [Kp, fan_out]
            fo, kp = L.index['dw%d' % k][1]
            self.half['dw%d' % k] = torch.zeros(fo, kp, dtype=torch.float16, device=self.device)
            self.half['dwT%d' % k] = torch.zeros(kp, fo, dtype=torch.float16, device=self.device)
        self._ws = {}

    # ---- parameters ---------------------------------------------------------------------------
    def view(self, buf, name):
        off, shp = self.layout.index[name]
        return buf[off:off + int(np.prod(shp))].view(*shp)

    def load_canonical(self, params):
        flat = self.layout.pack(params, np.float32)
        self.theta.copy_(torch.from_numpy(flat))
        self.refresh_half()

    def export_canonical(self):
        return self.layout.unpack(self.theta.detach().cpu().numpy())

    def export_canonical_grads(self, unscale=1.0):
        g = self.grad[:self.layout.n_params_padded].detach().cpu().numpy().astype(np.float64) * unscale
        return self.layout.unpack(g)

    def refresh_half(self):
        lib = _lib.load()
        st = _lib.stream_ptr()
        L = self.layout
        for l in range(L.n_layers):
            kp = L.layer_k(l)
            # forward copies (wih, whh) and the projection bias carry the 1/2 of the sigmoid gates; the transposed
            # copies read by the backward GEMMs / BPTT do not
            _lib.check(lib.avsi_cast_weights(_p(self.view(self.theta, 'wih%d' % l)), NG, kp, _p(self.half['wih%d' % l]),
                                             _p(self.half['wihT%d' % l]), 1, st), 'avsi_cast_weights')
            _lib.check(lib.avsi_cast_weights(_p(self.view(self.theta, 'whh%d' % l)), NG, HP, _p(self.half['whh%d' % l]),
                                             _p(self.half['whhT%d' % l]), 1, st), 'avsi_cast_weights')
            _lib.check(lib.avsi_gate_bias_prescale(_p(self.view(self.theta, 'b%d' % l)), NG, _p(self.bias_fwd[l]), st),
                       'avsi_gate_bias_prescale')
        _lib.check(lib.avsi_cast_weights(_p(self.view(self.theta, 'head_w')), L.nop, NY, _p(self.half['head']),
                                         _p(self.half['headT']), 0, st), 'avsi_cast_weights')
        for k in range(len(L.dense)):
            fo, kp = L.index['dw%d' % k][1]
            _lib.check(lib.avsi_cast_weights(_p(self.view(self.theta, 'dw%d' % k)), fo, kp, _p(self.half['dw%d' % k]),
                                             _p(self.half['dwT%d' % k]), 0, st), 'avsi_cast_weights')

    # ---- workspaces ---------------------------------------------------------------------------
    def workspace(self, T, B, training=True):
        key = (T, B, training)
        ws = self._ws.get(key)
        if ws is not None:
            return ws
        L = self.layout
        M = T * B
        dev = self.device
        ws = {'T': T, 'B': B, 'M': M}
        ws['x0'] = torch.zeros(M, L.k0p, dtype=torch.float16, device=dev)
        Mp = -(-M // 32) * 32
        # gate tensors and cell-state stash are INTERLEAVED (rows padded to 32, padding stays zero)
        ws['G'] = [torch.zeros(Mp, NG, dtype=torch.float16, device=dev) for _ in range(L.n_layers if training else 1)]
        # layer outputs sit between B leading and B trailing zero rows: h_{t-1} / h_{t+1} of the dW_hh
        # contractions are then plain row-shifted views (h_{-1} = h_T = 0)
        ws['Ybuf'] = [torch.zeros(M + 2 * B, NY, dtype=torch.float16, device=dev) for _ in range(L.n_layers)]
        ws['Y'] = [yb[B:B + M] for yb in ws['Ybuf']]
        # AVSI_Y_IL=1 (B a multiple of 32): the layer outputs are stored INTERLEAVED like G (the zero frames and the row
        # shifts by B are then whole 32-row blocks); the recurrence kernel stores 512-byte runs instead of 32 scattered
        # sectors per warp and the GEMMs read Y through their interleaved A / B operand paths.  Measured on the same
        # box (profiles/README.md): forward recurrence -0.10 ms per step, but the un-swizzled operand tiles cost the
        # dW / projection GEMMs 0.30 ms -- off by default
        ws['y_il'] = 1 if (B % 32 == 0 and os.environ.get('AVSI_Y_IL', '0') == '1') else 0
        ws['C'] = [torch.zeros(Mp, NY, dtype=torch.float32, device=dev) for _ in range(L.n_layers if training else 1)]
        ws['logits'] = torch.zeros(M, L.nop, dtype=torch.float32, device=dev)
        if training:
            ws['dlogits'] = torch.zeros(M, L.nop, dtype=torch.float16, device=dev)
            # dL/dy of a layer: INTERLEAVED like G (written by the dX GEMMs with C_IL, read by the BPTT kernel)
            ws['dY'] = [torch.zeros(Mp, NY, dtype=torch.float16, device=dev) for _ in range(2)]
            # per-tile bias-gradient sums of the BPTT kernels, reduced in tile order (None: atomics, AVSI_DETERMINISTIC=0)
            nbytes = int(_lib.load().avsi_lstm_bwd_scratch_bytes(B))
            ws['scratch'] = torch.empty(max(nbytes, 16) // 4, dtype=torch.float32, device=dev) if deterministic() else None
            ensure_reduce_scratch()
        if len(self._ws) > 4:
            self._ws = {}
        self._ws[key] = ws
        return ws

    # ---- forward ------------------------------------------------------------------------------
    def forward(self, ws, dropout=None, head=True):
        """ws['x0'] (time-major fp16 network input) -> ws['logits'] [T*B, nop] fp32.
        head=False stops ahead of the head GEMM and returns its input (x, interleaved flag): head_l1() takes it from there.

        dropout = (rate, seed, offset): tf.nn.dropout on the last layer's outputs ahead of the head (models.py:117);
        the mask is a function of (seed, offset) only, backward() regenerates it."""
        lib = _lib.load()
        L = self.layout
        T, B, M = ws['T'], ws['B'], ws['M']
        training = len(ws['G']) == L.n_layers
        yil = ws['y_il']
        x, ldx = ws['x0'], L.k0p
        for l in range(L.n_layers):
            G = ws['G'][l if training else 0]
            C = ws['C'][l if training else 0]
            kp = L.layer_k(l)
            gemm(_p(x), ldx, _p(self.half['wih%d' % l]), kp, _p(G), NG, None, M, NG, kp, 0, 0,
                 tag='gemm_proj_fwd', layout=C_IL | (A_IL if (l > 0 and yil) else 0))
            # HBM-bound at large batch: per row the kernel reads G (4096 B), writes the activated gates (4096 B),
            # c_t (2048 B) and h_t (1024 B); the recurrent product adds 2*M*2048*256 flops
            with _lib.span('lstm_fwd', nbytes=M * (2 * NG * 2 + NY * 4 + NY * 2), flops=2 * M * NG * HP):
                _lib.check(lib.avsi_lstm_fwd(_p(G), _p(self.half['whh%d' % l]), _p(self.bias_fwd[l]), _p(ws['Y'][l]), _p(C),
                                             T, B, yil, _lib.stream_ptr()), 'avsi_lstm_fwd')
            x, ldx = ws['Y'][l], NY
        ws['drop'] = None
        if dropout is not None and dropout[0] > 0.0:
            if 'Ydrop' not in ws:
                ws['Ydrop'] = torch.empty(M, NY, dtype=torch.float16, device=self.device)
            with _lib.span('dropout', nbytes=2 * M * NY * 2):
                _lib.check(lib.avsi_dropout_f16(_p(x), NY, _p(ws['Ydrop']), NY, M, NY, float(dropout[0]), int(dropout[1]),
                                                int(dropout[2]), None, 3 if yil else 0, _lib.stream_ptr()), 'avsi_dropout_f16')
            x = ws['Ydrop']
            ws['drop'] = (float(dropout[0]), int(dropout[1]), int(dropout[2]))
        if not head:
            return x, yil
        gemm(_p(x), NY, _p(self.half['head']), NY, _p(ws['logits']), L.nop, _p(self.view(self.theta, 'head_b')),
             M, L.n_out, NY, 0, 1, tag='gemm_head_fwd', layout=A_IL if yil else 0)
        return ws['logits']

    def head_l1(self, ws, x, yil, target, mask, seq_len, F, mode, grad_scale_dev, sums, logits_from_col):
        """Head GEMM with the masked-L1 loss and its gradient in the epilogue (avsi_head_l1): writes ws['dlogits'][:, :F],
        adds the six loss sums to `sums`, and writes ws['logits'] only from column `logits_from_col` on (what a CTC head
        still reads).  The training step's replacement for forward(head=True) + avsi_masked_l1."""
        lib = _lib.load()
        L = self.layout
        T, B, M = ws['T'], ws['B'], ws['M']
        if 'l1_partial' not in ws:
            ws['l1_partial'] = torch.zeros(int(lib.avsi_head_l1_workspace_bytes()) // 8, dtype=torch.float64, device=self.device)
        with _lib.span('head_l1', nbytes=M * (NY * 2 + 2 * F * 4 + L.nop * 2), flops=2 * M * L.n_out * NY):
            _lib.check(lib.avsi_head_l1(_p(x), NY, 1 if yil else 0, _p(self.half['head']), NY, _p(self.view(self.theta, 'head_b')),
                                        M, L.n_out, NY, _p(ws['logits']), L.nop, int(logits_from_col), _p(target), _p(mask),
                                        _p(seq_len), B, T, F, int(mode), 1.0, grad_scale_dev, _p(sums), _p(ws['l1_partial']),
                                        _p(ws['dlogits']), L.nop, _lib.stream_ptr()), 'avsi_head_l1')

    # ---- backward -----------------------------------------------------------------------------
    def backward(self, ws, zero_grad=True):
        """ws['dlogits'] (scaled dL/dlogits, fp16) -> self.grad (scaled, padded flat fp32)."""
        lib = _lib.load()
        L = self.layout
        T, B, M = ws['T'], ws['B'], ws['M']
        st = _lib.stream_ptr
        if zero_grad:
            self.grad.zero_()
        g = self.grad
        dl = ws['dlogits']
        drop = ws.get('drop')
        ylast = ws['Ydrop'] if drop else ws['Y'][L.n_layers - 1]
        # head: dW = dlogits^T . Y ; db = colsum(dlogits) ; dY = dlogits . Whead
        yil = ws['y_il']
        bil = B_IL if yil else 0
        gemm(_p(dl), L.nop, _p(ylast), NY, _p(self.view(g, 'head_w')), NY, None, L.n_out, NY, M, 1, 2,
             pick_split_k(L.n_out, NY, M), tag='gemm_dw_head', layout=bil)
        _lib.check(lib.avsi_colsum_f16(_p(dl), L.nop, M, 0, L.n_out, _p(self.view(g, 'head_b')), st()), 'avsi_colsum_f16')
        dY = ws['dY'][0]
        gemm(_p(dl), L.nop, _p(self.half['headT']), L.nop, _p(dY), NY, None, M, NY, L.nop, 0, 0, tag='gemm_dx', layout=C_IL)
        if drop:
            with _lib.span('dropout', nbytes=2 * M * NY * 2):
                _lib.check(lib.avsi_dropout_f16(_p(dY), NY, _p(dY), NY, M, NY, drop[0], drop[1], drop[2], None, 3, st()),
                           'avsi_dropout_f16')
        cur = 0
        for l in range(L.n_layers - 1, -1, -1):
            G, C, Y = ws['G'][l], ws['C'][l], ws['Y'][l]
            kp = L.layer_k(l)
            # per row: reads gates (4096 B), c (2048 B), dy (1024 B), writes dG (4096 B)
            with _lib.span('lstm_bwd', nbytes=M * (2 * NG * 2 + NY * 4 + NY * 2), flops=2 * M * NG * HP):
                _lib.check(lib.avsi_lstm_bwd(_p(G), _p(self.half['whhT%d' % l]), _p(C), _p(ws['dY'][cur]),
                                             _p(self.view(g, 'b%d' % l)), _p(ws['scratch']), T, B, st()), 'avsi_lstm_bwd')
            x, ldx = (ws['x0'], L.k0p) if l == 0 else (ws['Y'][l - 1], NY)
            # dWih = dG^T . X
            gemm(_p(G), NG, _p(x), ldx, _p(self.view(g, 'wih%d' % l)), kp, None, NG, kp, M, 1, 2,
                 pick_split_k(NG, kp, M), tag='gemm_dw_ih', layout=A_IL | (bil if l > 0 else 0))
            # dWhh[dir] = dG[dir]^T . h_prev  (fw: h_{t-1}, bw: h_{t+1}): row-shifted views of the zero-framed Y
            if T > 1:
                gw = self.view(g, 'whh%d' % l)
                sk = pick_split_k(GATES * HP, HP, M, 148)
                yb = ws['Ybuf'][l].data_ptr()
                # (HBM-bound: one direction's dG, 2048 B per row, and its half of Y, 512 B, are read once for 0.27 TFLOP)
                hh_bytes = M * (GATES * HP * 2 + HP * 2)
                gemm(G.data_ptr(), NG, yb, NY, _p(gw), HP, None, GATES * HP, HP, M, 1, 2, sk, tag='gemm_dw_hh',
                     layout=A_IL | bil, nbytes=hh_bytes)
                a_bw = G.data_ptr() + (GATES * HP // 8) * 512            # IL column offset: 512 B per 8-column chunk
                if yil:     # row 2B of Ybuf = 2B/32 row blocks of NY/8 chunks of 512 B; column HP = HP/8 chunks further
                    b_bw = yb + ((2 * B // 32) * (NY // 8) + HP // 8) * 512
                else:
                    b_bw = yb + (2 * B * NY + HP) * 2
                gemm(a_bw, NG, b_bw, NY, gw.data_ptr() + GATES * HP * HP * 4, HP, None, GATES * HP, HP, M, 1, 2, sk,
                     tag='gemm_dw_hh', layout=A_IL | bil, nbytes=hh_bytes)
            if l > 0:
                nxt = 1 - cur
                gemm(_p(G), NG, _p(self.half['wihT%d' % l]), NG, _p(ws['dY'][nxt]), NY, None, M, NY, NG, 0, 0,
                     tag='gemm_dx', layout=A_IL | C_IL)
                cur = nxt
        return g

    # ---- overflow guard ------------------------------------------------------------------------
    @property
    def guard_scale_ptr(self):
        """Device address of the guard's dynamic scale s (a float32): the grad_scale_dev of the loss kernels."""
        return ctypes_ptr(self.guard.data_ptr() + 16)

    def guard_state(self):
        """(steps skipped, current dynamic scale) -- synchronises; for logging and tests."""
        g = self.guard.cpu()
        return int(g[1]), float(g[4:5].view(torch.float32)[0])

    def _guard_check(self):
        _lib.check(_lib.load().avsi_grad_guard_check(_p(self.grad), self.layout.n_params_padded, _p(self.guard),
                                                     _lib.stream_ptr()), 'avsi_grad_guard_check')

    def _guard_update(self):
        _lib.check(_lib.load().avsi_grad_guard_update(_p(self.guard), int(self.guard_growth_interval), _lib.stream_ptr()),
                   'avsi_grad_guard_update')

    def sgd_step(self, lr, momentum=None, grad_unscale=1.0, unscale_dev=None, l2=0.0):
        """tf.train.GradientDescentOptimizer / MomentumOptimizer(0.9) update (models.py:169-173)."""
        lib = _lib.load()
        self.step_count += 1
        n = self.layout.n_params_padded
        acc = None
        if momentum is not None:
            if not hasattr(self, 'momentum_acc'):
                self.momentum_acc = torch.zeros(n, dtype=torch.float32, device=self.device)
            acc = self.momentum_acc
        with _lib.span('sgd'):
            self._guard_check()
            _lib.check(lib.avsi_sgd_momentum(_p(self.theta), _p(self.grad), _p(acc), n, lr, momentum or 0.0, grad_unscale,
                                             _p(unscale_dev), l2, _p(self.guard), _lib.stream_ptr()), 'avsi_sgd_momentum')
            self._guard_update()
            self.refresh_half()

    # ---- update -------------------------------------------------------------------------------
    def adam_step(self, lr=1e-3, grad_unscale=1.0, unscale_dev=None, l2=0.0, b1=0.9, b2=0.999, eps=1e-8,
                  device_step=False):
        """device_step: the update's number is read from the guard's call count on the device instead of being passed
        from the host (the form a captured CUDA graph replays); the host mirror `step_count` still advances."""
        lib = _lib.load()
        self.step_count += 1
        n = self.layout.n_params_padded
        with _lib.span('adam'):
            self._guard_check()
            _lib.check(lib.avsi_adam_tf(_p(self.theta), _p(self.grad), _p(self.adam_m), _p(self.adam_v), n, lr, b1, b2,
                                        eps, 0 if device_step else self.step_count, grad_unscale, _p(unscale_dev), l2, _p(self.guard),
                                        _lib.stream_ptr()), 'avsi_adam_tf')
            self._guard_update()
            self.refresh_half()
