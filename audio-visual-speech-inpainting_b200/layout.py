"""Parameter layouts: reference checkpoint names/shapes <-> padded kernel layouts.

The drop-in contract is the reference's canonical variable naming:

  cudnn_lstm/stack_bidirectional_rnn/cell_{l}/bidirectional_rnn/{fw,bw}/cudnn_compatible_lstm_cell/kernel
        [I_l + H, 4H], rows [x ; h], column blocks i, j(g), f, o        (models.py:106-115)
  .../bias   [4H]
  logits/weights [2H, F], logits/biases [F]                              (models.py:118-121)
  inpainting/weights|biases, asr/weights|biases                          (models.py:1902-1912)

The kernels use ONE flat fp32 buffer with these padded blocks (HP = 256, gate column
n = dir*1024 + unit*4 + gate, layer>0 input column = dir'*256 + unit'):

  wih{l} [2048, Kp_l]   whh{l} [2048, 256]   b{l} [2048]   head_w [NOp, 512]   head_b [NOp]

Padded rows/columns are zero and provably stay zero under training (zero gradient).
Pure numpy: covered by CPU tests.
"""
import numpy as np

HP = 256
GATES = 4


def round_up(x, m):
    return (x + m - 1) // m * m


def cell_prefix(layer, direction):
    return ('cudnn_lstm/stack_bidirectional_rnn/cell_%d/bidirectional_rnn/%s/cudnn_compatible_lstm_cell'
            % (layer, direction))


class ParamLayout(object):
    def __init__(self, in_dim, hidden, n_layers, out_dim=257, n_classes=0, dense=()):
        """dense: extra fully connected layers (name_w, name_b, fan_in, fan_out) living in the same flat buffer -- the
        speaker-embedding MLP of the SSNN models (models.py:804-809).  Canonical shape [fan_in, fan_out]; stored transposed
        as block `dw{k}` [fan_out, fan_in rounded up to 8] (the K-major B operand of the forward GEMM) + `db{k}` [fan_out]."""
        if hidden > HP:
            raise ValueError('hidden size %d > %d not supported by the recurrence kernel' % (hidden, HP))
        self.in_dim, self.hidden, self.n_layers = in_dim, hidden, n_layers
        self.out_dim, self.n_classes = out_dim, n_classes
        self.n_out = out_dim + n_classes
        self.k0p = round_up(in_dim, 64)                  # layer-0 K: 128-byte-aligned rows for the TMA operand boxes
        # (a 400-column pitch was measured 8 % slower in the projection GEMM: every 128-byte box row straddles two lines)
        self.nop = round_up(self.n_out, 64)              # head rows; also K of the dY GEMM
        self.blocks = []                                 # (name, offset, shape)
        off = 0
        for l in range(n_layers):
            kp = self.k0p if l == 0 else 2 * HP
            for name, shp in (('wih%d' % l, (2 * GATES * HP, kp)), ('whh%d' % l, (2 * GATES * HP, HP)),
                              ('b%d' % l, (2 * GATES * HP,))):
                self.blocks.append((name, off, shp))
                off += int(np.prod(shp))
        for name, shp in (('head_w', (self.nop, 2 * HP)), ('head_b', (self.nop,))):
            self.blocks.append((name, off, shp))
            off += int(np.prod(shp))
        self.dense = [tuple(d) for d in dense]
        for k, (_, _, fan_in, fan_out) in enumerate(self.dense):
            for name, shp in (('dw%d' % k, (fan_out, round_up(fan_in, 8))), ('db%d' % k, (fan_out,))):
                self.blocks.append((name, off, shp))
                off += int(np.prod(shp))
        self.n_params_padded = off
        self.index = {n: (o, s) for n, o, s in self.blocks}

    # ---- views ---------------------------------------------------------------------------
    def view(self, flat, name):
        off, shp = self.index[name]
        return flat[off:off + int(np.prod(shp))].reshape(shp)

    def layer_k(self, l):
        return self.k0p if l == 0 else 2 * HP

    def head_names(self):
        if self.n_classes:
            return [('inpainting', 0, self.out_dim), ('asr', self.out_dim, self.n_classes)]
        return [('logits', 0, self.out_dim)]

    def canonical_shapes(self):
        H = self.hidden
        shapes = {}
        for l in range(self.n_layers):
            i_l = self.in_dim if l == 0 else 2 * H
            for d in ('fw', 'bw'):
                shapes[cell_prefix(l, d) + '/kernel'] = (i_l + H, 4 * H)
                shapes[cell_prefix(l, d) + '/bias'] = (4 * H,)
        for scope, _, n in self.head_names():
            shapes[scope + '/weights'] = (2 * H, n)
            shapes[scope + '/biases'] = (n,)
        for name_w, name_b, fan_in, fan_out in self.dense:
            shapes[name_w] = (fan_in, fan_out)
            shapes[name_b] = (fan_out,)
        return shapes

    def n_params(self):
        return int(sum(np.prod(s) for s in self.canonical_shapes().values()))

    def _in_cols(self, l):
        """internal input column of canonical input row k, layer l."""
        H = self.hidden
        if l == 0:
            return np.arange(self.in_dim)
        k = np.arange(2 * H)
        return np.where(k < H, k, HP + (k - H))

    # ---- canonical -> internal ------------------------------------------------------------
    def pack(self, params, dtype=np.float32):
        """dict of canonical arrays -> flat padded buffer."""
        H = self.hidden
        flat = np.zeros(self.n_params_padded, dtype)
        shapes = self.canonical_shapes()
        for name, shp in shapes.items():
            if name not in params:
                raise KeyError('missing variable %s' % name)
            if tuple(np.shape(params[name])) != tuple(shp):
                raise ValueError('variable %s has shape %s, expected %s' % (name, np.shape(params[name]), shp))
        u = np.arange(H)
        for l in range(self.n_layers):
            wih, whh, b = self.view(flat, 'wih%d' % l), self.view(flat, 'whh%d' % l), self.view(flat, 'b%d' % l)
            i_l = self.in_dim if l == 0 else 2 * H
            cols = self._in_cols(l)
            for di, d in enumerate(('fw', 'bw')):
                kern = np.asarray(params[cell_prefix(l, d) + '/kernel'], dtype)
                bias = np.asarray(params[cell_prefix(l, d) + '/bias'], dtype)
                for q in range(GATES):
                    rows = di * GATES * HP + u * GATES + q
                    wih[np.ix_(rows, cols)] = kern[:i_l, q * H:(q + 1) * H].T
                    whh[np.ix_(rows, u)] = kern[i_l:, q * H:(q + 1) * H].T
                    b[rows] = bias[q * H:(q + 1) * H]
        hw, hb = self.view(flat, 'head_w'), self.view(flat, 'head_b')
        kcols = self._in_cols(1)
        for scope, r0, n in self.head_names():
            hw[np.ix_(np.arange(r0, r0 + n), kcols)] = np.asarray(params[scope + '/weights'], dtype).T
            hb[r0:r0 + n] = np.asarray(params[scope + '/biases'], dtype)
        for k, (name_w, name_b, fan_in, fan_out) in enumerate(self.dense):
            self.view(flat, 'dw%d' % k)[:, :fan_in] = np.asarray(params[name_w], dtype).T
            self.view(flat, 'db%d' % k)[:] = np.asarray(params[name_b], dtype)
        return flat

    # ---- internal -> canonical ------------------------------------------------------------
    def unpack(self, flat):
        """flat padded buffer (numpy) -> dict of canonical arrays (padding stripped)."""
        H = self.hidden
        flat = np.asarray(flat)
        out = {}
        u = np.arange(H)
        for l in range(self.n_layers):
            wih, whh, b = self.view(flat, 'wih%d' % l), self.view(flat, 'whh%d' % l), self.view(flat, 'b%d' % l)
            i_l = self.in_dim if l == 0 else 2 * H
            cols = self._in_cols(l)
            for di, d in enumerate(('fw', 'bw')):
                kern = np.zeros((i_l + H, 4 * H), flat.dtype)
                bias = np.zeros(4 * H, flat.dtype)
                for q in range(GATES):
                    rows = di * GATES * HP + u * GATES + q
                    kern[:i_l, q * H:(q + 1) * H] = wih[np.ix_(rows, cols)].T
                    kern[i_l:, q * H:(q + 1) * H] = whh[np.ix_(rows, u)].T
                    bias[q * H:(q + 1) * H] = b[rows]
                out[cell_prefix(l, d) + '/kernel'] = kern
                out[cell_prefix(l, d) + '/bias'] = bias
        hw, hb = self.view(flat, 'head_w'), self.view(flat, 'head_b')
        kcols = self._in_cols(1)
        for scope, r0, n in self.head_names():
            out[scope + '/weights'] = hw[np.ix_(np.arange(r0, r0 + n), kcols)].T.copy()
            out[scope + '/biases'] = hb[r0:r0 + n].copy()
        for k, (name_w, name_b, fan_in, fan_out) in enumerate(self.dense):
            out[name_w] = self.view(flat, 'dw%d' % k)[:, :fan_in].T.copy()
            out[name_b] = self.view(flat, 'db%d' % k).copy()
        return out

    def pad_mask(self):
        """1.0 where the flat buffer holds a real parameter, 0.0 on padding."""
        ones = {k: np.ones(s, np.float32) for k, s in self.canonical_shapes().items()}
        return (self.pack(ones) != 0).astype(np.float32)


def init_canonical(layout, seed=1, bias_scale=0.0):
    """Seeded synthetic weights in canonical layout: Glorot-uniform LSTM
    kernels (CudnnLSTM default), truncated-normal heads with stddev 1/sqrt(2H) (models.py:119),
    zero biases (or small random ones for tests)."""
    rng = np.random.default_rng(seed)
    params = {}
    for name, shp in layout.canonical_shapes().items():
        if name.endswith('/kernel'):
            lim = np.sqrt(6.0 / (shp[0] + shp[1]))
            params[name] = rng.uniform(-lim, lim, shp)
        elif '/weights' in name:
            w = rng.standard_normal(shp)
            bad = np.abs(w) > 2.0
            while bad.any():
                w[bad] = rng.standard_normal(int(bad.sum()))
                bad = np.abs(w) > 2.0
            # heads: stddev 1/sqrt(fan_in) (models.py:119); speaker_embedding/weights_1: 1/sqrt(audio_feat_dim) with
            # fan_in = 2 * audio_feat_dim (models.py:804)
            params[name] = w / np.sqrt(shp[0] / 2.0 if name.endswith('speaker_embedding/weights_1') else shp[0])
        else:
            params[name] = bias_scale * rng.standard_normal(shp)
    return params
