"""Golden vectors FROM TENSORFLOW ITSELF for everything the oracle restates from TF internals.

    pip install tensorflow-cpu            # any TF >= 1.13 with tf.compat.v1; py3.12 wheels exist for TF 2.16+
    python tests/golden/make_tf_golden.py # writes tests/golden/tf_golden.npz (+ tf_golden_ckpt.*, tf_golden.tfrecord)

NOT RUN YET: this image has no TensorFlow and no network (`pip install tensorflow-cpu` -> "No matching distribution
found", also with --find-links /opt/wheelhouse; the same on the GPU boxes, which have no network either) -- recorded in
DESIGN.md section 2.  The script is committed so that the first environment with TF can produce the file;
tests/test_tf_golden_cpu.py then checks the oracle against it (and skips while the file is absent).

What it records, on seeded inputs small enough to commit:
  stft            tf.signal.stft(frame_length=384, frame_step=192, fft_length=512, pad_end=True)   audio_processing.py:35-36
  inv_stft        tf.signal.inverse_stft(..., window_fn=inverse_stft_window_fn(192))               audio_processing.py:149-151
  mel             tf.signal.linear_to_mel_weight_matrix(80, 257, 16000, 125, 7600)                 audio_processing.py:63-65
  mfcc            tf.signal.mfccs_from_log_mel_spectrograms(...)[..., :13]                         audio_processing.py:75-82
  ctc_*           tf.compat.v1.nn.ctc_loss(labels sparse, logits time-major, blank = last) + grads models.py:1950-1953
  adam_*          tf.compat.v1.train.AdamOptimizer(1e-3), 3 steps                                  models.py:168,178
  lstm_*          tf.compat.v1.nn.rnn_cell.LSTMCell(forget_bias=0) (= CudnnCompatibleLSTMCell's base) forward over a
                  sequence via dynamic_rnn + tf.gradients; kernel / bias in the canonical layout     models.py:106-115
  checkpoint      tf.compat.v1.train.Saver bundle of two variables + Adam slots (names as TF writes them)
  tfrecord        one tf.train.SequenceExample laid out like tfrecord_utils.py:19-41, via tf.io.TFRecordWriter
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    import tensorflow as tf
    tf1 = tf.compat.v1
    tf1.disable_eager_execution()
    rng = np.random.default_rng(2024)
    out = {'tf_version': np.frombuffer(tf.__version__.encode(), np.uint8)}

    wav = np.round(rng.normal(0, 3000, (2, 2500))).astype(np.float32)
    logmel = rng.standard_normal((2, 7, 80)).astype(np.float32)
    T, B, C, L = 12, 3, 6, 4
    logits = rng.standard_normal((T, B, C)).astype(np.float32)
    labels = np.array([[0, 1, 1, 4], [2, 0, 0, 0], [3, 3, 2, 1]], np.int32)
    lab_len = np.array([4, 1, 3], np.int32)
    seq_len = np.array([12, 9, 12], np.int32)
    theta0 = rng.standard_normal(11).astype(np.float32)
    grads = rng.standard_normal((3, 11)).astype(np.float32)
    I, H, TL, BL = 5, 4, 6, 2
    x = rng.standard_normal((BL, TL, I)).astype(np.float32)
    kernel = (rng.standard_normal((I + H, 4 * H)) * 0.4).astype(np.float32)
    bias = (rng.standard_normal(4 * H) * 0.1).astype(np.float32)
    out.update(wav=wav, logmel=logmel, ctc_logits=logits, ctc_labels=labels, ctc_lab_len=lab_len, ctc_seq_len=seq_len,
               adam_theta0=theta0, adam_grads=grads, lstm_x=x, lstm_kernel=kernel, lstm_bias=bias)

    g = tf1.Graph()
    with g.as_default():
        w = tf1.constant(wav)
        stft = tf.signal.stft(w, frame_length=384, frame_step=192, fft_length=512, pad_end=True)
        inv = tf.signal.inverse_stft(stft, frame_length=384, frame_step=192, fft_length=512,
                                     window_fn=tf.signal.inverse_stft_window_fn(192))
        mel = tf.signal.linear_to_mel_weight_matrix(80, 257, 16000, 125.0, 7600.0)
        mfcc = tf.signal.mfccs_from_log_mel_spectrograms(tf1.constant(logmel))[..., :13]
        # CTC: dense labels -> sparse like tf.contrib.layers.dense_to_sparse / ctc_label_dense_to_sparse (models.py:1494)
        idx = np.array([[b, i] for b in range(B) for i in range(lab_len[b])], np.int64)
        val = np.array([labels[b, i] for b in range(B) for i in range(lab_len[b])], np.int32)
        sp = tf1.SparseTensor(idx, val, [B, L])
        lg = tf1.constant(logits)
        nll = tf1.nn.ctc_loss(sp, lg, seq_len, preprocess_collapse_repeated=False, ctc_merge_repeated=True, time_major=True)
        dlg = tf1.gradients(tf1.reduce_sum(nll), lg)[0]
        # Adam
        th = tf1.Variable(theta0, name='theta')
        gph = tf1.placeholder(tf.float32, [11])
        opt = tf1.train.AdamOptimizer(1e-3)
        step = opt.apply_gradients([(gph, th)])
        # LSTM cell, canonical weights, forward direction over the sequence + gradient of sum(h^2)
        cell = tf1.nn.rnn_cell.LSTMCell(H, forget_bias=0.0, name='cudnn_compatible_lstm_cell')
        xs = tf1.constant(x)
        hs, _ = tf1.nn.dynamic_rnn(cell, xs, dtype=tf.float32)
        lstm_vars = {v.op.name.split('/')[-1]: v for v in cell.variables}
        gk, gb = tf1.gradients(tf1.reduce_sum(hs * hs), [lstm_vars['kernel'], lstm_vars['bias']])
        saver = tf1.train.Saver()
        with tf1.Session(graph=g) as sess:
            sess.run(tf1.global_variables_initializer())
            sess.run([tf1.assign(lstm_vars['kernel'], kernel), tf1.assign(lstm_vars['bias'], bias)])
            r = sess.run(dict(stft=stft, inv_stft=inv, mel=mel, mfcc=mfcc, ctc_nll=nll, ctc_dlogits=dlg, lstm_h=hs,
                              lstm_dkernel=gk, lstm_dbias=gb))
            out.update(r)
            thetas = []
            for i in range(3):
                sess.run(step, {gph: grads[i]})
                thetas.append(sess.run(th))
            out['adam_thetas'] = np.stack(thetas)
            prefix = saver.save(sess, os.path.join(HERE, 'tf_golden_ckpt'), write_meta_graph=False)
            names = sorted(v.op.name for v in tf1.global_variables())
            out['ckpt_names'] = np.frombuffer('\n'.join(names).encode(), np.uint8)
            for v in tf1.global_variables():
                out['ckpt/' + v.op.name] = sess.run(v)
            print('checkpoint written:', prefix, names)

    ex = tf.train.SequenceExample()
    ex.context.feature['sequence_length'].int64_list.value.append(5)
    ex.context.feature['labels_length'].int64_list.value.append(3)
    ex.context.feature['target_audio_wav'].float_list.value.extend(wav[0, :64].tolist())
    ex.context.feature['sample_path'].bytes_list.value.append(b's1/bbaf2n')
    vid = rng.standard_normal((5, 136)).astype(np.float32)
    msk = (rng.uniform(size=(5, 257)) > 0.3).astype(np.float32)
    for v in vid:
        ex.feature_lists.feature_list['video_features'].feature.add().float_list.value.extend(v.tolist())
    for m in msk:
        ex.feature_lists.feature_list['mask'].feature.add().float_list.value.extend(m.tolist())
    for l in (3.0, 0.0, 7.0):
        ex.feature_lists.feature_list['labels'].feature.add().float_list.value.append(l)
    with tf.io.TFRecordWriter(os.path.join(HERE, 'tf_golden.tfrecord')) as wr:
        wr.write(ex.SerializeToString())
    out.update(rec_video=vid, rec_mask=msk, rec_wav=wav[0, :64])
    np.savez(os.path.join(HERE, 'tf_golden.npz'), **out)
    print('wrote tf_golden.npz with', sorted(out))


if __name__ == '__main__':
    try:
        import tensorflow  # noqa: F401
    except ImportError as e:
        print('TensorFlow is not installed (%s): nothing written.' % e)
        sys.exit(3)
    main()
