"""CPU oracle for the AV speech-inpainting hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it, and there only as
the checker (or the CPU arm being timed), never on the CUDA product path.

What it restates (reference = dr-pato/audio-visual-speech-inpainting, paths
relative to ``av_speech_inpainting/``):

* ``stft.py``    audio_processing.py:25-72,145-164 + masking.py:41-45
                 (tf.contrib.signal.stft / inverse_stft / mel matrix semantics)
* ``video.py``   av_sync.py:7-40, face_landmarks.py:30-39, tfrecord_utils.py:86-107
* ``maskgen.py`` dataset_generator.py:11-48
* ``blstm.py``   models.py:20-159 (SI) and models.py:1750-1963 (MTL), LSTM
                 equations of tf.contrib.cudnn_rnn.CudnnCompatibleLSTMCell
* ``ctc.py``     tf.nn.ctc_loss as called at models.py:1950-1953
* ``adam.py``    tf.train.AdamOptimizer as called at models.py:168,178
* also: ``blstm.forward_asr`` (models_asr.py:87-160), dropout with a GIVEN keep mask (models.py:117),
  ``ctc.ctc_beam_search`` (tf.nn.ctc_beam_search_decoder, models.py:1627) pinned to exhaustive enumeration,
  ``stft.preemphasis / get_mfcc / delta / add_delta_features`` (audio_processing.py:19-22, 74-103),
  ``feat_stats.py`` (audio_feat_preprocessing.py:76-115), ``cpu_port.py`` (the timed torch-CPU port),
  ``resample.py`` (downsampling, audio_processing.py:9-16 = scipy.signal.resample; pinned to scipy itself),
  ``phase.py`` (the consistency-projection stand-in for lws.run_lws, inference.py:143-154; unpinned against lws)

The arithmetic of the reference lives in TensorFlow 1.13-1.15 (un-vendored,
``requirements.txt:5-6`` gives only a lower bound) which is absent from this
image, so the reference itself cannot be run here.

PARITY PINNING STATUS
  * STFT -> mask -> iSTFT -> int16 chain: PINNED by the four known-answer
    fixtures the reference ships (docs/files/{800ms,1600ms}/ex{1,2}/
    {target,masked}.wav), reproduced to +-1 int16 LSB
    (tests/test_oracle_fixtures.py, vectors committed under tests/golden/).
  * BLSTM / L1 / CTC / Adam / mel: PARITY UNPINNED by the reference (it has no
    tests and TF cannot run here).  They are cross-checked against independent
    implementations instead (torch.nn.LSTM, torch ctc_loss, brute-force CTC
    path enumeration, finite differences), see tests/test_oracle_*.py.
"""
