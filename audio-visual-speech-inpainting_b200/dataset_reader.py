"""TFRecord data manager with the reference's interface (dataset_reader.py:12-60), without TensorFlow.

`DataManager.get_dataset(file_list, shuffle, seed)` / `get_iterator(dataset, batch_size, n_epochs)` keep their
names and argument meaning; instead of graph tensors the iterator is a Python generator of numpy batches in the
reference's order (dataset_reader.py:77-79):

    (sequence_length i32 [B], labels_length i32 [B], target_audio_wav i32 [B,N], sample_path [B] bytes,
     labels f32 [B,Lmax], video_features f32 [B,T,V], mask f32 [B,T,F])

Exhaustion raises StopIteration where the reference's loop catches tf.errors.OutOfRangeError.  Data parallel runs
pass `rank` / `world` so that each process reads files[rank::world]."""
import os
import random

import numpy as np

from . import parallel
from .tfrecord_io import parse_av_batch, parse_av_sample, parse_sequence_example, read_records


def _native_available():
    from . import tfrecord_io
    return tfrecord_io._native() is not None


class _Dataset(object):
    def __init__(self, files, parse, shuffle, buffer_size, seed, workers=1):
        self.files, self.parse, self.shuffle, self.buffer_size, self.seed = list(files), parse, shuffle, buffer_size, seed
        self.workers = workers

    def raw_records(self):
        """The record payloads in the order samples() parses them (file order, or through the shuffle buffer)."""
        def records():
            for f in self.files:
                for rec in read_records(f):
                    yield rec
        if not self.shuffle:
            for r in records():
                yield r
            return
        rng = random.Random(self.seed)
        buf = []
        for r in records():
            buf.append(r)
            if len(buf) >= self.buffer_size:
                yield buf.pop(rng.randrange(len(buf)))
        while buf:
            yield buf.pop(rng.randrange(len(buf)))

    def samples(self):
        def records():
            for f in self.files:
                for rec in read_records(f):
                    yield rec

        def raw():
            # records are parsed by a small thread pool, a bounded window ahead and in order (the native parser
            # releases the GIL): the consumer sees the same sequence as a serial loop
            if self.workers <= 1:
                for rec in records():
                    yield self.parse(rec)
                return
            from collections import deque
            from concurrent.futures import ThreadPoolExecutor
            with ThreadPoolExecutor(max_workers=self.workers) as pool:
                window = deque()
                for rec in records():
                    window.append(pool.submit(self.parse, rec))
                    if len(window) >= 4 * self.workers:
                        yield window.popleft().result()
                while window:
                    yield window.popleft().result()
        if not self.shuffle:
            for s in raw():
                yield s
            return
        rng = random.Random(self.seed)                       # tf.data shuffle: a buffer of buffer_size elements
        buf = []
        for s in raw():
            buf.append(s)
            if len(buf) >= self.buffer_size:
                yield buf.pop(rng.randrange(len(buf)))
        while buf:
            yield buf.pop(rng.randrange(len(buf)))


class DataManager(object):
    """Utilities to read TFRecords"""

    def __init__(self, num_audio_samples=48000, audio_feat_size=257, video_feat_size=136, buffer_size=1000, mode='fixed',
                 rank=0, world=1, num_parallel_calls=None, embedding_size=0, **unused):
        if mode != 'fixed':
            raise NotImplementedError("only the 'fixed' TFRecord mode works in the reference (SURVEY.md 2.4)")
        self.num_audio_samples, self.audio_feat_size, self.video_feat_size = num_audio_samples, audio_feat_size, video_feat_size
        self.buffer_size, self.mode, self.rank, self.world = buffer_size, mode, rank, world
        # dataset_reader_emb.py:12-81: records that also carry a per-utterance `embedding` (512 floats); the batches are
        # then the 8-tuple (seq_len, lab_len, wav, embedding, sample_path, labels, video, mask) of :79-81
        self.embedding_size = int(embedding_size or 0)
        # parser threads (tf.data's num_parallel_calls of dataset_reader.py:55); default: up to 8 host cores
        self.workers = num_parallel_calls if num_parallel_calls else max(1, min(8, os.cpu_count() or 1))

    def read_data_format_fixed(self, sample):
        if self.embedding_size:
            ctx, seq = parse_sequence_example(sample)
            wav = np.asarray(ctx['target_audio_wav'], np.float32)
            emb = np.asarray(ctx['embedding'], np.float32)
            if wav.shape[0] != self.num_audio_samples or emb.shape[0] != self.embedding_size:
                raise ValueError('target_audio_wav / embedding have %d / %d values, expected %d / %d'
                                 % (wav.shape[0], emb.shape[0], self.num_audio_samples, self.embedding_size))
            return (np.int32(ctx['sequence_length'][0]), np.int32(ctx['labels_length'][0]), wav.astype(np.int32), emb,
                    ctx['sample_path'][0] if ctx['sample_path'] else b'',
                    np.concatenate(seq['labels']).astype(np.float32) if seq.get('labels') else np.zeros(0, np.float32),
                    np.stack(seq['video_features']).astype(np.float32), np.stack(seq['mask']).astype(np.float32))
        fast = parse_av_sample(sample, self.num_audio_samples, self.audio_feat_size, self.video_feat_size)
        if fast is not None:
            seq_len, lab_len, wav, path, labels, video, mask = fast
            if wav.shape[0] != self.num_audio_samples:
                raise ValueError('target_audio_wav has %d samples, expected %d' % (wav.shape[0], self.num_audio_samples))
            return (np.int32(seq_len), np.int32(lab_len), wav.astype(np.int32), path, labels, video, mask)
        ctx, seq = parse_sequence_example(sample)
        wav = np.asarray(ctx['target_audio_wav'], np.float32)
        if wav.shape[0] != self.num_audio_samples:
            raise ValueError('target_audio_wav has %d samples, expected %d' % (wav.shape[0], self.num_audio_samples))
        return (np.int32(ctx['sequence_length'][0]), np.int32(ctx['labels_length'][0]), wav.astype(np.int32),
                ctx['sample_path'][0] if ctx['sample_path'] else b'',
                np.concatenate(seq['labels']).astype(np.float32) if seq.get('labels') else np.zeros(0, np.float32),
                np.stack(seq['video_features']).astype(np.float32), np.stack(seq['mask']).astype(np.float32))

    def get_dataset(self, file_list, shuffle=True, seed=None, keep_order=False):
        """keep_order: the caller passes the SAME (e.g. seed-shuffled) order on every rank; the shard is then taken from
        that order instead of the sorted list."""
        files = parallel.shard_list(file_list, self.rank, self.world, keep_order) if self.world > 1 else list(file_list)
        return _Dataset(files, self.read_data_format_fixed, shuffle, self.buffer_size, seed, self.workers)

    def get_iterator(self, dataset, batch_size=16, n_epochs=None, drop_remainder=False):
        def fast_batches():
            # whole batches parsed natively into their arrays by a thread pool, one batch ahead of the consumer
            import itertools
            import queue
            import threading
            from concurrent.futures import ThreadPoolExecutor
            q = queue.Queue(maxsize=2)

            def produce():
                try:
                    with ThreadPoolExecutor(max_workers=self.workers) as pool:
                        epoch = 0
                        while n_epochs is None or epoch < n_epochs:
                            recs = dataset.raw_records()
                            while True:
                                chunk = list(itertools.islice(recs, batch_size))
                                if not chunk or (drop_remainder and len(chunk) < batch_size):
                                    break
                                q.put(parse_av_batch(chunk, self.num_audio_samples, self.audio_feat_size,
                                                     self.video_feat_size, pool))
                            epoch += 1
                    q.put(None)
                except BaseException as e:                    # noqa: BLE001 -- re-raised in the consumer
                    q.put(e)
            threading.Thread(target=produce, daemon=True).start()
            while True:
                item = q.get()
                if item is None:
                    return
                if isinstance(item, BaseException):
                    raise item
                yield item

        def batches():
            epoch = 0
            while n_epochs is None or epoch < n_epochs:
                cur = []
                for s in dataset.samples():
                    cur.append(s)
                    if len(cur) == batch_size:
                        yield _collate(cur)
                        cur = []
                if cur and not drop_remainder:
                    yield _collate(cur)
                epoch += 1
        native = parse_av_sample is not None and _native_available() and not self.embedding_size
        return dataset, (fast_batches() if native else batches())


def _collate(samples):
    cols = list(zip(*samples))
    if len(cols) == 8:                                   # with embedding (dataset_reader_emb.py:79-81)
        return (np.asarray(cols[0], np.int32), np.asarray(cols[1], np.int32), np.stack(cols[2]), np.stack(cols[3]),
                list(cols[4]), np.stack(cols[5]), np.stack(cols[6]), np.stack(cols[7]))
    return (np.asarray(cols[0], np.int32), np.asarray(cols[1], np.int32), np.stack(cols[2]), list(cols[3]),
            np.stack(cols[4]), np.stack(cols[5]), np.stack(cols[6]))
