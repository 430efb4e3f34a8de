#!/usr/bin/env python
"""Benchmark of the hot path: AV-SI training utterances/s (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W              (N > 1: launched under torchrun)
    python bench.py --impl reference ...                       (CPU arm: the oracle port on host cores)

A step = one pass of the hot path over one synthetic GRID-shaped batch (3 s, 16 kHz, 75 landmark
frames): landmark upsampling + motion vectors -> fused STFT front end -> 3-layer BLSTM forward ->
masked-L1 -> BPTT -> (NCCL gradient all-reduce) -> TF-form Adam.  `value` times it with inputs
resident in HBM; `e2e` times the same step through the public model API from pinned HOST buffers
(H2D of wav / mask / landmarks and D2H of the loss inside the timed region) with the reference's fp32 placeholder
dtypes (training.py:69-71), `e2e_compact` the same with the samples / masks crossing PCIe as int16 / uint8.
`extra` holds few-step runs of the other BASELINE configurations on the same GPUs: AV-MTL-SI (configs[2]),
20 s utterances (configs[4]) and the batch sweep of configs[1].  Prints ONE JSON line.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'AV-SI training utterances/sec'
UNIT = 'utterances/s'
CPU_SAMPLE_B = 64          # utterances per CPU-arm step (the largest batch whose step stays within seconds on 16 cores)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=8)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--batch', type=int, default=int(os.environ.get('AVSI_BENCH_BATCH', 2048)), help='utterances per GPU per step')
    ap.add_argument('--audio-len', type=int, default=48000)
    ap.add_argument('--model', default='av-blstm')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-extras', action='store_true', help='skip the few-step runs of configs[2] / [4] / batch sweep')
    ap.add_argument('--long-batch', type=int, default=int(os.environ.get('AVSI_BENCH_LONG_BATCH', 2048)),
                    help='utterances per GPU of the 20 s extra (configs[4])')
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        p = json.load(open(path))
        return {'hbm_gbs': p.get('hbm_gbs', 6650.0), 'bf16_tflops': p.get('bf16_tflops', 1590.0),
                'bf16_tflops_sustained': p.get('bf16_tflops_sustained', 1400.0), 'source': 'measured'}
    return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1400.0, 'source': 'fallback'}


class ClockSampler(object):
    """nvidia-smi clock / throttle-reason samples during the timed region (B200_PROFILING.md recipe)."""
    FIELDS = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
              'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
              'clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.gpu), '--query-gpu=' + self.FIELDS,
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for ln in self.lines:
            f = [x.strip() for x in ln.split(',')]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for nm, val in zip(names, f[5:9]):
                if val.lower().startswith('active'):
                    reasons.add(nm)
        sm.sort()
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': max(smax) if smax else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


def cpu_baseline(args, steps=1, warmup=0):
    """The oracle port (kind = "port") timed on the host cores on a bounded sample of the workload."""
    import numpy as np
    from avsi_b200 import synth
    from oracle import cpu_port
    from oracle import video as ovideo
    B = CPU_SAMPLE_B
    b = synth.make_batch(B, audio_len=args.audio_len, seed=0)
    b['video'] = np.stack([ovideo.video_features(b['landmarks'][i].astype(np.float64), b['T'], b['vmean'][i], b['vstd'][i])
                           for i in range(B)]).astype(np.float32)
    ups, cores, dt = cpu_port.time_train_steps(b, steps=steps, warmup=warmup)
    return {'value': ups, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample_batch': B,
            'sample': '%d step(s) of a %d-utterance AV-SI train step (fwd+bwd+Adam, fp32 torch-CPU port of the '
                      'reference TF graph; TF 1.x itself is not installable here), %.2f s/step' % (steps, B, dt)}, dt


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cb, dt = cpu_baseline(args, steps=max(1, min(args.steps, 4)), warmup=min(args.warmup, 1))
    line = {'impl': 'reference', 'metric': METRIC, 'value': cb['value'], 'unit': UNIT, 'n_gpus': args.gpus,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': dt * 1e3, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            # the GPU arm's workload; each CPU step times a bounded sample of it (cpu_baseline.sample_batch utterances, not
            # batch_per_gpu: a 2048-utterance step would take ~2 minutes on these cores) -- `same_config` is by workload only
            'config': dict(workload_config(args, max(1, args.gpus)), cpu_sample_batch=CPU_SAMPLE_B),
            'cpu_baseline': cb,
            'e2e': {'value': cb['value'], 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(line))


def workload_config(args, world):
    """`config` of the JSON line, shared by both arms: names the workload, no model hyper-parameters."""
    B, N = args.batch, args.audio_len
    T = -(-N // 192)
    mtl = 'ctc' in args.model
    inp = args.model.split('-')[0]
    what = {'a': 'masked log-spectrogram', 'v': 'upsampled landmark motion vectors',
            'av': 'masked log-spectrogram ++ upsampled landmark motion vectors'}[inp]
    name = 'AV-MTL-SI train step (configs[2], joint CTC phone head)' if mtl else \
        '%s-SI train step (configs[%d])' % (inp.upper(), 4 if N > 48000 else 1)
    return {'workload': '%s: %g s 16 kHz utterances + 75-frame landmark streams, %s, 3x BLSTM-250, %s, TF-form Adam'
                        % (name, N / 16000.0, what, 'hole L1 + CTC loss' if mtl else 'L1 loss'),
            'model': args.model, 'batch_per_gpu': B, 'global_batch': B * world, 'frames': T,
            'parallelism': 'dp%d' % world,
            'l2_policy': 'inputs larger than L2 (wav+mask %.0f MB per step, activations %.1f GB)'
                         % (B * (N + T * 257) * 4 / 1e6, T * B * (3 * 2048 * 2 + 3 * 512 * 6) / 1e9)}


class Workload(object):
    """One synthetic GRID-shaped workload on this rank's GPU: host batch, resident copies, model, the resident step and
    the end-to-end step (inputs from pinned HOST buffers every step, loss read back every step)."""
    E2E_KEYS = ('wav', 'mask', 'landmarks', 'vmean', 'vstd', 'seq_len')

    def __init__(self, model_name, B, audio_len, pg, dev, rank, world, unique=None, want_e2e=False):
        import numpy as np
        import torch
        from avsi_b200 import _lib, av_sync, models, synth
        self.torch, self._lib, self.av_sync = torch, _lib, av_sync
        self.B, self.pg, self.dev, self.world = B, pg, dev, world
        n_unique = min(B, unique) if unique else B
        host = synth.make_batch(n_unique, audio_len=audio_len, seed=rank)      # rank-offset data, identical weights
        self.T = T = host['T']
        cfg = synth.default_config(model_name, batch_size=B * world, audio_len=audio_len)
        keys = ('wav', 'mask', 'landmarks', 'vmean', 'vstd', 'seq_len', 'mean', 'std')
        if n_unique == B:
            self.host = host
            self.pin = {k: torch.from_numpy(np.ascontiguousarray(host[k])).pin_memory() for k in keys} if want_e2e else None
            self.res = {k: (self.pin[k] if want_e2e else torch.from_numpy(np.ascontiguousarray(host[k]))).to(dev) for k in keys}
        else:
            # long utterances: a few distinct utterances tiled over the batch ON THE DEVICE (drawing 2048 x 320000 samples
            # on the host takes longer than the timed steps); every kernel still processes B distinct rows of memory
            assert not want_e2e
            self.host, self.pin = host, None
            idx = torch.arange(B, device=dev) % n_unique
            small = {k: torch.from_numpy(np.ascontiguousarray(host[k])).to(dev) for k in keys}
            self.res = {k: (small[k][idx].contiguous() if k not in ('mean', 'std') else small[k]) for k in keys}
        self.cls, self.inp = models.MODEL_REGISTRY[model_name]
        res = self.res
        video = self.video_of(res) if self.inp != 'a' else None
        if self.cls.MTL:                              # configs[2]: AV-MTL-SI, joint CTC phone head (labels stay resident)
            lab = np.ascontiguousarray(host['labels'][np.arange(B) % n_unique])
            ll = np.ascontiguousarray(host['lab_len'][np.arange(B) % n_unique])
            self.labels, self.lab_len = torch.from_numpy(lab).to(dev), torch.from_numpy(ll).to(dev)
            self.model = self.cls(res['seq_len'], self.lab_len, res['wav'], res['mask'], self.labels, res['mean'], res['std'],
                                  0.0, cfg, video_features=video, input=self.inp, is_training=True, device=dev, process_group=pg)
        else:
            self.model = self.cls(res['seq_len'], res['wav'], res['mask'], res['mean'], res['std'], 0.0, cfg,
                                  video_features=video, input=self.inp, is_training=True, device=dev, process_group=pg)

    def video_of(self, src):
        with self._lib.span('video_features'):
            return self.av_sync.video_pipeline(src['landmarks'], self.T, src['vmean'], src['vstd'], device=self.dev)

    def step_resident(self):
        res = self.res
        self.model.feed(sequence_lengths=res['seq_len'], target_sources=res['wav'], masks=res['mask'],
                        video_features=self.video_of(res) if self.inp != 'a' else None)
        self.model.train_op()

    def timed(self, fn, steps, profile=False):
        """K calls of fn bracketed by barrier + synchronize on both sides, CUDA events, MAX over ranks."""
        import torch.distributed as dist
        torch, _lib = self.torch, self._lib
        torch.cuda.synchronize()
        if self.pg is not None:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0 = _lib.launch_count()
        if profile:
            _lib.profile_start()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        prof = _lib.profile_stop() if profile else None
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=self.dev)
        if self.pg is not None:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), _lib.launch_count() - n0, prof

    def run_resident(self, steps, warmup, profile=True):
        for _ in range(max(warmup, 3)):
            self.step_resident()
        return self.timed(self.step_resident, steps, profile=profile)

    def run_e2e(self, steps, fp32):
        """End-to-end steps: every step's inputs come from pinned HOST buffers (H2D inside the timed region) and every
        step's loss is read back to the host.  The copy of step k+1 runs on a side stream while step k computes (two
        device staging sets) and the loss of step k is fetched one step late (pinned, non-blocking), so that neither
        transfer stalls the kernels.  fp32 = True stages the placeholders of training.py:69-71 as they are (fp32 samples,
        fp32 masks); False stages the storage dtypes (int16 samples as in target.wav / dataset_reader.py:78, uint8 {0,1}
        masks), widened to the fp32 feed tensors on the device (avsi_cast_to_f32, inside the timed step).
        Returns (ms, h2d bytes per step, host dtypes, losses)."""
        import numpy as np
        torch = self.torch
        model, dev, inp = self.model, self.dev, self.inp
        pin = dict(self.pin)
        if not fp32:
            pin['wav'] = torch.from_numpy(self.host['wav'].astype(np.int16)).pin_memory()
            pin['mask'] = torch.from_numpy(self.host['mask'].astype(np.uint8)).pin_memory()
        keys = self.E2E_KEYS
        copy_stream = torch.cuda.Stream(device=dev)
        stage = [{k: torch.empty(pin[k].shape, dtype=pin[k].dtype, device=dev) for k in keys} for _ in range(2)]
        ready = [torch.cuda.Event() for _ in range(2)]
        freed = [torch.cuda.Event() for _ in range(2)]
        loss_pin = torch.zeros(2, dtype=torch.float64).pin_memory()
        loss_evt = [torch.cuda.Event() for _ in range(2)]
        st = {'i': 0, 'losses': []}

        def prefetch(i):
            sl = i % 2
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(freed[sl])                 # the step that last used this staging set is done
                for k in keys:
                    stage[sl][k].copy_(pin[k], non_blocking=True)
                ready[sl].record(copy_stream)

        def reset():
            torch.cuda.synchronize()
            st['i'] = 0
            for sl in range(2):
                freed[sl].record(torch.cuda.current_stream())
            prefetch(0)

        def step():
            i = st['i']
            sl = i % 2
            cur = torch.cuda.current_stream()
            prefetch(i + 1)                                       # next step's H2D overlaps this step's kernels
            cur.wait_event(ready[sl])
            d = stage[sl]
            model.feed(sequence_lengths=d['seq_len'], target_sources=d['wav'], masks=d['mask'],
                       video_features=self.video_of(d) if inp != 'a' else None)
            model.train_op()
            freed[sl].record(cur)
            n = model.engine.layout.n_params_padded
            src = model.engine.grad[n + 4:n + 5].double() if self.pg is not None else \
                model._loss_pass(True, want_pred=False)['sums'][4:5]
            if i > 0:                                             # read the previous step's loss (already on the host)
                loss_evt[1 - sl].synchronize()
                st['losses'].append(float(loss_pin[1 - sl]))
            loss_pin[sl:sl + 1].copy_(src, non_blocking=True)
            loss_evt[sl].record(cur)
            st['i'] = i + 1

        def drain():
            i = st['i']
            if i > 0:
                loss_evt[(i - 1) % 2].synchronize()
                st['losses'].append(float(loss_pin[(i - 1) % 2]))

        reset()
        for _ in range(2):
            step()
        drain()
        reset()                                                   # the first H2D of the timed run is inside the region's lead-in
        st['losses'] = []

        def run():
            step()
            if st['i'] == steps:
                drain()
        ms, _, _ = self.timed(run, steps)
        torch.cuda.current_stream().wait_stream(copy_stream)
        torch.cuda.synchronize()
        if not all(np.isfinite(x) for x in st['losses']):
            raise SystemExit('bench: non-finite loss fetched in the end-to-end steps: %r' % st['losses'][-3:])
        h2d = sum(pin[k].numel() * pin[k].element_size() for k in keys)
        dtypes = {k: str(pin[k].dtype).replace('torch.', '') for k in ('wav', 'mask', 'landmarks')}
        return ms, h2d, dtypes, st['losses']

    def close(self):
        torch = self.torch
        self.model = self.res = self.pin = self.host = None
        torch.cuda.synchronize()
        torch.cuda.empty_cache()


def kernel_table(prof, steps):
    kernels = {}
    tot = sum(v['ms'] for v in prof.values()) or 1.0
    for name, v in sorted(prof.items(), key=lambda kv: -kv[1]['ms']):
        kernels[name] = {'ms_per_step': v['ms'] / steps, 'share': v['ms'] / tot, 'launches_per_step': v['launches'] / steps}
    return kernels


def library_digest():
    """sha256 of the CUDA sources the in-tree library was built from (the build stamp); ncu-derived figures carry it."""
    try:
        return open(os.path.join(ROOT, 'audio-visual-speech-inpainting_b200', 'csrc', 'build', 'stamp.txt')).read().strip()[:16]
    except Exception:
        return None


KERNEL_SOURCES = {      # bench span name -> the CUDA sources its dominant kernel is built from (ncu figures are valid for these)
    'lstm_bwd': ('lstm4_bwd.cu', 'common.cuh', 'sm100_ptx.cuh'),
    'lstm_fwd': ('lstm4.cu', 'common.cuh', 'sm100_ptx.cuh'),
    'frontend': ('frontend.cu', 'common.cuh'),
    'gemm_proj_fwd': ('gemm_sm100.cu', 'common.cuh', 'sm100_ptx.cuh'),
    'gemm_dw_ih': ('gemm_sm100.cu', 'common.cuh', 'sm100_ptx.cuh'),
    'gemm_dw_hh': ('gemm_sm100.cu', 'common.cuh', 'sm100_ptx.cuh'),
    'gemm_dw_head': ('gemm_sm100.cu', 'common.cuh', 'sm100_ptx.cuh'),
    'gemm_dx': ('gemm_sm100.cu', 'common.cuh', 'sm100_ptx.cuh'),
}


def kernel_source_digest(name):
    """sha256 (16 hex digits) of the sources of one kernel: ncu-derived per-kernel figures carry it and go stale with it."""
    import hashlib
    h = hashlib.sha256()
    try:
        for f in KERNEL_SOURCES[name]:
            h.update(open(os.path.join(ROOT, 'audio-visual-speech-inpainting_b200', 'csrc', f), 'rb').read())
    except Exception:
        return None
    return h.hexdigest()[:16]


def run_extras(args, pg, dev, rank, world, pk):
    """Few-step runs of the other BASELINE configurations on the same GPUs (resident inputs, CUDA events, max over ranks):
    configs[2] AV-MTL-SI with the CTC phone head, configs[4] 20 s utterances, and the batch sweep of configs[1]."""
    out = {}
    flop_utt = {48000: 6.62e9}

    def one(tag, model_name, B, audio_len, steps, unique=None):
        wl = Workload(model_name, B, audio_len, pg, dev, rank, world, unique=unique)
        ms, launches, prof = wl.run_resident(steps, 3)
        T = wl.T
        wl.close()
        utt = B * world * steps
        r = {'model': model_name, 'batch_per_gpu': B, 'global_batch': B * world, 'frames': T, 'steps': steps,
             'value': utt / (ms * 1e-3), 'unit': UNIT, 'ms_per_step': ms / steps, 'gpu_launches': launches,
             'kernels': {k: round(v['ms_per_step'], 3) for k, v in list(kernel_table(prof, steps).items())[:6]}}
        # whole-step fraction of the tensor roofline: 6.62 GFLOP per 3 s AV-SI utterance, linear in T
        r['frac_of_tensor_roofline'] = r['value'] / world * 6.62e9 * (T / 250.0) / (pk['bf16_tflops_sustained'] * 1e12)
        out[tag] = r
    one('mtl', 'av-blstm-ssnn-ctc', args.batch, 48000, 4)
    one('long_utterance', 'av-blstm', args.long_batch, 320000, 3, unique=64)
    sweep = {}
    for b in (8, 32, 128, 512):
        wl = Workload(args.model, b, 48000, pg, dev, rank, world)
        ms, launches, _ = wl.run_resident(6, 3, profile=False)
        sweep[str(b)] = {'value': b * world * 6 / (ms * 1e-3), 'ms_per_step': ms / 6, 'gpu_launches_per_step': launches / 6}
        if pg is None:
            # the same step replayed from ONE CUDA graph (models.capture_train_step): what the launch path costs at
            # the reference's own batch sizes
            step, res = wl.model.capture_train_step(), wl.res

            def graphed():
                step(sequence_lengths=res['seq_len'], target_sources=res['wav'], masks=res['mask'],
                     video_features=wl.video_of(res) if wl.inp != 'a' else None)
            for _ in range(3):
                graphed()
            msg, _, _ = wl.timed(graphed, 6)
            sweep[str(b)].update(value_cuda_graph=b * world * 6 / (msg * 1e-3), ms_per_step_cuda_graph=msg / 6)
        wl.close()
    out['batch_sweep'] = {'workload': 'configs[1] at other per-GPU batch sizes (B <= 224: mma.sync recurrence kernels); '
                                      '*_cuda_graph: the whole step replayed from one captured graph', 'unit': UNIT,
                          'per_gpu_batch': sweep}
    return out


def main():
    args = parse_args()
    if args.impl == 'reference':
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist
    import __graft_entry__
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local_rank)
    from avsi_b200 import parallel as _parallel
    numa_cores = 0 if os.environ.get('AVSI_NO_NUMA_BIND') else _parallel.bind_to_gpu_numa(local_rank)
    pg = None
    if world > 1:
        # NCCL prints its version banner to STDOUT at NCCL_DEBUG=VERSION/WARN; stdout carries the one JSON line
        if os.environ.get('NCCL_DEBUG', '').upper() in ('VERSION', 'WARN'):
            os.environ.pop('NCCL_DEBUG')
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    if rank == 0:
        __graft_entry__.build()                 # no-op when the in-tree library is current; the others wait for it
    if world > 1:
        dist.barrier()
        pg = dist.group.WORLD
    dev = torch.device('cuda', local_rank)
    B = args.batch
    wl = Workload(args.model, B, args.audio_len, pg, dev, rank, world, want_e2e=True)
    T = wl.T

    for _ in range(max(args.warmup, 3)):
        wl.step_resident()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.15)
    ms, launches, prof = wl.timed(wl.step_resident, args.steps, profile=True)
    clocks = sampler.stop() if rank == 0 else None
    # end to end, twice: with the reference's fp32 placeholders (the headline `e2e`) and with compact host staging
    ms_e2e, h2d, dtypes, losses = wl.run_e2e(args.steps, fp32=True)
    ms_e2c, h2d_c, dtypes_c, _ = wl.run_e2e(args.steps, fp32=False)
    wl.close()
    pk = peaks()
    extras = None if args.no_extras else run_extras(args, pg, dev, rank, world, pk)

    if rank != 0:
        if pg is not None:
            dist.destroy_process_group()
        return
    utt = B * world * args.steps
    value = utt / (ms * 1e-3)
    kernels = kernel_table(prof, args.steps)

    # roofline.traffic: dram bytes per launch from the committed ncu --set full captures, valid only for the kernel sources
    # they were taken on (profiles/ncu_traffic.json carries a digest of each kernel's sources; a stale entry yields null)
    traffic, traffic_note = {}, {}
    try:
        tj = json.load(open(os.path.join(ROOT, 'profiles', 'ncu_traffic.json')))
        if tj['workload'] == {'batch': B, 'audio_len': args.audio_len}:
            for name in KERNEL_SOURCES:
                if name not in tj:
                    continue
                if tj.get('source_digests', {}).get(name) == kernel_source_digest(name):
                    traffic[name] = tj[name]
                    traffic_note[name] = 'profiles/ncu_traffic.json (%s)' % tj.get('captures', 'ncu --set full')
                else:
                    traffic_note[name] = 'profiles/ncu_traffic.json was captured on other sources of this kernel: not used'
    except Exception:
        pass

    def roof(name):
        r = roof_(name)
        if r is not None:
            r['traffic'] = traffic.get(name)
            r['traffic_source'] = traffic_note.get(name, 'no ncu capture of this kernel on this workload committed')
        return r

    def roof_(name):
        v = prof.get(name)
        if not v or v['ms'] <= 0:
            return None
        sec = v['ms'] * 1e-3
        if v['bytes']:
            a = v['bytes'] / sec / 1e9
            r = {'kernel': name, 'bound': 'hbm', 'achieved': a, 'peak': pk['hbm_gbs'], 'unit': 'GB/s',
                 'frac': a / pk['hbm_gbs'], 'traffic': None, 'peak_source': pk['source'],
                 'frac_of_nominal_8000_gbs': a / 8000.0,          # BASELINE.md asks for both denominators
                 'algorithmic_bytes_per_launch': v['bytes'] / v['launches']}
            if v['flops']:
                r['tensor_tflops'] = v['flops'] / sec / 1e12
            if name in ('lstm_bwd', 'lstm_fwd'):
                r['note'] = ('the launch time of this kernel is independent of the number of clusters (the same at 4 and at 32): '
                             'it sits on its per-cluster dependent chain, not on the HBM roofline; frac is reported against HBM '
                             'because its DRAM traffic equals the algorithmic bytes (DESIGN.md section 4)')
            return r
        a = v['flops'] / sec / 1e12
        return {'kernel': name, 'bound': 'tensor', 'achieved': a, 'peak': pk['bf16_tflops_sustained'], 'unit': 'TFLOP/s',
                'frac': a / pk['bf16_tflops_sustained'], 'traffic': None, 'peak_source': pk['source'] + ' (sustained)',
                'algorithmic_flops_per_launch': v['flops'] / v['launches']}

    dominant = next(iter(kernels)) if kernels else None
    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': max(args.warmup, 3),
        'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f16', 'data': 'synthetic', 'arithmetic': 'fp16 operands, fp32 accumulate / state',
        'config': workload_config(args, world),
        'e2e': {'value': utt / (ms_e2e * 1e-3), 'unit': UNIT, 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': 8,
                'host_dtypes': dtypes, 'staging': 'fp32 placeholders of the reference feed contract (training.py:69-71)'},
        'e2e_compact': {'value': utt / (ms_e2c * 1e-3), 'unit': UNIT, 'h2d_bytes_per_step': h2d_c, 'd2h_bytes_per_step': 8,
                        'host_dtypes': dtypes_c,
                        'staging': 'samples int16 / masks uint8 on the host (narrowed outside the timed region), widened on the device'},
        # the e2e steps train on the same batch: the fetched losses must be finite and go down (work is not skipped)
        'e2e_loss_first_last': [losses[0] / (B * T * 257), losses[-1] / (B * T * 257)] if losses else None,
        'gpu_launches': launches,
        'clocks': clocks,
        'roofline': roof(dominant) if dominant else None,
        'roofline_lstm_bwd': roof('lstm_bwd'),
        'roofline_lstm_fwd': roof('lstm_fwd'),
        'roofline_frontend': roof('frontend'),
        'roofline_gemm_proj': roof('gemm_proj_fwd'),
        'roofline_gemm_dw_ih': roof('gemm_dw_ih'),
        'roofline_gemm_dw_hh': roof('gemm_dw_hh'),
        'roofline_gemm_dx': roof('gemm_dx'),
        # whole step against the tensor roofline: 6.62 GFLOP per AV-SI utterance
        'step_frac_of_tensor_roofline': value / world * 6.62e9 * (T / 250.0) / (pk['bf16_tflops_sustained'] * 1e12),
        'kernels': kernels,
        'library_digest': library_digest(),
        'host_cores_bound_to_gpu_numa_node': numa_cores,
    }
    if extras is not None:
        line['extra'] = extras
    if not args.no_cpu_baseline and world == 1:
        line['cpu_baseline'], _ = cpu_baseline(args, steps=2, warmup=0)
    elif world == 1:
        line['cpu_baseline'] = None
    print(json.dumps(line))
    if pg is not None:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
