"""BASELINE.json configs[3]: front-end throughput sweep over batch and utterance length.

Variants: `spec`   = wav + mask -> log-magnitude, normalised (fp32 target) + masked network input (A-SI training front end)
          `fbanks` = wav -> power-2 spectrogram -> log-mel-80 (models_asr.py:30-36, audio_feat_preprocessing.py:49-50)
Every launch is timed alone with CUDA events after a 256 MB write that evicts the 126 MB L2; achieved GB/s uses the
ALGORITHMIC fp32-interface bytes of SURVEY 8d (spec: 4N + 3*4*T*F; fbanks: 4N + 4*T*80).

    python profiles/bench_frontend.py [--out gpurun_out/frontend_sweep.json]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from avsi_b200 import audio_processing as ap  # noqa: E402


def peak_gbs():
    try:
        with open(os.path.join(os.path.dirname(__file__), '..', 'MEASURED_PEAKS.json')) as f:
            j = json.load(f)
        for k in ('hbm_gbs', 'hbm_gbps', 'hbm_gb_s'):
            if k in j:
                return float(j[k])
        for v in j.values():
            if isinstance(v, dict):
                for k2, v2 in v.items():
                    if 'hbm' in k2.lower() and isinstance(v2, (int, float)):
                        return float(v2)
    except Exception:
        pass
    return 6542.4


def time_launch(fn, flush, reps):
    ts = []
    for _ in range(reps):
        flush.add_(1.0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    p = argparse.ArgumentParser()
    p.add_argument('--out', default='gpurun_out/frontend_sweep.json')
    p.add_argument('--max-gb', type=float, default=40.0)
    a = p.parse_args()
    dev = torch.device('cuda:0')
    torch.cuda.set_device(dev)
    peak = peak_gbs()
    F, frame_len, hop = 257, 384, 192
    mean = torch.full((F,), 6.0, device=dev)
    std = torch.full((F,), 2.0, device=dev)
    mel = torch.tensor(ap.linear_to_mel_weight_matrix(), dtype=torch.float32, device=dev)
    flush = torch.zeros(64 << 20, device=dev)
    rows = []
    g = torch.Generator(device=dev).manual_seed(0)
    for secs in (1, 3, 10, 20, 60):
        for B in (1, 8, 64, 512, 4096):
            N = 16000 * secs
            T = -(-N // hop)
            ldx = 320
            need = 4.0 * B * N + 4.0 * B * T * F * 2 + 2.0 * B * T * ldx
            if need > a.max_gb * 1e9:
                continue
            wav = torch.randn(B, N, device=dev, generator=g).mul_(3500.0).round_().clamp_(-32767, 32767)
            mask = torch.ones(B, T, F, device=dev)
            mask[:, T // 3: T // 3 + max(1, T // 8)] = 0
            xh = torch.zeros(T * B, ldx, dtype=torch.float16, device=dev)
            reps = 5 if B * secs >= 512 else 20

            def spec():
                ap.fused_features(wav, frame_len, hop, T=T, F=F, mean=mean, std=std, mask=mask, power=1.0, log=True,
                                  want_spec=True, xh_out=xh, ldx=ldx, xh_skip_pad=True)

            def fbanks():
                ap.fused_features(wav, frame_len, hop, T=T, F=F, power=2.0, log=False, want_spec=False, mel=mel)
            for name, fn, nbytes in (('spec', spec, 4 * N + 3 * 4 * T * F), ('fbanks', fbanks, 4 * N + 4 * T * 80)):
                for _ in range(3):
                    fn()
                torch.cuda.synchronize()
                ms = time_launch(fn, flush, reps)
                gbs = B * nbytes / ms / 1e6
                rows.append(dict(variant=name, B=B, seconds=secs, T=T, ms=round(ms, 4), alg_bytes_per_utt=nbytes,
                                 gb_per_s=round(gbs, 1), frac_of_measured_hbm_peak=round(gbs / peak, 3),
                                 utt_per_s=round(B / ms * 1e3, 1), frames_per_us=round(B * T / ms / 1e3, 2)))
                print(json.dumps(rows[-1]), flush=True)
            del wav, mask, xh
            torch.cuda.empty_cache()
    os.makedirs(os.path.dirname(a.out) or '.', exist_ok=True)
    with open(a.out, 'w') as f:
        json.dump(dict(peak_gb_per_s=peak, l2_flush='256 MB write before every timed launch', rows=rows), f, indent=1)


if __name__ == '__main__':
    main()
