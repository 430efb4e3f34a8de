"""Relative-L2 error of the CUDA model against the float64 oracle as a function of the chain length T
(VERDICT round 1, item 1): outputs and gradients of AV-SI and AV-MTL-SI at T = 15 / 60 / 250 / 1667 through the
tcgen05 recurrence kernels (forced at B = 3) and through the small-batch mma.sync kernels.

    python profiles/parity_vs_T.py > profiles/r02_parity_vs_T.json       (on the B200 box)
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))


def main():
    from avsi_b200 import _lib
    from test_gpu_bench_shapes import chain_parity
    rows = []
    for path in ('l4', 'mma'):
        _lib.set_env(AVSI_LSTM_FWD=path, AVSI_LSTM_BWD=path)
        for model in ('av-blstm', 'av-blstm-ssnn-ctc'):
            for T in (15, 60, 250, 1667):
                if model.endswith('ctc') and T < 49:
                    continue                                   # fewer frames than CTC states of a 24-phone label
                r = chain_parity(model, 3, T * 192, seed=41)
                rows.append({'kernels': 'tcgen05 (lstm4_*)' if path == 'l4' else 'mma.sync (lstm_*)', 'model': model, 'B': 3,
                             'T': T, 'pred_rel_l2': r['pred'], 'loss_rel': r['loss'], 'grad_rel_l2': r['grad'],
                             'worst_variable_grad_rel_l2': r['worst'], 'worst_variable': r['worst_name']})
                print(json.dumps(rows[-1]), file=sys.stderr, flush=True)
    print(json.dumps({'tolerance': 2e-3, 'rows': rows}, indent=1))


if __name__ == '__main__':
    main()
