import sys, os, time, json
sys.path.insert(0, os.environ.get('GRAFT_REPO_ROOT', '/root/repo'))
import torch
import bench
from avsi_b200 import _lib
dev = torch.device('cuda:0')
wl = bench.Workload('av-blstm', 2048, 48000, None, dev, 0, 1)
for _ in range(3): wl.step_resident()
torch.cuda.synchronize()
def run(n, sleep):
    out = []
    for i in range(n):
        _lib.profile_start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); wl.step_resident(); e1.record()
        prof = _lib.profile_stop()
        out.append((round(e0.elapsed_time(e1), 2), round(prof['lstm_fwd']['ms'], 2), round(prof['lstm_bwd']['ms'], 2), round(prof['gemm_dw']['ms'], 2)))
        if sleep: time.sleep(sleep)
    return out
# profile_stop synchronises after every step: steps are NOT back to back here
print('sync each step     ', run(6, 0))
print('sync + 50 ms sleep ', run(6, 0.05))
# back to back: 8 steps, events per step, one sync at the end
ev = [torch.cuda.Event(enable_timing=True) for _ in range(9)]
ev[0].record()
for i in range(8):
    wl.step_resident(); ev[i + 1].record()
torch.cuda.synchronize()
print('back to back       ', [round(ev[i].elapsed_time(ev[i + 1]), 2) for i in range(8)])
import subprocess
print(subprocess.run('nvidia-smi --query-gpu=power.draw,power.limit,clocks.sm,clocks.mem,clocks_throttle_reasons.active --format=csv,noheader', shell=True, capture_output=True, text=True).stdout)
