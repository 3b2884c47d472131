"""Calibration of ncu's dram__bytes_* counters on this box: a plain 1 GiB copy and a 1 GiB read, whose algorithmic
traffic is known exactly.  Run under `ncu --set full` (tools/ncu_calib.sh) and compare."""
import torch
a = torch.empty(1 << 28, dtype=torch.float32, device='cuda').normal_()
b = torch.empty_like(a)
torch.cuda.synchronize()
for _ in range(2):
    b.copy_(a)
    s = a.sum()
torch.cuda.synchronize()
print('done', float(s))
