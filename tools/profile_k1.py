"""Driver for the ncu capture of the K1 lane kernel: two 4e6-prefix launches of the bench workload."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import demi_b200 as D
from demi_b200 import _native as N

eng = D.Engine(D.SchedulerConfig(N.MODEL_RAFT5, model_flags=1))
eng.set_externals(D.raft5_program())
n = 4_000_000
out = torch.empty(n * 32, dtype=torch.uint8, device="cuda")
for rep in range(2):
    eng.fuzz_batch_dev(1 + rep * n, n, 50, 5, out.data_ptr())
torch.cuda.synchronize()
print("done")
