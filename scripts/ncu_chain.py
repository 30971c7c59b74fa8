"""Workload for ncu: the bench op chain, eager, one stream (so every kernel appears as its own launch), batch 64,
one warm-up step + two steps.  A number printed by a run under ncu is never a bench value."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from de6d_b200 import chain as ch  # noqa: E402

batch = int(os.environ.get("DE6D_BATCH", "64"))
steps = int(os.environ.get("DE6D_STEPS", "2"))
cfg = ch.ChainConfig()
host = ch.make_inputs(cfg, batch, seed=0)
op = ch.OpChain(cfg, batch, use_graph=False, serial=True)
op.load(host)
op.capture()
for _ in range(steps):
    op.step()
torch.cuda.synchronize()
print("done")
