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
if os.environ.get("DE6D_TRACE"):   # entry-point sequence of one step, for scripts/ncu_traffic.py
    import json
    from de6d_b200 import _lib
    _lib.trace_begin()
    op.step()
    tr = _lib.trace_end()
    seq = []
    for name, a, _ in tr:
        shape = []
        for x in a:
            if x is None or (isinstance(x, int) and abs(x) >= (1 << 31)):
                break
            shape.append(round(x, 4) if isinstance(x, float) else x)
        seq.append([name, shape])
    json.dump(seq, open(os.environ["DE6D_TRACE"], "w"))
for _ in range(steps):
    op.step()
torch.cuda.synchronize()
print("done")
