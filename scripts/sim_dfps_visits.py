import sys, numpy as np
sys.path.insert(0, '/root/repo')
from de6d_b200 import synth
def part1by2(x):
    x = x & 0x3ff
    x = (x | (x << 16)) & 0x030000ff
    x = (x | (x << 8)) & 0x0300f00f
    x = (x | (x << 4)) & 0x030c30c3
    x = (x | (x << 2)) & 0x09249249
    return x
kind = sys.argv[1] if len(sys.argv) > 1 else "uniform"
N, M, NW = 16384, 4096, 16
xyz = (synth.clouds(1, N, seed=0) if kind == "uniform" else synth.lidar_clouds(1, N, seed=0))[0].astype(np.float64)
lo = xyz.min(0); ext = (xyz.max(0) - lo).max()
q = np.clip(((xyz - lo) * (1023.0 / ext)).astype(np.int64), 0, 1023)
key = part1by2(q[:, 0]) | (part1by2(q[:, 1]) << 1) | (part1by2(q[:, 2]) << 2)
order = np.lexsort((np.arange(N), key))
P = xyz[order]                         # sorted positions
nb = N // 32
bl = P.reshape(nb, 32, 3).min(1); bh = P.reshape(nb, 32, 3).max(1)
temp = np.full(N, 1e10)
cur = np.where(order == 0)[0][0]
visits = []   # per sample: active bucket ids
sel = [cur]
for it in range(1, M):
    s = P[cur]
    g = np.maximum(np.maximum(bl - s, s - bh), 0.0)
    lb = (g * g).sum(1)
    bmax = temp.reshape(nb, 32).max(1)
    act = np.where(lb < bmax)[0]
    visits.append(act)
    for b in act:
        sl = slice(b * 32, b * 32 + 32)
        d = ((P[sl] - s) ** 2).sum(1)
        temp[sl] = np.minimum(temp[sl], d)
    cur = int(np.argmax(temp)); sel.append(cur)
print(kind, "visits per sample: mean %.2f" % np.mean([len(v) for v in visits]))
# rounds of 4 consecutive samples (approximation of the multi-sample rounds): union of active buckets
def stats(mapping, name):
    mx = []; mean = []
    for r in range(0, len(visits) - 3, 4):
        act = np.unique(np.concatenate(visits[r:r + 4]))
        w = mapping(act)
        cnt = np.bincount(w, minlength=NW)
        mx.append(cnt.max()); mean.append(cnt.mean())
    print("%-28s per round: mean/warp %.2f  max/warp %.2f  (ideal %.2f)" % (name, np.mean(mean), np.mean(mx), np.mean(np.ceil(np.array(mean)))))
stats(lambda b: b % NW, "round-robin b % 16")
stats(lambda b: (b + b // NW) % NW, "skewed (b + b/16) % 16")
stats(lambda b: (b * 7) % NW, "stride 7")
stats(lambda b: (b // 2) % NW, "pairs (b/2) % 16")
rng = np.random.default_rng(0); perm = rng.integers(0, NW, nb)
stats(lambda b: perm[b], "random")
# single-sample rounds
def stats1(mapping, name):
    mx = [np.bincount(mapping(v), minlength=NW).max() if len(v) else 0 for v in visits]
    print("%-28s per sample: max/warp %.2f" % (name, np.mean(mx)))
stats1(lambda b: b % NW, "round-robin")
