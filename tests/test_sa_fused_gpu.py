"""Fused set-abstraction scale (csrc/sa_mlp.cu: grouping -> shared MLP on tcgen05 tf32 tensor cores -> idx_cnt mask -> max-pool)
against the reference composition groupers[i] -> mlps[i] -> mask -> max_pool2d of pointnet2_modules.py:461-478.

Two comparisons per shape:
  * EXACT ARITHMETIC CHECK: a float64 evaluation of the same network whose operands are reduced to tf32 exactly where the
    kernel / tensor core reduces them (BN-folded weights: cvt.rna when packed; inputs and hidden activations: truncation
    to the upper 19 bits, what tcgen05.mma kind::tf32 reads).  Products of tf32 numbers are exact in
    fp32 and the kernel accumulates in fp32, so it must agree to fp32 accumulation noise (median error <= 1e-5 of the largest
    activation; single layer: maximum <= 2e-5) plus, in deeper nets, rare tf32 rounding flips of hidden activations (maximum
    <= 5e-4).  This pins the gather, the layer chaining through tensor memory, the mask and the pooling.
  * REFERENCE CHECK: torch's own Conv2d / BatchNorm2d(eval) / ReLU / max_pool2d in fp32 (allow_tf32 off).  The kernel
    computes in tf32 like the reference's default cuDNN path (torch.backends.cudnn.allow_tf32 = True), so the bar is the
    tf32 one: 1e-3 of the largest activation per layer of depth (north_star allows tensor cores only here).
"""
import copy
import warnings

import numpy as np
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from de6d_b200 import synth

pytestmark = pytest.mark.gpu


def cu(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def tf32(x):
    """cvt.rna.tf32.f32 on a float32 tensor: round to nearest, ties away, 10-bit mantissa."""
    u = x.contiguous().view(torch.int32)
    r = ((u + 0x1000) & ~0x1FFF).view(torch.float32)
    return torch.where(torch.isfinite(x), r, x)


def tf32_rz(x):
    """What the tensor core reads from an fp32 operand: the upper 19 bits (truncation)."""
    return (x.contiguous().view(torch.int32) & ~0x1FFF).view(torch.float32)


def make_mlp(widths, seed):
    g = torch.Generator().manual_seed(seed)
    mods = []
    for cin, cout in zip(widths[:-1], widths[1:]):
        conv = nn.Conv2d(cin, cout, kernel_size=1, bias=False)
        bn = nn.BatchNorm2d(cout)
        with torch.no_grad():
            conv.weight.copy_(torch.randn(conv.weight.shape, generator=g) * (1.5 / cin ** 0.5))
            bn.weight.copy_(torch.rand(cout, generator=g) + 0.5)
            bn.bias.copy_(torch.randn(cout, generator=g) * 0.3)
            bn.running_mean.copy_(torch.randn(cout, generator=g) * 0.2)
            bn.running_var.copy_(torch.rand(cout, generator=g) + 0.5)
        mods += [conv, bn, nn.ReLU()]
    return nn.Sequential(*mods).cuda().eval()


def emulate(layers, grouped_kfirst, mask):
    """float64 network on tf32-rounded operands.  grouped_kfirst (B, K, M, ns) fp32 in the reference channel order."""
    x = tf32_rz(grouped_kfirst).double()                                     # activations: truncated by the tensor core
    for li, (w, b) in enumerate(layers):
        y = torch.einsum("ok,bkms->boms", tf32(w).double(), x) + b.double()[None, :, None, None]   # weights: rounded when packed
        y = torch.relu(y.float())          # the kernel adds the bias and applies ReLU in fp32
        x = tf32_rz(y).double() if li + 1 < len(layers) else y.double()
    x = x * mask[:, None, :, None].double()
    return x.max(dim=3).values.float()


SHAPES = [
    # (B, N, M, ns, widths [3+C, ...], radius)
    (2, 2048, 256, 32, [4, 16, 16, 32], 1.0),
    (2, 4096, 1024, 32, [3 + 64, 64, 64, 128], 1.2),
    (2, 4096, 1000, 64, [3 + 64, 64, 96, 128], 1.6),          # M not a multiple of 32, 96-wide hidden layer
    (3, 1024, 200, 16, [3 + 32, 32, 64], 2.0),                # two queries per warp
    (1, 1024, 96, 128, [3 + 16, 32, 32, 32, 48], 6.0),        # one query per tile, four layers
    (2, 512, 64, 8, [3 + 8, 16], 1.5),                        # single layer, 4 queries per warp
    (2, 3000, 500, 32, [3 + 0, 16, 32], 1.0),                 # no features: coordinates only
    (2, 2048, 300, 32, [3 + 5, 16, 32], 1.0),                 # C not a multiple of 4: scalar gather
    # weights too large for one SM (272 KB / 368 KB in tf32): a cta_group::2 pair holds half of every layer per CTA
    (2, 1024, 512, 32, [3 + 128, 128, 128, 256], 1.6),
    (3, 1024, 48, 16, [3 + 128, 128, 192, 256], 2.0),         # 9 work items: the last pair has one idle CTA
    (1, 2048, 1000, 64, [3 + 128, 128, 128, 256], 2.4),
    # weights beyond two SMs' shared memory (464 KB): two launches over 128-row blocks of the last layer, each a CTA pair
    (2, 1024, 512, 32, [3 + 128, 128, 256, 256], 4.8),
    (2, 1024, 100, 16, [3 + 64, 128, 256, 256], 2.0),
]


@pytest.mark.parametrize("B,N,M,ns,widths,radius", SHAPES, ids=[str(s[3]) + "x" + "-".join(map(str, s[4])) for s in SHAPES])
def test_fused_scale_vs_exact_tf32_and_torch(lib, B, N, M, ns, widths, radius):
    from de6d_b200 import pointnet2_utils as pu, sa_fused
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    C = widths[0] - 3
    xyz = cu(synth.clouds(B, N, seed=N)[:, :, :] * np.float32(0.25))         # denser cloud: balls of every fill level
    new_xyz = xyz[:, :M].contiguous() + 0.01
    new_xyz[:, -3:] += 500.0                                                   # empty balls: masked to zero
    feats = torch.randn(B, C, N, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1)) if C else None
    mlp = make_mlp(widths, seed=ns + len(widths))
    assert sa_fused.FusedSAScale.supported(mlp, ns)
    scale = sa_fused.FusedSAScale(radius, ns, mlp)
    got = scale(xyz, new_xyz, feats)
    torch.cuda.synchronize()
    assert int(scale.status.item()) == 0, "tensor-core wait timed out"
    idx_cnt, idx = pu.ball_query_cnt(radius, ns, xyz, new_xyz)
    assert int((idx_cnt == 0).sum()) >= 3 * B and int((idx_cnt > 0).sum()) > M // 2
    grouped = pu.group_concat(xyz, new_xyz, feats, idx)                         # (B, 3+C, M, ns), reference channel order
    mask = (idx_cnt > 0).float()
    with torch.no_grad():
        ref = F.max_pool2d(mlp(grouped) * mask[:, None, :, None], kernel_size=[1, ns]).squeeze(-1)
    exact = emulate(sa_fused.fold_mlp(mlp), grouped, mask)
    top = float(exact.abs().max())
    assert top > 0.1
    assert got.shape == ref.shape == (B, widths[-1], M)
    d_exact = (got - exact).abs()
    err_exact, med_exact = float(d_exact.max()), float(d_exact.median())
    err_ref = float((got - ref).abs().max())
    # single layer: fp32 accumulation noise only.  Deeper nets: a hidden activation that lands within fp32 noise of a tf32
    # rounding boundary may round the other way than in the float64 evaluation (one tf32 ulp = 2^-10 of that activation, in
    # ~1e-3 of the activations), so the MAXIMUM carries those flips while the median stays at accumulation noise.
    assert med_exact <= 1e-5 * top, "vs tf32-exact evaluation (median): %g of %g" % (med_exact, top)
    assert err_exact <= (2e-5 if len(widths) == 2 else 5e-4) * top, "vs tf32-exact evaluation (max): %g of %g" % (err_exact, top)
    assert err_ref <= 1e-3 * (len(widths) - 1) * top, "vs torch fp32: %g of %g" % (err_ref, top)
    assert float(got[:, :, -3:].abs().max()) == 0.0                             # empty balls
    assert bool((got >= 0).all())


def test_fused_sa_module_matches_reference_module(lib):
    """fuse_sa_module on the UNMODIFIED reference PointnetSAModuleFSMSG: same new_xyz bit for bit, new_features / scores
    to tf32 tolerance; scales whose weights do not fit run the reference composition."""
    from oracle import build_ref, ref_py
    if not ref_py.available():
        pytest.skip("oracle/_ref/py not staged")
    from de6d_b200 import compat, sa_fused
    warnings.filterwarnings("ignore")
    compat.install()
    mods = ref_py.load_tree("pcdet", None).pointnet2_modules
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    kw = dict(npoint_list=[256, 256], sample_range_list=[[0, 1024], [0, 1024]], sample_method_list=["f-fps", "d-fps"],
              radii=[0.8, 1.6, 3.2], nsamples=[16, 32, 32], mlps=[[64, 64, 64, 128], [64, 64, 96, 128], [64, 128, 256, 256]],
              aggregation_mlp=[128], confidence_mlp=[64])
    torch.manual_seed(3)
    m = mods.PointnetSAModuleFSMSG(**copy.deepcopy(kw)).cuda()
    for mod in m.modules():                                                      # non-trivial BatchNorm statistics
        if isinstance(mod, (nn.BatchNorm1d, nn.BatchNorm2d)):
            with torch.no_grad():
                mod.running_mean.normal_(0, 0.2); mod.running_var.uniform_(0.5, 1.5); mod.weight.uniform_(0.5, 1.5); mod.bias.normal_(0, 0.2)
    m.eval()
    fwd = sa_fused.fuse_sa_module(m)
    assert fwd.fused == [True, True, True]                                       # 67->128->256->256: the last layer in two 128-row launches
    assert [sc.parts for sc in fwd.scales] == [1, 1, 2]
    B, N = 3, 1024
    xyz = cu(synth.clouds(B, N, seed=5) * np.float32(0.2))
    feats = torch.randn(B, 64, N, device="cuda", generator=torch.Generator(device="cuda").manual_seed(2))
    with torch.no_grad():
        want = m(xyz, feats)
        got = fwd(xyz, feats)
    assert torch.equal(got[0], want[0])
    for a, b, name in ((got[1], want[1], "new_features"), (got[2], want[2], "new_scores")):
        top = float(b.abs().max())
        assert float((a - b).abs().max()) <= 5e-3 * top, name


def test_fused_scale_rejects_what_it_cannot_run(lib):
    from de6d_b200 import sa_fused
    assert sa_fused.FusedSAScale.supported(make_mlp([131, 128, 128, 256], 0), 32)          # via a CTA pair
    assert sa_fused.FusedSAScale.supported(make_mlp([131, 128, 256, 256], 0), 32)          # weights > two SMs' shared memory: last layer split
    assert sa_fused.FusedSAScale.plan([131, 128, 256, 256], 32) == (2, 2)
    assert not sa_fused.FusedSAScale.supported(make_mlp([259, 256, 256, 512], 0), 16)      # vote head: > 256 wide
    assert not sa_fused.FusedSAScale.supported(make_mlp([35, 24, 32], 0), 32)              # width not a multiple of 16
    assert not sa_fused.FusedSAScale.supported(make_mlp([35, 32, 32], 0), 24)              # nsample not a power of two
    with pytest.raises(ValueError):
        sa_fused.FusedSAScale(1.0, 32, make_mlp([259, 256, 256, 512], 0))
    train = make_mlp([35, 32], 0).train()
    assert not sa_fused.FusedSAScale.supported(train, 32)
