"""GPU parity tests, op level: every libdram_b200 kernel (called through the C ABI via dram_native.ops /
dram_native.functional) against the matching torch fp32 CPU computation — the same library calls the reference makes
(F.conv3d, F.batch_norm, F.max_pool3d, F.interpolate(trilinear, align_corners=True), ...) — or against the oracle.

Tolerances (max-abs error relative to max |ref|):
  fp32 CUDA-core kernels and elementwise kernels : 2e-5
  split-bf16 (bf16x3) tensor-core convolutions   : 1e-4   (operand error ~2^-16; north_star bound is 1e-3)
  single-pass bf16 tensor-core convolutions      : 2e-2   (fast mode, reported separately)
  index / mask work                              : bit-exact
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from util import assert_close, cuda_cl

pytestmark = pytest.mark.gpu

TOL_F32 = 2e-5
TOL_X3 = 1e-4
TOL_BF16 = 2e-2


@pytest.fixture(autouse=True)
def _seed():
    torch.manual_seed(1234)
    np.random.seed(1234)


def ops():
    from dram_native import ops as o
    return o


def DF():
    from dram_native import functional as f
    return f


# ------------------------------------------------------------------------------------------------ library / layout
def test_library_loads_on_device():
    from dram_native import lib
    L = lib.load()
    assert L.dram_sm_arch() == 100
    assert L.dram_device_check() == 0, L.dram_last_error()


def test_errors_are_loud():
    from dram_native import lib
    o = ops()
    with pytest.raises(lib.DramLibraryError):
        o.to_cl(torch.zeros(1, 2, 4, 4, 4))                       # CPU tensor: no fallback
    x = torch.zeros(1, 3, 4, 4, 4, device="cuda")
    with pytest.raises(lib.DramLibraryError):                      # kernel size 5 unsupported
        lib.check(lib.load().dram_conv3d_simt_fwd(x.data_ptr(), x.data_ptr(), 0, x.data_ptr(), 1, 4, 4, 4, 3, 3, 5, 0), "conv")


def test_layout_round_trip():
    o = ops()
    x = torch.randn(2, 5, 3, 4, 6)
    xc = o.to_cl(x.cuda())
    assert o.is_cl(xc) and torch.equal(xc.cpu(), x)
    assert torch.equal(xc.permute(0, 2, 3, 4, 1).contiguous().cpu(), x.permute(0, 2, 3, 4, 1).contiguous())
    back = o.to_ncdhw(xc)
    assert back.is_contiguous() and torch.equal(back.cpu(), x)


# ------------------------------------------------------------------------------------------------ CUDA-core convolution
@pytest.mark.parametrize("N,Cin,Cout,S,k,bias", [(2, 1, 32, (8, 9, 10), 3, False), (1, 5, 7, (6, 5, 7), 3, True),
                                                 (2, 8, 16, (8, 8, 8), 3, False), (2, 12, 8, (5, 6, 7), 1, True),
                                                 (1, 64, 1, (6, 6, 6), 1, True),
                                                 (2, 1, 32, (5, 6, 12), 3, True), (1, 1, 32, (4, 4, 4), 3, False)])   # quad-voxel first layer
def test_conv_simt_forward(N, Cin, Cout, S, k, bias):
    o = ops()
    x, w = torch.randn(N, Cin, *S), torch.randn(Cout, Cin, k, k, k) * 0.2
    b = torch.randn(Cout) if bias else None
    ref = F.conv3d(x, w, b, padding=k // 2)
    pack = o.pack_weight_f32(w.cuda(), 0)
    y = o.conv_simt(cuda_cl(x), pack, b.cuda() if bias else None, Cout, k)
    assert_close(y, ref, TOL_F32, "conv_simt fwd")


@pytest.mark.parametrize("N,Cin,Cout,S,k", [(2, 1, 32, (8, 9, 10), 3), (2, 6, 10, (6, 5, 7), 3), (2, 12, 8, (5, 6, 7), 1),
                                            (2, 1, 32, (6, 5, 12), 3), (1, 1, 32, (4, 4, 4), 3)])   # quad-voxel first layer
def test_conv_simt_dgrad_wgrad(N, Cin, Cout, S, k):
    o = ops()
    x = torch.randn(N, Cin, *S, requires_grad=True)
    w = (torch.randn(Cout, Cin, k, k, k) * 0.2).requires_grad_(True)
    y = F.conv3d(x, w, None, padding=k // 2)
    dy = torch.randn_like(y)
    y.backward(dy)
    dx = o.conv_simt(cuda_cl(dy), o.pack_weight_f32(w.detach().cuda(), 1), None, Cin, k)
    dw = o.conv_simt_wgrad(cuda_cl(x.detach()), cuda_cl(dy), k)
    assert_close(dx, x.grad, TOL_F32, "conv_simt dgrad")
    assert_close(dw, w.grad, TOL_F32, "conv_simt wgrad")


# ------------------------------------------------------------------------------------------------ tensor-core convolution
UMMA_CASES = [
    # N, Cin, Cout, spatial, k
    (1, 64, 64, (16, 16, 16), 3),      # tile (16,8,1), BN=64
    (2, 32, 64, (8, 8, 16), 3),        # Cin padded 32 -> 64 (ds0.c1 shape class)
    (1, 192, 64, (8, 8, 8), 3),        # 3 K blocks per tap (us2.c0 shape class)
    (1, 64, 192, (8, 8, 8), 3),        # BN=96 (dgrad of us2.c0 shape class)
    (2, 128, 256, (10, 10, 10), 3),    # ragged tile (5,5,5), 2 N tiles (bg shape class)
    (1, 64, 32, (20, 20, 20), 3),      # BN=32 (dgrad of ds0.c1 shape class)
    (1, 128, 16, (6, 7, 9), 1),        # 1x1x1, odd spatial sizes
    (1, 64, 64, (4, 24, 48), 3),       # kw-reuse mode, 3x3 tiles of (16w x 8h), halo on every side
    (2, 192, 64, (3, 16, 32), 3),      # kw-reuse mode, 3 channel blocks, batch 2
    (1, 64, 32, (5, 8, 16), 3),        # kw-reuse mode with BN=32 (dgrad of ds0.c1 at 80^3); odd tile count (single-tile last item)
    (1, 64, 128, (6, 8, 24), 3),       # tile pairs, (TW,TDD) = (8,2), BN=128 (three N=128 MMAs per K step)
    (1, 128, 384, (2, 8, 16), 3),      # tile pairs, 3 N tiles of 128 (dgrad of us1.c0 shape class)
    (2, 64, 64, (16, 40, 40), 3),      # tile pairs at the 40^3 tiling, 200 items on 148 persistent CTAs
    (1, 384, 128, (4, 16, 8), 3),      # tile pairs, 6 channel blocks
]


@pytest.mark.parametrize("N,Cin,Cout,S,k", UMMA_CASES)
@pytest.mark.parametrize("three", [True, False])
def test_conv_umma_forward(N, Cin, Cout, S, k, three):
    o = ops()
    x, w = torch.randn(N, Cin, *S), torch.randn(Cout, Cin, k, k, k) * (2.0 / (Cin * k ** 3)) ** 0.5
    ref = F.conv3d(x, w, None, padding=k // 2)
    xs = o.split_bf16(cuda_cl(x), three)
    w_hi, w_lo, _ = o.pack_weight_bf16(w.cuda(), 0, three)
    y = o.conv_umma(xs, w_hi, w_lo, Cout, k)
    assert_close(y, ref, TOL_X3 if three else TOL_BF16, "conv_umma fwd")


@pytest.mark.parametrize("N,Cin,Cout,S", [
    (2, 64, 128, (8, 8, 16)),       # channels-on-M, 128-channel tiles: a thread owns a channel
    (2, 32, 64, (8, 8, 16)),        # channels-on-M, stacked 64-channel tiles: two warps share a channel
    (1, 64, 64, (16, 40, 40)),      # channels-on-M, (TW,TDD) = (8,4)
    (2, 192, 64, (3, 16, 32)),      # weight-sharing tile pairs (us2.c0 shape class): warp transpose-reduce, odd tile count
    (2, 128, 256, (10, 10, 10)),    # ragged tiles: no statistics epilogue -> None, the caller runs bn_stats
])
def test_conv_umma_forward_bn_statistics_epilogue(N, Cin, Cout, S):
    """train-mode BatchNorm sums (sum y, sum y^2 per channel) from the convolution's epilogue == sums of the stored y"""
    o = ops()
    x, w = torch.randn(N, Cin, *S), torch.randn(Cout, Cin, 3, 3, 3) * (2.0 / (Cin * 27)) ** 0.5
    xs = o.split_bf16(cuda_cl(x), True)
    w_hi, w_lo, _ = o.pack_weight_bf16(w.cuda(), 0, True)
    y, sums = o.conv_umma(xs, w_hi, w_lo, Cout, 3, want_stats=True)
    assert torch.equal(y, o.conv_umma(xs, w_hi, w_lo, Cout, 3)), "the statistics epilogue must not change y"
    if S == (10, 10, 10):
        assert sums is None
        return
    assert sums is not None and sums.dtype == torch.float64 and sums.shape == (2 * Cout,)
    yd = y.double()
    ref = torch.cat([yd.sum(dim=(0, 2, 3, 4)), (yd * yd).sum(dim=(0, 2, 3, 4))])
    scale = torch.cat([yd.abs().sum(dim=(0, 2, 3, 4)), (yd * yd).sum(dim=(0, 2, 3, 4))])      # fp32 partials of <= 256 values each
    assert ((sums - ref).abs() <= 1e-5 * scale).all(), ((sums - ref).abs() / scale).max().item()
    assert ((sums - o.bn_stats(y)).abs() <= 1e-5 * scale).all()
    assert torch.equal(sums, o.conv_umma(xs, w_hi, w_lo, Cout, 3, want_stats=True)[1]), "fixed summation order: deterministic"


@pytest.mark.parametrize("N,C1,C2,Cout,S", [
    (1, 128, 64, 64, (8, 8, 16)),       # us2.c0 shape class: weight-sharing tile pairs forward, kw-reuse wgrad
    (1, 256, 128, 128, (4, 8, 8)),      # us1.c0 shape class: channels-on-M forward, generic wgrad
    (2, 128, 64, 256, (5, 5, 5)),       # us0.c0 shape class: ragged tiles, generic forward and wgrad
    (1, 64, 32, 64, (8, 8, 16)),        # second operand with channel padding (32 -> 64)
])
def test_conv_umma_virtual_concat_equals_materialised_concat(N, C1, C2, Cout, S):
    """UpsampleConvBlock5d's cat([up, skip]) (parts.py:151-155) read through two tensor maps == the same convolution on the
    materialised concat: forward, eval-mode plane epilogue and weight gradient, bit for bit"""
    o = ops()
    a, b = torch.randn(N, C1, *S), torch.randn(N, C2, *S)
    w = torch.randn(Cout, C1 + C2, 3, 3, 3) * (2.0 / ((C1 + C2) * 27)) ** 0.5
    dy = torch.randn(N, Cout, *S)
    cat = o.split_bf16(cuda_cl(torch.cat([a, b], 1)), True)
    xa, xb = o.split_bf16(cuda_cl(a), True), o.split_bf16(cuda_cl(b), True)
    virt = o.SplitPlanes(xa.hi, xa.lo, (N, C1 + C2, *S), xa.Cpad + xb.Cpad, second=xb)
    assert cat.Cpad == virt.Cpad
    w_hi, w_lo, _ = o.pack_weight_bf16(w.cuda(), 0, True)
    assert torch.equal(o.conv_umma(virt, w_hi, w_lo, Cout, 3), o.conv_umma(cat, w_hi, w_lo, Cout, 3))
    y1, s1 = o.conv_umma(virt, w_hi, w_lo, Cout, 3, want_stats=True)
    y2, s2 = o.conv_umma(cat, w_hi, w_lo, Cout, 3, want_stats=True)
    assert (s1 is None) == (s2 is None) and (s1 is None or torch.equal(s1, s2))
    if Cout % 64 == 0:
        scale, shift = torch.rand(Cout, device="cuda") + 0.5, torch.randn(Cout, device="cuda") * 0.1
        p1 = o.conv_umma(virt, w_hi, w_lo, Cout, 3, scale, shift, out_planes=True)
        p2 = o.conv_umma(cat, w_hi, w_lo, Cout, 3, scale, shift, out_planes=True)
        assert torch.equal(p1.hi, p2.hi) and torch.equal(p1.lo, p2.lo)
    dys = o.split_bf16(cuda_cl(dy), True)
    assert torch.equal(o.conv_umma_wgrad(dys, virt, C1 + C2, Cout, 3), o.conv_umma_wgrad(dys, cat, C1 + C2, Cout, 3))
    assert_close(o.merge_planes(virt), torch.cat([a, b], 1), 1e-4, "merge of a virtual concat")


def test_upsample_virtual_concat_planes():
    """the upsampled half written alone + the skip planes as second operand == the materialised upsample+concat planes"""
    o = ops()
    x, sk = torch.randn(2, 64, 3, 4, 5), torch.randn(2, 24, 6, 8, 10)
    xs, ss = o.split_bf16(cuda_cl(x), True), o.split_bf16(cuda_cl(sk), True)
    assert o.virtual_concat_ok(xs, ss)
    v = o.upsample2x_virtual_concat(xs, ss)
    m = o.upsample2x_concat_planes(xs, ss)
    assert v.second is ss and v.shape == m.shape and v.Cpad == m.Cpad
    assert torch.equal(v.hi, m.hi[:, :64]) and torch.equal(v.lo, m.lo[:, :64])
    assert torch.equal(o.merge_planes(v), o.merge_planes(m))
    assert not o.virtual_concat_ok(o.split_bf16(cuda_cl(torch.randn(1, 32, 2, 2, 2)), True), o.split_bf16(cuda_cl(torch.randn(1, 8, 4, 4, 4)), True))


def test_conv_umma_forward_folded_bn_relu():
    o = ops()
    N, Cin, Cout, S = 1, 64, 64, (8, 8, 8)
    x, w = torch.randn(N, Cin, *S), torch.randn(Cout, Cin, 3, 3, 3) * 0.03
    scale, shift = torch.rand(Cout) + 0.5, torch.randn(Cout) * 0.1
    ref = F.relu(F.conv3d(x, w, None, padding=1) * scale.view(1, -1, 1, 1, 1) + shift.view(1, -1, 1, 1, 1))
    xs = o.split_bf16(cuda_cl(x), True)
    w_hi, w_lo, _ = o.pack_weight_bf16(w.cuda(), 0, True)
    y = o.conv_umma(xs, w_hi, w_lo, Cout, 3, scale.cuda(), shift.cuda())
    assert_close(y, ref, TOL_X3, "conv_umma fwd + folded BN/ReLU epilogue")


# gradient operand as ONE bf16 plane x split weights / split layer input (DRAM_BWD_PRECISION=bf16x2, ops.grad_planes_three):
# dy is rounded to bf16 (2^-9 per element, unbiased), everything else keeps 16+ bits
TOL_DY1 = 3e-3


@pytest.mark.parametrize("N,Cin,Cout,S,k", UMMA_CASES)
@pytest.mark.parametrize("dy_planes", [2, 1])
def test_conv_umma_dgrad(N, Cin, Cout, S, k, dy_planes):
    o = ops()
    if Cin % 16:
        pytest.skip("dgrad output channels must be a multiple of 16")
    x = torch.randn(N, Cin, *S, requires_grad=True)
    w = torch.randn(Cout, Cin, k, k, k) * (2.0 / (Cin * k ** 3)) ** 0.5
    y = F.conv3d(x, w, None, padding=k // 2)
    dy = torch.randn_like(y)
    y.backward(dy)
    dys = o.split_bf16(cuda_cl(dy), dy_planes == 2)
    w_hi, w_lo, _ = o.pack_weight_bf16(w.cuda(), 1, True)
    dx = o.conv_umma(dys, w_hi, w_lo, Cin, k)
    assert_close(dx, x.grad, TOL_X3 if dy_planes == 2 else TOL_DY1, "conv_umma dgrad")
    if dy_planes == 1:          # exactly the product with the bf16-rounded gradient: nothing else is lost
        ref = torch.nn.grad.conv3d_input(x.shape, w, dy.bfloat16().float(), padding=k // 2)
        assert_close(dx, ref, TOL_X3, "conv_umma dgrad vs bf16-rounded dy")


@pytest.mark.parametrize("N,Cin,Cout,S,k", UMMA_CASES + [
    (8, 256, 512, (10, 10, 10), 3), (3, 64, 64, (12, 20, 16), 3),
    (2, 128, 48, (6, 16, 24), 3),      # kw-reuse wgrad: Cout padded 48 -> 64, even number of (kd,kh,channel-block) sources
    (4, 64, 64, (8, 40, 40), 3),       # kw-reuse wgrad: several voxel slabs, odd number of sources (single-source last item)
])
@pytest.mark.parametrize("three", [True, False])
def test_conv_umma_wgrad(N, Cin, Cout, S, k, three):
    o = ops()
    x = torch.randn(N, Cin, *S)
    w = (torch.randn(Cout, Cin, k, k, k) * 0.05).requires_grad_(True)
    y = F.conv3d(x, w, None, padding=k // 2)
    dy = torch.randn_like(y)
    y.backward(dy)
    dw = o.conv_umma_wgrad(o.split_bf16(cuda_cl(dy), three), o.split_bf16(cuda_cl(x), three), Cin, Cout, k)
    assert_close(dw, w.grad, TOL_X3 if three else TOL_BF16, "conv_umma wgrad")
    if three:                   # single-plane gradient x split input (default backward)
        dw1 = o.conv_umma_wgrad(o.split_bf16(cuda_cl(dy), False), o.split_bf16(cuda_cl(x), True), Cin, Cout, k)
        assert_close(dw1, w.grad, TOL_DY1, "conv_umma wgrad, one-plane dy")
        ref = torch.nn.grad.conv3d_weight(x, w.shape, dy.bfloat16().float(), padding=k // 2)
        assert_close(dw1, ref, TOL_X3, "conv_umma wgrad vs bf16-rounded dy")


@pytest.mark.parametrize("N,Cin,Cout,S", [
    (1, 64, 64, (16, 16, 16)),      # BN = 64, two M tiles per CTA
    (1, 64, 64, (1, 8, 32)),        # BN = 64, 2 M tiles in all: one per CTA
    (1, 192, 64, (8, 8, 8)),        # (TW,TDD) = (8,2), 3 K blocks per tap (us2.c0 shape class)
    (1, 64, 192, (8, 8, 8)),        # Cout = 192: a launch of one 128-channel tile + a launch of the 64-channel tail (dgrad of us2.c0)
    (2, 64, 128, (8, 8, 16)),       # BN = 128
    (1, 128, 384, (2, 8, 16)),      # three 128-channel N tiles (dgrad of us1.c0 shape class)
    (2, 64, 64, (16, 40, 40)),      # 40^3 tiling, more work items than clusters
    (2, 64, 32, (8, 8, 16)),        # 32-channel tile (dgrad of ds0.c1): N = 64 / 32 MMAs
    (1, 64, 96, (4, 8, 16)),        # 64-channel + 32-channel tail tiles
    (1, 64, 224, (2, 8, 16)),       # 128 + 64 + 32
])
def test_conv_umma_sm_pair_kernel_matches_single_sm_kernels(monkeypatch, N, Cin, Cout, S):
    """k_conv_umma_fwd4 (cta_group::2) and the single-SM kernels it replaces (DRAM_CONV_V4=0: tile pairs / channels on M) compute
    the same three split products; only the order of the fp32 additions differs.  Also keeps the single-SM kernels tested."""
    o = ops()
    x, w = torch.randn(N, Cin, *S), torch.randn(Cout, Cin, 3, 3, 3) * (2.0 / (Cin * 27)) ** 0.5
    ref = F.conv3d(x, w, None, padding=1)
    xs = o.split_bf16(cuda_cl(x), True)
    w_hi, w_lo, _ = o.pack_weight_bf16(w.cuda(), 0, True)
    y4, s4 = o.conv_umma(xs, w_hi, w_lo, Cout, 3, want_stats=True)
    monkeypatch.setenv("DRAM_CONV_V4", "0")
    y1, s1 = o.conv_umma(xs, w_hi, w_lo, Cout, 3, want_stats=True)
    monkeypatch.delenv("DRAM_CONV_V4")
    assert_close(y4, ref, TOL_X3, "conv_umma fwd, SM pairs")
    assert_close(y1, ref, TOL_X3, "conv_umma fwd, single SM")
    assert_close(y4, y1.cpu(), 2e-5, "SM pairs vs single SM")
    assert s4 is not None and s1 is not None
    yd = y4.double()
    scale = torch.cat([yd.abs().sum(dim=(0, 2, 3, 4)), (yd * yd).sum(dim=(0, 2, 3, 4))])
    assert ((s4 - s1).abs() <= 1e-5 * scale).all()
    assert torch.equal(y4, o.conv_umma(xs, w_hi, w_lo, Cout, 3)), "deterministic, and the statistics epilogue does not change y"


@pytest.mark.parametrize("N,Cin,Cout,S", [(1, 64, 128, (6, 8, 24)), (2, 128, 256, (10, 10, 10)), (1, 384, 128, (4, 16, 8)),
                                          (2, 96, 128, (5, 6, 7))])       # 96 channels: M blocks 54 = 13 groups of four + two
def test_conv_umma_wgrad_sm_pair_kernel_matches_single_sm_kernel(monkeypatch, N, Cin, Cout, S):
    """k_conv_umma_wgrad2 (cta_group::2) against the generic single-SM wgrad (DRAM_WGRAD_V2=0) and against torch"""
    o = ops()
    x = torch.randn(N, Cin, *S)
    w = (torch.randn(Cout, Cin, 3, 3, 3) * 0.05).requires_grad_(True)
    y = F.conv3d(x, w, None, padding=1)
    dy = torch.randn_like(y)
    y.backward(dy)
    dys, xs = o.split_bf16(cuda_cl(dy), True), o.split_bf16(cuda_cl(x), True)
    d2 = o.conv_umma_wgrad(dys, xs, Cin, Cout, 3)
    monkeypatch.setenv("DRAM_WGRAD_V2", "0")
    d1 = o.conv_umma_wgrad(dys, xs, Cin, Cout, 3)
    monkeypatch.delenv("DRAM_WGRAD_V2")
    assert_close(d2, w.grad, TOL_X3, "wgrad, SM pairs")
    assert_close(d1, w.grad, TOL_X3, "wgrad, single SM")
    assert_close(d2, d1.cpu(), 2e-5, "SM pairs vs single SM")
    assert torch.equal(d2, o.conv_umma_wgrad(dys, xs, Cin, Cout, 3)), "deterministic"


@pytest.mark.parametrize("name,Cin,Cout,S", [("us2.c0", 192, 64, 80), ("us2.c1", 64, 64, 80), ("ds0.c1", 32, 64, 80),
                                             ("us1.c0", 384, 128, 40), ("ds1.c1", 64, 128, 40), ("us0.c1", 256, 256, 20)])
def test_conv_umma_full_size_batch8_adjoint_identities(name, Cin, Cout, S):
    """BASELINE's full size (batch 8, the benchmark's exact tile / cluster schedules), checked through size-independent
    properties instead of a CPU oracle: forward, dgrad and wgrad of one layer are three views of ONE trilinear form,
    <conv(x, w), g> == <x, dgrad(g, w)> == <w, wgrad(x, g)>, and the forward is linear in x."""
    o = ops()
    B = 8
    gen = torch.Generator(device="cuda").manual_seed(7)
    x = o.new_volume(B, Cin, S, S, S, "cuda").normal_(generator=gen)
    g = o.new_volume(B, Cout, S, S, S, "cuda").normal_(generator=gen)
    w = torch.randn(Cout, Cin, 3, 3, 3, device="cuda", generator=gen) * (2.0 / (Cin * 27)) ** 0.5
    xs, gs = o.split_bf16(x), o.split_bf16(g)
    wf, wd = o.pack_weight_bf16(w, 0), o.pack_weight_bf16(w, 1)
    y = o.conv_umma(xs, wf[0], wf[1], Cout, 3)
    dx = o.conv_umma(gs, wd[0], wd[1], Cin, 3)
    dw = o.conv_umma_wgrad(gs, xs, Cin, Cout, 3)
    assert torch.isfinite(y).all() and torch.isfinite(dx).all() and torch.isfinite(dw).all()
    a = (y.double() * g.double()).sum().item()
    b = (x.double() * dx.double()).sum().item()
    c = (w.double() * dw.double()).sum().item()
    scale = (y.double().abs() * g.double().abs()).sum().item()          # sum of |terms|: the three sums agree to ~1e-6 of it
    assert abs(a - b) <= 2e-6 * scale and abs(a - c) <= 2e-6 * scale, (name, a, b, c, scale)
    # linearity of the forward in x (split-bf16 operands: each term is rounded to ~2^-16 on its own)
    x2 = o.new_volume(B, Cin, S, S, S, "cuda").normal_(generator=gen)
    y2 = o.conv_umma(o.split_bf16(x2), wf[0], wf[1], Cout, 3)
    y12 = o.conv_umma(o.split_bf16(0.75 * x - 1.5 * x2), wf[0], wf[1], Cout, 3)
    ref = 0.75 * y - 1.5 * y2
    assert (y12 - ref).abs().max().item() <= 1e-4 * ref.abs().max().item()


@pytest.mark.parametrize("Cin,Cout,S", [(32, 64, (16, 16, 16)), (24, 64, (8, 8, 16)), (32, 64, (5, 5, 5)), (16, 32, (8, 8, 8))])
def test_conv_umma_skips_only_zero_padding(monkeypatch, Cin, Cout, S):
    """With one 64-channel K block the MMAs over the all-zero channel padding are not issued: same bits as issuing them"""
    o = ops()
    x, w = torch.randn(2, Cin, *S), torch.randn(Cout, Cin, 3, 3, 3) * 0.1
    xs = o.split_bf16(cuda_cl(x), True)
    w_hi, w_lo, _ = o.pack_weight_bf16(w.cuda(), 0, True)
    y = o.conv_umma(xs, w_hi, w_lo, Cout, 3)
    assert_close(y, F.conv3d(x, w, None, padding=1), TOL_X3, "conv_umma fwd, padded K block")
    monkeypatch.setenv("DRAM_CONV_FULL_K", "1")
    assert torch.equal(y, o.conv_umma(xs, w_hi, w_lo, Cout, 3))


@pytest.mark.parametrize("N,Cin,Cout,S", [(2, 64, 64, (16, 16, 16)),      # channels-on-M kernel, stacked 64-channel tile
                                          (1, 64, 128, (8, 16, 16)),     # channels-on-M kernel, 128-channel tile
                                          (1, 192, 64, (16, 16, 16)),    # tile-pair kernel
                                          (2, 128, 256, (5, 5, 5)),      # generic kernel, ragged tiles
                                          (1, 256, 512, (10, 10, 10))])
@pytest.mark.parametrize("three", [True, False])
def test_conv_umma_eval_epilogue_writes_planes(N, Cin, Cout, S, three):
    """inference epilogue: folded BatchNorm + ReLU written as split planes == conv -> bn_relu_apply_planes, bit for bit"""
    o = ops()
    x, w = torch.randn(N, Cin, *S), torch.randn(Cout, Cin, 3, 3, 3) * (2.0 / (27 * Cin)) ** 0.5
    scale, shift = (torch.rand(Cout) + 0.5).cuda(), torch.randn(Cout).cuda()
    xs = o.split_bf16(cuda_cl(x), three)
    w_hi, w_lo, _ = o.pack_weight_bf16(w.cuda(), 0, three)
    y = o.conv_umma(xs, w_hi, w_lo, Cout, 3)
    ref, _ = o.bn_relu_apply_planes(y, scale, shift, False)
    got = o.conv_umma(xs, w_hi, w_lo, Cout, 3, scale, shift, out_planes=True)
    assert got.Cpad == ref.Cpad == Cout and tuple(got.shape) == tuple(ref.shape)
    assert torch.equal(got.hi.view(torch.int16), ref.hi.view(torch.int16)), "hi plane"
    if three:
        assert torch.equal(got.lo.view(torch.int16), ref.lo.view(torch.int16)), "lo plane"
    else:
        assert got.lo is None
    assert_close(o.merge_planes(got), F.relu(F.conv3d(x, w, None, padding=1) * scale.cpu().view(1, -1, 1, 1, 1) + shift.cpu().view(1, -1, 1, 1, 1)),
                 TOL_X3 if three else TOL_BF16, "eval epilogue vs torch")


def test_conv_umma_is_deterministic():
    o = ops()
    x, w = torch.randn(2, 64, 8, 8, 8), torch.randn(64, 64, 3, 3, 3) * 0.05
    dy = torch.randn(2, 64, 8, 8, 8)
    xs, dys = o.split_bf16(cuda_cl(x), True), o.split_bf16(cuda_cl(dy), True)
    w_hi, w_lo, _ = o.pack_weight_bf16(w.cuda(), 0, True)
    a, b = o.conv_umma(xs, w_hi, w_lo, 64, 3), o.conv_umma(xs, w_hi, w_lo, 64, 3)
    assert torch.equal(a, b)
    g1, g2 = o.conv_umma_wgrad(dys, xs, 64, 64, 3), o.conv_umma_wgrad(dys, xs, 64, 64, 3)
    assert torch.equal(g1, g2)


# ------------------------------------------------------------------------------------------------ fused conv unit
def _unit_reference(x, w, b, gamma, beta, rm, rv, training, pool):
    y = F.conv3d(x, w, b, padding=w.shape[2] // 2)
    z = F.relu(F.batch_norm(y, rm, rv, gamma, beta, training, 0.1, 1e-5))
    return (z, F.max_pool3d(z, 2, 2, 0)) if pool else z


@pytest.mark.parametrize("Cin,Cout,S,k,bias,pool,training", [
    (1, 32, (8, 8, 8), 3, False, False, True),       # CUDA-core path (first layer)
    (64, 64, (8, 8, 8), 3, False, True, True),       # tensor-core path + pool
    (64, 64, (8, 8, 8), 3, False, True, False),      # eval mode
    (6, 10, (7, 9, 5), 3, False, True, True),        # ragged pool (floor) on the CUDA-core path
    (64, 8, (6, 6, 6), 1, True, False, True),        # 1x1x1 reshape head with conv bias (models.py:490)
])
def test_conv_bn_relu_unit(Cin, Cout, S, k, bias, pool, training):
    N = 2
    x = torch.randn(N, Cin, *S, requires_grad=True)
    w = (torch.randn(Cout, Cin, k, k, k) * (2.0 / (Cin * k ** 3)) ** 0.5).requires_grad_(True)
    b = torch.randn(Cout).requires_grad_(True) if bias else None
    gamma, beta = (torch.rand(Cout) + 0.5).requires_grad_(True), (torch.randn(Cout) * 0.1).requires_grad_(True)
    rm, rv = torch.randn(Cout) * 0.1, torch.rand(Cout) + 0.5
    rm_ref, rv_ref = rm.clone(), rv.clone()
    ref = _unit_reference(x, w, b, gamma, beta, rm_ref, rv_ref, training, pool)
    outs = ref if pool else (ref,)
    gouts = [torch.randn_like(t) for t in outs]
    torch.autograd.backward(outs, gouts)

    xg = cuda_cl(x.detach()).requires_grad_(True)
    wg = w.detach().cuda().requires_grad_(True)
    bg = b.detach().cuda().requires_grad_(True) if bias else None
    gg, beg = gamma.detach().cuda().requires_grad_(True), beta.detach().cuda().requires_grad_(True)
    rmg, rvg = rm.cuda(), rv.cuda()
    got = DF().conv_bn_relu(xg, wg, bg, gg, beg, rmg, rvg, training, 0.1, 1e-5, 1, pool)
    gots = got if pool else (got,)
    for t, r, name in zip(gots, outs, ("a", "pooled")):
        assert_close(t, r, TOL_X3, f"unit fwd {name}")
    torch.autograd.backward(gots, [g.cuda() for g in gouts])
    gs = (x.grad.abs().max() + 1e-12)
    assert_close(xg.grad, x.grad, 2e-4, "unit dx")
    assert_close(wg.grad, w.grad, 2e-4, "unit dw")
    assert_close(gg.grad, gamma.grad, 2e-4, "unit dgamma")
    assert_close(beg.grad, beta.grad, 2e-4, "unit dbeta")
    if bias and not training:
        assert_close(bg.grad, b.grad, 2e-4, "unit dbias")
    if training:
        assert_close(rmg, rm_ref, TOL_X3, "running_mean")
        assert_close(rvg, rv_ref, TOL_X3, "running_var")


def _ref_unit(x, p, training, pool=False):
    y = F.conv3d(x, p["w"], None, padding=1)
    z = F.relu(F.batch_norm(y, p["rm"], p["rv"], p["g"], p["b"], training, 0.1, 1e-5))
    return (z, F.max_pool3d(z, 2, 2, 0)) if pool else z


def _mk_unit(ci, co):
    return {"w": (torch.randn(co, ci, 3, 3, 3) * (2.0 / (ci * 27)) ** 0.5).requires_grad_(True),
            "g": (torch.rand(co) + 0.5).requires_grad_(True), "b": (torch.randn(co) * 0.1).requires_grad_(True),
            "rm": torch.randn(co) * 0.1, "rv": torch.rand(co) + 0.5}


def _to_gpu(p):
    return {k: (v.detach().cuda().requires_grad_(True) if v.requires_grad else v.clone().cuda()) for k, v in p.items()}


@pytest.mark.parametrize("training", [True, False])
@pytest.mark.parametrize("S", [(8, 8, 8), (6, 10, 8)])
def test_plane_carried_encoder_decoder(training, S):
    """A two-level U-Net slice whose activations exist only as bf16 split planes between the tensor-core convolutions:
    [unit+pool] -> [unit] -> upsample x2 + concat -> [unit] -> [unit + fused RAM head].  Exercises the plane-writing
    BN/ReLU(+pool) kernels, the fused max-pool + BN backward, the upsample+concat on planes, the in-place (pitched) read
    of the skip gradient and the RAM-fused last unit, against (a) the same torch ops the reference calls and (b) the
    same slice run through the fp32-activation kernels (identical split-bf16 operands, so identical ReLU masks).

    Gradients through train-mode BatchNorm at these tiny sizes hinge on single ReLU decisions: one activation within
    ~1e-5 of zero that lands on the other side in split-bf16 arithmetic moves a weight gradient by ~1/sqrt(voxels)
    = 3e-2.  The seeded (6,10,8) training case contains such an element after the upsample (the fp32-activation path
    deviates from torch by the same 2e-2 and from the plane path by 8e-3; the CUDA-core fp32 path does not), so that one
    case is bounded by 5e-2; every other case is held to 5e-4 against torch and 2e-4 against (b)."""
    df = DF()
    N = 2
    u1, u2, u3, u4 = _mk_unit(16, 32), _mk_unit(32, 64), _mk_unit(96, 32), _mk_unit(32, 32)
    wt, bt = (torch.randn(1, 32, 1, 1, 1) * 0.2).requires_grad_(True), torch.randn(1).requires_grad_(True)
    x = torch.randn(N, 16, *S, requires_grad=True)
    ref_p = [{k: (v.clone() if not v.requires_grad else v) for k, v in u.items()} for u in (u1, u2, u3, u4)]
    a, p = _ref_unit(x, ref_p[0], training, pool=True)
    q = _ref_unit(p, ref_p[1], training)
    cat = torch.cat([F.interpolate(q, scale_factor=2, mode="trilinear", align_corners=True), a], dim=1)
    f = _ref_unit(_ref_unit(cat, ref_p[2], training), ref_p[3], training)
    ram = F.conv3d(f, wt, bt)
    gout = torch.randn_like(ram)
    ram.backward(gout)
    ref = {"dx": x.grad, "dwt": wt.grad, "dbt": bt.grad}
    for i, u in enumerate((u1, u2, u3, u4)):
        ref.update({f"dw{i}": u["w"].grad, f"dg{i}": u["g"].grad, f"db{i}": u["b"].grad})

    def run_gpu(planes):
        gu = [_to_gpu(u) for u in (u1, u2, u3, u4)]
        wtg, btg = wt.detach().cuda().requires_grad_(True), bt.detach().cuda().requires_grad_(True)
        xg = cuda_cl(x.detach()).requires_grad_(True)
        unit = lambda t, u, **kw: df.conv_bn_relu(t, u["w"], None, u["g"], u["b"], u["rm"], u["rv"], training, 0.1, 1e-5, 1, **kw)
        ag, pg = unit(xg, gu[0], pool=True, out_planes=planes)
        qg = unit(pg, gu[1], out_planes=planes)
        catg = df.upsample_concat(qg, ag)
        hg = unit(catg, gu[2], out_planes=planes)
        if planes:
            assert isinstance(ag, df.Act) and isinstance(pg, df.Act) and isinstance(catg, df.Act)
            assert ag.t.untyped_storage().nbytes() <= 4, "the autograd handle of a plane-carried activation has no storage"
            assert_close(df.as_tensor(ag), a, TOL_X3, "a (planes)")
            assert_close(df.as_tensor(pg), p, TOL_X3, "pooled (planes)")
            assert_close(df.as_tensor(catg), cat, TOL_X3, "upsample+concat (planes)")
            assert df.ram_fusable(gu[3]["w"], wtg, None)
            ramg = df.conv_bn_relu_ram(hg, gu[3]["w"], gu[3]["g"], gu[3]["b"], gu[3]["rm"], gu[3]["rv"], training, 0.1, 1e-5,
                                       1, wtg, btg)
        else:
            ramg = df.RamReduce.apply(unit(hg, gu[3]), wtg, btg)
        ramg.backward(gout.cuda())
        out = {"ram": ramg.detach(), "dx": xg.grad, "dwt": wtg.grad, "dbt": btg.grad}
        for i, u in enumerate(gu):
            out.update({f"dw{i}": u["w"].grad, f"dg{i}": u["g"].grad, f"db{i}": u["b"].grad, f"rm{i}": u["rm"], f"rv{i}": u["rv"]})
        return out

    got, base = run_gpu(True), run_gpu(False)
    assert_close(got["ram"], ram, 2e-4, "fused RAM")
    flip = training and S == (6, 10, 8)
    for k, r in ref.items():
        assert_close(got[k], r, 5e-2 if flip else 5e-4, f"{k} vs torch")
        assert_close(got[k], base[k], 5e-2 if flip else 2e-4, f"{k} vs the fp32-activation path")
    if training:
        for i in range(4):
            assert_close(got[f"rm{i}"], ref_p[i]["rm"], TOL_X3, f"unit{i} running_mean")
            assert_close(got[f"rv{i}"], ref_p[i]["rv"], TOL_X3, f"unit{i} running_var")


def test_pool_backward_on_planes_routes_to_first_maximum():
    """all-zero windows after ReLU: the fused max-pool + BN backward must route like ATen (first element of the window)"""
    o = ops()
    N, C, S = 1, 8, (4, 6, 4)
    y = torch.randn(N, C, *S)
    y[:, :, :2] = -1.0                                  # whole windows clipped by the ReLU
    a = F.relu(y).requires_grad_(True)
    pl = F.max_pool3d(a, 2, 2, 0)
    ga, gp = torch.randn_like(a), torch.randn_like(pl)
    torch.autograd.backward([a, pl], [ga, gp])          # a.grad = ga + routed gp
    ref = a.grad * (y > 0)                              # ReLU backward; eval-mode BN with scale 1 / shift 0
    one, zero = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
    dys = o.bn_pool_bwd_apply_planes(cuda_cl(ga), C, cuda_cl(gp), cuda_cl(y), one, zero, None, None, None, None, 1.0)
    got = o.merge_planes(dys)
    assert_close(got, ref, 1e-5, "dy")


def test_bn_running_stats_double_update():
    """checkpointed blocks update the running statistics twice per step (models.py:123-143)."""
    o = ops()
    y = torch.randn(2, 8, 4, 4, 4)
    rm, rv = torch.zeros(8), torch.ones(8)
    for _ in range(2):
        F.batch_norm(y, rm, rv, None, None, True, 0.1, 1e-5)
    yg = cuda_cl(y)
    rmg, rvg = torch.zeros(8, device="cuda"), torch.ones(8, device="cuda")
    o.bn_finalize(o.bn_stats(yg), y.numel() // 8, None, None, rmg, rvg, 0.1, 1e-5, 2)
    assert_close(rmg, rm, TOL_F32, "running_mean x2")
    assert_close(rvg, rv, TOL_F32, "running_var x2")


def test_maxpool_backward_first_max_on_ties():
    """all-zero windows (common after ReLU) route the gradient to the first element like ATen."""
    o = ops()
    a = torch.zeros(1, 4, 4, 4, 4)
    a[0, :, 1, 1, 1] = 2.0
    a.requires_grad_(True)
    p = F.max_pool3d(a, 2, 2, 0)
    g = torch.randn_like(p)
    p.backward(g)
    da = torch.zeros(1, 4, 4, 4, 4, device="cuda").contiguous(memory_format=torch.channels_last_3d)
    o.maxpool2_bwd(cuda_cl(a.detach()), cuda_cl(g), da)
    assert torch.equal(da.cpu(), a.grad)


# ------------------------------------------------------------------------------------------------ decoder glue / resize
@pytest.mark.parametrize("C1,C2,s,skip", [(8, 4, (4, 5, 6), (8, 10, 12)), (16, 8, (3, 3, 3), (7, 6, 8)), (3, 5, (4, 4, 4), (8, 8, 8)),
                                          (32, 8, (4, 5, 6), (8, 10, 12)),          # tiled adjoint: one partial 8 x 8 tile
                                          (64, 32, (10, 10, 10), (20, 20, 20)),     # tiled adjoint: 2 x 2 tiles, 2 channel chunks, z segments
                                          (32, 4, (9, 3, 17), (18, 7, 34))])        # tiled adjoint: odd extents, cropped skip
def test_upsample_concat(C1, C2, s, skip):
    from oracle_import import O
    x = torch.randn(2, C1, *s, requires_grad=True)
    sk = torch.randn(2, C2, *skip, requires_grad=True)
    up = F.interpolate(x, scale_factor=(2, 2, 2), mode="trilinear", align_corners=True)
    ref = O.crop_concat(up, sk)
    g = torch.randn_like(ref)
    ref.backward(g)
    xg, skg = cuda_cl(x.detach()).requires_grad_(True), cuda_cl(sk.detach()).requires_grad_(True)
    got = DF().upsample_concat(xg, skg)
    assert_close(got, ref, TOL_F32, "upsample+concat fwd")
    got.backward(g.cuda())
    assert_close(xg.grad, x.grad, TOL_F32, "upsample+concat dx")
    assert torch.equal(skg.grad.cpu(), sk.grad), "skip gradient is a pure copy: bit-exact"


@pytest.mark.parametrize("C1,C2,s", [(32, 8, (4, 5, 6)), (128, 64, (12, 16, 20)), (32, 32, (2, 8, 8))])
def test_upsample_adjoint_tiled_kernel_matches_register_blocked_kernel(monkeypatch, C1, C2, s):
    """k_up2x_adjoint_tiled (shared-memory tiles, marching along z) == k_up2x_adjoint (DRAM_UP2X_ADJ_BLOCKED=1) up to the order
    of the fp32 additions, and both are the exact transpose of the forward: <up(x), g> == <x, adjoint(g)>"""
    o = ops()
    B = 2
    D, H, W = (2 * v for v in s)
    g = cuda_cl(torch.randn(B, C1 + C2, D, H, W))
    dx_t, _ = o.upsample2x_concat_bwd(g, (B, C1, *s), (B, C2, D, H, W), want_dskip=False)
    monkeypatch.setenv("DRAM_UP2X_ADJ_BLOCKED", "1")
    dx_b, _ = o.upsample2x_concat_bwd(g, (B, C1, *s), (B, C2, D, H, W), want_dskip=False)
    monkeypatch.delenv("DRAM_UP2X_ADJ_BLOCKED")
    assert_close(dx_t, dx_b.cpu(), TOL_F32, "tiled vs register-blocked adjoint")
    x = torch.randn(B, C1, *s)
    up = F.interpolate(x, scale_factor=(2, 2, 2), mode="trilinear", align_corners=True)
    lhs = (up.double() * g[:, :C1].cpu().double()).sum().item()
    rhs = (x.double() * dx_t.cpu().double()).sum().item()
    assert abs(lhs - rhs) <= 1e-6 * (up.double().abs() * g[:, :C1].cpu().double().abs()).sum().item(), (lhs, rhs)


@pytest.mark.parametrize("C,src,dst", [(1, (20, 20, 20), (16, 16, 16)), (1, (16, 16, 16), (20, 20, 20)), (8, (10, 12, 9), (16, 16, 16)),
                                        (17, (5, 6, 7), (5, 6, 7)), (4, (6, 6, 6), (1, 13, 6))])
def test_trilinear_resize(C, src, dst):
    x = torch.randn(2, C, *src, requires_grad=True)
    ref = F.interpolate(x, size=dst, mode="trilinear", align_corners=True)
    g = torch.randn_like(ref)
    ref.backward(g)
    xg = cuda_cl(x.detach()).requires_grad_(True)
    got = DF().TrilinearResize.apply(xg, dst)
    assert_close(got, ref, TOL_F32, "trilinear fwd")
    got.backward(g.cuda())
    assert_close(xg.grad, x.grad, TOL_F32, "trilinear bwd")


# ------------------------------------------------------------------------------------------------ RAM head
@pytest.mark.parametrize("C,O", [(64, 1), (8, 1), (64, 3), (6, 2)])
def test_ram_reduce(C, O):
    feat = torch.randn(2, C, 6, 7, 8, requires_grad=True)
    w = (torch.randn(O, C, 1, 1, 1) * 0.2).requires_grad_(True)
    b = torch.randn(O).requires_grad_(True)
    ref = F.conv3d(feat, w, b)
    g = torch.randn_like(ref)
    ref.backward(g)
    fg = cuda_cl(feat.detach()).requires_grad_(True)
    wg, bg = w.detach().cuda().requires_grad_(True), b.detach().cuda().requires_grad_(True)
    got = DF().RamReduce.apply(fg, wg, bg)
    assert_close(got, ref, TOL_F32, "ram reduce fwd")
    got.backward(cuda_cl(g))
    assert_close(fg.grad, feat.grad, TOL_F32, "ram reduce dfeat")
    assert_close(wg.grad, w.grad, TOL_F32, "ram reduce dw")
    assert_close(bg.grad, b.grad, TOL_F32, "ram reduce db")


def test_ram_reduce_fused_bn_relu():
    o = ops()
    y = torch.randn(2, 64, 5, 6, 7)
    scale, shift = torch.rand(64) + 0.5, torch.randn(64) * 0.2
    w, b = torch.randn(1, 64) * 0.2, torch.randn(1)
    a = F.relu(y * scale.view(1, -1, 1, 1, 1) + shift.view(1, -1, 1, 1, 1))
    ref = F.conv3d(a, w.view(1, 64, 1, 1, 1), b)
    got = o.ram_reduce(cuda_cl(y), w.cuda(), b.cuda(), scale.cuda(), shift.cuda())
    assert_close(got, ref, TOL_F32, "fused BN+ReLU+RAM reduce")


@pytest.mark.parametrize("use_sigmoid,gt0", [(False, False), (True, True)])
def test_masked_mean(use_sigmoid, gt0):
    from oracle_import import O
    x = torch.randn(3, 1, 8, 9, 10, requires_grad=True)
    lobes = O.ellipsoid_lobe(3, (8, 9, 10), seed=3)
    fx = torch.sigmoid(x) if use_sigmoid else x
    m = (lobes > 0).float() if gt0 else lobes
    ref = (fx * m).reshape(3, -1).sum(-1) / m.reshape(3, -1).sum(-1)
    g = torch.randn(3)
    ref.backward(g)
    xg = x.detach().cuda().requires_grad_(True)
    mean, cnt = DF().MaskedMean.apply(xg, lobes.cuda(), use_sigmoid, gt0)
    assert_close(mean, ref, TOL_F32, "masked mean")
    assert torch.equal(cnt.cpu(), m.reshape(3, -1).sum(-1)), "mask count is exact"
    mean.backward(g.cuda())
    assert_close(xg.grad, x.grad, TOL_F32, "masked mean bwd")


def test_fused_interval_regression_loss_matches_tensor_op_form(monkeypatch):
    """dram_int_reg_loss (IntRegHinge: two masked-pool reductions + one tiny kernel) against the tensor-op form of the same
    loss (DRAM_FUSED_LOSS=0: metrics.py:121-137, 158-177 op by op) and against the CPU oracle: value and gradient w.r.t. the
    RAM logits.  The samples cover every branch of get_labels: the ratio inside its band, the band cut by rub +- band_width on
    either side, an empty intersection above and below the band, ctss 0, and an inactive hinge (zero gradient)."""
    import metrics
    from oracle_import import O
    torch.manual_seed(11)
    B, S = 8, (8, 8, 8)
    lobes = O.ellipsoid_lobe(B, S, seed=5).cuda()
    dense = (torch.randn(B, 1, *S) * 2).cuda()
    dense[7] = -9.0                                              # mean probability ~1e-4: inside band 0 -> hinge inactive
    lesions = torch.zeros(B, 1, *S)
    for b, frac in enumerate([0.2, 0.02, 0.04, 0.6, 0.0, 0.3, 0.9, 0.0]):      # lesion-candidate ratio of each sample
        lesions[b].view(-1)[: int(frac * S[0] * S[1] * S[2])] = 1.0
    lesions = lesions.cuda()
    ctsses = ["3", "2", "2", "5", "0", "1", "3", "0"]            # bands 0.05-0.35, 0.01-0.05, ..., 0-0.001
    freq = {k: 0.1 + 0.12 * k for k in range(6)}

    class Host:
        ctss_frequency_map, debug_path, epoch_n = freq, "/tmp/x", 0

    loss_obj = metrics.IntRegRefineLoss(band_width=1e-2)
    out = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("DRAM_FUSED_LOSS", mode)
        d = dense.clone().requires_grad_(True)
        loss = loss_obj.compute_reg_loss_with_logits(d, lobes, lesions, ctsses, obj=Host())
        loss.backward()
        out[mode] = (loss.detach().cpu(), d.grad.cpu())
    monkeypatch.delenv("DRAM_FUSED_LOSS")
    assert out["1"][0].shape == out["0"][0].shape == ()
    assert abs(out["1"][0].item() - out["0"][0].item()) <= 1e-6 * abs(out["0"][0].item())
    assert_close(out["1"][1], out["0"][1], 1e-6, "d reg loss / d RAM, fused vs tensor ops")
    assert out["1"][1][7].abs().max().item() == 0.0 and out["1"][1][0].abs().max().item() > 0.0
    dc = dense.detach().cpu().requires_grad_(True)
    ref = O.reg_loss(dc, lobes.cpu(), lesions.cpu(), ctsses, freq, band_width=1e-2)
    ref.backward()
    assert abs(out["1"][0].item() - ref.item()) <= 1e-5 * abs(ref.item())
    assert_close(out["1"][1], dc.grad, TOL_F32, "d reg loss / d RAM vs oracle")


def test_fused_boot_bce_matches_unfused_loss():
    """dram_boot_bce_fwd/_bwd against the tensor-op BootBinCrossEntropy + pseudo_labels of metrics.py (same arithmetic as
    the reference's metrics.py:10-51,325-354): value and gradient w.r.t. the refined logits, incl. a sample with
    keep = 0, saturated logits (the eps clamp) and voxels outside every lobe."""
    import metrics
    from dram_native import functional as DF
    torch.manual_seed(5)
    B, S = 3, (6, 7, 8)
    dense = (torch.randn(B, 1, *S) * 2).cuda()
    refined = (torch.randn(B, 1, *S) * 3).cuda()
    refined[0, 0, 0, 0, :4] = torch.tensor([40.0, -40.0, 17.0, -17.0])          # p saturates: pt hits the clamp
    lobes = (torch.rand(B, 1, *S) > 0.3).float().cuda()
    lesions = (torch.rand(B, 1, *S) > 0.5).float().cuda()
    ctsses = ["3", "0", "5"]
    loss_obj = metrics.IntRegRefineLoss()
    keep = torch.tensor([1.0, 0.0, 1.0], device="cuda")
    r1 = refined.clone().requires_grad_(True)
    t = loss_obj.pseudo_labels(dense, lobes, lesions, ctsses, keep)
    ref = loss_obj.bootstrap_loss(torch.sigmoid(r1), t, lobes > 0)
    ref.backward()
    r2 = refined.clone().requires_grad_(True)
    got = DF.BootBce.apply(dense, r2, lobes, lesions, keep, 0.1, 1e-7)
    got.backward()
    assert abs(got.item() - ref.item()) <= 2e-6 * abs(ref.item()), (got.item(), ref.item())
    assert_close(r2.grad, r1.grad, 2e-5, "d seg_loss / d refined")
    # weighting by an upstream gradient
    r3 = refined.clone().requires_grad_(True)
    (DF.BootBce.apply(dense, r3, lobes, lesions, keep, 0.1, 1e-7) * 2.5).backward()
    assert_close(r3.grad, 2.5 * r1.grad, 2e-5, "scaled gradient")


def test_pooling_dense_features_matches_oracle():
    from oracle_import import O
    import models
    d = torch.randn(2, 1, 8, 8, 8)
    lobes = O.ellipsoid_lobe(2, (8, 8, 8), seed=5)
    got = models.pooling_dense_features(d.cuda(), lobes.cuda())
    assert_close(got, O.masked_pool(d, lobes), TOL_F32, "pooling_dense_features")


@pytest.mark.parametrize("act", [1, 2])
def test_ram_upsample_mask_scatter(act):
    """inference epilogue: values within fp32 tolerance, lobe masking / indexing bit-exact."""
    o = ops()
    ram = torch.randn(1, 1, 8, 8, 8)
    crop = (13, 9, 17)
    mask = (torch.rand(*crop) > 0.4).to(torch.uint8)
    scan_shape, off = (20, 16, 24), (3, 2, 5)
    if act == 1:
        up = F.interpolate(torch.sigmoid(ram), size=crop, mode="trilinear", align_corners=True)[0, 0]
    else:
        up = F.relu(F.interpolate(ram, size=crop, mode="trilinear", align_corners=True)[0, 0])
    ref = torch.full(scan_shape, -7.0)
    view = ref[off[0]:off[0] + crop[0], off[1]:off[1] + crop[1], off[2]:off[2] + crop[2]]
    view[mask > 0] = up[mask > 0]
    heat = torch.full(scan_shape, -7.0, device="cuda")
    mx = torch.zeros(1, device="cuda")
    o.ram_upsample_mask_scatter(ram[0, 0].cuda(), mask.cuda(), heat, off, act, 1.0, mx)
    touched_ref = ref != -7.0
    assert torch.equal((heat.cpu() != -7.0), touched_ref), "lobe mask / paste indices must be bit-exact"
    assert_close(heat.cpu()[touched_ref], ref[touched_ref], TOL_F32, "scatter values")
    assert abs(mx.item() - up.max().item()) <= 1e-5 * max(1.0, abs(up.max().item()))


# ------------------------------------------------------------------------------------------------ 1x1x1 reshape heads on planes
@pytest.mark.parametrize("N,C,S", [(2, 64, (5, 6, 7)), (1, 128, (4, 5, 9)), (3, 40, (3, 3, 3)), (1, 104, (2, 5, 31))])
@pytest.mark.parametrize("three", [True, False])
def test_pointwise8_planes(N, C, S, three):
    """Conv3d(C, 8, 1) straight from split planes (models.py:488-494) and its weight / bias gradient"""
    o = ops()
    x = torch.randn(N, C, *S)
    w, b = torch.randn(8, C, 1, 1, 1) / C ** 0.5, torch.randn(8)
    xs = o.split_bf16(cuda_cl(x), three)
    assert o.pointwise8_ok(xs, 8, 1) and not o.pointwise8_ok(xs, 16, 1) and not o.pointwise8_ok(xs, 8, 3)
    xr = o.merge_planes(xs).cpu()                               # what the planes hold (== x to 2^-16 when three)
    assert_close(xr, x, TOL_X3 if three else TOL_BF16, "planes")
    ref = F.conv3d(xr, w, b)
    y = o.pointwise8_planes(xs, w.cuda(), b.cuda())
    assert_close(y, ref, TOL_F32, "pointwise8 fwd")
    assert_close(o.pointwise8_planes(xs, w.cuda(), None), F.conv3d(xr, w), TOL_F32, "pointwise8 fwd, no bias")
    dy = torch.randn_like(ref)
    dw_ref = torch.nn.grad.conv3d_weight(xr, w.shape, dy)
    dw, db = o.pointwise8_planes_wgrad(xs, cuda_cl(dy))
    assert_close(dw, dw_ref, TOL_F32, "pointwise8 dw")
    assert_close(db, dy.sum(dim=(0, 2, 3, 4)), TOL_F32, "pointwise8 dbias")
    dw2, db2 = o.pointwise8_planes_wgrad(xs, cuda_cl(dy))
    assert torch.equal(dw, dw2) and torch.equal(db, db2), "pointwise8 wgrad must be deterministic"


@pytest.mark.parametrize("training", [True, False])
def test_reshape_head_unit_on_planes(training):
    """[Conv3d(64, 8, 1) + bias -> BatchNorm3d -> ReLU] on a detached `Act` (planes) input: DC3DATGeneric._reshape_head"""
    o, f = ops(), DF()
    N, C, S = 2, 64, (6, 5, 8)
    x = torch.randn(N, C, *S)
    conv, bn = torch.nn.Conv3d(C, 8, 1), torch.nn.BatchNorm3d(8)
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5); bn.bias.normal_(); bn.running_mean.normal_(); bn.running_var.uniform_(0.5, 2.0)
    bn.train(training)
    rm, rv = bn.running_mean.clone().cuda(), bn.running_var.clone().cuda()
    xs = o.split_bf16(cuda_cl(x), True)
    xr = o.merge_planes(xs).cpu()
    ref = F.relu(bn(conv(xr)))
    g = torch.randn_like(ref)
    ref.backward(g)
    act = f.Act(f._handle(xs.shape, "cuda"), xs)
    wg, bg = conv.weight.detach().cuda().requires_grad_(True), conv.bias.detach().cuda().requires_grad_(True)
    gg, beg = bn.weight.detach().cuda().requires_grad_(True), bn.bias.detach().cuda().requires_grad_(True)
    from dram_native import lib
    lib.PROFILE.reset()
    got = f.conv_bn_relu(act.detach(), wg, bg, gg, beg, rm, rv, training, 0.1, 1e-5, 1, False)
    assert lib.PROFILE.calls.get("dram_pointwise8_planes_fwd") == 1 and "dram_merge_planes" not in lib.PROFILE.calls
    assert_close(got, ref, TOL_X3, "reshape head fwd")
    got.backward(g.cuda())
    assert lib.PROFILE.calls.get("dram_pointwise8_planes_wgrad") == 1
    for t, r, name in ((wg, conv.weight, "dw"), (bg, conv.bias, "dbias"), (gg, bn.weight, "dgamma"), (beg, bn.bias, "dbeta")):
        scale = max(r.grad.abs().max().item(), 1e-3 * conv.weight.grad.abs().max().item())
        assert (t.grad.cpu() - r.grad).abs().max().item() <= 2e-4 * max(scale, 1e-6), f"reshape head {name}"
    if training:
        assert_close(rm, bn.running_mean, TOL_F32, "running mean")
        assert_close(rv, bn.running_var, TOL_F32, "running var")


# ------------------------------------------------------------------------------------------------ PCM stencil attention
@pytest.mark.parametrize("merge,self_loop,conn,grid,Fd", [
    ("scaled_dot_product_relu", False, 2, (6, 7, 8), 8), ("sm", True, 1, (5, 5, 5), 8),
    ("scaled_dot_product", False, 3, (4, 6, 5), 8), ("smrelu", False, 2, (2, 1, 3), 8),
    ("scaled_dot_product_relu", False, 2, (3, 4, 64), 8), ("scaled_dot_product_relu", False, 2, (2, 3, 37), 8),
    ("scaled_dot_product_relu", False, 2, (3, 20, 20), 8),          # plane larger than one 256-thread block
    ("scaled_dot_product_relu", False, 2, (9, 6, 70), 8),           # fused inference kernel: two x tiles, halo columns between them
    ("scaled_dot_product_relu", False, 2, (40, 9, 8), 8),           # fused inference kernel: several z segments, ragged y tiles
    ("scaled_dot_product_relu", True, 3, (4, 5, 6), 4), ("smscaled", False, 2, (5, 4, 7), 16)])
def test_pcm_attention(merge, self_loop, conn, grid, Fd):
    from oracle_import import O
    import models
    B, Cf = 2, 17
    torch.manual_seed(7)
    pcm = models.PCM(grid, Cf, 1, Fd, 0, 8, 1, 3, merge, self_loop, connectivity=conn, p_enc_dim=0)
    sd = {"attention_module." + k: v.detach().clone().requires_grad_(True) for k, v in pcm.state_dict().items()}
    cam = torch.randn(B, 1, *grid, requires_grad=True)
    f = torch.randn(B, Cf, *grid, requires_grad=True)
    ref = O.pcm_forward(sd, cam, f, merge_type=merge, self_loop=self_loop, connectivity=conn)
    g = torch.randn_like(ref)
    ref.backward(g)
    pcm = pcm.cuda()
    camg, fg = cam.detach().cuda().requires_grad_(True), cuda_cl(f.detach()).requires_grad_(True)
    got = pcm(camg, fg)
    assert_close(got, ref, 5e-5, "pcm fwd")
    with torch.no_grad():                                   # inference launch: no softmax stats kept
        assert_close(pcm(camg.detach(), fg.detach()), ref, 5e-5, "pcm fwd (inference kernel)")
    got.backward(g.cuda())
    assert_close(camg.grad, cam.grad, 1e-4, "pcm dcam")
    assert_close(fg.grad, f.grad, 1e-4, "pcm df")
    # softmax is shift invariant: without ReLU the phi-bias gradient is analytically 0 (the reference value is rounding
    # noise), so parameter gradients are compared on the scale of the largest parameter gradient
    scale = max(v.grad.abs().max().item() for k, v in sd.items())
    for name, p in pcm.named_parameters():
        ref = sd["attention_module." + name].grad
        assert (p.grad.cpu() - ref).abs().max().item() <= 2e-4 * scale, f"pcm d{name}"


# ------------------------------------------------------------------------------------------------ peer-memory exchanges
@pytest.mark.parametrize("world", [2, 8])
def test_peer_allreduce_protocol_all_ranks_played_on_one_gpu(world):
    """csrc/peer.cu with nvirt = world: block b of ONE launch plays rank b, the "peer" mailboxes are buffers of this GPU.
    Covers the push / flag / wait / rank-ordered reduce protocol, slot reuse over many calls and the fused BatchNorm finalize
    (the multi-process NVLink path runs the same kernel with nvirt = 1: tests/dp_parity.py under torchrun)."""
    import ctypes
    from dram_native import lib
    L = lib.load()
    o = ops()
    boxes = []
    for _ in range(world):
        p = ctypes.c_void_p()
        lib.check(L.dram_peer_alloc(ctypes.byref(p)), "peer_alloc")
        boxes.append(p)
    box_arr = (ctypes.c_void_p * world)(*[b.value for b in boxes])
    arr = lambda ts: (ctypes.c_void_p * len(ts))(*[0 if t is None else t.data_ptr() for t in ts])
    st = torch.cuda.current_stream().cuda_stream
    torch.manual_seed(7)
    try:
        for call in range(11):                                            # > 2 x slots: every slot is reused
            n = [5, 1040, 64, 129][call % 4]
            ins = [torch.randn(n, device="cuda", dtype=torch.float64) for _ in range(world)]
            outs = [torch.empty(n, device="cuda", dtype=torch.float64) for _ in range(world)]
            lib.check(L.dram_peer_allreduce_f64(box_arr, arr(ins), arr(outs), n, 0, world, world, st), "peer_allreduce")
            ref = ins[0].clone()
            for t in ins[1:]:
                ref = ref + t                                             # rank order, like the kernel
            for r in range(world):
                assert torch.equal(outs[r], ref), (call, r)
        # in place
        ins = [torch.randn(33, device="cuda", dtype=torch.float64) for _ in range(world)]
        ref = torch.stack(ins).sum(0)
        lib.check(L.dram_peer_allreduce_f64(box_arr, arr(ins), arr(ins), 33, 0, world, world, st), "peer_allreduce")
        for r in range(world):
            assert torch.allclose(ins[r], ref, rtol=0, atol=1e-12)
        # fused BatchNorm finalize == bn_finalize on the summed statistics with the summed count
        C = 96
        sums = [torch.rand(2 * C, device="cuda", dtype=torch.float64) * 100 for _ in range(world)]
        for s_ in sums:
            s_[C:] += s_[:C] ** 2 / 50.0                                   # keep the variance positive
        counts = [50.0 + r for r in range(world)]
        gamma, beta = torch.rand(C, device="cuda") + 0.5, torch.randn(C, device="cuda")
        rms = [torch.zeros(C, device="cuda") for _ in range(world)]
        rvs = [torch.ones(C, device="cuda") for _ in range(world)]
        gs = [torch.empty(2 * C + 1, device="cuda", dtype=torch.float64) for _ in range(world)]
        o4 = [torch.empty((4, C), device="cuda") for _ in range(world)]
        cnt = (ctypes.c_double * world)(*counts)
        lib.check(L.dram_bn_finalize_peer(box_arr, arr(sums), cnt, arr(gs), 0, world, world, arr([gamma] * world), arr([beta] * world),
                                          arr(rms), arr(rvs), 0.1, 1e-5, 2, arr([t[0] for t in o4]), arr([t[1] for t in o4]),
                                          arr([t[2] for t in o4]), arr([t[3] for t in o4]), C, st), "bn_finalize_peer")
        total = sums[0].clone()
        for t in sums[1:]:
            total = total + t
        rm_ref, rv_ref = torch.zeros(C, device="cuda"), torch.ones(C, device="cuda")
        mean, rstd, scale, shift = o.bn_finalize(total, sum(counts), gamma, beta, rm_ref, rv_ref, 0.1, 1e-5, 2)
        for r in range(world):
            assert torch.equal(gs[r][:2 * C], total) and gs[r][2 * C].item() == sum(counts)
            for got, ref_ in zip(o4[r], (mean, rstd, scale, shift)):
                assert torch.equal(got, ref_)
            assert torch.equal(rms[r], rm_ref) and torch.equal(rvs[r], rv_ref)
        torch.cuda.synchronize()
    finally:
        for b in boxes:
            L.dram_peer_free(b)
