"""Kernel-level parity of the fused convolution kernels (tcgen05 and SIMT) against a float64 torch
reference of the same op, through the C-ABI test hook storm_debug_conv."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _ref(x0, w, x1=None, ss=None, silu=False, x2=None, w2=None, bias=None, res=None, out_scale=1.0):
    x = x0 if x1 is None else torch.cat([x0, x1], dim=-1)
    x = x.double()
    if ss is not None:
        x = x * ss[:, None, None, :, 0].double() + ss[:, None, None, :, 1].double()
    if silu:
        x = F.silu(x)
    y = F.conv2d(x.permute(0, 3, 1, 2), w.double(), padding=w.shape[-1] // 2)
    if x2 is not None:
        y = y + F.conv2d(x2.double().permute(0, 3, 1, 2), w2.double())
    y = y.permute(0, 2, 3, 1)
    if bias is not None:
        y = y + bias.double()
    if res is not None:
        y = y + res.double()
    return (y * out_scale).float()


def _rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())


CASES = [
    # B, H, W, C0, C1, Cout, taps, ss, silu, skip(C2), res
    (1, 16, 16, 16, 0, 16, 9, False, False, 0, False),
    (2, 16, 8, 32, 0, 64, 9, True, True, 0, False),
    (2, 24, 40, 128, 0, 128, 9, True, True, 0, True),
    (1, 32, 32, 256, 128, 256, 9, True, True, 384, False),
    (2, 20, 12, 128, 128, 128, 9, True, True, 256, False),
    (3, 32, 32, 256, 0, 256, 1, True, False, 0, True),
    (1, 64, 64, 128, 0, 256, 9, True, True, 128, False),
    (1, 8, 8, 96, 96, 192, 9, True, True, 192, False),
    (2, 33, 17, 64, 0, 48, 9, False, True, 0, False),
]


@pytest.mark.parametrize("engine", [1, 0])
@pytest.mark.parametrize("case", CASES)
def test_fused_conv(engine, case):
    from storm_b200 import _native
    B, H, W, C0, C1, Cout, taps, use_ss, silu, C2, use_res = case
    g = torch.Generator(device="cuda").manual_seed(hash(case) & 0xFFFF)
    rn = lambda *s: torch.randn(*s, device="cuda", generator=g)   # noqa: E731
    k = 3 if taps == 9 else 1
    x0 = rn(B, H, W, C0)
    x1 = rn(B, H, W, C1) if C1 else None
    Cin = C0 + C1
    w = rn(Cout, Cin, k, k) / (Cin * taps) ** 0.5
    ss = None
    if use_ss:
        ss = torch.stack([1.0 + 0.3 * rn(B, Cin), 0.2 * rn(B, Cin)], dim=-1).contiguous()
    x2 = rn(B, H, W, C2) if C2 else None
    w2 = rn(Cout, C2, 1, 1) / C2 ** 0.5 if C2 else None
    bias = 0.1 * rn(Cout)
    res = rn(B, H, W, Cout) if use_res else None
    kw = dict(x1=x1, ss=ss, silu=silu, x2=x2, w2=w2, bias=bias, res=res, out_scale=0.70710678)
    ref = _ref(x0, w, **kw)
    out = _native.debug_conv(engine, x0, w, **kw)
    err = _rel(out, ref)
    print(f"engine={engine} case={case} rel-L2={err:.3e}")
    assert torch.isfinite(out).all()
    # SIMT fp32 and the bf16 x 3 tensor-core split reach ~5e-6; the "f16 + 2 x fp8" split (every segment a multiple of 32
    # channels) ~2e-5 per layer
    assert err < (2e-5 if engine == 1 else 8e-5)


@pytest.mark.parametrize("engine", [1, 0])
@pytest.mark.parametrize("B,N,C", [(2, 1024, 256), (1, 2048, 256), (3, 128, 192), (2, 256, 64)])
def test_attention_core(engine, B, N, C):
    from storm_b200 import _native
    g = torch.Generator(device="cuda").manual_seed(N + C)
    q, k, v = (torch.randn(B, N, C, device="cuda", generator=g) * s for s in (1.5, 1.5, 1.0))
    w = torch.softmax(torch.einsum("bqc,bkc->bqk", q.double(), k.double()) * C ** -0.5, dim=-1)
    ref = torch.einsum("bqk,bkc->bqc", w, v.double()).float()
    out = _native.debug_attention(engine, q, k, v)
    err = _rel(out, ref)
    print(f"attention engine={engine} B={B} N={N} C={C} rel-L2={err:.3e}")
    assert torch.isfinite(out).all() and err < 2e-5


@pytest.mark.parametrize("B,H,W,C0,C1,Cout,sep", [(2, 40, 24, 128, 0, 128, False), (1, 32, 32, 256, 128, 256, False),
                                                    (2, 16, 16, 128, 0, 128, True), (1, 64, 16, 256, 0, 256, True),
                                                    (1, 24, 8, 96, 96, 192, False)])
def test_fused_skip_conv(B, H, W, C0, C1, Cout, sep):
    """conv3x3(act(cat[x0,x1])) and conv1x1(raw) from ONE launch (ResnetBlock Conv_0 + Conv_2)."""
    from storm_b200 import _native
    g = torch.Generator(device="cuda").manual_seed(H * W + Cout)
    rn = lambda *s: torch.randn(*s, device="cuda", generator=g)   # noqa: E731
    Cin = C0 + C1
    x0 = rn(B, H, W, C0)
    x1 = rn(B, H, W, C1) if C1 else None
    x2 = rn(B, H, W, Cin) if sep else None
    w = rn(Cout, Cin, 3, 3) / (Cin * 9) ** 0.5
    w2 = rn(Cout, Cin, 1, 1) / Cin ** 0.5
    ss = torch.stack([1.0 + 0.3 * rn(B, Cin), 0.2 * rn(B, Cin)], dim=-1).contiguous()
    bias = 0.1 * rn(Cout)
    out, skip = _native.debug_conv_fused(x0, w, w2, x1=x1, ss=ss, silu=True, x2=x2, bias=bias)
    ref = _ref(x0, w, x1=x1, ss=ss, silu=True, bias=bias)
    raw = x2 if sep else (x0 if x1 is None else torch.cat([x0, x1], dim=-1))
    ref_skip = F.conv2d(raw.double().permute(0, 3, 1, 2), w2.double()).permute(0, 2, 3, 1).float()
    e1, e2 = _rel(out, ref), _rel(skip, ref_skip)
    print(f"fused skip {B,H,W,C0,C1,Cout,sep}: main {e1:.3e} skip {e2:.3e}")
    assert e1 < 2e-5 and e2 < 2e-5
