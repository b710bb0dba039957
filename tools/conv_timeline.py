"""Per-layer timing + CTA-0 cycle counters of the tcgen05 conv for the dominant NCSN++ shapes (B=8)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from storm_b200 import _native  # noqa: E402

B = 8
SHAPES = [  # name, H, W, C0, C1, Cout, prologue(ss+silu), skip C2, res (True | "up" = half-resolution residual)
    ("b31.conv0 up 256->256 @256", 256, 256, 256, 0, 256, False, 0, False),
    ("b31.conv1 256->256 +skip256 @256", 256, 256, 256, 0, 256, True, 256, False),
    ("b32.conv0 384->128 @256", 256, 256, 256, 128, 128, True, 0, False),
    ("b32.conv1 128->128 +skip384 @256", 256, 256, 128, 0, 128, True, 384, False),
    ("b4.conv0 128->128 @256", 256, 256, 128, 0, 128, True, 0, False),
    ("b4.conv1 128->128 +res @256", 256, 256, 128, 0, 128, True, 0, True),
    ("b27.conv0 512->256 @128", 128, 128, 256, 256, 256, True, 0, False),
    ("b22.conv0 512->256 @64", 64, 64, 256, 256, 256, True, 0, False),
    ("b13.conv0 256->256 @32", 32, 32, 256, 0, 256, True, 0, False),
    ("b17.conv1 256->256 +skip512 @32", 32, 32, 256, 0, 256, True, 512, False),
    ("b33.conv1 128->128 +skip256 @256", 256, 256, 128, 0, 128, True, 256, False),
    ("b27.conv1 256->256 +skip512 @128", 128, 128, 256, 0, 256, True, 512, False),
    ("b31.conv1 256->256 +res_up @256", 256, 256, 256, 0, 256, True, 0, "up"),
    ("out-branch GEMM 128->36 1x1 @256", 256, 256, 128, 0, 36, True, 0, "1x1"),
    ("out-branch GEMM 256->36 1x1 @128", 128, 128, 256, 0, 36, True, 0, "1x1"),
    ("NIN 256->256 1x1 @32", 32, 32, 256, 0, 256, True, 0, "1x1"),
]
g = torch.Generator(device="cuda").manual_seed(0)
rn = lambda *s: torch.randn(*s, device="cuda", generator=g)   # noqa: E731
for name, H, W, C0, C1, Cout, pro, C2, use_res in SHAPES:
    x0 = rn(B, H, W, C0)
    x1 = rn(B, H, W, C1) if C1 else None
    Cin = C0 + C1
    k = 1 if use_res == "1x1" else 3
    w = rn(Cout, Cin, k, k) / (Cin * k * k) ** 0.5
    ss = torch.stack([1.0 + 0.3 * rn(B, Cin), 0.2 * rn(B, Cin)], dim=-1).contiguous() if pro else None
    x2 = rn(B, H, W, C2) if C2 else None
    w2 = rn(Cout, C2, 1, 1) / C2 ** 0.5 if C2 else None
    res = rn(B, H, W, Cout) if use_res is True else None
    res_up = rn(B, H // 2, W // 2, Cout) if use_res == "up" else None
    out, ms, d = _native.debug_conv(0, x0, w, x1=x1, ss=ss, silu=pro, x2=x2, w2=w2, bias=None, res=res, res_up=res_up, timed=True,
                                    max_prec=0 if use_res == "1x1" else 1)
    macs = B * H * W * Cout * (k * k * Cin + C2)
    tf = 2 * macs / (ms * 1e-3) / 1e12
    print(f"{name:36s} {ms * 1e3:8.1f} us  {tf:6.1f} TFLOP/s alg  cta0: issue {d[0]} wait_act {d[1]} wait_w {d[2]} "
          f"retire {d[3]} prod_idle {d[4]} epi {d[5]} | producer t0: issue_loads {d[6]} empty_wait {d[8]} ld+convert seg0 {d[9]} seg1 {d[10]}")
