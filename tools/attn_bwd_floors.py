"""Floor measurements of attn_bwd_tc_kernel (probe build): kernel time with parts of the work skipped."""
import ctypes as C
import math
import sys

import torch

sys.path.insert(0, ".")
from dinov2_distillation_b200 import _lib as L, ops  # noqa: E402

lib = L.load()


def run(hd, heads, half, B=64, N=256):
    D = heads * hd
    dt = torch.float16 if half else torch.bfloat16
    scale = (5.0 if hd != 64 else 1.0) / math.sqrt(hd)
    q = (torch.randn(B, N, D, device="cuda") * 0.5).to(dt)
    k = (torch.randn(B, N, D, device="cuda") * 0.5).to(dt)
    v = torch.randn(B, N, D, device="cuda").to(dt)
    d_o = torch.randn(B, N, D, device="cuda").bfloat16()
    alts = tuple(t.bfloat16() for t in (q, k, v)) if half else None
    o, lse = ops.attention_fwd(q, k, v, heads, scale)
    for mask, what in ((0, "full"), (1, "no dQ"), (2, "no dV/dK"), (4, "no dP"), (8, "no exp math"), (16, "no dS smem store"),
                       (7, "only S MMA"), (8 + 16, "no exp, no dS store"), (7 + 8 + 16, "skeleton")):
        ops.set_option("attn_probe_skip", mask)
        for _ in range(2):
            ops.attention_bwd(q, k, v, o, lse, d_o, heads, scale, alts=alts)
        torch.cuda.synchronize()
        lib.b200_profile_enable(1)
        for _ in range(10):
            ops.attention_bwd(q, k, v, o, lse, d_o, heads, scale, alts=alts)
        torch.cuda.synchronize()
        ms, fl, n = (C.c_double * 3)(), (C.c_double * 3)(), (C.c_longlong * 3)()
        L.check(lib.b200_profile_read(3, ms, fl, n), "profile_read")
        lib.b200_profile_enable(0)
        print(f"hd={hd} heads={heads} {'fp16' if half else 'bf16'} skip={mask:2d} ({what}): {ms[2] * 1e3 / max(n[2], 1):.1f} us")
    ops.set_option("attn_probe_skip", 0)


run(64, 6, False)
run(16, 24, True)
