"""Attention backward: tcgen05 kernel (attention_bwd_tc.cu) against the fp32 torch reference and the mma.sync kernels,
with per-launch timings. Run on a B200:  python tools/attn_bwd_check.py [--time]"""
import math
import sys
import time

import torch

sys.path.insert(0, ".")
from dinov2_distillation_b200 import ops  # noqa: E402


def ref(q, k, v, d_o, heads, scale):
    B, Nq, D = q.shape
    Nk = k.shape[1]
    hd = D // heads
    qr, kr, vr = (t.float().detach().requires_grad_(True) for t in (q, k, v))
    qh = qr.view(B, Nq, heads, hd).transpose(1, 2)
    kh = kr.view(B, Nk, heads, hd).transpose(1, 2)
    vh = vr.view(B, Nk, heads, hd).transpose(1, 2)
    p = torch.softmax(qh @ kh.transpose(-1, -2) * scale, dim=-1)
    o = (p @ vh).transpose(1, 2).reshape(B, Nq, D)
    o.backward(d_o.float())
    return qr.grad, kr.grad, vr.grad


def rel(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def run(B, heads, hd, Nq, Nk, half, shared_q=False, fused_qkv=False, colsum=False, do_scale=1.0):
    D = heads * hd
    scale = (5.0 if hd != 64 else 1.0) / math.sqrt(hd)
    dt = torch.float16 if half else torch.bfloat16
    g = torch.Generator(device="cuda").manual_seed(B * 1000 + heads * 10 + hd + Nq + Nk)
    if fused_qkv:
        qkv = (torch.randn(B, Nq, 3 * D, device="cuda", generator=g) * 0.5).to(dt)
        q, k, v = qkv[..., :D], qkv[..., D:2 * D], qkv[..., 2 * D:]
    else:
        q = (torch.randn(1 if shared_q else B, Nq, D, device="cuda", generator=g) * 0.5).to(dt)
        if shared_q:
            q = q.expand(B, Nq, D)
        k = (torch.randn(B, Nk, D, device="cuda", generator=g) * 0.5).to(dt)
        v = torch.randn(B, Nk, D, device="cuda", generator=g).to(dt)
    d_o = (torch.randn(B, Nq, D, device="cuda", generator=g) * do_scale).bfloat16()
    o, lse = ops.attention_fwd(q, k, v, heads, scale)
    gq, gk, gv = ref(q, k, v, d_o, heads, scale)
    out = {}
    for name, on in (("tc", 1), ("mma", 0)):
        ops.set_option("attn_tc_bwd", on)
        cs = tuple(torch.zeros(D, device="cuda") for _ in range(3)) if colsum else None
        dq, dk, dv = ops.attention_bwd(q, k, v, o, lse, d_o, heads, scale, colsums=cs)
        torch.cuda.synchronize()
        out[name] = (dq, dk, dv, cs)
    ops.set_option("attn_tc_bwd", 1)
    tag = f"B={B} h={heads} hd={hd} Nq={Nq} Nk={Nk} {'fp16' if half else 'bf16'}{' sharedq' if shared_q else ''}{' fusedqkv' if fused_qkv else ''}"
    ok = True
    for name in ("tc", "mma"):
        dq, dk, dv, cs = out[name]
        e = (rel(dq, gq), rel(dk, gk), rel(dv, gv))
        line = f"  {name:3s} dq {e[0]:.2e} dk {e[1]:.2e} dv {e[2]:.2e}"
        if cs is not None:
            # (the k-bias gradient vanishes identically -- softmax-gradient rows sum to zero -- so it is scored against
            # the magnitude of the q-bias gradient)
            ec = (rel(cs[0], gq.sum((0, 1))), ((cs[1] - gk.sum((0, 1))).norm() / gq.sum((0, 1)).norm()).item(),
                  rel(cs[2], gv.sum((0, 1))))
            line += f" | colsum dq {ec[0]:.2e} dk {ec[1]:.2e} dv {ec[2]:.2e}"
            e = e + ec
        bad = any(not (x < 2e-2) for x in e)
        if name == "tc" and bad:
            ok = False
        print(tag if name == "tc" else " " * len(tag), line, "FAIL" if bad else "")
    same = all(torch.equal(a, b) for a, b in zip(out["tc"][:3], out["mma"][:3]))
    dmax = max((a.float() - b.float()).abs().max().item() / b.float().abs().max().item() for a, b in zip(out["tc"][:3], out["mma"][:3]))
    print(" " * len(tag), f"  tc vs mma: bitwise equal {same}, max |diff| / max |x| {dmax:.2e}; launches tc "
          f"{ops.get_option('stat_attn_tc_bwd')} mma {ops.get_option('stat_attn_mma_bwd')}")
    return ok


def timeit(B, heads, hd, N, half, iters=20):
    """Kernel-only time of the backward proper (library profile hooks: CUDA events around the launch, delta kernel and
    operand copies excluded), with a 256 MB L2 flush between launches."""
    import ctypes as C
    from dinov2_distillation_b200 import _lib as L
    lib = L.load()
    D = heads * hd
    scale = (5.0 if hd != 64 else 1.0) / math.sqrt(hd)
    dt = torch.float16 if half else torch.bfloat16
    q = (torch.randn(B, N, D, device="cuda") * 0.5).to(dt)
    k = (torch.randn(B, N, D, device="cuda") * 0.5).to(dt)
    v = torch.randn(B, N, D, device="cuda").to(dt)
    d_o = torch.randn(B, N, D, device="cuda").bfloat16()
    alts = tuple(t.bfloat16() for t in (q, k, v)) if half else None
    o, lse = ops.attention_fwd(q, k, v, heads, scale)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    flops = 8.0 * B * heads * N * N * hd
    for name, on in (("tc", 1), ("mma", 0)):
        ops.set_option("attn_tc_bwd", on)
        for _ in range(3):
            ops.attention_bwd(q, k, v, o, lse, d_o, heads, scale, alts=alts)
        torch.cuda.synchronize()
        lib.b200_profile_enable(1)
        for _ in range(iters):
            flush.zero_()
            ops.attention_bwd(q, k, v, o, lse, d_o, heads, scale, alts=alts)
        torch.cuda.synchronize()
        ms, fl, n = (C.c_double * 3)(), (C.c_double * 3)(), (C.c_longlong * 3)()
        L.check(lib.b200_profile_read(3, ms, fl, n), "profile_read")
        lib.b200_profile_enable(0)
        us = ms[2] * 1e3 / max(n[2], 1)
        print(f"time B={B} h={heads} hd={hd} N={N} {'fp16' if half else 'bf16'} {name}: {us:.1f} us/launch ({n[2]} launches, cold L2)"
              f" {flops / us * 1e-6:.0f} TFLOP/s")
    ops.set_option("attn_tc_bwd", 1)


if __name__ == "__main__":
    torch.manual_seed(0)
    t0 = time.time()
    ok = True
    cases = [
        # B, heads, hd, Nq, Nk, half
        dict(B=1, heads=1, hd=64, Nq=64, Nk=128, half=False),
        dict(B=1, heads=1, hd=64, Nq=128, Nk=128, half=False),
        dict(B=1, heads=1, hd=64, Nq=256, Nk=256, half=False),
        dict(B=2, heads=6, hd=64, Nq=256, Nk=256, half=False, fused_qkv=True),
        dict(B=3, heads=6, hd=64, Nq=200, Nk=200, half=False),
        dict(B=2, heads=3, hd=64, Nq=100, Nk=77, half=False),
        dict(B=70, heads=6, hd=64, Nq=256, Nk=256, half=False, fused_qkv=True),
        dict(B=1, heads=1, hd=32, Nq=128, Nk=128, half=False),
        dict(B=2, heads=4, hd=32, Nq=256, Nk=256, half=False),
        dict(B=2, heads=4, hd=16, Nq=256, Nk=256, half=False),
        dict(B=2, heads=4, hd=24, Nq=256, Nk=256, half=False),
        dict(B=2, heads=4, hd=48, Nq=256, Nk=256, half=False),
        dict(B=1, heads=1, hd=32, Nq=128, Nk=128, half=True),
        dict(B=2, heads=24, hd=16, Nq=256, Nk=256, half=True, colsum=True, do_scale=1e-4),
        dict(B=2, heads=16, hd=24, Nq=256, Nk=256, half=True, colsum=True, do_scale=1e-4),
        dict(B=2, heads=16, hd=24, Nq=256, Nk=256, half=True, shared_q=True, colsum=True),
        dict(B=2, heads=16, hd=48, Nq=256, Nk=256, half=True, colsum=True),
        dict(B=2, heads=4, hd=64, Nq=150, Nk=150, half=True, colsum=True),
        dict(B=3, heads=24, hd=32, Nq=100, Nk=77, half=True, colsum=True),
        dict(B=2, heads=16, hd=24, Nq=64, Nk=64, half=True, colsum=True),
        # streaming mode (fp32 dQ accumulator)
        dict(B=1, heads=1, hd=64, Nq=384, Nk=384, half=False),
        dict(B=2, heads=3, hd=64, Nq=300, Nk=257, half=False),
        dict(B=2, heads=16, hd=64, Nq=1369, Nk=1369, half=False, fused_qkv=True),
        dict(B=2, heads=16, hd=64, Nq=1369, Nk=1369, half=True, colsum=True, do_scale=1e-4),
        dict(B=2, heads=16, hd=64, Nq=1369, Nk=1369, half=True, shared_q=True, colsum=True),
        dict(B=1, heads=4, hd=32, Nq=1024, Nk=520, half=True, colsum=True),
        dict(B=1, heads=2, hd=48, Nq=200, Nk=600, half=False),
    ]
    for c in cases:
        try:
            ok &= run(**c)
        except Exception as e:  # noqa: BLE001
            print("EXC", c, repr(e)[:300])
            ok = False
            break
    print("ALL OK" if ok else "SOME FAILED", f"({time.time() - t0:.1f}s)")
    if "--time" in sys.argv and ok:
        timeit(64, 6, 64, 256, False)
        timeit(64, 24, 16, 256, True)
        timeit(64, 16, 24, 256, True)
        timeit(32, 16, 48, 256, True)
        timeit(32, 16, 64, 1369, False, iters=5)
        timeit(32, 16, 64, 1369, True, iters=5)
