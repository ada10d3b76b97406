"""Timing sweep of the self-attention kernels (development tool).  python tools/attn_bench.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmgl_b200 import _capi as K  # noqa: E402

BF16 = torch.bfloat16


def run(b, s, heads, d, causal, pad, bwd=True, iters=20):
    h = heads * d
    qkv = torch.randn(b * s, 3 * h, device="cuda").to(BF16)
    km = None
    if pad:
        km = torch.ones(b, s, dtype=torch.uint8, device="cuda")
        km[:, int(s * 0.47):int(s * 0.8)] = 0
    o = torch.empty(b * s, h, dtype=BF16, device="cuda")
    stats = torch.empty(b, heads, s, 2, dtype=torch.float32, device="cuda")
    d_o = torch.randn(b * s, h, device="cuda").to(BF16)
    dqkv = torch.empty_like(qkv)
    q, k, v = qkv[:, :h], qkv[:, h:2 * h], qkv[:, 2 * h:]

    def fwd():
        K.attn_fwd(q, k, v, km, None, o, stats, b, s, s, heads, d, d ** -0.5, causal)

    def bw():
        K.attn_bwd(d_o, q, k, v, km, None, o, stats, dqkv[:, :h], dqkv[:, h:2 * h], dqkv[:, 2 * h:], b, s, s, heads, d,
                   d ** -0.5, causal)

    res = []
    for fn in (fwd, bw) if bwd else (fwd,):
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        res.append(e0.elapsed_time(e1) / iters * 1e3)
    nt = (s + 127) // 128
    blocks = (nt * (nt + 1) // 2 if causal else nt * nt) * b * heads
    flops = 4.0 * blocks * 128 * 128 * d
    print(f"b={b:3d} s={s:5d} nh={heads:3d} d={d:3d} causal={int(causal)} pad={int(pad)}: fwd {res[0]:8.1f} us "
          f"({flops / res[0] / 1e6:6.1f} TF/s block-granular, {res[0] * 1e3 / (blocks / 148):7.1f} ns per block per SM)"
          + (f"  bwd {res[1]:8.1f} us" if bwd else ""))


if __name__ == "__main__":
    run(8, 640, 32, 64, True, True)      # frozen OPT layer of the cfg2 step
    run(8, 640, 32, 64, True, False)
    run(42, 512, 12, 64, False, False)   # RoBERTa neighbors
    run(24, 197, 12, 64, False, False)   # CLIP
    run(64, 128, 32, 64, True, False)    # one block per CTA: fixed cost per CTA
    run(64, 256, 32, 64, False, False)   # one CTA = 2 tiles x 2 blocks
    run(8, 1280, 32, 64, True, False)
    run(8, 2560, 32, 64, False, False)   # long rows: steady state
    run(4, 1152, 32, 128, True, True)    # cfg5 shape
