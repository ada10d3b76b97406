"""cross-attention core forward + backward at the cfg2 / cfg4 shape (timing + ncu target).  python tools/xattn_one.py [batch] [nk]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmgl_b200 import _capi as K

b = int(sys.argv[1]) if len(sys.argv) > 1 else 16
nk = int(sys.argv[2]) if len(sys.argv) > 2 else 64
d = int(sys.argv[3]) if len(sys.argv) > 3 else 64
s, heads = (640, 32) if d == 64 else (1152, 32)
h = heads * d
BF = torch.bfloat16
q = torch.randn(b * s, h, device="cuda").to(BF)
kv = torch.randn(b * nk, 2 * h, device="cuda").to(BF)
k, v = kv[:, :h], kv[:, h:]
mask = (torch.rand(b, nk, device="cuda") > 0.3).to(torch.uint8); mask[:, 0] = 1
o = torch.empty_like(q); stats = torch.empty(b, heads, s, 2, dtype=torch.float32, device="cuda")
d_o = torch.randn_like(q); dq = torch.empty_like(q); dkv = torch.empty_like(kv)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def fwd(): K.xattn_fwd(q, k, v, mask, o, stats, b, s, nk, heads, d)
def bwd(): K.xattn_bwd(d_o, q, k, v, o, stats, mask, dq, dkv[:, :h], dkv[:, h:], b, s, nk, heads, d)
for name, fn, nbytes in (("fwd", fwd, b * ((2 * s * h + 2 * nk * h) * 2 + nk)), ("bwd", bwd, b * ((4 * s * h + 2 * nk * h) * 2 + (s * h + 2 * nk * h) * 2 + nk))):
    for _ in range(3): fn()
    ts = []
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ts.sort(); ms = ts[len(ts) // 2]
    print(f"xattn_{name} b={b} s={s} nk={nk} d={d}: {ms * 1e3:7.1f} us  {nbytes / ms / 1e6:7.1f} GB/s of algorithmic bytes ({nbytes / 1e6:.1f} MB)")
