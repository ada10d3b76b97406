"""Timeline of CTA 0 of the single-pass attention backward (needs -DMMGL_TRACE:
MMGL_EXTRA_FLAGS=-DMMGL_TRACE python -m mmgl_b200.build --force).  Development tool.
tags: softmax thread 0: 1SSSSxx (step SSSS; 00 start, 10/11 S+dP half ready, 20/21 half computed, 30 prev MMAs done,
40 prev finished, 50 P/dS published); mma: 2000+x S/dP half issued, 3000+s P/dS seen, 4000+s second group issued; tma: 1000+s"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmgl_b200 import _capi as K  # noqa: E402

b, s, heads, d = int(sys.argv[1]) if len(sys.argv) > 1 else 8, int(sys.argv[2]) if len(sys.argv) > 2 else 640, 32, 64
causal = (sys.argv[3] != "0") if len(sys.argv) > 3 else True
h = heads * d
qkv = torch.randn(b * s, 3 * h, device="cuda").to(torch.bfloat16)
o = torch.empty(b * s, h, dtype=torch.bfloat16, device="cuda")
stats = torch.empty(b, heads, s, 2, dtype=torch.float32, device="cuda")
d_o = torch.randn(b * s, h, device="cuda").to(torch.bfloat16)
dqkv = torch.empty_like(qkv)
q, k, v = qkv[:, :h], qkv[:, h:2 * h], qkv[:, 2 * h:]
K.attn_fwd(q, k, v, None, None, o, stats, b, s, s, heads, d, d ** -0.5, causal)
for _ in range(3):
    K.attn_bwd(d_o, q, k, v, None, None, o, stats, dqkv[:, :h], dqkv[:, h:2 * h], dqkv[:, 2 * h:], b, s, s, heads, d, d ** -0.5, causal)
torch.cuda.synchronize()
buf = (C.c_ulonglong * 4096)()
fn = C.CDLL(K.LIB_PATH).mmgl_debug_trace_bwd
fn.argtypes = [C.c_void_p]
fn(buf)
ev = []
for role in range(4):
    for i in range(511):
        tag, clk = buf[role * 1024 + 2 * i], buf[role * 1024 + 2 * i + 1]
        if tag == 0:
            break
        ev.append((clk, role, tag))
ev.sort()
t0 = ev[0][0]
names = {0: "softmax0", 2: "mma", 3: "tma"}
for clk, role, tag in ev[:int(sys.argv[4]) if len(sys.argv) > 4 else 400]:
    print(f"{clk - t0:8d}  {names.get(role, role):9s} {tag}")
