"""Timeline of thread 0 of CTA 0 of the cross-attention forward (needs -DMMGL_TRACE).  Development tool.
tags: 1 start, 2 first loads landed; per tile it: 100+10it S done, +1 softmax done, +2 block barrier passed, +3 P V issued,
+4 next Q landed, +5 P V done, +6 epilogue done"""
import ctypes as C, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmgl_b200 import _capi as K
b, s, nk, heads, d = 16, 640, 64, 32, 64
h = heads * d
q = torch.randn(b * s, h, device="cuda").to(torch.bfloat16)
kv = torch.randn(b * nk, 2 * h, device="cuda").to(torch.bfloat16)
mask = torch.ones(b, nk, dtype=torch.uint8, device="cuda")
o = torch.empty_like(q); stats = torch.empty(b, heads, s, 2, dtype=torch.float32, device="cuda")
for _ in range(3):
    K.xattn_fwd(q, kv[:, :h], kv[:, h:], mask, o, stats, b, s, nk, heads, d)
torch.cuda.synchronize()
buf = (C.c_ulonglong * 1024)()
fn = C.CDLL(K.LIB_PATH).mmgl_debug_trace_xattn
fn.argtypes = [C.c_void_p]
fn(buf)
t0 = None
for i in range(511):
    tag, clk = buf[2 * i], buf[2 * i + 1]
    if tag == 0:
        break
    t0 = t0 or clk
    print(f"{clk - t0:8d}  {tag}")
