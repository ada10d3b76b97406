import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmgl_b200 import ops
b, s, heads, d = 8, 640, 32, 64
h = heads * d
qkv = torch.randn(b, s, 3 * h, device="cuda").to(torch.bfloat16).requires_grad_(True)
km = torch.ones(b, s, dtype=torch.uint8, device="cuda")
km[:, 300:512] = 0
for it in range(3):
    o = ops.self_attention(qkv, km, heads)
    o.backward(o.detach())
torch.cuda.synchronize()
ts = []
for fn in ("fwd", "fwd+bwd"):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        o = ops.self_attention(qkv, km, heads)
        if fn != "fwd":
            o.backward(o.detach())
    e1.record(); torch.cuda.synchronize()
    print(fn, e0.elapsed_time(e1) / 10, "ms")
q = qkv.detach()[..., :h].reshape(b, s, heads, d).transpose(1, 2)
k = qkv.detach()[..., h:2*h].reshape(b, s, heads, d).transpose(1, 2)
v = qkv.detach()[..., 2*h:].reshape(b, s, heads, d).transpose(1, 2)
allowed = torch.ones(s, s, dtype=torch.bool, device="cuda").tril_()[None, None] & km.bool()[:, None, None, :]
for _ in range(3): torch.nn.functional.scaled_dot_product_attention(q, k, v, attn_mask=allowed)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): torch.nn.functional.scaled_dot_product_attention(q, k, v, attn_mask=allowed)
e1.record(); torch.cuda.synchronize()
print("sdpa fwd", e0.elapsed_time(e1) / 10, "ms")
