"""Small end-to-end exercise of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck):
the gated cross-attention layer of smoke(), the self-attention forward + single-pass backward (causal + padding,
two query tiles, two (sample, head) items per CTA), the cross-attention core, LayerNorm / RMSNorm, CE, RoPE, SwiGLU.
    compute-sanitizer --tool memcheck python tools/sanitize_target.py"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as G
from mmgl_b200 import ops

G.smoke()
BF = torch.bfloat16
gen = torch.Generator().manual_seed(0)
b, s, heads, d = 2, 200, 2, 64
h = heads * d
qkv = (torch.randn(b, s, 3 * h, generator=gen)).cuda().to(BF).requires_grad_(True)
km = torch.ones(b, s, dtype=torch.uint8); km[0, 150:] = 0
o = ops.self_attention(qkv, km.cuda(), heads, causal=True)
o.backward(torch.randn(b, s, h, generator=gen).cuda().to(BF))
q = torch.randn(b, s, h, generator=gen).cuda().to(BF).requires_grad_(True)
k = torch.randn(b, 40, h, generator=gen).cuda().to(BF).requires_grad_(True)
v = torch.randn(b, 40, h, generator=gen).cuda().to(BF).requires_grad_(True)
mask = torch.ones(b, 40, dtype=torch.bool); mask[1, 30:] = False
ox = ops.xattn_core(q, k, v, mask.cuda(), heads)
ox.backward(torch.randn_like(ox))
x = torch.randn(b * s, h, generator=gen).cuda().to(BF).requires_grad_(True)
w = torch.ones(h, device="cuda", requires_grad=True)
y = ops.rms_norm(ops.layer_norm(x, w, torch.zeros(h, device="cuda", requires_grad=True)), w)
logits = ops.linear(y, torch.randn(512, h, generator=gen).cuda().to(BF))
loss = ops.cross_entropy(logits, torch.randint(0, 512, (b * s,), generator=gen).cuda())
loss.backward()
gu = torch.randn(64, 512, generator=gen).cuda().to(BF).requires_grad_(True)
ops.swiglu(gu).sum().backward()
torch.cuda.synchronize()
print("sanitize target finished, loss", float(loss))
