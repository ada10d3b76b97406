"""ncu target: ONE cold launch (L2 flushed first) of each hot kernel at the shapes of the cfg2 step, per-GPU batch 16, in a
fixed order so that the launches in the report can be matched to shapes by index:

  0 gemm  fc2 forward        M=10240 N=2048 K=8192  bias + dropout + residual
  1 gemm  fc1 forward        M=10240 N=8192 K=2048  bias + ReLU
  2 gemm  fc2 dgrad          M=10240 N=8192 K=2048  b_t, ReLU mask
  3 gemm  encoder QKV        M=19149 N=2304 K=768   bias
  4 gemm  encoder fc1        M=19149 N=3072 K=768   bias + GELU
  5 sattn_fwd                B=16 nh=32 S=640 D=64 causal + key padding
  6 sattn_delta, 7 sattn_bwd (same shape)
  8 xattn_fwd, 9 xattn_bwd   B=16 S=640 Nk=64 nh=32 D=64, ragged mask
 10 layernorm_fwd, 11 layernorm_bwd   10240 x 2048

    ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:mmgl -o gpurun_out/r02_targets \
        python tools/ncu_targets.py
Without ncu it prints CUDA-event times of the same launches (cold, single launch: upper bounds)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmgl_b200 import _capi as K  # noqa: E402

BF = torch.bfloat16
torch.manual_seed(0)
dev = "cuda"
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)


def rn(*shape, scale=1.0):
    return (torch.randn(*shape, device=dev) * scale).to(BF)


targets = []
m = 10240
a8, w2 = rn(m, 8192), rn(2048, 8192, scale=0.02)
a2, w1 = rn(m, 2048), rn(8192, 2048, scale=0.02)
res, out2, out8 = rn(m, 2048), torch.empty(m, 2048, dtype=BF, device=dev), torch.empty(m, 8192, dtype=BF, device=dev)
b2, b8 = torch.randn(2048, device=dev), torch.randn(8192, device=dev)
targets.append(("gemm fc2 fwd M10240 N2048 K8192 bias+dropout+residual",
                lambda: K.gemm(a8, w2, out2, bias=b2, residual=res, dropout_p=0.1, dropout_seed=7)))
targets.append(("gemm fc1 fwd M10240 N8192 K2048 bias+relu", lambda: K.gemm(a2, w1, out8, bias=b8, relu=True)))
targets.append(("gemm fc2 dgrad M10240 N8192 K2048 b_t relu_mask", lambda: K.gemm(a2, w2, out8, b_t=True, relu_mask=a8)))
me = 19149
ae, wq, wf = rn(me, 768), rn(2304, 768, scale=0.03), rn(3072, 768, scale=0.03)
oq, of = torch.empty(me, 2304, dtype=BF, device=dev), torch.empty(me, 3072, dtype=BF, device=dev)
bq, bf = torch.randn(2304, device=dev), torch.randn(3072, device=dev)
targets.append(("gemm encoder qkv M19149 N2304 K768 bias", lambda: K.gemm(ae, wq, oq, bias=bq)))
targets.append(("gemm encoder fc1 M19149 N3072 K768 bias+gelu", lambda: K.gemm(ae, wf, of, bias=bf, relu=2)))

b, s, heads, d = 16, 640, 32, 64
h = heads * d
qkv = rn(b * s, 3 * h)
km = torch.ones(b, s, dtype=torch.uint8, device=dev)
km[:, int(s * 0.47):int(s * 0.8)] = 0
o = torch.empty(b * s, h, dtype=BF, device=dev)
stats = torch.empty(b, heads, s, 2, dtype=torch.float32, device=dev)
d_o = rn(b * s, h)
dqkv = torch.empty_like(qkv)
q, k, v = qkv[:, :h], qkv[:, h:2 * h], qkv[:, 2 * h:]
targets.append(("sattn_fwd B16 nh32 S640 D64 causal+pad", lambda: K.attn_fwd(q, k, v, km, None, o, stats, b, s, s, heads, d, d ** -0.5, True)))
targets.append(("sattn_delta + sattn_bwd (same shape)", lambda: K.attn_bwd(d_o, q, k, v, km, None, o, stats, dqkv[:, :h], dqkv[:, h:2 * h],
                                                                          dqkv[:, 2 * h:], b, s, s, heads, d, d ** -0.5, True)))
nk = 64
qx, kvx = rn(b * s, h), rn(b * nk, 2 * h)
xm = (torch.rand(b, nk, device=dev) > 0.3).to(torch.uint8)
xm[:, 0] = 1
ox, sx = torch.empty_like(qx), torch.empty(b, heads, s, 2, dtype=torch.float32, device=dev)
dqx, dkvx = torch.empty_like(qx), torch.empty_like(kvx)
targets.append(("xattn_fwd B16 S640 Nk64", lambda: K.xattn_fwd(qx, kvx[:, :h], kvx[:, h:], xm, ox, sx, b, s, nk, heads, d)))
targets.append(("xattn_bwd B16 S640 Nk64", lambda: K.xattn_bwd(d_o, qx, kvx[:, :h], kvx[:, h:], ox, sx, xm, dqx, dkvx[:, :h], dkvx[:, h:], b, s, nk, heads, d)))
x = rn(m, 2048)
g, bt = torch.ones(2048, device=dev), torch.zeros(2048, device=dev)
y, mean, rstd = torch.empty_like(x), torch.empty(m, device=dev), torch.empty(m, device=dev)
dx, dg, db = torch.empty_like(x), torch.empty(2048, device=dev), torch.empty(2048, device=dev)
targets.append(("layernorm_fwd 10240x2048", lambda: K.layernorm_fwd(x, g, bt, y, mean, rstd, 1e-5)))
targets.append(("layernorm_bwd 10240x2048 (+residual grad, +affine)", lambda: K.layernorm_bwd(y, x, g, mean, rstd, x, dx, dg, db)))

for _, fn in targets:     # warm-up: tensor maps, workspaces, function attributes
    fn()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
for name, fn in targets:
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn()
    e1.record()
    torch.cuda.synchronize()
    print(f"{e0.elapsed_time(e1) * 1e3:9.1f} us  {name}", flush=True)
torch.cuda.cudart().cudaProfilerStop()
