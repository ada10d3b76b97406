"""Short-K (K = 768) encoder GEMMs at the packed RoBERTa row count of workload cfg2, batch 16: time + check against torch.
Development tool.  MMGL_GEMM_STAGED=0/1 switches the transposed (64-byte row segment) epilogue stores."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmgl_b200 import _capi as K  # noqa: E402
from microbench import time_it  # noqa: E402

BF16 = torch.bfloat16
torch.manual_seed(0)
m = int(os.environ.get("ROWS", 46000))
for n, k, kw_name in [(2304, 768, "bias"), (768, 768, "bias+residual"), (3072, 768, "bias+gelu"), (768, 3072, "bias+residual"),
                      (2048, 2048, "bias+residual"), (8192, 2048, "bias+relu")]:
    a = (torch.randn(m, k, device="cuda") * 0.5).to(BF16)
    b = (torch.randn(n, k, device="cuda") * 0.05).to(BF16)
    bias = torch.randn(n, device="cuda")
    res = torch.randn(m, n, device="cuda").to(BF16)
    out = torch.empty(m, n, dtype=BF16, device="cuda")
    kw = dict(bias=bias)
    if "residual" in kw_name:
        kw["residual"] = res
    if "gelu" in kw_name:
        kw["relu"] = 2
    if "relu" in kw_name:
        kw["relu"] = 1
    K.gemm(a, b, out, **kw)
    ref = a.float() @ b.float().t() + bias
    if kw.get("relu") == 2:
        ref = torch.nn.functional.gelu(ref)
    if kw.get("relu") == 1:
        ref = torch.relu(ref)
    if "residual" in kw:
        ref = ref + res.float()
    err = (out.float() - ref).abs().max().item() / ref.abs().max().item()
    med, best = time_it(lambda: K.gemm(a, b, out, **kw))
    print(json.dumps({"m": m, "n": n, "k": k, "epi": kw_name, "us": round(med * 1e3, 1), "tflops": round(2.0 * m * n * k / med / 1e9, 1),
                      "rel_err": round(err, 5), "staged": os.environ.get("MMGL_GEMM_STAGED", "1")}), flush=True)
