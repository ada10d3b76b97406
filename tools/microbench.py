"""Per-kernel timing on one B200 (CUDA events, L2 flushed between iterations).  Development tool: prints one JSON
line per case; bench.py is the judged entry point.  Usage: python tools/microbench.py [gemm] [xattn] [rowops]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmgl_b200 import _capi as K  # noqa: E402

BF16 = torch.bfloat16
_flush = None


def flush_l2():
    global _flush
    if _flush is None:
        _flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    _flush.zero_()


def time_it(fn, iters=20, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush_l2()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def bench_gemm():
    shapes = [  # (M, N, K, a_t, b_t, label)
        (10240, 2048, 2048, 0, 0, "q/out proj B=16"),
        (10240, 6144, 2048, 0, 0, "fused qkv B=16"),
        (10240, 8192, 2048, 0, 0, "fc1 B=16"),
        (10240, 2048, 8192, 0, 0, "fc2 B=16"),
        (1024, 2048, 2048, 0, 0, "k/v proj (bank) B=16"),
        (10240, 2048, 8192, 0, 1, "dgrad fc1 B=16"),
        (10240, 2048, 6144, 0, 1, "dgrad fused qkv B=16"),
        (8192, 2048, 10240, 1, 1, "wgrad fc1 B=16"),
        (2048, 8192, 10240, 1, 1, "wgrad fc2 B=16"),
        (2048, 2048, 10240, 1, 1, "wgrad q/out B=16"),
        (8192, 8192, 8192, 0, 0, "square 8k"),
        (10240, 50272, 2048, 0, 0, "lm_head B=16"),
        (10240, 2048, 50272, 0, 1, "lm_head dgrad B=16"),
    ]
    for m, n, k, a_t, b_t, label in shapes:
        a = torch.randn((k, m) if a_t else (m, k), device="cuda").to(BF16)
        b = torch.randn((k, n) if b_t else (n, k), device="cuda").to(BF16)
        out = torch.empty((m, n), dtype=BF16, device="cuda")
        flops = 2.0 * m * n * k
        res = {"kernel": "gemm", "label": label, "m": m, "n": n, "k": k, "a_t": a_t, "b_t": b_t}
        for bn in (0, 192):
            med, best = time_it(lambda: K.gemm(a, b, out, a_t=bool(a_t), b_t=bool(b_t), block_n=bn))
            res[f"ours_bn{bn}_tflops"] = round(flops / med / 1e9, 1)
        for sk in (1, 2):
            med, best = time_it(lambda: K.gemm(a, b, out, a_t=bool(a_t), b_t=bool(b_t), block_n=256, pair=2, stream_k=sk))
            res[f"pair256_{'streamk' if sk == 2 else 'dp'}_tflops"] = round(flops / med / 1e9, 1)
        am = a.t() if a_t else a
        bm = b if b_t else b.t()
        med, best = time_it(lambda: torch.matmul(am, bm, out=out))
        res["cublas_tflops"] = round(flops / med / 1e9, 1)
        print(json.dumps(res), flush=True)


def bench_epilogue():
    """cost of the fused epilogue on the short-K, N = H GEMM (out_proj of every layer)"""
    m, n, k = 5120, 2048, 2048
    a = torch.randn(m, k, device="cuda").to(BF16)
    b = torch.randn(n, k, device="cuda").to(BF16)
    out = torch.empty((m, n), dtype=BF16, device="cuda")
    aux = torch.empty((m, n), dtype=BF16, device="cuda")
    res = torch.randn(m, n, device="cuda").to(BF16)
    bias = torch.randn(n, device="cuda")
    gate = torch.tensor([0.5], device="cuda")
    flops = 2.0 * m * n * k
    variants = {
        "plain": {},
        "bias": dict(bias=bias),
        "bias+residual": dict(bias=bias, residual=res),
        "bias+dropout": dict(bias=bias, dropout_p=0.1, dropout_seed=123),
        "bias+dropout+residual": dict(bias=bias, residual=res, dropout_p=0.1, dropout_seed=123),
        "bias+dropout+aux+gate+residual": dict(bias=bias, residual=res, aux=aux, gate=gate, dropout_p=0.1, dropout_seed=123),
        "bias+relu (fc1-like)": dict(bias=bias, relu=True),
    }
    for name, kw in variants.items():
        r = {"kernel": "gemm_epilogue", "variant": name}
        for label, extra in (("auto", {}), ("single192", dict(block_n=192, pair=1)), ("pair256", dict(block_n=256, pair=2))):
            med, _ = time_it(lambda: K.gemm(a, b, out, **kw, **extra))
            r[label + "_tflops"] = round(flops / med / 1e9, 1)
        print(json.dumps(r), flush=True)


def bench_xattn():
    for b, s, nk, heads, d in [(4, 640, 64, 32, 64), (4, 640, 128, 32, 64), (8, 640, 64, 32, 64), (2, 1152, 128, 32, 128)]:
        h = heads * d
        q = torch.randn(b * s, h, device="cuda").to(BF16)
        kv = torch.randn(b * nk, 2 * h, device="cuda").to(BF16)
        mask = torch.ones(b, nk, dtype=torch.uint8, device="cuda")
        o = torch.empty_like(q)
        stats = torch.empty(b, heads, s, 2, dtype=torch.float32, device="cuda")
        do = torch.randn_like(q)
        dq = torch.empty_like(q)
        dkv = torch.empty_like(kv)
        fwd = lambda: K.xattn_fwd(q, kv[:, :h], kv[:, h:], mask, o, stats, b, s, nk, heads, d)
        bwd = lambda: K.xattn_bwd(do, q, kv[:, :h], kv[:, h:], o, stats, mask, dq, dkv[:, :h], dkv[:, h:], b, s, nk, heads, d)
        fm, _ = time_it(fwd)
        bm, _ = time_it(bwd)
        bytes_fwd = (2 * b * s * h + 2 * b * nk * h) * 2 + b * nk
        bytes_bwd = (4 * b * s * h + 2 * b * nk * h) * 2 + (b * s * h + 2 * b * nk * h) * 2
        print(json.dumps({"kernel": "xattn", "b": b, "s": s, "nk": nk, "heads": heads, "d": d,
                          "fwd_us": round(fm * 1e3, 1), "fwd_GBs": round(bytes_fwd / fm / 1e6, 1),
                          "bwd_us": round(bm * 1e3, 1), "bwd_GBs": round(bytes_bwd / bm / 1e6, 1)}), flush=True)


def bench_rowops():
    rows, hidden = 2560, 2048
    x = torch.randn(rows, hidden, device="cuda").to(BF16)
    g = torch.ones(hidden, device="cuda")
    bta = torch.zeros(hidden, device="cuda")
    y = torch.empty_like(x)
    mean = torch.empty(rows, device="cuda")
    rstd = torch.empty(rows, device="cuda")
    med, _ = time_it(lambda: K.layernorm_fwd(x, g, bta, y, mean, rstd, 1e-5))
    print(json.dumps({"kernel": "layernorm_fwd", "rows": rows, "hidden": hidden, "us": round(med * 1e3, 1),
                      "GBs": round(2 * rows * hidden * 2 / med / 1e6, 1)}), flush=True)
    dx = torch.empty_like(x)
    dg = torch.empty(hidden, device="cuda")
    db = torch.empty(hidden, device="cuda")
    med, _ = time_it(lambda: K.layernorm_bwd(y, x, g, mean, rstd, x, dx, dg, db))
    print(json.dumps({"kernel": "layernorm_bwd(+affine)", "rows": rows, "hidden": hidden, "us": round(med * 1e3, 1)}), flush=True)
    out = torch.empty(8192, device="cuda")
    f = torch.randn(rows, 8192, device="cuda").to(BF16)
    med, _ = time_it(lambda: K.colsum(f, out))
    print(json.dumps({"kernel": "colsum", "m": rows, "n": 8192, "us": round(med * 1e3, 1),
                      "GBs": round(rows * 8192 * 2 / med / 1e6, 1)}), flush=True)


def bench_norm():
    """LayerNorm forward / backward(dx) at the step's shapes; run with MMGL_NORM_STAGED=0 for the register-resident backward."""
    for rows, hidden in [(10240, 2048), (5120, 2048), (4608, 768)]:
        x = torch.randn(rows, hidden, device="cuda").to(BF16)
        g = torch.ones(hidden, device="cuda")
        bta = torch.zeros(hidden, device="cuda")
        y = torch.empty_like(x)
        mean = torch.empty(rows, device="cuda")
        rstd = torch.empty(rows, device="cuda")
        dx = torch.empty_like(x)
        res = torch.randn(rows, hidden, device="cuda").to(BF16)
        r = {"kernel": "norm", "staged": os.environ.get("MMGL_NORM_STAGED", "1"), "rows": rows, "hidden": hidden}
        med, _ = time_it(lambda: K.layernorm_fwd(x, g, bta, y, mean, rstd, 1e-5))
        r["fwd_us"], r["fwd_GBs"] = round(med * 1e3, 1), round(2 * rows * hidden * 2 / med / 1e6)
        med, _ = time_it(lambda: K.layernorm_bwd(y, x, g, mean, rstd, None, dx))
        r["bwd_us"], r["bwd_GBs"] = round(med * 1e3, 1), round(3 * rows * hidden * 2 / med / 1e6)
        med, _ = time_it(lambda: K.layernorm_bwd(y, x, g, mean, rstd, res, dx))
        r["bwd_res_us"], r["bwd_res_GBs"] = round(med * 1e3, 1), round(4 * rows * hidden * 2 / med / 1e6)
        print(json.dumps(r), flush=True)


def bench_skinny_gemm():
    """Rank-64 LoRA weight gradients of T5-base at batch 8 (4608 encoder tokens): K-sliced (stream_k=2) vs data-parallel (default)."""
    gen = torch.Generator(device="cuda").manual_seed(0)
    for label, m, n, k in [("lora dA", 64, 768, 4608), ("lora dB", 768, 64, 4608), ("lora dA dec", 64, 768, 1024),
                           ("opt lora dA", 64, 2048, 5120)]:
        a = torch.randn(k, m, device="cuda", generator=gen).to(BF16)
        b = torch.randn(k, n, device="cuda", generator=gen).to(BF16)
        out = torch.empty(m, n, dtype=torch.float32, device="cuda")
        r = {"kernel": "skinny_gemm", "label": label, "m": m, "n": n, "k": k}
        for name, sk in (("k_sliced", 2), ("data_parallel", 1)):
            med, _ = time_it(lambda: K.gemm(a, b, out, a_t=True, b_t=True, stream_k=sk))
            r[name + "_us"] = round(med * 1e3, 1)
        print(json.dumps(r), flush=True)


def bench_encoder_gemm():
    """short-K GEMMs of the frozen RoBERTa / CLIP encoders: activation cost in the epilogue, tile choices"""
    for m, n, k, label in [(21504, 3072, 768, "roberta fc1"), (21504, 2304, 768, "roberta qkv"), (21504, 768, 3072, "roberta fc2"),
                           (21504, 768, 768, "roberta out"), (4728, 3072, 768, "clip fc1")]:
        a = torch.randn(m, k, device="cuda").to(BF16)
        b = torch.randn(n, k, device="cuda").to(BF16)
        out = torch.empty((m, n), dtype=BF16, device="cuda")
        bias = torch.randn(n, device="cuda")
        flops = 2.0 * m * n * k
        r = {"kernel": "encoder_gemm", "label": label, "m": m, "n": n, "k": k}
        for act in (0, 1, 2, 3):
            for name, extra in (("auto", {}), ("s192", dict(block_n=192, pair=1)), ("s256", dict(block_n=256, pair=1)),
                                ("p256", dict(block_n=256, pair=2))):
                med, _ = time_it(lambda: K.gemm(a, b, out, bias=bias, relu=act, **extra))
                r[f"act{act}_{name}"] = round(flops / med / 1e9, 1)
        med, _ = time_it(lambda: torch.nn.functional.linear(a, b, bias.to(BF16)))
        r["cublas"] = round(flops / med / 1e9, 1)
        print(json.dumps(r), flush=True)


def bench_tiles():
    """tile / pair choices on the step's forward shapes (B K-major)"""
    for m, n, k, label in [(5120, 2048, 8192, "fc2"), (5120, 2048, 2048, "out_proj"), (5120, 6144, 2048, "qkv"),
                           (5120, 8192, 2048, "fc1"), (21504, 3072, 768, "roberta fc1"), (21504, 2304, 768, "roberta qkv"),
                           (21504, 768, 3072, "roberta fc2"), (21504, 768, 768, "roberta out"), (4728, 3072, 768, "clip fc1"),
                           (4608, 768, 768, "t5 proj"), (5120, 50272, 2048, "lm_head")]:
        a = torch.randn(m, k, device="cuda").to(BF16)
        b = torch.randn(n, k, device="cuda").to(BF16)
        out = torch.empty((m, n), dtype=BF16, device="cuda")
        bias = torch.randn(n, device="cuda")
        flops = 2.0 * m * n * k
        r = {"kernel": "tiles", "label": label, "m": m, "n": n, "k": k}
        for name, extra in (("auto", {}), ("s128", dict(block_n=128, pair=1)), ("s192", dict(block_n=192, pair=1)),
                            ("s256", dict(block_n=256, pair=1)), ("p128", dict(block_n=128, pair=2)),
                            ("p192", dict(block_n=192, pair=2)), ("p256", dict(block_n=256, pair=2))):
            med, _ = time_it(lambda: K.gemm(a, b, out, bias=bias, relu=1, **extra))
            r[name] = round(flops / med / 1e9, 1)
        print(json.dumps(r), flush=True)


if __name__ == "__main__":
    which = sys.argv[1:] or ["gemm", "xattn", "rowops"]
    if "gemm" in which:
        bench_gemm()
    if "norm" in which:
        bench_norm()
    if "skinny" in which:
        bench_skinny_gemm()
    if "epilogue" in which:
        bench_epilogue()
    if "encoder" in which:
        bench_encoder_gemm()
    if "tiles" in which:
        bench_tiles()
    if "xattn" in which:
        bench_xattn()
    if "rowops" in which:
        bench_rowops()
