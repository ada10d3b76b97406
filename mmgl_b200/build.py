"""Build libmmgl_b200.so (sm_100a) in-tree with nvcc.  Usage: python -m mmgl_b200.build [--force]"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIB_DIR, "libmmgl_b200.so")
SOURCES = ["capi.cu", "gemm_sm100.cu", "xattn.cu", "xattn_sm100.cu", "sattn_sm100.cu", "sattn_bwd_sm100.cu", "rowops.cu", "llama_ops.cu", "optim.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
         "-Xcompiler", "-fPIC", "-Xptxas", "-v"] + os.environ.get("MMGL_EXTRA_FLAGS", "").split()
# --use_fast_math only where the epilogue / softmax issue rate matters and every transcendental on the path is an explicit
# intrinsic anyway (ex2 / rcp in the attention and GELU code; the tanh of the scalar gate is common.cuh:tanh_precise).
# rowops.cu (LayerNorm / RMSNorm statistics, cross-entropy log-sum-exp, gate-gradient reductions) is HBM-bound and is
# compiled with IEEE division / sqrt / logf and denormals kept.
FAST_MATH = {"gemm_sm100.cu", "xattn_sm100.cu", "sattn_sm100.cu", "sattn_bwd_sm100.cu"}


def _digest():
    h = hashlib.sha256()
    for root in (CSRC, os.path.join(os.path.dirname(HERE), "include")):
        for name in sorted(os.listdir(root)):
            with open(os.path.join(root, name), "rb") as f:
                h.update(name.encode())
                h.update(f.read())
    h.update(" ".join(FLAGS + sorted(FAST_MATH)).encode())
    return h.hexdigest()


def _stable_order(out: str) -> str:
    """ptxas -v reports the kernels of a file in a different order on every run; sort the per-kernel blocks by name so
    the committed lib/ptxas.log only changes when register / shared-memory use does."""
    head, blocks = [], []
    for line in out.splitlines():
        if "Compile time" in line:
            continue
        if "Compiling entry function" in line:
            blocks.append([line])
        elif blocks and line.startswith(("ptxas info", "    ")):
            blocks[-1].append(line)
        else:
            head.append(line)
    return "\n".join(head + [ln for b in sorted(blocks, key=lambda b: b[0]) for ln in b]) + "\n"


def build(force=False, verbose=False):
    os.makedirs(LIB_DIR, exist_ok=True)
    stamp = os.path.join(LIB_DIR, "build.sha256")
    digest = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read().strip() == digest:
        return LIB
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(LIB_DIR, src.replace(".cu", ".o"))
        objs.append(obj)
        cmd = [NVCC, *FLAGS, *(["--use_fast_math"] if src in FAST_MATH else []), "-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = []
    for src, p in procs:
        out, _ = p.communicate()
        log.append(f"==== {src}\n{_stable_order(out)}")
        if p.returncode != 0:
            sys.stderr.write("\n".join(log))
            raise RuntimeError(f"nvcc failed on {src}")
    with open(os.path.join(LIB_DIR, "ptxas.log"), "w") as f:
        f.write("\n".join(log))
    cmd = [NVCC, "-shared", "-o", LIB, *objs, "-lcudart_static", "-lpthread", "-ldl", "-lrt"]
    subprocess.check_call(cmd)
    with open(stamp, "w") as f:
        f.write(digest)
    if verbose:
        print("\n".join(log))
    return LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(path)
