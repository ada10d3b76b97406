"""Recipe for ``oracle/_ref``: the REAL reference modules, compiled -- TEST / BASELINE INFRASTRUCTURE ONLY.

The reference is pure Python, so "compiling it from the sources where they lie" means byte-compiling
``/root/reference/model/{modelling_cross_attention,modelling_self_attention,graph}.py`` with ``py_compile`` into
``oracle/_ref/model/*.bytecode`` (source-less, unchecked-hash pyc files under another extension).  No reference source text enters the repository;
``oracle/_ref/`` is git-ignored but travels to the GPU box with the snapshot, like the built ``.so`` (same image, same
CPython 3.12 magic number on both sides).  ``__graft_entry__.build()`` runs this when ``/root/reference`` exists.

    python oracle/build_ref.py            # -> oracle/_ref/model/*.bytecode + oracle/_ref/MANIFEST.json

``oracle/ref_loader.py`` imports the result; bench.py's ``--impl reference`` (CPU arm, kind "reference") and
``--impl eager`` (the same modules on the B200: the honest GPU baseline) run it.  Nothing in the product
package imports it.
"""
from __future__ import annotations

import hashlib
import json
import os
import py_compile
import sys

REF_ROOT = os.environ.get("MMGL_REFERENCE_ROOT", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
FILES = ["model/modelling_cross_attention.py", "model/modelling_self_attention.py", "model/graph.py"]


def build(force: bool = False) -> str | None:
    """Returns the output directory, or None when the reference tree is absent (GPU box: uses the prebuilt files)."""
    if not os.path.isdir(REF_ROOT):
        return OUT if os.path.exists(os.path.join(OUT, "MANIFEST.json")) else None
    manifest = {"python": sys.version.split()[0], "magic": __import__("importlib.util").util.MAGIC_NUMBER.hex(), "files": {}}
    for rel in FILES:
        src = os.path.join(REF_ROOT, rel)
        dst = os.path.join(OUT, rel[:-3] + ".bytecode")   # a .pyc by content; gpurun snapshots skip *.pyc
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        with open(src, "rb") as f:
            digest = hashlib.sha256(f.read()).hexdigest()
        manifest["files"][rel] = {"sha256_of_source": digest, "pyc": os.path.relpath(dst, OUT)}
        if force or not os.path.exists(dst) or os.path.getmtime(dst) < os.path.getmtime(src):
            py_compile.compile(src, cfile=dst, dfile=src, doraise=True, optimize=0,
                               invalidation_mode=py_compile.PycInvalidationMode.UNCHECKED_HASH)
    with open(os.path.join(OUT, "MANIFEST.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
