"""one forward + backward of the cfg2 OPT-layer attention shape (ncu target).  python tools/attn_one.py [batch]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.attn_bench import run
run(int(sys.argv[1]) if len(sys.argv) > 1 else 16, 640, 32, 64, True, True, iters=2)
