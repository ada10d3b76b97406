"""quick backward timing of three shapes (development tool)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.attn_bench import run
run(16, 640, 32, 64, True, True)
run(8, 640, 32, 64, True, True)
run(8, 2560, 32, 64, False, False)
run(4, 1152, 32, 128, True, True)
