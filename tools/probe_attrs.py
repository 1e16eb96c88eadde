import ctypes, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
torch.cuda.init()
from raglite_b200 import _lib
lib = _lib.load()
for which in (0, 1, 2):
    out = (ctypes.c_int * 10)()
    rc = lib.rl_debug_scan_kernel_attrs(which, out)
    print(which, rc, dict(zip(["numRegs", "maxThreadsPerBlock", "staticSmem", "localBytes", "maxDynSmem", "occBlocks", "occErr", "regsPerBlock", "regsPerSM", "smemOptin"], list(out))))
