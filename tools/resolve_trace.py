"""Diagnostic: event trace of the running-sum resolve warp inside the fused tail (library built with
-DQMCL_RESOLVE_TRACE, e.g. quickmcl_b200/variants/libtrace.so via QMCL_B200_LIB).  Prints the time
between consecutive events of one update of a bench workload.

events: 1 start, 2 tile summaries loaded (32 tiles), 3 tile scan done, 4 chunk summaries staged,
5 a tile is opened (tiles before it walked), 6 predicted crossing confirmed, 7 literal additions begin,
8 literal chunk done, 9 end"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import quickmcl_b200 as q  # noqa: E402
from quickmcl_b200 import _abi  # noqa: E402

NAMES = {1: "start", 2: "tiles loaded", 3: "tile scan", 4: "chunks staged", 5: "tile opened",
         6: "crossing confirmed", 7: "literal begins", 8: "literal done", 9: "end"}
wl = sys.argv[1] if len(sys.argv) > 1 else "C2"
mode = int(sys.argv[2]) if len(sys.argv) > 2 else 0
stream = torch.cuda.current_stream()
h = bench.Harness(q, torch, wl, 0, stream)
h.f.set_prefix_mode(mode)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for it in range(4):
    h.reset(it, flush)
    h.step(False)
torch.cuda.synchronize()
L = _abi.lib()
fn = L.qmcl_debug_resolve_trace
fn.restype = C.c_int
fn.argtypes = [C.POINTER(C.c_uint64), C.c_size_t, C.POINTER(C.c_size_t)]
buf = np.zeros(4096, np.uint64)
n = C.c_size_t(0)
assert fn(buf.ctypes.data_as(C.POINTER(C.c_uint64)), 4096, C.byref(n)) == 0
ev = buf[:n.value].reshape(-1, 2)
t0 = int(ev[0, 1])
prev = t0
tot = {}
for k, t in ev:
    k, t = int(k), int(t)
    print(f"{(t - t0) / 1e3:8.2f} us  +{(t - prev) / 1e3:6.2f}  {NAMES.get(k, k)}")
    tot[k] = tot.get(k, 0) + (t - prev)
    prev = t
print("time attributed to the step that ENDS with each event (us):",
      {NAMES[k]: round(v / 1e3, 2) for k, v in sorted(tot.items())})
