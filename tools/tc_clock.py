"""In-kernel globaltimer timeline of the tensor-core layer kernel (CTA 0).  Needs a library built with the timeline hooks:
    PVRAFT_NVCC_FLAGS=-DPVRAFT_TC_TIMELINE python -m pvraft_b200.build --force   (then rebuild without it)
    DBG=1 python tools/tc_clock.py"""
import os, sys, ctypes as C
os.environ['PVRAFT_TC_DBG'] = os.environ.get('DBG', '1')
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pvraft_b200 import ops, _lib
dev = torch.device('cuda:0')
b, n = 8, 8192
x = torch.randn(b, n, 64, device=dev); w = torch.randn(64, 64, device=dev)
tw = ops.tc_weights(w); out = torch.empty(b, n, 64, device=dev)
lib = C.CDLL(_lib.LIB_PATH)
for dbg in [int(v) for v in os.environ.get('DBG', '1').split(',')] * 2:
    os.environ['PVRAFT_TC_DBG'] = str(dbg)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); ops.tc_linear([x], tw, out=out); e.record(); torch.cuda.synchronize()
    buf = (C.c_ulonglong * 64)()
    lib.pvraft_tc_debug_clock(buf)
    t0 = buf[0]
    rel = lambda a, b: [int(buf[i] - t0) for i in range(a, b)]
    print('dbg', dbg, 'event us %.1f' % (s.elapsed_time(e) * 1e3), 'end', rel(1, 2), 'xform', rel(8, 16), 'mma', rel(16, 24), 'epi_start', rel(24, 28), 'epi_end', rel(28, 32), 'epi_detail', rel(32, 40), 'iso', rel(40, 44), 'iso_cycles', [int(buf[i] - buf[44]) for i in range(44, 48)])
