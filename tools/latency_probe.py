import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jacks_b200 import _native
ctx = _native.context(0)
L = _native.lib()
a, b, c = C.c_double(), C.c_double(), C.c_double()
L.jacks_debug_latency_probe.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)]
print('rc', L.jacks_debug_latency_probe(ctx.h, C.byref(a), C.byref(b), C.byref(c)))
print('dependent DFMA cycles', a.value, ' dependent MUFU.RCP64H cycles', b.value, ' 4 independent DFMA chains: cycles per DFMA', c.value)
f = C.c_double()
L.jacks_debug_3op_probe.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
print('rc', L.jacks_debug_3op_probe(ctx.h, C.byref(f)), ' DFMA with 3 distinct register operands: %.2f TFLOP/s' % (f.value / 1e12), ' vs shared-operand peak %.2f' % (ctx.fp64_peak_flops() / 1e12))
