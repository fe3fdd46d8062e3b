import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from reak_b200 import _abi
lib = _abi.load_library()
tf, clk = C.c_double(), C.c_double()
print("rc", lib.rkb_measure_fp64_peak(0, 1.0, C.byref(tf), C.byref(clk)), "DFMA peak %.2f TFLOP/s, nominal clock %.0f MHz" % (tf.value, clk.value))
