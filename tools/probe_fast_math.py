import ctypes as C, struct, sys, os
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import uhdr_testlib as T
g=T.Gpu(); lib=g.lib
first=struct.unpack("<I",struct.pack("<f",0.0031308))[0]; last=struct.unpack("<I",struct.pack("<f",1.0))[0]
w=C.c_float(-1); assert lib.uhdr_b200_probe_pow_fast(C.c_uint(first),C.c_uint(last-first+1),C.byref(w))==0
print("pow approx worst abs error", w.value)
w=C.c_float(-1); f=(127-40)<<23; n=((127+40)<<23)-f
assert lib.uhdr_b200_probe_log2_fast(C.c_uint(f),C.c_uint(n),C.byref(w))==0
print("lg2 worst error/bound", w.value)
