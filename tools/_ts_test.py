import sys; sys.path.insert(0,'/root/repo')
import numpy as np, nmrgnn_b200
m=nmrgnn_b200.load_model()
rng=np.random.default_rng(3)
A=rng.normal(size=(128,64)).astype(np.float32); W=rng.normal(size=(64,128)).astype(np.float32)
ref=A.astype(np.float16).astype(np.float64)@W.astype(np.float16).astype(np.float64)
d3=m.handle.selftest_gemm(A,W,3); d4=m.handle.selftest_gemm(A,W,4)
print('mode3 err',np.abs(d3-ref).max()/np.abs(ref).max(),'mode4 (TS) err',np.abs(d4-ref).max()/np.abs(ref).max(), 'max|d3-d4|',np.abs(d3-d4).max())
