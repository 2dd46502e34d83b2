import sys,time,ctypes
sys.path.insert(0,'/root/repo')
import numpy as np
import gwdepth_b200
from gwdepth_b200 import capi
rng=np.random.default_rng(0)
mats=[(5*rng.random((100,12+5*(b%8)))-rng.random((100,1))).astype(np.float32) for b in range(96)]
flat=np.concatenate([m.reshape(-1) for m in mats]); offs=np.cumsum([0]+[m.size for m in mats])[:-1].astype(np.int64); T=np.array([m.shape[1] for m in mats],dtype=np.int32)
qi=np.empty((96,100),np.int32); ti=np.empty((96,100),np.int32); cnt=np.empty(96,np.int32)
L=capi.lib()
for th in (1,2,4,8):
    t=time.perf_counter()
    for _ in range(50): L.gwd_lsap_batch(flat.ctypes.data, offs.ctypes.data, T.ctypes.data, 100, 96, qi.ctypes.data, ti.ctypes.data, cnt.ctypes.data, th)
    print("C call threads=%d: %.3f ms"%(th,(time.perf_counter()-t)/50*1e3))
