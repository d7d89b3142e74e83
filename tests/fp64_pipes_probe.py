import ctypes, torch, sys
sys.path.insert(0,'/root/repo')
from gapro_b200 import _lib
lib=_lib.load()
dev=torch.device('cuda:0'); scratch=torch.zeros(8,dtype=torch.float64,device=dev)
st=torch.cuda.current_stream(dev).cuda_stream
for mode in (0,1,2,2,1,0):
    v=ctypes.c_double()
    _lib.check(lib.gapro_fp64_peak(mode,20000,ctypes.byref(v),scratch.data_ptr(),st),'peak')
    print(mode, round(v.value,2))
