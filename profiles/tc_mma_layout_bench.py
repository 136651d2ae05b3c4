import sys, ctypes as C, torch
sys.path.insert(0,'/root/repo')
from mlff_b200 import _lib
lib=_lib.load().dll
lib.so3k_tc_bench.argtypes=[C.c_int32,C.c_int32,C.c_int32,C.c_void_p,C.c_void_p]
out=torch.zeros(2,dtype=torch.int64,device='cuda')
for N in (144,128,256,64):
    for layout in (0,2):
        for reps in (100,1000):
            lib.so3k_tc_bench(N,layout,reps,out.data_ptr(),None); torch.cuda.synchronize()
            o=out.cpu().tolist(); print(f'N={N} layout={layout} reps={reps}: {o[0]/o[1]:.1f} cycles/MMA (ideal {128*N/256:.0f})')
