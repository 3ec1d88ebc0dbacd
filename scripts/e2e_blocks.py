import sys, os, time, ctypes
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, torch, bench, problems
from cholespy_b200 import CholeskySolverF, _capi
name='cfg4'; p=problems.build_problem(name); perm,cp,ri,v=bench.get_factor(name,p); lib=_capi.load()
h=ctypes.c_void_p(); _capi.check(lib.chb_init(0))
_capi.check(lib.chb_solver_create_from_factor(ctypes.byref(h),0,len(perm),len(ri),cp.ctypes.data,ri.ctypes.data,v.ctypes.data,perm.ctypes.data,0))
s=CholeskySolverF._from_handle(h,len(perm),0)
b=torch.rand((p['n'],128),dtype=torch.float32).pin_memory(); x=torch.empty_like(b).pin_memory()
bd=b.cuda(); xd=torch.zeros_like(bd)
for k in (32,64,128):
    bk=bd[:, :k].contiguous(); xk=torch.zeros_like(bk)
    for _ in range(3): s.solve(bk,xk)
    torch.cuda.synchronize(); t=time.perf_counter()
    for _ in range(5): s.solve(bk,xk)
    torch.cuda.synchronize(); print("device solve k=%d: %.2f ms"%(k,(time.perf_counter()-t)/5*1e3),flush=True)
for nb in ("1","2","3","4"):
    os.environ["CHB_HOST_BLOCKS"]=nb
    for _ in range(2): s.solve_host(b,x)
    t=time.perf_counter()
    for _ in range(4): s.solve_host(b,x)
    print("solve_host blocks=%s: %.2f ms"%(nb,(time.perf_counter()-t)/4*1e3),flush=True)
