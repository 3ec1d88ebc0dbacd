import sys, os, numpy as np
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0,ROOT); sys.path.insert(0,ROOT+"/tests")
import torch, problems
from oracle import oracle
from helpers import factorize
from cholespy_b200 import _capi
lib=_capi.load()
for name in ["ico3","ico5","grid2d-64"]:
    p=problems.build_problem(name)
    perm,cp,ri,v=factorize(lib,p["n"],p["ii"],p["jj"],p["x"],p["fmt"])
    rp,ci,_=oracle.csc_to_csr(cp,ri,v)
    lo,nl,lo_order=oracle.levels(rp,ci,False); up,nu,up_order=oracle.levels(cp,ri,True)
    for trial in range(3):
        for cubin in (False, True):
            ref=oracle.ReferenceSolver(p["n"],p["ii"],p["jj"],p["x"],p["fmt"],np.float32,cpu=False,use_sm100_cubin=cubin)
            l,u,pm=ref.schedule()
            dl=(l!=lo_order).sum(); du=(u!=up_order).sum()
            # is ref's lower order a permutation? position of each row
            okperm = np.array_equal(np.sort(l),np.arange(p["n"]))
            # check topological validity: every dep appears earlier
            pos=np.empty(p["n"],np.int64); pos[l]=np.arange(p["n"]) if okperm else 0
            rows=np.repeat(np.arange(p["n"]),np.diff(rp)); off=ci<rows
            viol=(pos[ci[off]]>pos[rows[off]]).sum() if okperm else -1
            print(name,"trial",trial,"sm100cubin" if cubin else "ptx-jit","lower mismatches",dl,"upper mismatches",du,"perm ok",okperm,"topo violations",viol, "perm equal", np.array_equal(pm,perm))
            # level sequence implied: levels of ref order under oracle's levels must be nondecreasing
            if dl: 
                seq=lo[l]; print("   oracle-level sequence nondecreasing:", bool(np.all(np.diff(seq)>=0)), "first mismatch idx", int(np.flatnonzero(l!=lo_order)[0]))
            ref.close()
