import sys, os
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, torch
import problems
from helpers import factorize, max_rel_err
from cholespy_b200 import CholeskySolverF, CholeskySolverD, MatrixType, _capi
from oracle import oracle
lib = _capi.load()
name = sys.argv[1]
p = problems.build_problem(name)
perm, cp, ri, v = factorize(lib, p['n'], p['ii'], p['jj'], p['x'], p['fmt'])
for cls, dt in ((CholeskySolverF, np.float32), (CholeskySolverD, np.float64)):
    s = cls(p['n'], torch.tensor(p['ii'], device='cuda'), torch.tensor(p['jj'], device='cuda'), torch.tensor(p['x'], device='cuda'), getattr(MatrixType, p['fmt']))
    for k in (1, 3, 4, 32, 128):
        b = np.random.default_rng(0).random((p['n'], k)).astype(dt)
        bt = torch.tensor(b, device='cuda'); xt = torch.zeros_like(bt)
        s.solve(bt, xt); torch.cuda.synchronize()
        x = xt.cpu().numpy()
        xo = oracle.solve(cp, ri, v, perm, b)
        print(name, dt.__name__, 'k', k, 'err', max_rel_err(x, xo), 'zeros', float((x == 0).mean()), flush=True)
