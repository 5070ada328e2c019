import sys; import os; sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import numpy as np
from mwr_fast_forward_operators_and_lbls_b200.forward import math_selftest
rng=np.random.default_rng(1)
def ulps(g,r): return np.max(np.abs(g-r)/np.spacing(np.abs(r)))
x=np.concatenate([rng.uniform(-700,700,400000),rng.uniform(-1,1,400000)])
for op in ("exp_nb","exp_tb"): print(op, ulps(math_selftest(op,x).cpu().numpy(),np.exp(x)))
x=np.concatenate([np.exp(rng.uniform(-600,600,400000)),rng.uniform(0.5,2,400000),1+rng.uniform(-1e-6,1e-6,100000),1+rng.uniform(-.02,.02,100000)])
for op in ("log_nb","log_tb"):
    g=math_selftest(op,x).cpu().numpy(); r=np.log(x); print(op, ulps(g,r))

x=np.exp(rng.uniform(-300,300,1000000))
s=math_selftest("rcp_seed",x).cpu().numpy(); print("rcp seed max rel err 2^%.2f"%np.log2(np.max(np.abs(s*x-1))))
s=math_selftest("rcp_nr1",x).cpu().numpy(); print("rcp nr1 max rel err %.3e"%np.max(np.abs(s*x-1)))
x=np.concatenate([rng.uniform(-50,50,400000)]); g=math_selftest("exp_tb",x).cpu().numpy(); r=np.exp(x); print("exp_tb |x|<50 max rel", np.max(np.abs(g-r)/r))
x=np.concatenate([rng.uniform(-5,5,400000)]); g=math_selftest("exp_tb",x).cpu().numpy(); r=np.exp(x); print("exp_tb |x|<5 max rel", np.max(np.abs(g-r)/r))
