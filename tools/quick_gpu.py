import time, numpy as np, sys
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import demi_b200 as D
from demi_b200 import _native as N
from oracle import binding as O
ext = D.pack_externals(D.raft5_program())
eng = D.Engine(D.SchedulerConfig(N.MODEL_RAFT5, model_flags=1)); eng.set_externals(ext)
g = eng.fuzz_batch(1, 20000, 50, 5); c = O.fuzz_batch(N.MODEL_RAFT5, ext, 1, 20000, 50, 5, model_flags=1)
print('parity', (g==c).all(), (g!=c).sum())
if not (g==c).all():
    i=np.nonzero(g!=c)[0][0]; print(i, g[i], c[i])
for n in (100000, 1000000, 4000000):
    t=time.time(); a=eng.fuzz_batch(1,n,50,5); dt=time.time()-t; s=eng.stats()
    print(n, 'wall %.3f'%dt, 'kernel_ms %.2f'%s.kernel_ms, 'prefixes/s kernel %.3e'%(n/s.kernel_ms*1e3), 'deliv/s %.3e'%(s.deliveries/s.kernel_ms*1e3))
