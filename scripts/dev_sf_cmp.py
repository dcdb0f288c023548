import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests import emu_lib
from tests.oracle_lib import OracleSanFermin
from wittgenstein_b200 import SanFerminSignature, SanFerminSignatureParameters
N=int(sys.argv[1]); step=int(sys.argv[2]); T=int(sys.argv[3]); nb=sys.argv[4] if len(sys.argv)>4 else None; nl=sys.argv[5] if len(sys.argv)>5 else None
seed=int(sys.argv[6]) if len(sys.argv)>6 and sys.argv[6]!='-' else None
k=int(sys.argv[7]) if len(sys.argv)>7 else 1
p=SanFerminSignature(SanFerminSignatureParameters(N,N,2,48,300,k,False,None if nb=='-' else nb,None if nl=='-' else nl), _api=emu_lib.api())
o=OracleSanFermin(N,N,2,48,300,k,None if nb=='-' else nb,None if nl=='-' else nl)
if seed is not None: p.network().set_seed(seed); o.set_seed(seed)
p.init(); o.init()
def cmp(tag):
    ok=True
    if p.network().rng_state()!=o.rng_state(): print(tag,"rng differ"); ok=False
    if p.network().msgs_size()!=o.msgs_live(): print(tag,"msgs differ",p.network().msgs_size(),o.msgs_live()); ok=False
    if not (p.network().counters()==o.counters()).all():
        d=(p.network().counters()!=o.counters()); print(tag,"counters differ rows",np.argwhere(d.any(axis=1)).ravel(),"nodes",np.argwhere(d.any(axis=0))[:5].ravel()); ok=False
    a=p.scalars(); b=o.scalars()
    for k in a:
        if not (a[k]==b[k]).all(): print(tag,"scalar differs",k,np.argwhere(a[k]!=b[k])[:5].ravel()); ok=False
    return ok
if not cmp("init"): sys.exit(1)
while p.network().time<T:
    r1=p.network().run_ms(step); r2=o.run_ms(step)
    if r1!=r2: print("ret differs"); sys.exit(1)
    if not cmp("t=%d"%o.time): sys.exit(1)
s=p.scalars(); print("OK", o.time, "done", s['done'].sum(), "agg min/max", s['agg'].min(), s['agg'].max(), p.network().stats()['draws'])
