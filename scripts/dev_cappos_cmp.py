# dev helper (test infrastructure): SanFerminCappos on the host emulation of the device logic vs the oracle
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests.oracle_lib import OracleCappos
from wittgenstein_b200 import SanFerminCappos, SanFerminCapposParameters
N=int(sys.argv[1]); k=int(sys.argv[2]); step=int(sys.argv[3]); T=int(sys.argv[4])
nb=sys.argv[5] if len(sys.argv)>5 and sys.argv[5]!='-' else None; nl=sys.argv[6] if len(sys.argv)>6 and sys.argv[6]!='-' else None
seed=int(sys.argv[7]) if len(sys.argv)>7 and sys.argv[7]!='-' else None
force=int(sys.argv[8]) if len(sys.argv)>8 else 0
api=None
if os.environ.get("WTG_TEST_EMU","1")=="1":
    from tests import emu_lib
    api=emu_lib.api()
p=SanFerminCappos(SanFerminCapposParameters(N,N//2,2,48,150,k,nb,nl), _api=api, tunables={"force_shuffle_serial":1} if force else None)
o=OracleCappos(N,N//2,2,48,150,k,nb,nl,seed=seed)
if seed is not None: p.network().set_seed(seed)
p.init(); o.init()
def cmp(tag):
    ok=True
    if p.network().rng_state()!=o.rng_state(): print(tag,"rng differ"); ok=False
    if p.network().msgs_size()!=o.msgs_live(): print(tag,"msgs differ",p.network().msgs_size(),o.msgs_live()); ok=False
    if not (p.network().counters()==o.counters()).all():
        d=(p.network().counters()!=o.counters()); print(tag,"counters differ rows",np.argwhere(d.any(axis=1)).ravel(),"nodes",np.argwhere(d.any(axis=0))[:5].ravel()); ok=False
    a=p.scalars(); b=o.scalars()
    for kk in a:
        if not (a[kk]==b[kk]).all(): print(tag,"scalar differs",kk,np.argwhere(a[kk]!=b[kk])[:5].ravel(), a[kk][a[kk]!=b[kk]][:5], b[kk][a[kk]!=b[kk]][:5]); ok=False
    return ok
if not cmp("init"): sys.exit(1)
while p.network().time<T:
    r1=p.network().run_ms(step); r2=o.run_ms(step)
    if r1!=r2: print("ret differs"); sys.exit(1)
    if not cmp("t=%d"%o.time): sys.exit(1)
s=p.scalars(); print("OK", o.time, "done", s['done'].sum(), "sigs min/max", s['sigs'].min(), s['sigs'].max(), "draws", p.network().stats()['draws'])
