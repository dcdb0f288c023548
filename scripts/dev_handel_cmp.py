import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests import emu_lib
from tests.oracle_lib import OracleHandel
from wittgenstein_b200 import Handel, HandelParameters
a=sys.argv
N=int(a[1]); thr=int(a[2]); pairing=int(a[3]); lw=int(a[4]); extra=int(a[5]); period=int(a[6]); fp=int(a[7]); down=int(a[8])
nb=a[9]; nl=a[10]; desync=int(a[11]); byz=int(a[12]); step=int(a[13]); T=int(a[14]); seed=int(a[15]) if len(a)>15 and a[15]!='-' else None; hid=bool(int(a[16])) if len(a)>16 else False
p=Handel(HandelParameters(N,thr,pairing,lw,extra,period,fp,down,nb,nl,desync,bool(byz),hid), _api=emu_lib.api())
o=OracleHandel(N,thr,pairing,lw,extra,period,fp,down,nb,nl,desync,bool(byz),seed=seed,hidden_byzantine=hid)
if seed is not None: p.network().set_seed(seed)
p.init(); o.init()
def cmp(tag, full=True):
    ok=True
    if p.network().rng_state()!=o.rng_state(): print(tag,"rng differ"); ok=False
    if p.network().msgs_size()!=o.msgs_live(): print(tag,"msgs differ",p.network().msgs_size(),o.msgs_live()); ok=False
    if not (p.network().counters()==o.counters()).all():
        d=(p.network().counters()!=o.counters()); print(tag,"counters differ rows",np.argwhere(d.any(axis=1)).ravel(),"nodes",np.argwhere(d.any(axis=0))[:5].ravel()); ok=False
    s1=p.scalars(); s2=o.scalars()
    for k in s1:
        if not (s1[k]==s2[k]).all(): print(tag,"scalar differs",k,np.argwhere(s1[k]!=s2[k])[:5].ravel(), s1[k][s1[k]!=s2[k]][:5], s2[k][s1[k]!=s2[k]][:5]); ok=False
    if full:
        for w in range(6):
            if not (p.rows(w)==o.rows(w)).all(): print(tag,"rows differ",w, np.argwhere((p.rows(w)!=o.rows(w)).any(axis=1))[:5].ravel()); ok=False
        l1=p.level_scalars(); l2=o.level_scalars()
        for k in l1:
            if not (l1[k]==l2[k]).all(): print(tag,"level differs",k,np.argwhere(l1[k]!=l2[k])[:5]); ok=False
    return ok
for n in (0,1,N//2,N-1):
    if not (p.ranks(n)==o.ranks(n)).all(): print("ranks differ",n); sys.exit(1)
    for l in range(p.levels):
        if not (p.peers(n,l)==o.peers(n,l)).all(): print("peers differ",n,l,p.peers(n,l)[:8],o.peers(n,l)[:8]); sys.exit(1)
a1=p.network().attrs(); a2=o.attrs()
for k in a1:
    if not (a1[k]==a2[k]).all(): print("attrs differ",k); sys.exit(1)
if not cmp("init"): sys.exit(1)
while o.time<T:
    r1=p.network().run_ms(step); r2=o.run_ms(step)
    if r1!=r2: print("ret differs"); sys.exit(1)
    if not cmp("t=%d"%o.time): sys.exit(1)
print("OK",o.time,"cont",p.continue_if(),o.continue_if(),"checked",p.scalars()['sigs_checked'].sum(), "filtered", p.scalars()['msg_filtered'].sum(), "draws", p.network().stats()['draws'])
