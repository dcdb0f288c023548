import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sys, numpy as np
from tests import emu_lib
from tests.oracle_lib import OracleGSF
from wittgenstein_b200 import GSFSignature, GSFSignatureParameters
api = emu_lib.api()
N=int(sys.argv[1]); thr=float(sys.argv[2]); dead=float(sys.argv[3]); nb=sys.argv[4]; nl=sys.argv[5]; step=int(sys.argv[6]); T=int(sys.argv[7])
pair=int(sys.argv[8]) if len(sys.argv)>8 else 4
tmo=int(sys.argv[9]) if len(sys.argv)>9 else 50
per=int(sys.argv[10]) if len(sys.argv)>10 else 20
acc=int(sys.argv[11]) if len(sys.argv)>11 else 10
prm = GSFSignatureParameters(N, thr, pair, tmo, per, acc, dead, nb, nl)
p = GSFSignature(prm, _api=api); p.init()
o = OracleGSF(N, prm.threshold, pair, tmo, per, acc, prm.nodes_down, nb, nl); o.init()
def cmp(tag):
    ok=True
    a=p.network().attrs(); b=o.attrs()
    for k in a:
        if not (a[k]==b[k]).all(): print(tag,"attrs differ",k); ok=False
    if p.network().rng_state()!=o.rng_state(): print(tag,"rng differ", p.network().rng_state(), o.rng_state()); ok=False
    if p.network().msgs_size()!=o.msgs_live(): print(tag,"msgs differ",p.network().msgs_size(),o.msgs_live()); ok=False
    if not (p.verified()==o.verified()).all(): print(tag,"verified differ", np.argwhere((p.verified()!=o.verified()).any(axis=1))[:5].ravel()); ok=False
    if not (p.network().counters()==o.counters()).all():
        d=(p.network().counters()!=o.counters()); print(tag,"counters differ rows",np.argwhere(d.any(axis=1)).ravel(), "nodes", np.argwhere(d.any(axis=0))[:5].ravel()); ok=False
    s1=p.scalars(); s2=o.scalars()
    for k in s1:
        if not (s1[k]==s2[k]).all(): print(tag,"scalars differ",k, np.argwhere(s1[k]!=s2[k])[:5].ravel()); ok=False
    l1=p.level_scalars(); l2=o.level_scalars()
    for k in l1:
        if not (l1[k]==l2[k]).all(): print(tag,"level scalars differ",k, np.argwhere(l1[k]!=l2[k])[:5]); ok=False
    for w in (1,2):
        if not (p.rows(w)==o.level_rows(w)).all(): print(tag,"rows differ",w); ok=False
    return ok
# peers
for n in range(0,N,max(1,N//16)):
    for l in range(p.levels):
        if not (p.peers(n,l)==o.peers(n,l)).all(): print("peers differ",n,l); sys.exit(1)
if not cmp("init"): sys.exit(1)
while p.network().time < T:
    p.network().run_ms(step); o.run_ms(step)
    if not cmp("t=%d"%p.network().time): sys.exit(1)
print("OK up to", p.network().time, p.network().stats(), "cont", p.continue_if(), o.continue_if())
