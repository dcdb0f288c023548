# dev helper (test infrastructure): CasperIMD on the host emulation of the device logic vs the oracle, step by step
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests.oracle_lib import OracleCasper
from wittgenstein_b200 import CasperIMD, CasperParemeters
cyc=int(sys.argv[1]); bpc=int(sys.argv[2]); apr=int(sys.argv[3]); step=int(sys.argv[4]); T=int(sys.argv[5])
delay=int(sys.argv[6]) if len(sys.argv)>6 else 0
nb=sys.argv[7] if len(sys.argv)>7 and sys.argv[7]!='-' else None; nl=sys.argv[8] if len(sys.argv)>8 and sys.argv[8]!='-' else None
seed=int(sys.argv[9]) if len(sys.argv)>9 else None
use_emu=os.environ.get("WTG_TEST_EMU","1")=="1"
api=None
if use_emu:
    from tests import emu_lib
    api=emu_lib.api()
p=CasperIMD(CasperParemeters(cyc,False,bpc,apr,1000,1,nb,nl), _api=api)
o=OracleCasper(cyc,False,bpc,apr,1000,1,nb,nl)
if seed is not None: p.network().set_seed(seed); o.set_seed(seed)
p.network().set_tunable('casper_votes', T//(8000*cyc)+3)
kind=os.environ.get('BYZ','WF')
p.init(delay, kind); o.init(delay, kind)
def cmp(tag):
    ok=True
    if p.network().rng_state()!=o.rng_state(): print(tag,"rng differ"); ok=False
    if p.network().msgs_size()!=o.msgs_live(): print(tag,"msgs differ",p.network().msgs_size(),o.msgs_live()); ok=False
    if not (p.network().counters()==o.counters()).all():
        d=(p.network().counters()!=o.counters()); print(tag,"counters differ rows",np.argwhere(d.any(axis=1)).ravel(),"nodes",np.argwhere(d.any(axis=0))[:5].ravel()); ok=False
    a=p.node_state(); b=o.node_state()
    for k in a:
        if not (a[k]==b[k]).all(): print(tag,"node state differs",k,np.argwhere(a[k]!=b[k])[:5].ravel(), a[k][a[k]!=b[k]][:5], b[k][a[k]!=b[k]][:5]); ok=False
    a=p.blocks(); b=o.blocks()
    for k in a:
        if len(a[k])!=len(b[k]) or not (a[k]==b[k]).all(): print(tag,"blocks differ",k,a[k][-5:],b[k][-5:]); ok=False
    if p.byz()!=o.byz(): print(tag,"byz differs",p.byz(),o.byz()); ok=False
    return ok
if not cmp("init"): sys.exit(1)
while p.network().time<T:
    r1=p.network().run_ms(step); r2=o.run_ms(step)
    if r1!=r2: print("ret differs", r1, r2, o.time); sys.exit(1)
    if not cmp("t=%d"%o.time): sys.exit(1)
b=p.blocks()
for i in range(1,len(b['height'])):
    if p.block_attestations(i)!=o.block_attestations(i): print("block atts differ", i); sys.exit(1)
print("OK", o.time, "blocks", len(b['height']), "heights", b['height'][-6:], "observer head", p.node_state()['head'][0], "byz", p.byz(), "deliveries", o.deliveries())
