import sys, time
sys.path.insert(0, '/root/repo')
import __graft_entry__ as ge
ge.build()
import xsparse_b200 as xsb
import torch
nx=128; layers=127; world=2
nzn=layers*world+1
N=nx*nx*nzn; plane=nx*nx
splits=[plane*layers*r for r in range(world)]+[N]
h=xsb.Handle(N,N,slab=(world,0,splits))
g=xsb.Handle(nx**3,nx**3)
def T(f):
    torch.cuda.synchronize(); h.synchronize(); t=time.perf_counter(); r=f(); h.synchronize(); g.synchronize(); return (time.perf_counter()-t)*1e3, r
for it in range(4):
    h.reset(); g.reset()
    te,_=T(lambda: h.emit_p1fem(nx,nx,nzn,flavour=xsb.RAW,cz_range=(0,layers)))
    tg,_=T(lambda: g.emit_p1fem(nx,nx,nx,flavour=xsb.RAW))
    tc,c=T(lambda: h.route_count())
    buf=torch.empty(2*max(sum(c)-c[0],1),dtype=torch.int64,device='cuda')
    tp,_=T(lambda: h.route_prepare(buf,sum(c)-c[0]))
    tf,_=T(lambda: h.flush())
    tf2,_=T(lambda: g.flush())
    print(f"emit slab {te:.3f} plain {tg:.3f}  route_count {tc:.3f} prepare {tp:.3f} flush slab {tf:.3f} plain {tf2:.3f}", c)
