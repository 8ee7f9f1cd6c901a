"""How often two sibling targets share an M2L candidate, from the oracle's lists (CPU only): the basis of the two-targets-per-warp
M2L kernel that was written, shown not to pay (two-wide FP32 costs two issue cycles on B200) and removed.    python tests/tools/m2l_pairing.py N CAPACITY
Prints, for the three ways of pairing the eight children, evaluations needed when a pair of targets shares one evaluation, and the
issue-slot ratio against one evaluation per (target, candidate) when a shared evaluation runs at the higher of the two orders."""
import sys, numpy as np
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import oracle
from nbody_b200 import workloads
n=int(sys.argv[1]); cap=int(sys.argv[2])
P=workloads.plummer(n)
sk,perm=oracle.sort_keys(oracle.morton_keys(P[:,0:3],(1.,1.,1.)))
t=oracle.Tree(sk,(1.,1.,1.),cap,21)
m2l,p2p=t.traverse(0.5)
cnt=np.asarray(t.leaf_count,np.int64)
nn=len(cnt)
par=np.arange(nn)+np.asarray(t.parent_off,np.int64)
a=np.concatenate([m2l[:,0],m2l[:,1]]).astype(np.int64); b=np.concatenate([m2l[:,1],m2l[:,0]]).astype(np.int64)
# child slot of a within its parent: children contiguous: first child index = par + child_off[par]
sib=np.asarray(t.sibling,np.int64)
slot=sib[a]
ok=(a!=0)
print('directed m2l pairs',len(a),'with sibling slot',ok.sum())
a=a[ok]; b=b[ok]; slot=slot[ok]; pa=par[a]
for name,grp in (('x-adjacent (slot>>1)',slot>>1),('z-adjacent (slot&3)',slot&3),('y-adjacent',(slot&1)|((slot>>2)<<1))):
    key=(pa*nn+b)*4+grp
    u,c=np.unique(key,return_counts=True)
    print(name,': packed evaluations',len(u),'for',len(a),'pairs -> utilisation',len(a)/(2*len(u)), 'issue ratio vs scalar', len(u)/len(a))
g=np.asarray(t.geom,np.float64)
d2=((g[a,:3]-g[b,:3])**2).sum(1); ext=g[a,3]+g[b,3]; lo=(0.75*ext*ext)<0.13*d2
print('low-order fraction of pairs',lo.mean())
key=(pa*nn+b)*4+(slot>>1)
u,inv=np.unique(key,return_inverse=True)
any_hi=np.zeros(len(u),bool); np.logical_or.at(any_hi,inv,~lo)
F4,F3=546.,229.
scalar=( (~lo).sum()*F4+lo.sum()*F3 )
packed=( any_hi.sum()*F4+(~any_hi).sum()*F3 )   # issue cost of a packed evaluation ~ one scalar evaluation of that order
print('packed evals: hi',any_hi.sum(),'lo',(~any_hi).sum(),' issue-slot ratio packed/scalar (flop-weighted)',packed/scalar)
