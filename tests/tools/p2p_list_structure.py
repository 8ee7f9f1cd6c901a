"""Structure of the P2P lists of the reference-rule traversal, from the oracle (CPU only): how much of a sibling group's union list
each leaf uses, and how far a leaf's source entries merge into contiguous particle runs. Input to the leaf-kernel plan in DESIGN
section 10.    python tests/tools/p2p_list_structure.py N CAPACITY [plummer|uniform|two_galaxies]"""
import sys, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import oracle
from nbody_b200 import workloads
n=int(sys.argv[1]); cap=int(sys.argv[2]); kind=sys.argv[3] if len(sys.argv)>3 else 'plummer'
P=workloads.GENERATORS[kind](n)
sk,perm=oracle.sort_keys(oracle.morton_keys(P[:,0:3],(1.,1.,1.)))
t=oracle.Tree(sk,(1.,1.,1.),cap,21)
m2l,p2p=t.traverse(0.5)
cnt=np.asarray(t.leaf_count,np.int64); par=np.arange(len(cnt))+np.asarray(t.parent_off,np.int64)
a=np.concatenate([p2p[:,0],p2p[:,1]]).astype(np.int64); b=np.concatenate([p2p[:,1],p2p[:,0]]).astype(np.int64)
# directed unique (self pairs appear twice after doubling -> unique)
key=np.unique(a*len(cnt)+b); a=key//len(cnt); b=key%len(cnt)
# per target leaf: sources particles
src_particles=np.bincount(a,weights=cnt[b],minlength=len(cnt))
pairs_total=(cnt[a]*cnt[b]).sum()
# union per parent of target: unique (parent(a), b)
pk=np.unique(par[a]*len(cnt)+b); pa=pk//len(cnt); pb=pk%len(cnt)
union_particles=np.bincount(pa,weights=cnt[pb],minlength=len(cnt))
# targets per parent
leaves=np.unique(a)
tg_per_parent=np.bincount(par[leaves],weights=cnt[leaves],minlength=len(cnt))
nleaf_per_parent=np.bincount(par[leaves],minlength=len(cnt))
groups=np.nonzero(nleaf_per_parent)[0]
union_pairs=(union_particles[groups]*tg_per_parent[groups]).sum()
print(f"{kind} N={n} cap={cap}: leaves {len(leaves)}, avg targets/leaf {cnt[leaves].mean():.1f}, sibling groups {len(groups)}, avg leaves/group {nleaf_per_parent[groups].mean():.2f}, avg targets/group {tg_per_parent[groups].mean():.1f}")
print(f" exact pair evaluations {pairs_total:.3e}; with union lists per sibling group {union_pairs:.3e} (x{union_pairs/pairs_total:.2f})")
print(f" source particles staged: per-leaf lists {src_particles[leaves].sum():.3e}; per-group union {union_particles[groups].sum():.3e} (x{union_particles[groups].sum()/src_particles[leaves].sum():.2f})")
# contiguous-run merging of each target leaf's source list (sources sorted by first particle)
begin=np.asarray(t.leaf_index,np.int64)
order=np.lexsort((begin[b],a))
aa=a[order]; bb=b[order]
same=(aa[1:]==aa[:-1])&(begin[bb[1:]]==begin[bb[:-1]]+cnt[bb[:-1]])
runs=len(aa)-same.sum()
print(f" P2P entries {len(aa):.3e} (avg {cnt[bb].mean():.1f} particles); merged into contiguous runs {runs:.3e} (avg {cnt[bb].sum()/runs:.1f} particles = {16*cnt[bb].sum()/runs:.0f} B)")
