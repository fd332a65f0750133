import sys, numpy as np, time
sys.path.insert(0,'/root/repo')
from oracle import gen_inputs as g, pool_ref
dem=g.pool_demand(); dist=g.stand_distances(50)
recs,st=pool_ref.find(dem,dist,4,0,8,dedup=False,cap=1<<22)
print(st, recs.shape)
# per-tuple best: group by pickup tuple (cols 0..3), keep min (cost, order of appearance)
key_t = (recs[:,0].astype(np.int64)<<30)|(recs[:,1].astype(np.int64)<<20)|(recs[:,2].astype(np.int64)<<10)|recs[:,3]
order=np.lexsort((np.arange(len(recs)), recs[:,8], key_t))
kt=key_t[order]; first=np.ones(len(order),bool); first[1:]=kt[1:]!=kt[:-1]
R=recs[order[first]]
seq=order[first]  # enumeration index as rank
print('records after per-tuple best', len(R))
cost=R[:,8].astype(np.int64); rank=seq.astype(np.int64)
key=cost*(1<<32)+rank
cust=R[:,:4]
n=722
def rounds_on(idx, alive):
    """dominance rounds on record subset idx; returns kept list, total visits, rounds"""
    visits=0; rounds=0; kept=[]
    live=idx[alive[cust[idx]].all(1)]
    while len(live):
        rounds+=1; visits+=len(live)
        best=np.full(n, np.iinfo(np.int64).max)
        np.minimum.at(best, cust[live].ravel(), np.repeat(key[live],4))
        dom=(best[cust[live]]==key[live][:,None]).all(1)
        sel=live[dom]; kept+=list(sel)
        alive[cust[sel].ravel()]=False
        live=live[alive[cust[live]].all(1)]
    return kept, visits, rounds
alive=np.ones(n,bool)
k,v,r=rounds_on(np.arange(len(R)), alive)
print('global: kept',len(k),'rounds',r,'record-visits',v, 'x3 passes')
# banded by cost
alive=np.ones(n,bool); tv=0; tr=0; kept=[]; filt=0
levels=np.unique(cost)
srt=np.argsort(cost,kind='stable')
bounds=np.searchsorted(cost[srt], levels)
hist=np.diff(np.append(bounds,len(R)))
print('cost levels', len(levels), 'hist head', list(zip(levels[:12],hist[:12])))
for li,l in enumerate(levels):
    idx=srt[bounds[li]:bounds[li]+hist[li]]
    filt+=len(idx)
    k2,v2,r2=rounds_on(idx, alive); kept+=k2; tv+=v2; tr+=r2
print('banded per level: kept',len(kept),'rounds',tr,'visits in rounds',tv,'filter visits',filt)
print(sorted(k)==sorted(kept))

# banded with geometric growth (what the kernel does)
for growth, band0 in ((4,4096),(2,4096),(2,1024),(1.5,2048)):
    alive=np.ones(n,bool); tv=0; tr=0; kept=[]; mx=0; nb=0
    cum=np.cumsum(hist); target=band0; lo=0; bands=[]
    for li in range(len(levels)):
        if cum[li]>=target and li<len(levels)-1:
            bands.append((lo,li+1)); lo=li+1; target=cum[li]*growth
    bands.append((lo,len(levels)))
    per=[]
    for (a0,a1) in bands:
        idx=srt[bounds[a0]:(bounds[a1] if a1<len(levels) else len(R))]
        live0=int(alive[cust[idx]].all(1).sum())
        k2,v2,r2=rounds_on(idx, alive); kept+=k2; tv+=v2; tr+=r2; mx=max(mx,live0); per.append((len(idx),live0,r2,v2))
    print('growth',growth,'band0',band0,'bands',len(bands),'rounds',tr,'visits',tv,'max live',mx, sorted(k)==sorted(kept)); print('   per band (size, live, rounds, visits):',per)
