"""CPU simulation behind DESIGN 4.1 (chunk-compact sampler layout): how many points a pick has to touch when the exact
spatial pruning of fps_bucket.cu is applied per warp (512 consecutive Morton-sorted points), per half of a warp's chunks (256,
what the kernel does since round 2), per 128-point chunk, per 64 points and per lane (16), with the group's own maximum or
the warp's posted maximum as the threshold.  numpy only; runs in ~1 min:  python scripts/fps_prune_sim.py
(The sort is the kernel's: 15 Morton bits handed out greedily to the axis with the largest cell extent; the pruning test
here uses exact float64 box distances, so the counts are a lower bound of the kernel's by a hair.)"""
import sys, numpy as np
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import bench
xyz_np = bench.make_inputs(16, 0)[0]
def run(x, m=4096, bits=15):
    n = x.shape[0]
    lo, hi = x.min(0), x.max(0); e = hi-lo
    nb=[0,0,0]; c=e.copy(); order=[]
    for i in range(bits):
        ax=0
        if c[1]>c[0] and c[1]>=c[2]: ax=1
        if c[2]>c[0] and c[2]>c[1]: ax=2
        order.append(ax); nb[ax]+=1; c[ax]*=0.5
    q=[np.clip(np.floor((x[:,a]-lo[a])*((1<<nb[a])/e[a])).astype(np.int64),0,(1<<nb[a])-1) for a in range(3)]
    r=nb[:]; key=np.zeros(n,np.int64)
    for ax in order:
        r[ax]-=1; key=(key<<1)|((q[ax]>>r[ax])&1)
    p=np.argsort(key,kind='stable'); xs=x[p].astype(np.float32)
    def boxes(g):
        v=xs.reshape(n//g,g,3); return v.min(1), v.max(1)
    G={'warp512':512,'half256':256,'quarter128':128,'q64':64,'lane16':16}
    WM={'half256':'half_warpmax','quarter128':'quarter_warpmax','lane16':'lane_warpmax','q64':'q64_warpmax'}
    B={k:boxes(g) for k,g in G.items()}
    md=np.full(n,1e10,np.float32)
    cur=int(np.where(p==0)[0][0])
    stats={k:0 for k in G}; stats['half_warpmax']=0; stats['quarter_warpmax']=0; stats['lane_warpmax']=0; stats['q64_warpmax']=0
    for j in range(1,m):
        pk=xs[cur]
        d=((xs-pk)**2).sum(1).astype(np.float32)
        # tests with the state BEFORE the update (as the kernel does)
        wmax=md.reshape(-1,512).max(1)
        for k,g in G.items():
            blo,bhi=B[k]
            dd=np.maximum(np.maximum(blo-pk,pk-bhi),0); bd=(dd**2).sum(1)
            gmax=md.reshape(-1,g).max(1)
            stats[k]+= int((bd<gmax).sum())*g
            if k != 'warp512':  # the same boxes against the WARP's maximum (what the helper warps test)
                stats[WM[k]] += int((bd < np.repeat(wmax, 512 // g)).sum()) * g
        md=np.minimum(md,d)
        cur=int(md.argmax())
    return {k:v/(m-1) for k,v in stats.items()}
for ci in (0,9):
    s=run(xyz_np[ci].astype(np.float32))
    print(ci, {k:round(v) for k,v in s.items()})
