"""Prototype (numpy, CPU, statistics only -- float scores, not the integer spec): TWO-LEVEL weighted rendezvous.
Nodes fall into G groups by a hash of their own address; level 1 is a weighted rendezvous over the groups (weight = sum
of the live members), level 2 over the members of the winning group.  P(node) stays w_j / W exactly, the work per object
drops from M to G + M/G pair hashes, and the price is ~2x the minimal movement when a node leaves (its group's weight
changes).  Evidence for DESIGN.md section 8 item 7; nothing in the product uses this."""
import numpy as np
rng=np.random.default_rng(1)
def mix64(x):
    x=np.asarray(x,dtype=np.uint64)
    x^=x>>np.uint64(30); x*=np.uint64(0xBF58476D1CE4E5B9)
    x^=x>>np.uint64(27); x*=np.uint64(0x94D049BB133111EB)
    x^=x>>np.uint64(31); return x
def pair_u(keys,seeds):
    h=mix64(keys^np.uint64(0xD6E8FEB86659FD93)); a=(h&np.uint64(0xFFFFFFFF)).astype(np.uint32); b=((h>>np.uint64(32)).astype(np.uint32))|np.uint32(1); ab=a*b
    s0=(seeds&np.uint64(0xFFFFFFFF)).astype(np.uint32); m=((seeds>>np.uint64(32)).astype(np.uint32))|np.uint32(1); s2=(mix64(seeds^np.uint64(0xA0761D6478BD642F))&np.uint64(0xFFFFFFFF)).astype(np.uint32)
    p=s0[None,:]*b[:,None]+ab[:,None]
    return p*m[None,:]+s2[None,:]
def hrw(keys,seeds,w):
    u=pair_u(keys,seeds).astype(np.float64)+0.5
    with np.errstate(divide="ignore"):
        sc=-np.log(u/2**32)/w[None,:]
    sc[:,w==0]=np.inf
    return sc.argmin(1)
def hrw2(keys,seeds,w,G):
    grp=(mix64(seeds^np.uint64(0x1234567))>>np.uint64(40)).astype(np.int64)%G
    Wg=np.array([w[grp==g].sum() for g in range(G)],dtype=np.float64)
    gseeds=mix64(np.arange(G,dtype=np.uint64)*np.uint64(0x9E3779B97F4A7C15)+np.uint64(77))
    g=hrw(keys,gseeds,Wg)
    out=np.empty(len(keys),dtype=np.int64)
    for gg in range(G):
        sel=np.nonzero(g==gg)[0]
        members=np.nonzero(grp==gg)[0]
        if len(sel)==0: continue
        out[sel]=members[hrw(keys[sel],seeds[members],w[members].astype(np.float64))]
    return out
M,N=1024,200000
seeds=mix64(np.arange(M,dtype=np.uint64)*np.uint64(0x100000001b3)+np.uint64(5))
keys=mix64(np.arange(N,dtype=np.uint64)*np.uint64(0x9E3779B97F4A7C15)+np.uint64(9))
for wname,w in (("uniform",np.ones(M)),("1..16",rng.integers(1,17,M).astype(np.float64))):
    for G in (32,64):
        a=hrw2(keys,seeds,w,G)
        cnt=np.bincount(a,minlength=M); e=N*w/w.sum()
        chi=((cnt-e)**2/e).sum()
        # leave of node 17
        w2=w.copy(); w2[17]=0
        b=hrw2(keys,seeds,w2,G)
        moved=(a!=b).sum(); minimal=(a==17).sum()
        # join back is symmetric; flat HRW for comparison
        print("weights %-7s G=%2d: chi2 %.0f (df %d), leave(17): moved %d, minimal %d, ratio %.2f; hashes/object %d vs %d"%(wname,G,chi,M-1,moved,minimal,moved/max(minimal,1),G+M//G,M))
