"""CPU prototype of the K3 far-column bound (written before the kernel): with the exact P of the C oracle, how many far
tiles / cells survive the tilted corner bound at 128x128, 32x32 and 4x8?  python tools/k3_bound_prototype.py 1|3"""
import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np
from pasio_b200 import synth
from oracle import c_oracle, pasio_oracle as po

which = int(sys.argv[1])
if which == 1:
    counts = synth.piecewise_poisson(100000, 0); cands = np.arange(len(counts)+1)
else:
    n = 400000
    counts = synth.piecewise_poisson(n, 1); cands = synth.random_candidates(n, n//10, 1)
fo = c_oracle.FlatOracle(counts, 1.0, 1.0)
t0=time.time()
score, splits, P, prev = fo.square_split(cands)
print('oracle', time.time()-t0, 's; N', len(cands), 'splits', len(splits))
C = fo.Cg[cands].astype(np.int64); L = cands.astype(np.int64)
G = fo.g; Lg = fo.lg
N = len(cands)
def self_score(i, j):
    s = C[j]-C[i]+1
    return G[s] - s*Lg[L[j]-L[i]]
best = P - 0.0   # pen = 0
# distance of argmax
d = np.arange(N) - prev
print('argmax distance percentiles', np.percentile(d[1:], [50,90,99,99.9,100]))
RB = 128
def tilt_fit(idx):
    x = C[idx]-C[idx[0]]; y = L[idx]-L[idx[0]]; p = P[idx]-P[idx[0]]
    A = np.stack([x,y,np.ones(len(idx))],1).astype(float)
    sol,_,_,_ = np.linalg.lstsq(A, -p, rcond=None)
    return sol[0], sol[1]
def box_max(u_lo,u_hi,l_lo,l_hi,a,b):
    out=-np.inf
    for u in (u_lo,u_hi):
        for l in (l_lo,l_hi):
            s=u+1
            out=max(out, G[s]-s*Lg[l]+a*u+b*l)
    return out
rs = np.random.RandomState(0)
nb = (N-1)//RB
tot_cells=0; surv0=0; surv1=0; evalcells=0
for b in rs.choice(np.arange(8, nb), size=12, replace=False):
    r0 = 1+RB*b; r1=min(r0+RB, N)
    rows = np.arange(r0,r1)
    e = 1+RB*(b-2)-1          # last final row when far is released (end of block b-3)
    anchors = set([e]+[int(prev[e-k]) for k in range(7)])
    LB = np.full(len(rows), -np.inf)
    for a_ in anchors:
        LB = np.maximum(LB, np.array([self_score(a_, j) for j in rows]) + P[a_])
    gap = best[rows]-LB
    # far columns: blocks <= b-3
    ncolb = b-2
    s0=0;s1=0;ev=0
    for c in range(ncolb):
        c0 = 0 if c==0 else 1+RB*c; c1 = 1+RB*(c+1)
        cols = np.arange(c0,c1)
        a,bb = tilt_fit(cols)
        mpt = np.max(P[cols]+a*C[cols]+bb*L[cols])
        ub = mpt + box_max(C[rows[0]]-C[cols[-1]], C[rows[-1]]-C[cols[0]], L[rows[0]]-L[cols[-1]], L[rows[-1]]-L[cols[0]], a, bb)
        rmin = np.min(LB + a*C[rows]+bb*L[rows])
        if ub - rmin >= 0:
            s0+=1
            # level 1: 32x32
            for rq in range(0,len(rows),32):
                rr = rows[rq:rq+32]
                for cq in range(0,len(cols),32):
                    cc = cols[cq:cq+32]
                    a1,b1 = tilt_fit(cc)
                    mpt1 = np.max(P[cc]+a1*C[cc]+b1*L[cc])
                    ub1 = mpt1+box_max(C[rr[0]]-C[cc[-1]], C[rr[-1]]-C[cc[0]], L[rr[0]]-L[cc[-1]], L[rr[-1]]-L[cc[0]], a1,b1)
                    rmin1 = np.min(LB[rq:rq+32]+a1*C[rr]+b1*L[rr])
                    if ub1-rmin1>=0:
                        s1+=1
                        # level 2: 4x8
                        for r4 in range(0,len(rr),4):
                            r_ = rr[r4:r4+4]
                            for c8 in range(0,len(cc),8):
                                c_ = cc[c8:c8+8]
                                mp = np.max(P[c_]+a1*C[c_]+b1*L[c_])
                                ub2 = mp+box_max(C[r_[0]]-C[c_[-1]], C[r_[-1]]-C[c_[0]], L[r_[0]]-L[c_[-1]], L[r_[-1]]-L[c_[0]], a1,b1)
                                rm2 = np.min(LB[rq+r4:rq+r4+4]+a1*C[r_]+b1*L[r_])
                                if ub2-rm2>=0: ev += len(r_)*len(c_)
    far_cells = len(rows)*(1+RB*(b-2))
    print('block %d: anchors %d gap(LB) med %.1f max %.1f | far tiles %d surv128 %d surv32 %d (of %d) eval cells %d = %.3f%% of far, per row %.1f'
          % (b, len(anchors), np.median(gap), gap.max(), ncolb, s0, s1, s0*16, ev, 100.0*ev/far_cells, ev/len(rows)))
