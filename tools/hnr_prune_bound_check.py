"""Measurement behind the decision NOT to prune the harmonicity refinement: for every correlation
maximum of forward-cross-correlation frames of two synthetic clips, the sinc70/700 + Brent refined strength (CPU oracle) is
compared with the candidate bound r[i] + d2r + 1e-3.  Result on the committed generator: 371 of 13,024 maxima exceed the
bound (worst by 0.22), all on noise-like rows -- the bound is not rigorous, so k_hnr_refine refines every maximum.

    python tools/hnr_prune_bound_check.py
"""
import sys, numpy as np
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import mshds_oracle as orc
from robust_speech_analysis_framework_b200.synth import synth_clip
fs=16000.0; dx=1/fs
worst=0; viol=0; tot=0
for idx in (100,101):
    x=orc.pcm_to_float(synth_clip(idx,3.0).numpy())
    floor=60.0 if idx%2==0 else 100.0
    ppw=4.5; dtw=ppw/floor; W=int(np.floor(dtw/dx)); half=W//2-1; W=2*half
    maxlag=int(np.floor(W/ppw))+2
    nper=int(np.floor(fs/floor))
    B=W
    for t in np.arange(0.3,2.7,0.013):
        start=int(np.floor((t-0.5*(1/floor+dtw))/dx))
        seg=x[start:start+maxlag+W].copy()
        c=int(round(t/dx)); lm=x[c-nper:c+nper].mean()
        seg-=lm
        sx2=(seg[:W]**2).sum()
        r=np.zeros(B+1); r[0]=1
        for lag in range(1,maxlag+1):
            sy2=(seg[lag:lag+W]**2).sum()
            r[lag]=(seg[:W]*seg[lag:lag+W]).sum()/np.sqrt(sx2*sy2)
        y=np.concatenate([r[:0:-1],r])   # symmetric, centre index B (0-based) => 1-based centre B+1
        for i in range(2,maxlag):
            if r[i]>0 and r[i]>r[i-1] and r[i]>=r[i+1]:
                d2r=2*r[i]-r[i-1]-r[i+1]; dr=0.5*(r[i+1]-r[i-1])
                f=1/dx/(i+dr/d2r)
                mode=4 if f>0.3/dx else 3
                v,xr=orc.improve_extremum(y, i+B+1, mode, True)
                ub=r[i]+d2r+1e-3
                tot+=1
                if v>ub:
                    viol+=1; worst=max(worst,v-ub)
                    if viol<10: print("viol idx",idx,"t",t,"lag",i,"r",r[i-1:i+2],"refined",v,"ub",ub,"x",xr-(B+1))
print("total",tot,"violations",viol,"worst",worst)
