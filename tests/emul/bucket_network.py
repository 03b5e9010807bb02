"""Python restatement of ss_bucket_kernel's network (csrc/sort_kernels.cu: ss_bucket_network + the tie check + the
A / B merge positions): the keys-only comparator first; when the bucket turns out to hold equal keys it is loaded again
and sorted with the (key, index) comparator.  Run by tests/test_host.py — it pins the properties the shortcut rests on:
with equal KEYS the two lanes of a shuffle pair disagree (the upper takes the lower's element, the lower keeps it), so
a row index can be lost — but never a key, the keys still come out sorted, and whenever an index was lost two equal
keys end up side by side, which is what the kernel checks for."""
import numpy as np
def gt_full(ka,va,kb,vb): return ka>kb or (ka==kb and va>vb)
def network(sk, sv, N2, Na, Nb, full):
    def cmpx(t,j,k):
        i=((t & ~(j-1))<<1)|(t&(j-1)); l=i|j; up=(i&k)==0
        gt = gt_full(sk[i],sv[i],sk[l],sv[l]) if full else sk[i]>sk[l]
        if gt==up:
            sk[i],sk[l]=sk[l],sk[i]; sv[i],sv[l]=sv[l],sv[i]
    def warp_stages(blk,k,j0):
        g0=[(blk<<6)+lane for lane in range(32)]; g1=[g+32 for g in g0]
        k0=[sk[g] for g in g0]; k1=[sk[g] for g in g1]; v0=[sv[g] for g in g0]; v1=[sv[g] for g in g1]
        jj=j0
        if jj==32:
            for lane in range(32):
                upa=(g0[lane]&k)==0
                gt = gt_full(k0[lane],v0[lane],k1[lane],v1[lane]) if full else k0[lane]>k1[lane]
                if gt==upa:
                    k0[lane],k1[lane]=k1[lane],k0[lane]; v0[lane],v1[lane]=v1[lane],v0[lane]
            jj=16
        while jj>0:
            nk0,nv0,nk1,nv1=list(k0),list(v0),list(k1),list(v1)
            for lane in range(32):
                o=lane^jj; lower=(lane&jj)==0
                upa=(g0[lane]&k)==0; upb=(g1[lane]&k)==0
                for (K,V,NK,NV,up) in ((k0,v0,nk0,nv0,upa),(k1,v1,nk1,nv1,upb)):
                    keep_min=(lower==up)
                    if full: take = gt_full(K[lane],V[lane],K[o],V[o])==keep_min
                    else: take = ((K[lane]>K[o])==keep_min)   # asymmetric on equal keys: an index may be lost
                    if take: NK[lane],NV[lane]=K[o],V[o]
            k0,v0,k1,v1=nk0,nv0,nk1,nv1
            jj>>=1
        for lane in range(32):
            sk[g0[lane]]=k0[lane]; sk[g1[lane]]=k1[lane]; sv[g0[lane]]=v0[lane]; sv[g1[lane]]=v1[lane]
    for blk in range(N2>>6):
        # phases 2..64 in registers: emulate by loading/storing each phase (equivalent)
        k=2
        while k<=64:
            warp_stages(blk,k,k>>1); k<<=1
    k=128
    while k<=Na:
        lim = N2 if k<=Nb else Na; npair=lim>>1
        j=k>>1
        while j>32:
            for t in range(npair): cmpx(t,j,k)
            j>>=1
        for blk in range(lim>>6): warp_stages(blk,k,32)
        k<<=1
def bucket(keys, idx):
    cnt=len(keys); Na=64
    while Na*2<=cnt: Na<<=1
    Nb=0
    if cnt>Na:
        Nb=64
        while Nb<cnt-Na: Nb<<=1
        if Nb>=Na: Na<<=1; Nb=0
    N2=Na+Nb
    SENT=(1<<64)-1
    sk=[int(keys[i]) if i<cnt else SENT for i in range(N2)]
    sv=[int(idx[i]) if i<cnt else 0xffffffff for i in range(N2)]
    def load():
        return ([int(keys[i]) if i<cnt else SENT for i in range(N2)],
                [int(idx[i]) if i<cnt else 0xffffffff for i in range(N2)])
    network(sk,sv,N2,Na,Nb,False)
    assert sorted(sk)==sorted(load()[0]), "the key multiset must survive the keys-only network"
    assert all(sk[i]<=sk[i+1] for i in range(Na-1)) and all(sk[i]<=sk[i+1] for i in range(Na,N2-1))
    tie=any(((i>0 and i!=Na and sk[i]==sk[i-1]) or sv[i]==0xffffffff) for i in range(cnt))
    if tie:
        sk,sv=load()
        network(sk,sv,N2,Na,Nb,True)
    else:
        assert sorted(zip(sk,sv))==sorted(zip(*load())), "no tie found, yet an element was lost"
    out=[None]*cnt
    for i in range(cnt):
        pos=i
        if Nb:
            inA=i<Na; lo,hi=(Na,cnt) if inA else (0,Na); first=lo
            while lo<hi:
                mid=(lo+hi)>>1
                if sk[mid]<sk[i] or (sk[mid]==sk[i] and sv[mid]<sv[i]): lo=mid+1
                else: hi=mid
            pos=(i if inA else i-Na)+(lo-first)
        assert out[pos] is None
        out[pos]=(sk[i],sv[i])
    return out, tie
def run_cases():
    rng=np.random.default_rng(0)
    for cnt in [1,5,63,64,65,100,128,187,192,200,300,500,513,700]:
        for mode in ("distinct","ties","const","fewties"):
            if mode=="distinct": keys=rng.permutation(10*cnt)[:cnt]
            elif mode=="ties": keys=rng.integers(0,max(2,cnt//8),size=cnt)
            elif mode=="const": keys=np.full(cnt,7)
            else:
                keys=rng.permutation(10*cnt)[:cnt]
                if cnt>3: keys[cnt//2]=keys[1]
            idx=rng.permutation(100000)[:cnt]
            out,tie=bucket(keys,idx)
            ref=sorted(zip([int(k) for k in keys],[int(v) for v in idx]))
            assert out==ref,(cnt,mode)


if __name__ == '__main__':
    run_cases()
    print('ok')
