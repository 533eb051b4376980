import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, lzbench_b200 as L
from tests import oracle_lib as O
def seqs(blk):
    ip=0; out=[]
    while True:
        t=blk[ip]; ip+=1; l=t>>4
        if l==15:
            while True:
                b=blk[ip]; ip+=1; l+=b
                if b!=255: break
        ip+=l
        if ip>=len(blk): out.append((l,0,0)); return out
        off=blk[ip]|(blk[ip+1]<<8); ip+=2; m=t&15
        if m==15:
            while True:
                b=blk[ip]; ip+=1; m+=b
                if b!=255: break
        out.append((l,off,m+4))
parts=[('repeat', 11, 83), ('mixed', 6663, 451), ('zero', 9132, 863), ('mixed', 20, 361)]
d=np.concatenate([L.corpus(k,n,seed=s) for k,n,s in parts])
cs=L.chunk_list(len(d),16384)
ctx=L.Lz4B200(0)
res={}
for g in ('v2','v3'):
    ctx.set_option('compress',g)
    packed,compr,total=ctx.compress_batch(d,cs)
    res[g]=seqs(bytes(packed[:total])); print(g,total,len(res[g]),'sequences, literals',sum(s[0] for s in res[g]))
ref=O.oracle_compress(d); r=seqs(bytes(ref)); print('ref',len(ref),len(r),'sequences, literals',sum(s[0] for s in r))
pos=0
a,b=res['v3'],r
# first divergence by position
def spans(sq):
    p=0; out=[]
    for l,o,m in sq: out.append((p,l,o,m)); p+=l+m
    return out
sa,sb=spans(a),spans(b)
big=[s for s in sa if s[1]>40]; print('v3 literal runs > 40:',big[:20])
big=[s for s in sb if s[1]>40]; print('ref literal runs > 40:',big[:20])
