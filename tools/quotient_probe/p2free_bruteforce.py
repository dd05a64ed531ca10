import sys, os, itertools, time
sys.path.insert(0,'/tmp'); sys.path.insert(0,'/root/repo')
from basefield import *
from tools import golden_gates as GG
proofs=[load_base(13)]+[load_base(t,f"/tmp/rvk/vk_base_{t}.json") for t in (3,5,6)]
ALTS={"Selection":[0,1,2,3],"ParallelSelection4":list(range(8)),"ZeroCheck":[0,1,2,3],"FmaBaseNoConst":[0,1]}
MULTI={"ZeroCheck":2,"ParallelSelection4":4}
def tables(order, ux_alt, nr, l0n):
    """per proof: (fixed dict of group->list of candidate values keyed by variant), unknown columns"""
    out=[]
    for (c,ch,o,pr) in proofs:
        alpha=ch["alpha"]; cells=o["perm"][:c["n_copy"]]; ap=ONE
        cand={}; unk={}
        def gs(terms,ap):
            acc=ZERO
            for t in terms: acc=eadd(acc,emul(ap,t)); ap=emul(ap,alpha)
            return acc,ap
        for part in order:
            if part=="lookup":
                t=lookup_terms(c,ch,o); unk["lookupA"],ap=gs(t[:-1],ap); unk["lookupB"],ap=gs(t[-1:],ap)
            elif part=="bool":
                b=o["perm"][c["n_copy"]]; v,ap=gs([esub(emul(b,b),b)],ap); cand["bool"]={0:v}
            elif part=="gates":
                for name,nc,deg,path in c["gates"]:
                    sel=ONE
                    for bi,bit in enumerate(path): sel=emul(sel,o["const"][bi] if bit else esub(ONE,o["const"][bi]))
                    alts=[ux_alt] if name=="UIntXAdd" else ALTS.get(name,[0])
                    d={}; nxt=None
                    for alt in alts:
                        rel=GG.eval_gate(name,c,cells,o["const"][len(path):],nc,alt<<1)
                        if not rel: continue
                        rel=[emul(r,sel) for r in rel]
                        d[(alt,0)],nxt=gs(rel,ap)
                        if name in MULTI:
                            R_=MULTI[name]; I=len(rel)//R_
                            d[(alt,1)],_=gs([rel[i*R_+r] for r in range(R_) for i in range(I)],ap)
                    if d: cand[name]=d; ap=nxt
            elif part=="cp":
                if l0n: os.environ["L0_NORMALIZED"]="1"
                else: os.environ.pop("L0_NORMALIZED",None)
                t=copy_perm_terms(c,ch,o,nr); unk["cpL0"],ap=gs(t[:1],ap); unk["cp"],ap=gs(t[1:],ap)
        out.append((cand,unk,rhs(c,ch,o)))
    return out
GATES=["ConstantsAllocator","ZeroCheck","FmaBaseNoConst","UIntXAdd","Selection","ParallelSelection4","Reduction4","bool"]
n=0; t0=time.time()
for order in (("lookup","bool","gates","cp"),("gates","bool","lookup","cp"),("bool","lookup","gates","cp"),("lookup","gates","bool","cp")):
  for ux in (0,1,2):
    for nr in ("boojum",):
        T=tables(order,ux,nr,0)
        keysets=[list(T[0][0][g].keys()) for g in GATES]
        for combo in itertools.product(*keysets):
            for signs in itertools.product((1,-1),repeat=len(GATES)):
                A=[];b=[]
                for cand,unk,r in T:
                    fixed=ZERO
                    for g,k,s in zip(GATES,combo,signs):
                        v=cand[g][k]; fixed=eadd(fixed,v if s==1 else eneg(v))
                    rr=esub(r,fixed); labels=["lookupA","lookupB","cpL0","cp"]
                    A.append([unk[l][0] for l in labels]); b.append(rr[0]); A.append([unk[l][1] for l in labels]); b.append(rr[1])
                sol,ok,rank=solve_base(A,b); n+=1
                if ok: print("CONSISTENT",order,ux,nr,combo,signs,sol); sys.stdout.flush()
        print(order,ux,n,time.time()-t0); sys.stdout.flush()
print("done",n)
