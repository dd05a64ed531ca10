import sys, os, itertools, time
sys.path.insert(0,'/tmp'); sys.path.insert(0,'/root/repo')
from basefield import *
from tools import golden_gates as GG
proofs=[load(f"{R}/test_proofs/recursion_layer/node_layer_proof_{t}_0_0.json", f"{R}/setup/recursion_layer/vk_node.json", "recursion") for t in range(3,16)]
for t in range(3,16):
    f=f"/tmp/rvk/vk_leaf_{t}.json"
    if os.path.exists(f): proofs.append(load(f"{R}/test_proofs/recursion_layer/leaf_layer_proof_{t}_0.json", f, "recursion"))
ALTS={"Selection":4,"ParallelSelection4":8,"ZeroCheck":4,"UIntXAdd":[1,2],"Poseidon2Flattened":[0,2],"FmaBaseNoConst":2}
MULTI={"ZeroCheck":2,"FmaExt":2,"ParallelSelection4":4,"UIntXAdd":2}
rounds=lambda i:(i//12 if i<48 else 4+(i-48) if i<70 else 26+(i-70)//12)
# per proof: list of groups; each group: dict key->(dict label->value), all with same term count
table=[]
for (c,ch,o,pr) in proofs:
    alpha=ch["alpha"]; cells=o["perm"][:c["n_copy"]]; ap=ONE; groups=[]
    def cls_sum(terms, ap, labeler):
        d={}
        for i,t in enumerate(terms):
            l=labeler(i); d[l]=eadd(d.get(l,ZERO),emul(ap,t)); ap=emul(ap,alpha)
        return d,ap
    b=o["perm"][c["n_copy"]]; d,ap=cls_sum([esub(emul(b,b),b)],ap,lambda i:"bool"); groups.append({0:d})
    for name,nc,deg,path in c["gates"]:
        sel=ONE
        for bi,bit in enumerate(path): sel=emul(sel,o["const"][bi] if bit else esub(ONE,o["const"][bi]))
        a=ALTS.get(name,1); alts=a if isinstance(a,list) else list(range(a))
        g={}; nxt=None
        for alt in alts:
            rel=GG.eval_gate(name,c,cells,o["const"][len(path):],nc,alt<<1)
            if not rel: continue
            rel=[emul(r,sel) for r in rel]
            if name=="Poseidon2Flattened": lab=lambda i:f"P2:{rounds(i)}"
            elif name in MULTI and not (name=="UIntXAdd" and alt==0): lab=lambda i,n=name:f"{n}:{i%MULTI[n]}"
            else: lab=lambda i,n=name:n
            g[(alt,0)],nxt=cls_sum(rel,ap,lab)
            if name in MULTI and name!="UIntXAdd":
                R_=MULTI[name]; I=len(rel)//R_
                rm=[rel[i*R_+r] for r in range(R_) for i in range(I)]
                lab2=lambda i,n=name,I=I:f"{n}:{i//I}"
                g[(alt,1)],_=cls_sum(rm,ap,lab2)
        if g: groups.append(g); ap=nxt
    g={}
    for nr in ("boojum","pow7"):
        t=copy_perm_terms(c,ch,o,nr)
        g[nr],nxt=cls_sum(t,ap,lambda i:"cpL0" if i==0 else "cp")
    groups.append(g); table.append((groups,rhs(c,ch,o)))
keys=[list(g.keys()) for g in table[0][0]]
print([len(k) for k in keys]); sys.stdout.flush()
n=0;t0=time.time()
for combo in itertools.product(*keys):
    labels=None; A=[];b=[]
    for groups,r in table:
        row={}
        for g,k in zip(groups,combo): row.update(g[k])
        if labels is None: labels=sorted(row)
        A.append([row[l][0] for l in labels]); b.append(r[0]); A.append([row[l][1] for l in labels]); b.append(r[1])
    sol,ok,rank=solve_base(A,b); n+=1
    if ok: print("CONSISTENT",combo,dict(zip(labels,sol))); sys.stdout.flush()
    if n%200==0: print(n,time.time()-t0); sys.stdout.flush()
print("done",n)
