import sys, os, itertools
sys.path.insert(0,'/tmp'); sys.path.insert(0,'/root/repo')
from basefield import *
from tools import golden_gates as GG
proofs=[load(f"{R}/test_proofs/recursion_layer/node_layer_proof_{t}_0_0.json", f"{R}/setup/recursion_layer/vk_node.json", "recursion") for t in range(3,16)]
for t in range(3,16):
    f=f"/tmp/rvk/vk_leaf_{t}.json"
    if os.path.exists(f): proofs.append(load(f"{R}/test_proofs/recursion_layer/leaf_layer_proof_{t}_0.json", f, "recursion"))
print(len(proofs),"proofs")
def monomials(nv,d):
    out=[()]
    for k in range(1,d+1): out+=list(itertools.combinations_with_replacement(range(nv),k))
    return out
GENERIC={"Selection":(4,0,2,1),"ZeroCheck":(3,0,2,2),"ConstantsAllocator":(1,1,1,1),"ParallelSelection4":None}  # width, consts per inst, degree, relations
def rows_for(p2alt, generic, uintx=2):
    rows=[]
    for (c,ch,o,pr) in proofs:
        alpha=ch["alpha"]; cells=o["perm"][:c["n_copy"]]; ap=ONE; row={}
        def add(label,terms):
            nonlocal ap
            for t in terms:
                row[label]=eadd(row.get(label,ZERO),emul(ap,t)); ap=emul(ap,alpha)
        b=o["perm"][c["n_copy"]]
        if "bool" in generic:
            row["bool:1"]=ap; row["bool:b"]=emul(ap,b); row["bool:bb"]=emul(ap,emul(b,b)); ap=emul(ap,alpha)
        else: add("bool",[esub(emul(b,b),b)])
        for name,nc,deg,path in c["gates"]:
            sel=ONE
            for bi,bit in enumerate(path): sel=emul(sel,o["const"][bi] if bit else esub(ONE,o["const"][bi]))
            kc=o["const"][len(path):]
            if name in generic and GENERIC.get(name):
                w,cpi,d,nrel=GENERIC[name]
                inst=GG.instances(name,c,nc)
                mons=monomials(w+cpi,d)
                for t in range(inst):
                    vars_=list(cells[w*t:w*t+w])+[kc[t*cpi+i] for i in range(cpi)]
                    mv=[]
                    for m in mons:
                        v=ONE
                        for idx in m: v=emul(v,vars_[idx])
                        mv.append(emul(v,sel))
                    for r in range(nrel):
                        for m,v in zip(mons,mv):
                            l=f"{name}:{r}:{m}"; row[l]=eadd(row.get(l,ZERO),emul(ap,v))
                        ap=emul(ap,alpha)
                continue
            if name=="ParallelSelection4" and name in generic:
                inst=GG.instances(name,c,nc); mons=monomials(4,2)
                for t in range(inst):
                    x=cells[13*t:13*t+13]
                    for i in range(4):
                        vars_=[x[0],x[1+3*i],x[2+3*i],x[3+3*i]]
                        for m in mons:
                            v=ONE
                            for idx in m: v=emul(v,vars_[idx])
                            l=f"ParSel:{m}"; row[l]=eadd(row.get(l,ZERO),emul(ap,emul(v,sel)))
                        ap=emul(ap,alpha)
                continue
            var={"UIntXAdd":uintx,"Poseidon2Flattened":p2alt<<1}.get(name,0)
            rel=GG.eval_gate(name,c,cells,kc,nc,var)
            if not rel: continue
            rel=[emul(r,sel) for r in rel]
            if name=="UIntXAdd" and uintx==2:
                for i,r in enumerate(rel): add(f"UIntXAdd:{i%2}",[r])
            elif name=="FmaExt":
                for i,r in enumerate(rel): add(f"FmaExt:{i%2}",[r])
            else: add(name,rel)
        t=copy_perm_terms(c,ch,o,"boojum"); add("cpL0",t[:1]); add("cp",t[1:])
        rows.append((row,rhs(c,ch,o)))
    return rows
def solve_rows(rows,tag):
    labels=sorted({l for row,_ in rows for l in row})
    A=[];b=[]
    for row,r in rows:
        A.append([row.get(l,ZERO)[0] for l in labels]); b.append(r[0]); A.append([row.get(l,ZERO)[1] for l in labels]); b.append(r[1])
    sol,ok,rank=solve_base(A,b)
    print(tag,"eq",len(A),"unknowns",len(labels),"rank",rank,"consistent",ok); sys.stdout.flush()
    if ok and rank<len(A):
        for l,s in zip(labels,sol):
            if s: print("    ",l,s if s<(1<<63) else s-P)
for p2alt in (0,2):
    for gen in (["Selection"],["ZeroCheck"],["ConstantsAllocator","bool"],["ParallelSelection4"],["Selection","ZeroCheck","ConstantsAllocator","bool"],["Selection","ParallelSelection4","ConstantsAllocator","bool"]):
        solve_rows(rows_for(p2alt,gen),f"P2 alt {p2alt} generic {gen}")
