import sys, os
sys.path.insert(0,'/root/repo')
from tools.golden_quotient import *
from tools.golden_linsolve import solve
from tools import golden_gates as GG
R="/root/reference"
def gsum_into(row,label,terms,ap,alpha):
    for t in terms:
        row[label]=eadd(row.get(label,ZERO),emul(ap,t)); ap=emul(ap,alpha)
    return ap
def build_row(c,ch,o,variants,order=("lookup","bool","gates","cp"),nr="boojum",split=None):
    alpha=ch["alpha"]; cells=o["perm"][:c["n_copy"]]+o["plain"]; ap=ONE; row={}
    for part in order:
        if part=="lookup" and c["LR"]:
            t=lookup_terms(c,ch,o)
            ap=gsum_into(row,"lookupA",t[:-1],ap,alpha); ap=gsum_into(row,"lookupB",t[-1:],ap,alpha)
        elif part=="bool" and c["has_bool"]:
            b=o["perm"][c["n_copy"]]; ap=gsum_into(row,"bool",[esub(emul(b,b),b)],ap,alpha)
        elif part=="gates":
            for name,nc,deg,path in c["gates"]:
                sel=ONE
                for bi,bit in enumerate(path): sel=emul(sel,o["const"][bi] if bit else esub(ONE,o["const"][bi]))
                rel=GG.eval_gate(name,c,cells,o["const"][len(path):],nc,variants.get(name,0))
                if rel:
                    sp=(split or {}).get(name)
                    if sp:
                        for i,r in enumerate(rel): ap=gsum_into(row,f"{name}:{sp(i)}",[emul(r,sel)],ap,alpha)
                    else: ap=gsum_into(row,name,[emul(r,sel) for r in rel],ap,alpha)
        elif part=="cp":
            t=copy_perm_terms(c,ch,o,nr)
            ap=gsum_into(row,"cpL0",t[:1],ap,alpha); ap=gsum_into(row,"cp",t[1:],ap,alpha)
    return row
def run(proofs,variants,unknown,order=("lookup","bool","gates","cp"),nr="boojum",split=None,tag=""):
    rows=[(build_row(c,ch,o,variants,order,nr,split),rhs(c,ch,o)) for c,ch,o,pr in proofs]
    labels=sorted({l for row,_ in rows for l in row if any(l==u or l.startswith(u+":") for u in unknown)})
    A=[];b=[]
    for row,r in rows:
        fixed=ZERO
        for l,v in row.items():
            if l not in labels: fixed=eadd(fixed,v)
        A.append([row.get(l,ZERO) for l in labels]); b.append(esub(r,fixed))
    sol,ok,rank=solve(A,b)
    print(tag,"eq",len(A),"unknowns",len(labels),"rank",rank,"consistent",ok)
    if ok and rank==len(labels) and len(A)>len(labels):
        for l,s in zip(labels,sol): print("    ",l,s,"=1" if s==ONE else "=-1" if s==eneg(ONE) else "")
    return ok
def load_base(t, vk=None):
    return load(f"{R}/test_proofs/base_layer/basic_circuit_proof_{t}_0.json", vk or f"{R}/setup/base_layer/vk_{t}.json", f"base_{t}")
if __name__=="__main__":
    proofs=[load_base(13)]
    for t in (3,5,6):
        f=f"/tmp/rvk/vk_base_{t}.json"
        if os.path.exists(f): proofs.append(load_base(t,f))
    print(len(proofs),"P2-free base proofs")
    for order in (("lookup","bool","gates","cp"),("gates","bool","lookup","cp"),("bool","lookup","gates","cp"),("lookup","gates","bool","cp")):
        for u in (0,2):
            run(proofs,{"UIntXAdd":u},["lookupA","lookupB"],order=order,tag=f"{order} uintx {u}: unknown lookup A,B")
            run(proofs,{"UIntXAdd":u},["lookupA","lookupB","cpL0"],order=order,tag=f"{order} uintx {u}: unknown lookup A,B,cpL0")
