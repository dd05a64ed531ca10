import sys, os
sys.path.insert(0,'/tmp'); sys.path.insert(0,'/root/repo')
from multi import *
P=(1<<64)-(1<<32)+1
def solve_base(A,b):
    n,m=len(A),len(A[0]); M=[list(r)+[bb] for r,bb in zip(A,b)]; piv=[]; r=0
    for col in range(m):
        pr=next((i for i in range(r,n) if M[i][col]%P),None)
        if pr is None: continue
        M[r],M[pr]=M[pr],M[r]; inv=pow(M[r][col],P-2,P); M[r]=[x*inv%P for x in M[r]]
        for i in range(n):
            if i!=r and M[i][col]%P:
                f=M[i][col]; M[i]=[(x-f*y)%P for x,y in zip(M[i],M[r])]
        piv.append(col); r+=1
    sol=[None]*m
    for i,col in enumerate(piv): sol[col]=M[i][m]
    return sol, all(M[i][m]%P==0 for i in range(r,n)), len(piv)
def run_base(proofs,variants,unknown,order=("lookup","bool","gates","cp"),nr="boojum",split=None,tag=""):
    rows=[(build_row(c,ch,o,variants,order,nr,split),rhs(c,ch,o)) for c,ch,o,pr in proofs]
    labels=sorted({l for row,_ in rows for l in row if any(l==u or l.startswith(u+":") for u in unknown)})
    A=[];b=[]
    for row,r in rows:
        fixed=ZERO
        for l,v in row.items():
            if l not in labels: fixed=eadd(fixed,v)
        rr=esub(r,fixed)
        A.append([row.get(l,ZERO)[0] for l in labels]); b.append(rr[0])
        A.append([row.get(l,ZERO)[1] for l in labels]); b.append(rr[1])
    sol,ok,rank=solve_base(A,b)
    print(tag,"eq",len(A),"unknowns",len(labels),"rank",rank,"consistent",ok)
    if ok:
        for l,s in zip(labels,sol): print("    ",l,s,"=1" if s==1 else "=-1" if s==P-1 else "")
    return ok
if __name__=="__main__":
    proofs=[load(f"{R}/test_proofs/recursion_layer/node_layer_proof_{t}_0_0.json", f"{R}/setup/recursion_layer/vk_node.json", "recursion") for t in range(3,16)]
    for t in range(3,16):
        f=f"/tmp/rvk/vk_leaf_{t}.json"
        if os.path.exists(f): proofs.append(load(f"{R}/test_proofs/recursion_layer/leaf_layer_proof_{t}_0.json", f, "recursion"))
    allg=["bool","cp","cpL0","ConstantsAllocator","Poseidon2Flattened","ZeroCheck","FmaBaseNoConst","FmaExt","UIntXAdd","Selection","ParallelSelection4","Reduction4"]
    rounds=lambda i:(i//12 if i<48 else 4+(i-48) if i<70 else 26+(i-70)//12)
    for alt in (0,1,2,3):
        run_base(proofs,{"UIntXAdd":2,"Poseidon2Flattened":alt<<1},allg,order=("bool","gates","cp"),split={"Poseidon2Flattened":rounds},tag=f"P2 alt {alt} by 30 rounds")
