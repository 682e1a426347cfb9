"""The SAH-optimal BVH8 collapse of csrc/bvh_build.cu (k_collapse_dp + dp_expand), restated in Python on random binary trees: the decisions
the bottom-up pass stores (flags byte + argmin split per slot budget) must expand, top-down, into nodes of at most 8 entries whose leaf
entries hold at most 3 triangles, cover every triangle exactly once, and realise exactly the optimum the recurrence computed. The CUDA code is
checked end to end on the GPU (frames bit-identical across collapse=0 / collapse=1, tests/test_gpu_golden.py); this CPU test pins the
recurrence and the decision encoding."""
import random

INF=float('inf')
def build(n):
    # random binary tree over n leaves; returns nodes dict: id -> (left,right); leaves are ('L',i)
    items=[('L',i) for i in range(n)]; nodes={}; cnt={}; area={}
    for i in range(n): area[('L',i)]=random.random()*0.1+0.01
    nid=0
    while len(items)>1:
        i=random.randrange(len(items)-1)
        a,b=items[i],items[i+1]
        nodes[nid]=(a,b)
        cnt[nid]=(1 if a[0]=='L' else cnt[a[1]])+(1 if b[0]=='L' else cnt[b[1]])
        area[('N',nid)]=area[a]+area[b]*random.uniform(0.3,1.0)+random.random()*0.05
        items[i:i+2]=[('N',nid)]; nid+=1
    return nodes,cnt,area,items[0]
def run(n,cTri):
    nodes,cnt,area,root=build(n)
    F={}; dec={}
    def Fget(r,j): return area[r]*cTri if r[0]=='L' else F[r[1]][j]
    for nid in sorted(nodes):   # children have smaller ids
        l,r=nodes[nid]; A=area[('N',nid)]; T=cnt[nid]
        S={};K={}
        for j in range(2,9):
            best=INF;bk=1
            for k in range(1,j):
                v=Fget(l,k)+Fget(r,j-k)
                if v<best: best=v;bk=k
            S[j]=best;K[j]=bk
        inner=A*1.0+S[8]; leaf=A*T*cTri if T<=3 else INF
        f={1:min(leaf,inner)}; flags=1 if leaf<=inner else 0
        for j in range(2,9):
            if S[j]<f[j-1]: f[j]=S[j]; flags|=1<<(j-1)
            else: f[j]=f[j-1]
        F[nid]=f; dec[nid]=(flags,K)
    # expand recursively into BVH8 nodes, computing realised cost
    total=[0.0]; covered=[]
    def leaves(r):
        if r[0]=='L': return [r[1]]
        l,rr=nodes[r[1]]; return leaves(l)+leaves(rr)
    def expand(rootref):
        stack=[(rootref,8,True)]; ents=[]
        while stack:
            ref,j,split=stack.pop()
            if ref[0]=='L': ents.append((ref,True)); continue
            flags,K=dec[ref[1]]
            if not split:
                while j>1 and not (flags>>(j-1))&1: j-=1
                if j==1: ents.append((ref,bool(flags&1))); continue
            k=K[j]; l,r=nodes[ref[1]]
            stack.append((r,j-k,False)); stack.append((l,k,False))
        return ents
    def make_node(ref):
        total[0]+=area[ref]*1.0
        ents=expand(ref); assert 1<=len(ents)<=8
        for e,isleaf in ents:
            if isleaf:
                ls=leaves(e); assert len(ls)<=3; covered.extend(ls); total[0]+=area[e]*len(ls)*cTri
            else: make_node(e)
    if root[0]=='L': return
    make_node(root)
    assert sorted(covered)==list(range(n)),"coverage"
    opt=area[root]*1.0+ (min(Fget(nodes[root[1]][0],k)+Fget(nodes[root[1]][1],8-k) for k in range(1,8)))
    assert abs(total[0]-opt)<1e-9*max(1,opt),(total[0],opt)
    return True


def test_dp_collapse_decisions_realise_the_optimum():
    random.seed(1)
    for n in (2, 3, 4, 5, 7, 9, 17, 40, 133, 1000):
        for c in (0.3, 0.6, 1.0, 2.0):
            run(n, c)
