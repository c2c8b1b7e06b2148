import ctypes as C, numpy as np, sys, time, itertools
sys.path.insert(0,'.')
from qibochem_b200 import synthetic
lib=C.CDLL('qibochem_b200/libqibochem_b200.so')
u64p,f64p=C.POINTER(C.c_uint64),C.POINTER(C.c_double)
def stats(n,hx,hz,hc):
    st=np.zeros(16,dtype=np.uint64)
    hx,hz,hc=(np.ascontiguousarray(a) for a in (hx,hz,hc))
    lib.qc_k2_plan_stats(n,C.c_size_t(hx.size),hx.ctypes.data_as(u64p),hz.ctypes.data_as(u64p),hc.ctypes.data_as(f64p),st.ctypes.data_as(u64p),None)
    return st
def compress(m,G,n):
    out=np.zeros_like(m); p=0
    for b in range(n):
        if b in G: continue
        out |= ((m>>np.uint64(b))&np.uint64(1))<<np.uint64(p); p+=1
    return out
def cost(st): # ms at 30 local qubits
    return 2.69*st[2]+0.446*st[3]+0.139*st[4]
n=int(sys.argv[1]); g=int(sys.argv[2])
hx,hz,hc,_=synthetic.molecular_like_hamiltonian(n,100000)
def run(choose, G0):
    rem=np.ones(hx.size,bool); G=list(G0); tot=np.zeros(3); layouts=0; c=0
    while rem.any():
        gm=np.uint64(sum(1<<b for b in G))
        ready=rem & ((hx&gm)==0)
        if not ready.any():
            G=choose(hx[rem],G); continue
        st=stats(n-g,compress(hx[ready],G,n),compress(hz[ready],G,n),hc[ready])
        tot+=np.array([st[2],st[3],st[4]],float); c+=cost(st); layouts+=1
        rem&=~ready
        if rem.any(): G=choose(hx[rem],G)
    return layouts,tot,c
def by_usage(xs,G):
    usage=sorted((int(((xs>>np.uint64(b))&np.uint64(1)).sum()),b) for b in range(n))
    return [b for _,b in usage[:g]]
print('usage', run(by_usage, list(range(n-g,n))))
def by_plan(xs_rem_unused,G):
    return None
# planner-guided: candidates from the 2g+2 lowest-usage bits
def make_guided(k):
    def choose(xs,G):
        usage=sorted((int(((xs>>np.uint64(b))&np.uint64(1)).sum()),b) for b in range(n))
        cand=[b for _,b in usage[:k]]
        best=None
        remmask=None
        for Gc in itertools.combinations(cand,g):
            gm=np.uint64(sum(1<<b for b in Gc))
            sel=(cur_rem_x & gm)==0
            if not sel.any(): continue
            st=stats(n-g,compress(cur_rem_x[sel],Gc,n),compress(cur_rem_z[sel],Gc,n),cur_rem_c[sel])
            score=cost(st)/max(1,len(np.unique(cur_rem_x[sel])))
            if best is None or score<best[0]: best=(score,Gc,sel.sum())
        return list(best[1])
    return choose

def run_guided(k, G0, objective):
    rem=np.ones(hx.size,bool); G=list(G0); tot=np.zeros(3); layouts=0; c=0
    first=True
    while rem.any():
        # choose next G (also for the first layout: the rotations leave an arbitrary one)
        usage=sorted((int(((hx[rem]>>np.uint64(b))&np.uint64(1)).sum()),b) for b in range(n))
        cand=[b for _,b in usage[:k]]
        best=None
        for Gc in itertools.combinations(cand,g):
            gm=np.uint64(sum(1<<b for b in Gc))
            sel=rem & ((hx & gm)==0)
            if not sel.any(): continue
            st=stats(n-g,compress(hx[sel],Gc,n),compress(hz[sel],Gc,n),hc[sel])
            nx=len(np.unique(hx[sel]))
            swaps=len(set(Gc)-set(G))
            score=objective(cost(st),nx,swaps,st)
            if best is None or score<best[0]: best=(score,Gc,st)
        _,Gc,st=best
        gm=np.uint64(sum(1<<b for b in Gc)); sel=rem & ((hx & gm)==0)
        tot+=np.array([st[2],st[3],st[4]],float); c+=cost(st)+56*4*len(set(Gc)-set(G)); layouts+=1
        rem&=~sel; G=list(Gc)
    return layouts,tot,c
t=time.time()
print('guided cost/xmask k=2g', run_guided(2*g, list(range(n-g,n)), lambda c,nx,sw,st:(c+224*sw)/nx), time.time()-t)
print('guided max ready k=2g', run_guided(2*g, list(range(n-g,n)), lambda c,nx,sw,st:-nx))
