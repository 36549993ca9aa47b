"""Randomised model of the bulk-mode pipeline of tc::gemm_gn_kernel (2 loader warps, 2 converter groups, in-order MMA
issuer, double-buffered accumulator): mbarrier parity semantics, asserts on slot ownership.  Odd staging depths fail."""
import random, itertools
class Bar:
    def __init__(s, count): s.count=count; s.pending=count; s.phase=0
    def arrive(s):
        s.pending-=1
        assert s.pending>=0
        if s.pending==0: s.phase+=1; s.pending=s.count
    def test(s, parity): return (s.phase & 1) != parity
def run(nraw, na, nu, ntile, seed, NG=2, NL=2):
    rnd=random.Random(seed)
    rfull=[Bar(1) for _ in range(nraw)]; rempty=[Bar(1) for _ in range(nraw)]   # one arrival per agent (warp-level abstraction)
    full=[Bar(1) for _ in range(na)]; empty=[Bar(1) for _ in range(na)]
    tfull=[Bar(1),Bar(1)]; tempty=[Bar(1),Bar(1)]
    U=nu*ntile
    raw_owner=[None]*nraw; a_owner=[None]*na  # data checks
    log=[]
    def loader(l):
        u=l
        while u<U:
            rs=u%nraw; ph=(u//nraw)&1
            while not rempty[rs].test(ph^1): yield
            assert raw_owner[rs] is None, ("raw overwrite", u, raw_owner[rs])
            raw_owner[rs]=u
            yield
            rfull[rs].arrive()
            u+=NL
    def conv(g):
        u=g
        while u<U:
            rs=u%nraw; rph=(u//nraw)&1; a=u%na; aph=(u//na)&1
            while not rfull[rs].test(rph): yield
            assert raw_owner[rs]==u, ("raw mismatch", u, raw_owner[rs])
            raw_owner[rs]=None
            rempty[rs].arrive()
            while not empty[a].test(aph^1): yield
            assert a_owner[a] is None, ("A overwrite", u, a_owner[a])
            a_owner[a]=u
            yield
            full[a].arrive()
            u+=NG
    def mma():
        st=0; ph=0; acc=0; aph=0; u=0
        for t in range(ntile):
            while not tempty[acc].test(aph^1): yield
            for kb in range(nu):
                while not full[st].test(ph): yield
                assert a_owner[st]==u, ("A mismatch", u, a_owner[st])
                a_owner[st]=None
                yield
                empty[st].arrive()
                u+=1; st+=1
                if st==na: st=0; ph^=1
            tfull[acc].arrive()
            acc^=1
            if acc==0: aph^=1
    def epi():
        acc=0; aph=0
        for t in range(ntile):
            while not tfull[acc].test(aph): yield
            yield
            tempty[acc].arrive()
            acc^=1
            if acc==0: aph^=1
    agents=[loader(l) for l in range(NL)]+[conv(g) for g in range(NG)]+[mma(),epi()]
    alive=list(range(len(agents)))
    steps=0
    while alive:
        i=rnd.choice(alive)
        try: next(agents[i])
        except StopIteration: alive.remove(i)
        steps+=1
        if steps>200000: return "DEADLOCK/livelock"
    return "ok"
if __name__ == '__main__':
    bad=0
    for nraw,na,nu,ntile in itertools.product((2,4,6),range(2,5),range(1,11),range(1,5)):
        for seed in range(6):
            r=run(nraw,na,nu,ntile,seed)
            if r!="ok": print(nraw,na,nu,ntile,seed,r); bad+=1
    print("bad",bad)
