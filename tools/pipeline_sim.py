"""Randomised-schedule replay of the plane-gather producer / MMA pipelines with mbarrier PARITY-wait semantics (host-only,
no GPU): shows that "every warp observes every slot" deadlocks (a slow observer falls one revolution behind a barrier)
while split ownership with stride <= ring size never does.  python tools/pipeline_sim.py"""
import random
def passed(cnt, parity): return (cnt & 1) != parity

def sim_wgrad(Wu, nunits, stages, rts, seed, scheme):
    rnd=random.Random(seed)
    full=[0]*stages; empty=[0]*stages; gfull=[0,0]; gempty=[0,0]
    total_units=rts*nunits
    # agents: list of generators yielding when blocked
    def g_filler():
        for rti in range(rts):
            gb=rti&1
            while not passed(gempty[gb], ((rti>>1)&1)^1): yield
            gfull[gb]+=1
            yield
    def unit_owner(w):
        o=w
        while o<total_units:
            st=o%stages; ph=(o//stages)&1
            while not passed(empty[st], ph^1): yield
            full[st]+=1
            o+=Wu
            yield
    def observe_all(w, W=8):
        spr=1+nunits; s=0; stage=0; phase=0; rti=0; j=0
        while s<rts*spr:
            own=(s%W)==w
            if j==0:
                gb=rti&1
                while not passed(gempty[gb], ((rti>>1)&1)^1): yield
                if own: gfull[gb]+=1
            else:
                while not passed(empty[stage], phase^1): yield
                if own: full[stage]+=1
                stage+=1
                if stage==stages: stage=0; phase^=1
            j+=1
            if j==spr: j=0; rti+=1
            s+=1
            yield
    def mma():
        st=0; ph=0
        for git in range(rts):
            gb=git&1
            while not passed(gfull[gb], (git>>1)&1): yield
            for u in range(nunits):
                while not passed(full[st], ph): yield
                empty[st]+=1
                st+=1
                if st==stages: st=0; ph^=1
                yield
            gempty[gb]+=1
        return
    if scheme=='split':
        agents=[g_filler()]+[unit_owner(w) for w in range(Wu)]+[mma()]
    else:
        agents=[observe_all(w) for w in range(8)]+[mma()]
    alive=list(range(len(agents)))
    idle=0
    while alive:
        i=rnd.choice(alive)
        # run agent i for a random number of steps
        before=(tuple(full),tuple(empty),tuple(gfull),tuple(gempty))
        for _ in range(rnd.randint(1,6)):
            try: next(agents[i])
            except StopIteration:
                alive.remove(i); break
        after=(tuple(full),tuple(empty),tuple(gfull),tuple(gempty))
        idle = 0 if after!=before else idle+1
        if idle>20000: return 'DEADLOCK'
    return 'ok'
for scheme in ('observe','split'):
    for nunits,stages in ((7,6),(4,5),(3,5),(2,3),(4,8),(2,8),(1,8)):
        res=[sim_wgrad(min(7,stages),nunits,stages,12,seed,scheme) for seed in range(60)]
        print(scheme,nunits,stages,res.count('ok'),'/',len(res))
