#!/usr/bin/env python3
"""Discrete simulation of the mbarrier protocol of attn_bwd_umma_kernel (deadlock / phase-aliasing check on CPU).
Each actor is a generator yielding ('wait', bar, parity) | ('arrive', bar) | ('commit', bar) | ('tma', bar)."""
import itertools, random, sys

class Bar:
    def __init__(self, count): self.count=count; self.pending=count; self.phase=0
    def arrive(self):
        self.pending-=1
        if self.pending==0: self.pending=self.count; self.phase+=1
    def done(self, parity):  # try_wait.parity semantics: true iff the phase with this parity has completed most recently
        return (self.phase & 1) != parity if False else ((self.phase-1) & 1)==parity and self.phase>0 or False

def try_wait(bar, parity):
    # mbarrier phase bit starts at 0; wait(parity) succeeds when current phase parity != parity  (i.e. phase `parity` completed)
    return (bar.phase & 1) != parity

def run(T, NST, NSLOT, seed):
    rnd=random.Random(seed)
    slow=rnd.choice(['prod','mma','wgA','wgB',None])
    B={'kv':Bar(1),'fin':Bar(1),'dq_full':Bar(1),'dq_empty':Bar(2)}
    for i in range(NST): B[f'qf{i}']=Bar(1); B[f'qe{i}']=Bar(1)
    for i in range(3): B[f'sf{i}']=Bar(1)
    for i in range(4): B[f'ud{i}']=Bar(1)
    for i in range(2): B[f'pe{i}']=Bar(1)
    U=2*T
    log=[]
    def producer():
        yield ('tma','kv')
        for i in range(T):
            st=i%NST
            if i>=NST: yield ('wait',f'qe{st}',((i//NST)-1)&1)
            yield ('tma',f'qf{st}')
    def mma():
        def issue_s(u):
            i,hf=u>>1,u&1; st=i%NST; slot=u%NSLOT
            if hf==0: yield ('wait',f'qf{st}',(i//NST)&1)
            yield ('commit',f'sf{slot}')
        yield ('wait','kv',0)
        for u in range(min(NSLOT,U)):
            yield from issue_s(u)
        for u in range(U):
            i,hf=u>>1,u&1; st=i%NST; pb=i&1
            yield ('wait',f'ud{hf*2+pb}',(i>>1)&1)
            if u+NSLOT<U: yield from issue_s(u+NSLOT)
            if hf==1:
                if i>=1: yield ('wait','dq_empty',(i-1)&1)
                yield ('commit',f'qe{st}'); yield ('commit',f'pe{pb}'); yield ('commit','dq_full')
        yield ('commit','fin')
    def wg(w):
        def drain(i):
            yield ('wait','dq_full',i&1)
            yield ('arrive','dq_empty')
        for i in range(T):
            u=2*i+w; slot=u%NSLOT
            yield ('wait',f'sf{slot}',(u//NSLOT)&1)
            if i>=2: yield ('wait',f'pe{i&1}',((i>>1)-1)&1)
            yield ('arrive',f'ud{w*2+(i&1)}')
            if i>=1: yield from drain(i-1)
        yield from drain(T-1)
        yield ('wait','fin',0)
    actors={'prod':producer(),'mma':mma(),'wgA':wg(0),'wgB':wg(1)}
    cur={k:None for k in actors}
    alive=set(actors)
    async_q=[]  # pending async completions (commits / tma), delivered in order with random delay
    steps=0
    while alive:
        steps+=1
        progressed=False
        order=list(alive); rnd.shuffle(order)
        # deliver async completions randomly (in order)
        if async_q and rnd.random()<0.5:
            B[async_q.pop(0)].arrive(); progressed=True
        for a in order:
            if a==slow and rnd.random()<0.9: continue   # adversarial: one actor is much slower than the others
            if cur[a] is None:
                try: cur[a]=next(actors[a])
                except StopIteration: alive.discard(a); progressed=True; continue
            op=cur[a]
            if op[0]=='wait':
                if try_wait(B[op[1]],op[2]): cur[a]=None; progressed=True
            elif op[0]=='arrive':
                B[op[1]].arrive(); cur[a]=None; progressed=True
            else:  # commit / tma: asynchronous completion
                async_q.append(op[1]); cur[a]=None; progressed=True
        if not progressed:
            if async_q: B[async_q.pop(0)].arrive(); continue
            if slow is not None: slow=None; continue
            return False,{a:cur[a] for a in alive},{k:(b.phase,b.pending) for k,b in B.items()}
    return True,None,None

bad=0
for T in range(1,12):
    for NST in (2,4):
        for NSLOT in (2,3):
            for seed in range(60):
                ok,blocked,state=run(T,NST,NSLOT,seed)
                if not ok:
                    bad+=1
                    if bad<4: print('DEADLOCK T',T,'NST',NST,'NSLOT',NSLOT,'seed',seed,blocked,state)
print('bad',bad)
