"""Discrete-event model of a SHARDED sequence call on two streams (why urnn_v2.cu keeps sharded runs on one stream).

Every rank runs the same program: stream E issues launches E0, E1, ..., stream D issues D0, D1, ...  A launch is `grid`
persistent CTAs of one SM each (a CTA of gemm_v2 takes the whole shared memory of an SM).  When the body of a launch is
done, its last CTA keeps its SM and spins until the same launch has reached that point on every peer (the in-kernel
statistic exchange of one lane); then the launch completes.  With programmatic dependent launch (PDL) the next launch
of a stream becomes resident as soon as the current one has all its CTAs running: its CTAs take free SMs and block in
griddepcontrol.wait until the current launch completes.  Without PDL a launch gets SMs only after its predecessor in the
stream completed.  CTAs are placed on free SMs in the order in which they became eligible.

run(...) returns "ok" when every rank finishes, "deadlock" when nothing can move.  What the model shows (tests/
test_pipeline_model.py): one stream never deadlocks; two streams without PDL never deadlock (a spinning CTA holds one SM,
everything else of the other stream still runs, in waves if need be); two streams WITH PDL deadlock for some timings --
rank A: E_k spins for B, E_k+1 is resident on every other SM, D_j cannot start; rank B: the mirror image.  That is the
hang seen at 4 GPUs; at 2 GPUs it takes a rarer timing, which is why the 2-GPU runs passed."""
import random


class Launch:
    def __init__(self, stream, idx, grid, work):
        self.stream, self.idx, self.grid = stream, idx, grid
        self.work = work                 # time units per CTA body
        self.placed = 0                  # CTAs that own an SM
        self.running = []                # remaining body time of placed CTAs that have been released to run
        self.blocked = 0                 # resident CTAs waiting for the predecessor (PDL)
        self.done_ctas = 0
        self.at_exchange = False         # body complete, last CTA spinning
        self.complete = False


def run(world=2, sms=6, grid=6, nlaunch=6, streams=2, pdl=True, seed=0, max_steps=100000):
    rnd = random.Random(seed)
    ranks = []
    for r in range(world):
        prog = {s: [Launch(s, i, grid, 1 + rnd.randint(0, 3)) for i in range(nlaunch)] for s in range(streams)}
        ranks.append({"prog": prog, "free": sms, "head": {s: 0 for s in range(streams)}, "speed": 1 + rnd.randint(0, 2)})

    def exchange_ready(s, i):
        return all(rk["prog"][s][i].at_exchange or rk["prog"][s][i].complete for rk in ranks)

    for step in range(max_steps):
        moved = False
        for r, rk in enumerate(ranks):
            if step % rk["speed"]:
                continue                                          # ranks drift against each other
            order = list(range(streams))
            rnd.shuffle(order)
            for s in order:
                i = rk["head"][s]
                if i >= nlaunch:
                    continue
                cur = rk["prog"][s][i]
                # place CTAs of the current launch
                while cur.placed < cur.grid and rk["free"] > 0 and not cur.at_exchange:
                    rk["free"] -= 1; cur.placed += 1; cur.running.append(cur.work); moved = True
                # PDL: the successor becomes resident once the current launch has all its CTAs placed
                if pdl and i + 1 < nlaunch and cur.placed == cur.grid:
                    nxt = rk["prog"][s][i + 1]
                    while nxt.placed < nxt.grid and rk["free"] > 0:
                        rk["free"] -= 1; nxt.placed += 1; nxt.blocked += 1; moved = True
                # advance bodies
                still = []
                for t in cur.running:
                    if t > 1:
                        still.append(t - 1)
                    else:
                        cur.done_ctas += 1
                        if cur.done_ctas < cur.grid:
                            rk["free"] += 1                       # all but the last CTA give their SM back
                    moved = moved or True
                cur.running = still
                if not cur.at_exchange and cur.done_ctas == cur.grid:
                    cur.at_exchange = True; moved = True          # last CTA keeps its SM and spins
                if cur.at_exchange and not cur.complete and exchange_ready(s, i):
                    cur.complete = True; rk["free"] += 1; rk["head"][s] += 1; moved = True
                    if i + 1 < nlaunch:                           # release the resident successor
                        nxt = rk["prog"][s][i + 1]
                        nxt.running.extend([nxt.work] * nxt.blocked); nxt.blocked = 0
        if all(rk["head"][s] >= nlaunch for rk in ranks for s in range(streams)):
            return "ok"
        if not moved and all(step % rk["speed"] == 0 for rk in ranks):
            return "deadlock"
    return "deadlock"


if __name__ == "__main__":
    for streams, pdl in ((1, True), (2, False), (2, True)):
        for world in (2, 4):
            res = [run(world=world, streams=streams, pdl=pdl, seed=s) for s in range(200)]
            print(f"streams={streams} pdl={pdl} world={world}: {res.count('deadlock')} deadlocks in {len(res)} timings")
