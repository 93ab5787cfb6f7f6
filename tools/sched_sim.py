"""Discrete-event model of one tile's value-mode steps inside rollout_tc_kernel: an in-order tensor pipe fed by one issuing
thread that walks the job list, two accumulator buffers, two epilogue groups (one per buffer) that drain an accumulator
(early release) and then do the activation math.  Used to compare job orders before spending GPU time on them.
  python tools/sched_sim.py            (durations calibrated on profiles/r2_tile_timeline.txt)"""
import sys

D = dict(L1=450, L2=2750, L3=2700, drain=550, math=2600, e3=2600, e3_early=False, cp=0)

def simulate(jobs, D=D, verbose=False):
    """jobs: list of dicts: kind L1/L2/L3, name, needs: list of data tokens, makes: token produced by its epilogue, sync: bool"""
    pipe_free = 0.0; issue_t = 0.0
    acc_empty = [0.0, 0.0]; acc_full = {}
    ready = {"za0": 0.0}
    grp_free = [0.0, 0.0]
    nbuf = 0
    # epilogue groups walk in order too; a group can only start its job k after its previous jobs
    out = []
    # two passes are entangled (MMA needs epilogue outputs and vice versa) -> process jobs in list order, resolving the
    # epilogue of job j right after its MMA (valid because every dependency points backwards in the list)
    for j, job in enumerate(jobs):
        b = nbuf; nbuf ^= 1
        t = max(issue_t, acc_empty[b])
        for tok in job.get("needs", []): t = max(t, ready[tok])
        dur = D[job["kind"]] + (D["cp"] if job.get("cp") else 0)
        # fine-grained operand arrival: chunk i of 4 needed before quarter i of the MMAs
        ex0 = max(t, pipe_free)
        ex1 = ex0 + dur
        for i, tok in enumerate(job.get("needs4", [])):
            ex1 = max(ex1, ready[tok] + dur * (4 - i) / 4)
            ex0 = max(ex0, 0)
        if job.get("needs4"): t = max(t, ready[job["needs4"][0]])
        ex1 = max(ex1, max(t, pipe_free) + dur)
        issue_t = max(t, ex1 - dur) + 60          # the issuing thread moves on once the MMAs are queued (queue depth ignored)
        issue_t = max(issue_t, ex1 - 1500)        # ... but it cannot run further ahead than the pipe's queue
        pipe_free = ex1
        g = b
        s = max(grp_free[g], ex1)
        if job["kind"] == "L3":
            e = s + D["drain"] + D["e3"]
            acc_empty[b] = s + D["drain"] if D["e3_early"] else e
        else:
            acc_empty[b] = s + D["drain"]
            e = acc_empty[b] + D["math"]
        grp_free[g] = e
        if job.get("makes"): ready[job["makes"]] = e
        if job.get("sync"):
            m = max(grp_free); grp_free = [m + 100, m + 100]
        out.append((job["name"], t, ex1 - dur, ex1, g, s, e))
        if verbose: print(f"{j:3d} {job['name']:10s} issue {t:8.0f} exec {ex1 - dur:8.0f}-{ex1:8.0f} | epi g{g} {s:8.0f}-{e:8.0f}")
    return out

def plain(H):
    J = []
    for t in range(H):
        for c in range(4): J.append(dict(kind="L1", name=f"L1d{c}t{t}", needs=[f"za{t}"], makes=f"h1d{c}t{t}"))
        for c in range(4): J.append(dict(kind="L2", name=f"L2d{c}t{t}", needs4=[f"h1d{i}t{t}" for i in range(4)] if c == 0 else [], makes=f"xs{c}t{t}"))
        for c in range(4): J.append(dict(kind="L1", name=f"L1r{c}t{t}", makes=f"h1r{c}t{t}"))
        J.append(dict(kind="L3", name=f"L3dt{t}", needs4=[f"xs{i}t{t}" for i in range(4)], makes=f"za{t+1}"))
        for c in range(4): J.append(dict(kind="L2", name=f"L2r{c}t{t}", needs4=[f"h1r{i}t{t}" for i in range(4)] if c == 0 else [], sync=(c == 3)))
    return J

def staged(H, order):
    J = []
    for c in range(4): J.append(dict(kind="L1", name=f"L1d{c}t0", needs=["za0"], makes=f"h1d{c}t0"))
    for t in range(H):
        for c in range(4): J.append(dict(kind="L2", name=f"L2d{c}t{t}", needs4=[f"h1d{i}t{t}" for i in range(4)] if c == 0 else [], makes=f"xs{c}t{t}", cp=(c == 0 and t > 0)))
        for c in range(4): J.append(dict(kind="L1", name=f"L1r{c}t{t}", makes=f"h1r{c}t{t}"))
        J.append(dict(kind="L3", name=f"L3dt{t}", needs4=[f"xs{i}t{t}" for i in range(4)], makes=f"za{t+1}"))
        R = lambda c: dict(kind="L2", name=f"L2r{c}t{t}", needs4=[f"h1r{i}t{t}" for i in range(4)] if c == 0 else [], sync=(c == 3))
        N = lambda c: dict(kind="L1", name=f"L1d{c}t{t+1}", needs=[f"za{t+1}"], makes=f"h1d{c}t{t+1}")
        for ch in order: J.append(R(int(ch[1])) if ch[0] == "R" else N(int(ch[1])))
    return J

def period(J, key="L2d0t"):
    out = simulate(J)
    ts = [o[2] for o in out if o[0].startswith(key)]
    return [round(b - a) for a, b in zip(ts, ts[1:])]

if __name__ == "__main__":
    H = 6
    print("plain           ", period(plain(H)))
    for order in ["R0 R1 N0 N1 N2 N3 R2 R3", "R0 R1 N0 N1 R2 N2 N3 R3", "R0 N0 R1 N1 R2 N2 R3 N3", "R0 R1 N0 R2 N1 R3 N2 N3", "R0 R1 N0 R2 N1 N2 R3 N3",
                  "R0 R1 N0 R2 N1 N2 N3 R3", "R0 R1 R2 N0 N1 N2 N3 R3", "R0 R1 R2 R3 N0 N1 N2 N3"]:
        for early in (False, True):
            D["e3_early"] = early
            print(f"{order:26s} e3_early={int(early)}", period(staged(H, order.split())))
    if len(sys.argv) > 1:
        D["e3_early"] = True
        simulate(staged(3, sys.argv[1].split()), verbose=True)
