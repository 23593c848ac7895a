"""Level-schedule statistics of the C0 trade plan: rounds per null for the ASAP schedule k_trade uses, and for a
capacity-aware list schedule (slack-based) that fills rounds of 32 trades.  CPU only (oracle pair stream)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import oracle

R, T, CAP = 384, 1920, int(sys.argv[1]) if len(sys.argv) > 1 else 32
tot_asap, tot_list, tot_depth, tot_depth2 = 0, 0, 0, 0
N = 20
for null in range(N):
    pairs = [oracle.trade_pair(1764, null, t, R) for t in range(T)]
    last = np.zeros(R, dtype=int)
    asap = np.zeros(T, dtype=int)
    for t, (a, b) in enumerate(pairs):
        l = max(last[a], last[b]) + 1
        last[a] = last[b] = l
        asap[t] = l
    D = asap.max()
    cnt = np.bincount(asap, minlength=D + 1)[1:]
    rounds_asap = int(np.ceil(cnt / CAP).sum())
    # ALAP (latest level that keeps depth D): backward pass
    nxt = np.full(R, D + 1, dtype=int)
    alap = np.zeros(T, dtype=int)
    for t in range(T - 1, -1, -1):
        a, b = pairs[t]
        l = min(nxt[a], nxt[b]) - 1
        nxt[a] = nxt[b] = l
        alap[t] = l
    # list scheduling in program order per row: a trade is ready when both rows' previous trades are placed in earlier levels
    prev_a = np.full(T, -1); prev_b = np.full(T, -1)
    lastt = np.full(R, -1)
    for t, (a, b) in enumerate(pairs):
        prev_a[t], prev_b[t] = lastt[a], lastt[b]
        lastt[a] = lastt[b] = t
    placed = np.zeros(T, dtype=int)          # level of each trade, 0 = not yet
    remaining = set(range(T))
    level, rounds_list = 0, 0
    while remaining:
        level += 1
        ready = [t for t in remaining if (prev_a[t] < 0 or 0 < placed[prev_a[t]] < level) and (prev_b[t] < 0 or 0 < placed[prev_b[t]] < level)]
        ready.sort(key=lambda t: alap[t])                       # least slack first
        urgent = [t for t in ready if alap[t] <= level]         # must go now to keep the depth
        k = max(len(urgent), (len(ready) // CAP) * CAP)         # fill whole rounds; never delay critical trades
        if k == 0:
            k = len(ready)
        take = ready[:k]
        for t in take:
            placed[t] = level
            remaining.discard(t)
        rounds_list += int(np.ceil(len(take) / CAP))
    tot_asap += rounds_asap; tot_list += rounds_list; tot_depth += D; tot_depth2 += level
print("capacity %d trades per CTA-round: ASAP schedule %.1f rounds / null (depth %.1f); list schedule %.1f rounds / null (depth %.1f); ideal %.1f" % (
    CAP, tot_asap / N, tot_depth / N, tot_list / N, tot_depth2 / N, T / CAP))


def variant(pairs, CAP, rule):
    """Event-driven list schedule without the backward pass.  rule: 'order' = defer the youngest ready trades,
    'chain' = prefer trades with successors on both rows."""
    T = len(pairs)
    prev = np.full((T, 2), -1)
    succ = np.full((T, 2), -1)
    lastt = np.full(R, -1)
    for t, (a, b) in enumerate(pairs):
        for side, row in enumerate((a, b)):
            p = lastt[row]
            prev[t, side] = p
            if p >= 0:
                succ[p, 0 if pairs[p][0] == row else 1] = t
            lastt[row] = t
    npred = (prev >= 0).sum(1)
    ready = [t for t in range(T) if npred[t] == 0]
    level, rounds, placed_n = 0, 0, 0
    while placed_n < T:
        level += 1
        n = len(ready)
        k = n if n < CAP else (n // CAP) * CAP
        if rule == "order":
            ready.sort()
        else:
            ready.sort(key=lambda t: (-(succ[t] >= 0).sum(), t))
        take, rest = ready[:k], ready[k:]
        nxt = list(rest)
        for t in take:
            for s in succ[t]:
                if s >= 0:
                    npred[s] -= 1
                    if npred[s] == 0:
                        nxt.append(s)
        placed_n += len(take)
        rounds += int(np.ceil(len(take) / CAP))
        ready = nxt
    return rounds, level


for rule in ("order", "chain"):
    tr, tl = 0, 0
    for null in range(N):
        pairs = [oracle.trade_pair(1764, null, t, R) for t in range(T)]
        r_, l_ = variant(pairs, CAP, rule)
        tr += r_; tl += l_
    print("no backward pass, rule %-6s: %.1f rounds / null, depth %.1f" % (rule, tr / N, tl / N))


def first_fit(pairs, CAP):
    """Online, in program order: earliest level after both predecessors that still has room for one more trade."""
    last = np.zeros(R, dtype=int)
    cnt = np.zeros(4096, dtype=int)
    D = 0
    for a, b in pairs:
        l = max(last[a], last[b]) + 1
        while cnt[l] >= CAP:
            l += 1
        cnt[l] += 1
        last[a] = last[b] = l
        D = max(D, l)
    return D, cnt[1:D + 1]


for cap in (CAP, 2 * CAP):
    tl, fill = 0, []
    for null in range(N):
        pairs = [oracle.trade_pair(1764, null, t, R) for t in range(T)]
        D_, c_ = first_fit(pairs, cap)
        tl += int(np.ceil(c_ / CAP).sum()); fill.append(D_)
    print("online first fit, at most %d trades per level: %.1f rounds / null, %.1f levels" % (cap, tl / N, np.mean(fill)))
