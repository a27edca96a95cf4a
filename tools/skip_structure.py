"""dev: what a SkipPer warp of the LevelRyser kernel could skip.  Reads a plan dumped with SP_LEVEL_DUMP (rows by
level, columns in plan order) and replays the zero structure of the cold rows for sampled warps: blocks with a zero
in every lane (what the kernel skips today), how far such a zero is guaranteed to last (a row of level L keeps its
value until a column >= L flips), and the loop trips left if the warp jumped over those spans.
usage: skip_structure.py plan.txt [B] [c] [warps]"""
import sys, numpy as np
path = sys.argv[1]; B = int(sys.argv[2]) if len(sys.argv) > 2 else 3; c = int(sys.argv[3]) if len(sys.argv) > 3 else 11
W = int(sys.argv[4]) if len(sys.argv) > 4 else 48
L = open(path).read().split("\n")
n = int(L[0].split()[0])
rows = [list(map(float, l.split())) for l in L[1:n + 1]]
lvl = np.array([int(r[0]) for r in rows]); xb = np.array([r[1] for r in rows]); D = np.array([r[2:] for r in rows])   # D[j][k]
cold = np.where(lvl >= B)[0]
tc = cold[lvl[cold] >= c]                      # tile-constant rows: the tile filter
inner = cold[lvl[cold] < c]
nblk = 1 << (c - B)
rng = np.random.default_rng(1)
ntiles = 1 << (n - 1 - c)
def gray_bits(i):
    g = i ^ (i >> 1)
    return np.array([(g >> k) & 1 for k in range(n - 1)], dtype=float)
# survivors of the tile filter
# survivors of the tile filter: a warp takes 32 consecutive survivors, as the kernel's queue hands them out
surv = []
tested = 0
for w in range(W):
    t = int(rng.integers(0, ntiles - 4096)); got = 0
    while got < 32:
        tested += 1
        g = gray_bits(t << c)
        x = xb[tc] + D[tc][:, :n - 1] @ g
        if np.all(x != 0): surv.append(t); got += 1
        t += 1
print("n=%d B=%d c=%d  cold rows %d (tile-constant %d)  tile filter keeps %.1f%%" % (n, B, c, len(cold), len(tc), 100 * len(surv) / tested))
tot_blocks = hot_now = trips_jump = hot_jump = lane_nonzero = 0
maxlane = 0
pers_hist = np.zeros(c - B + 1)
for w in range(W):
    tiles = surv[32 * w:32 * w + 32]
    Z = np.zeros((32, nblk), bool); P = np.zeros((32, nblk), int)
    for l, t in enumerate(tiles):
        for blk in range(nblk):
            g = gray_bits((t << c) | (blk << B))
            x = xb[inner] + D[inner][:, :n - 1] @ g
            z = x == 0
            Z[l, blk] = z.any()
            if z.any():
                zl = int(lvl[inner][z].max()) - B        # the zero lasts until a column >= B + zl flips
                nxt = ((blk >> zl) + 1) << zl if zl > 0 else blk + 1
                P[l, blk] = min(nxt, nblk)
    allz = Z.all(axis=0)
    tot_blocks += nblk; hot_now += int((~allz).sum()); lane_nonzero += int((~Z).sum())
    maxlane += int((~Z).sum(axis=1).max())
    blk = 0
    while blk < nblk:
        trips_jump += 1
        if allz[blk]:
            j = int(P[:, blk].min()); pers_hist[int(np.log2(max(j - blk, 1)))] += 1
            blk = max(j, blk + 1)
        else:
            hot_jump += 1; blk += 1
print("blocks per tile round %d: evaluated today %.1f%% (lanes with a non-zero product: %.1f%% of lane-blocks)" %
      (nblk, 100 * hot_now / tot_blocks, 100 * lane_nonzero / (32 * tot_blocks)))
print("with span jumps: loop trips %.1f%% of the blocks (hot %.1f%%, jumps %.1f%%);  jump length histogram (log2): %s" %
      (100 * trips_jump / tot_blocks, 100 * hot_jump / tot_blocks, 100 * (trips_jump - hot_jump) / tot_blocks, pers_hist.astype(int)))
print("lane-decoupled bound: the busiest lane of a warp evaluates %.1f%% of the blocks" % (100 * maxlane / tot_blocks))
