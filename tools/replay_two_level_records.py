"""What-if replay (numpy, CPU): how many gather requests / distinct 128 B lines / 32 B sectors per warp would a
two-levels-per-gather node layout need, against the current depth-ordered 8-byte nodes?  Walks the bench
forest over consecutive C24 rows in warps of 32, with levels 0..3 served from constant memory as in the
kernel.  Result quoted in DESIGN.md section 7.  Usage: python tools/replay_two_level_records.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from quickchem_b200 import synth, capi
b = capi.Booster(os.path.join(ROOT, 'build', 'oh_booster_100x18.model'), parse_only=True)
nodes, off, depth, orig = b.flat()
x = synth.quick_features(synth.raw_fields(24))
x = x[100000:100000+32*512]
n = x.shape[0]
xs = np.concatenate([x, np.full((n,1), -np.inf, np.float32)], axis=1)
thr = nodes[:,0].view(np.float32); meta = nodes[:,1]
feat = (meta >> 26).astype(np.int64); rel = (meta & ((1<<23)-1)).astype(np.int64)
ar = np.arange(n)
CTOP=4
def distinct(a):
    s = np.sort(a, axis=1)
    return (np.diff(s, axis=1) != 0).sum(1) + 1 - (s[:,0] == -1)
tot = dict(cur_lines=0, cur_req=0, tri_lines=0, tri_req=0, cur_bytes=0, tri_bytes=0, tri_sec=0, cur_sec=0)
for t in range(0,100,5):
    n0, n1 = int(off[t]), int(off[t+1]); m = n1-n0
    r = rel[n0:n1]; dep = np.zeros(m, np.int32)
    for i in np.nonzero(r)[0]:
        dep[i+r[i]] = dep[i]+1; dep[i+r[i]+1] = dep[i]+1
    # triplet slot of each even-depth>=CTOP node
    slot = np.full(m, -1, np.int64); nslots = 0
    lvl = [np.nonzero(dep==d)[0] for d in range(int(dep.max())+1)]
    # roots at depth CTOP: children pairs of depth CTOP-1 internal nodes, in BFS order -> pairs contiguous
    for i in lvl[CTOP-1] if CTOP-1 < len(lvl) else []:
        if r[i]:
            slot[i+r[i]] = nslots; slot[i+r[i]+1] = nslots+1; nslots += 2
    d = CTOP
    while d < len(lvl):
        for i in lvl[d]:
            if r[i] == 0: continue
            base = nslots; nslots += 4   # block LL LR RL RR
            for c in (0,1):
                ch = i + r[i] + c
                if r[ch]:
                    slot[ch + r[ch]] = base + 2*c; slot[ch + r[ch] + 1] = base + 2*c + 1
        d += 2
    tot['cur_bytes'] += m*8; tot['tri_bytes'] += nslots*16
    # walk
    idx = np.full(n, 0, np.int64); done = np.zeros(n, bool); d = 0
    gx = lambda i: xs[ar, feat[n0+i]]
    while not done.all():
        act = ~done
        if d >= CTOP:
            line = np.where(act, (n0+idx)*8//128, -1).reshape(-1,32)
            tot['cur_lines'] += distinct(line).sum(); tot['cur_req'] += (line!=-1).any(1).sum()
            tot['cur_sec'] += distinct(np.where(act, (n0+idx)*8//32, -1).reshape(-1,32)).sum()
            if (d - CTOP) % 2 == 0:
                s = slot[idx]; assert (s[act] >= 0).all()
                tl = np.where(act, s*16//128, -1).reshape(-1,32)
                tot['tri_lines'] += distinct(tl).sum(); tot['tri_req'] += (tl!=-1).any(1).sum()
                tot['tri_sec'] += distinct(np.where(act, s*16//32, -1).reshape(-1,32)).sum()
        v = gx(idx); right = ~(v < thr[n0+idx]); rr = r[idx]
        idx = np.where(act, idx + rr + right, idx); done |= act & (rr == 0); d += 1
rows = n*20
for k,v in tot.items(): print(k, v, round(v/rows,3))
print('lines ratio tri/cur', tot['tri_lines']/tot['cur_lines'], 'req ratio', tot['tri_req']/tot['cur_req'])
