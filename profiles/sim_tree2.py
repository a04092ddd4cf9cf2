"""CPU estimate: how many 4-ary steps per ray would a longest-axis median-split tree need, against the reference Morton tree with the big
leaves hoisted?  (exact boxes, immediate leaf tests; not product code)"""
import sys, time
import numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
sys.setrecursionlimit(100000)
from oracle import oracle as O
from raytracergpu_mastersproject_b200 import scenes
exec(open('/tmp/sim_tree.py').read().split("rng = np.random.default_rng(0)")[0].split("spec = sys.argv[1]")[0])   # imports only
spec = sys.argv[1]; nr = int(sys.argv[2])
sc = scenes.load_scene(spec)
b = O.build_bvh(sc["models"], sc["triangles"], sc["spheres"])
nodes = b["nodes"]; T = len(sc["triangles"]); S = len(sc["spheres"]); N = T + S
box = nodes["aabb"].astype(np.float64)
lo = box[:, 0::2].copy(); hi = box[:, 1::2].copy()
left = nodes["leftIndex"].astype(np.int64); right = nodes["rightIndex"].astype(np.int64)
leafOffset = N - 1
tris = b["tris"]
v0 = tris["v0"][:, :3].astype(np.float64); v1 = tris["v1"][:, :3].astype(np.float64); v2 = tris["v2"][:, :3].astype(np.float64)
def area(l, h):
    d = np.maximum(h - l, 0); return d[..., 0] * d[..., 1] + d[..., 1] * d[..., 2] + d[..., 2] * d[..., 0]
leaf_area = area(lo[leafOffset:], hi[leafOffset:]); root_area = area(lo[0], hi[0])
big = np.nonzero(leaf_area > root_area / 256.0)[0]
small = np.nonzero(leaf_area <= root_area / 256.0)[0]
llo = lo[leafOffset:]; lhi = hi[leafOffset:]
cen = 0.5 * (llo + lhi)

def tri_hit(g, o, d, tmin, tmax):
    a, bb, c = v0[g], v1[g], v2[g]
    u = bb - a; v = c - a; n = np.cross(u, v); nn = np.dot(n, n)
    if nn == 0: return None
    nh = n / np.sqrt(nn); den = np.dot(nh, d)
    if abs(den) < 1e-4: return None
    t = (np.dot(nh, a) - np.dot(nh, o)) / den
    if t < tmin or t > tmax: return None
    P = o + t * d; pp = P - a; w = n / nn
    aa = np.dot(w, np.cross(pp, v)); b2 = np.dot(w, np.cross(u, pp))
    if aa < 0 or b2 < 0 or aa + b2 > 1: return None
    return t, nh
def slab(l, h, o, rinv):
    t1 = (l - o) * rinv; t2 = (h - o) * rinv
    return np.max(np.minimum(t1, t2)), np.min(np.maximum(t1, t2))

class Tree:   # binary tree as arrays; leaves are ids >= nint (prim = id - nint)
    pass
def build_median(prims):
    n = len(prims)
    L = np.zeros(n - 1, np.int64); R = np.zeros(n - 1, np.int64)
    blo = np.zeros((2 * n - 1, 3)); bhi = np.zeros((2 * n - 1, 3))
    blo[n - 1:] = llo[prims]; bhi[n - 1:] = lhi[prims]
    nxt = [0]
    order = np.arange(n)
    def rec(idx):          # idx: positions into prims ; returns node id
        if len(idx) == 1: return n - 1 + idx[0]
        me = nxt[0]; nxt[0] += 1
        c = cen[prims[idx]]
        ext = c.max(0) - c.min(0); ax = int(np.argmax(ext))
        k = len(idx) // 2
        part = np.argpartition(c[:, ax], k)
        a = rec(idx[part[:k]]); bb = rec(idx[part[k:]])
        L[me] = a; R[me] = bb
        blo[me] = np.minimum(blo[a], blo[bb]); bhi[me] = np.maximum(bhi[a], bhi[bb])
        return me
    rec(order)
    return L, R, blo, bhi, n

def records_for(L, R, blo, bhi, nint):
    cache = {}
    def rec(i):
        r = cache.get(i)
        if r is None:
            ent = [R[i], L[i]]
            while len(ent) < 4:
                best, pick = -1.0, -1
                for k, e in enumerate(ent):
                    if e >= nint: continue
                    a = area(blo[e], bhi[e])
                    if a > best: best, pick = a, k
                if pick < 0: break
                e = ent[pick]; ent[pick:pick + 1] = [R[e], L[e]]
            cache[i] = r = ent
        return r
    return rec

def walk(o, d, rec, blo, bhi, nint, prims, bigs):
    rinv = 1.0 / d; closest, hitn = 1e7, None; steps = lb = 0
    def test(g):
        nonlocal closest, hitn, lb
        lb += 1
        tn, tf = slab(llo[g], lhi[g], o, rinv)
        if not (tn < tf) or g >= T: return
        r = tri_hit(g, o, d, 0.001, closest)
        if r: closest, hitn = r
    for g in bigs: test(g)
    stack = [0]
    while stack:
        cur = stack.pop(); steps += 1; cand = []
        for e in rec(cur):
            tn, tf = slab(blo[e], bhi[e], o, rinv)
            if not (tn < tf) or tn > closest or tf < 0.001: continue
            if e >= nint: test(prims[e - nint])
            else: cand.append((tn, e))
        cand.sort(reverse=True); stack.extend(e for _, e in cand)
    return steps, lb, closest, hitn

t0 = time.time()
L, R, blo, bhi, n = build_median(small)
print("median tree built", f"{time.time()-t0:.0f}s", n)
recM = records_for(L, R, blo, bhi, n - 1)
# reference tree with hoisting (tight boxes)
lo2 = lo.copy(); hi2 = hi.copy(); lo2[big + leafOffset] = np.inf; hi2[big + leafOffset] = -np.inf
parent = np.full(2 * N - 1, -1, np.int64); parent[left[:leafOffset]] = np.arange(leafOffset); parent[right[:leafOffset]] = np.arange(leafOffset)
dirty = set()
for e in big + leafOffset:
    p = parent[e]
    while p >= 0: dirty.add(int(p)); p = parent[p]
done = {}
def fix(i):
    if i >= leafOffset or i not in dirty: return lo2[i], hi2[i]
    if i in done: return done[i]
    l1, h1 = fix(left[i]); l2, h2 = fix(right[i])
    lo2[i] = np.minimum(l1, l2); hi2[i] = np.maximum(h1, h2); done[i] = (lo2[i], hi2[i]); return done[i]
fix(0)
def recRef_factory():
    cache = {}
    bigset = set(int(x) for x in big + leafOffset)
    def valid(e): return (e not in bigset) and (e >= leafOffset or lo2[e][0] <= hi2[e][0])
    def rec(i):
        r = cache.get(i)
        if r is None:
            ent = [e for e in (right[i], left[i]) if valid(e)]
            for _ in range(64):
                best, pick = -1.0, -1
                for k, e in enumerate(ent):
                    if e >= leafOffset: continue
                    a = area(lo2[e], hi2[e])
                    if a > best: best, pick = a, k
                if pick < 0: break
                e = ent[pick]; ch = [c for c in (right[e], left[e]) if valid(c)]
                if len(ent) - 1 + len(ch) > 4: break
                ent[pick:pick + 1] = ch
            cache[i] = r = ent
        return r
    return rec
recR = recRef_factory()
allprims = np.arange(N)
rng = np.random.default_rng(0)
cam = np.array([275., 275., -800.]); rays = []
while len(rays) < nr:
    px = rng.uniform(-0.35, 0.35, 2); d = np.array([px[0], px[1], 1.0]); d /= np.linalg.norm(d); o = cam
    for depth in range(4):
        st, lb_, t, nrm = walk(o, d, recR, lo2, hi2, leafOffset, allprims, big)
        if nrm is None: break
        P = o + t * d
        if np.dot(nrm, d) > 0: nrm = -nrm
        r = rng.normal(size=3); r /= np.linalg.norm(r); nd = nrm + r * rng.uniform(); nd /= np.linalg.norm(nd)
        o, d = P, nd; rays.append((o.copy(), d.copy()))
A = np.array([walk(o, d, recR, lo2, hi2, leafOffset, allprims, big)[:2] for o, d in rays])
B = np.array([walk(o, d, recM, blo, bhi, n - 1, small, big)[:2] for o, d in rays])
print(spec, "rays", len(rays))
print("reference Morton tree, big leaves hoisted: steps/ray %.2f leafboxes %.2f" % tuple(A.mean(0)))
print("longest-axis median tree over the small leaves: steps/ray %.2f leafboxes %.2f" % tuple(B.mean(0)))
