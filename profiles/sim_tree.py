"""CPU simulation of the 4-ary nearest-first t-culled walk on C2 (exact boxes, no quantisation) to estimate what hoisting the big
primitives out of the hierarchy would save.  Not product code."""
import sys, time
import numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from oracle import oracle as O
from raytracergpu_mastersproject_b200 import scenes, make_ubo

spec = sys.argv[1] if len(sys.argv) > 1 else "meshRoom:660:1"
sc = scenes.load_scene(spec)
b = O.build_bvh(sc["models"], sc["triangles"], sc["spheres"])
nodes = b["nodes"]; T = len(sc["triangles"]); S = len(sc["spheres"]); N = T + S
box = nodes["aabb"].astype(np.float64)          # minX maxX minY maxY minZ maxZ
lo = box[:, 0::2].copy(); hi = box[:, 1::2].copy()
left = nodes["leftIndex"].astype(np.int64); right = nodes["rightIndex"].astype(np.int64)
leafOffset = N - 1
tris = b["tris"]
v0 = tris["v0"][:, :3].astype(np.float64); v1 = tris["v1"][:, :3].astype(np.float64); v2 = tris["v2"][:, :3].astype(np.float64)
print("N", N, "root box", lo[0], hi[0])

def area(l, h):
    d = np.maximum(h - l, 0); return d[..., 0] * d[..., 1] + d[..., 1] * d[..., 2] + d[..., 2] * d[..., 0]

def build_records(lo, hi):
    """greedy cut per internal node, computed lazily (dict)"""
    cache = {}
    def rec(i):
        r = cache.get(i)
        if r is None:
            ent = [right[i], left[i]]
            while len(ent) < 4:
                best, pick = -1.0, -1
                for k, e in enumerate(ent):
                    if e >= leafOffset: continue
                    a = area(lo[e], hi[e])
                    if a > best: best, pick = a, k
                if pick < 0: break
                e = ent[pick]
                ent[pick:pick + 1] = [right[e], left[e]]
            cache[i] = r = ent
        return r
    return rec

def tri_hit(g, o, d, tmin, tmax):
    a, bb, c = v0[g], v1[g], v2[g]
    u = bb - a; v = c - a; n = np.cross(u, v)
    nn = np.dot(n, n)
    if nn == 0: return None
    nh = n / np.sqrt(nn)
    den = np.dot(nh, d)
    if abs(den) < 1e-4: return None
    t = (np.dot(nh, a) - np.dot(nh, o)) / den
    if t < tmin or t > tmax: return None
    P = o + t * d; pp = P - a; w = n / nn
    aa = np.dot(w, np.cross(pp, v)); b2 = np.dot(w, np.cross(u, pp))
    if aa < 0 or b2 < 0 or aa + b2 > 1: return None
    return t, nh

def slab(l, h, o, rinv):
    t1 = (l - o) * rinv; t2 = (h - o) * rinv
    tn = np.max(np.minimum(t1, t2)); tf = np.min(np.maximum(t1, t2))
    return tn, tf

def walk(o, d, rec, lo, hi, skip=None, extra=None):
    """returns (steps, leafboxes, tests, t, normal)"""
    rinv = 1.0 / d
    closest, hitn = 1e7, None
    steps = lb = tests = 0
    def test_leaf(e):
        nonlocal closest, hitn, lb, tests
        g = e - leafOffset
        lb += 1
        tn, tf = slab(lo[e], hi[e], o, rinv)
        if not (tn < tf): return
        if g >= T: return
        tests += 1
        r = tri_hit(g, o, d, 0.001, closest)
        if r: closest, hitn = r
    if extra is not None:
        for e in extra: test_leaf(e)
    stack = [0]
    while stack:
        cur = stack.pop()
        steps += 1
        cand = []
        for e in rec(cur):
            if skip is not None and e in skip: continue
            tn, tf = slab(lo[e], hi[e], o, rinv)
            if not (tn < tf) or tn > closest or tf < 0.001: continue
            if e >= leafOffset: test_leaf(e)
            else: cand.append((tn, e))
        cand.sort(reverse=True)                     # nearest popped first
        stack.extend(e for _, e in cand)
    return steps, lb, tests, closest, hitn

rng = np.random.default_rng(0)
# rays: camera rays -> first hit -> diffuse bounce (cosine-ish) ; a few bounces
cam = np.array([275., 275., -800.]); 
rec = build_records(lo, hi)
rays = []
t0 = time.time()
while len(rays) < int(sys.argv[2] if len(sys.argv) > 2 else 600):
    px = rng.uniform(-0.35, 0.35, 2)
    d = np.array([px[0], px[1], 1.0]); d /= np.linalg.norm(d)
    o = cam
    for depth in range(4):
        st, lb_, ts, t, n = walk(o, d, rec, lo, hi)
        if n is None: break
        P = o + t * d
        if np.dot(n, d) > 0: n = -n
        r = rng.normal(size=3); r /= np.linalg.norm(r)
        nd = n + r * rng.uniform() ; nd /= np.linalg.norm(nd)
        o, d = P, nd
        rays.append((o.copy(), d.copy()))
print("rays", len(rays), f"{time.time()-t0:.1f}s")
# variant A: current
A = np.array([walk(o, d, rec, lo, hi)[:3] for o, d in rays])
print("current tree: steps/ray %.2f leafboxes %.2f tests %.2f" % tuple(A.mean(0)))
# variant B: big leaves hoisted
leaf_area = area(lo[leafOffset:], hi[leafOffset:])
root_area = area(lo[0], hi[0])
big = np.nonzero(leaf_area > root_area / 256.0)[0] + leafOffset
print("big leaves", len(big), big[:20] - leafOffset)
bigset = set(int(x) for x in big)
lo2 = lo.copy(); hi2 = hi.copy()
lo2[big] = np.inf; hi2[big] = -np.inf
# refit bottom-up: process internal nodes in an order where children come first: iterate until stable using parent pointers
parent = np.full(2 * N - 1, -1, np.int64)
parent[left[:leafOffset]] = np.arange(leafOffset); parent[right[:leafOffset]] = np.arange(leafOffset)
dirty = set()
for e in big:
    p = parent[e]
    while p >= 0:
        dirty.add(int(p)); p = parent[p]
print("inflated internal nodes", len(dirty))
# recompute dirty nodes bottom-up: sort by subtree size? use recursion with memo
import functools
sys.setrecursionlimit(10000)
done = {}
def fix(i):
    if i >= leafOffset or i not in dirty: return lo2[i], hi2[i]
    if i in done: return done[i]
    l1, h1 = fix(left[i]); l2, h2 = fix(right[i])
    lo2[i] = np.minimum(l1, l2); hi2[i] = np.maximum(h1, h2)
    done[i] = (lo2[i], hi2[i]); return done[i]
fix(0)
rec2 = build_records(lo2, hi2)
B = np.array([walk(o, d, rec2, lo2, hi2, skip=bigset, extra=None)[:3] for o, d in rays])
print("hoisted (tree part only): steps/ray %.2f leafboxes %.2f tests %.2f" % tuple(B.mean(0)))
# cost of the big list: box tests passing
def big_pass(o, d):
    rinv = 1.0 / d; c = 0
    for e in big:
        tn, tf = slab(lo[e], hi[e], o, rinv)
        c += bool(tn < tf)
    return c
print("big leaf boxes passed per ray %.2f of %d" % (np.mean([big_pass(o, d) for o, d in rays]), len(big)))
