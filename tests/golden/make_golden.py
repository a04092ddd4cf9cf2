"""Generates tests/golden/*.npz from the CPU oracle (oracle/rt_oracle.c).

These fixtures pin the ORACLE's own outputs on larger scenes: they guard the oracle (and through it the CUDA path) against
drift, they do not add reference authority -- that comes from make_spirv_golden.py (outputs of the reference's compiled shaders).
Inputs are the seeded scenes of tests/scene_util.py plus the host's complexScene; regenerate with
    python tests/golden/make_golden.py
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))

import scene_util as SU  # noqa: E402
from oracle import oracle as O  # noqa: E402

CASES = {
    "room_seed1": dict(scene=dict(seed=1, n_tris=200, n_spheres=20), W=64, H=48, spp=3, random_state=12345),
    "sorted_seed2": dict(scene=dict(seed=2, n_tris=1500, n_spheres=100, sort_morton=True), W=64, H=48, spp=2, random_state=777),
    "complexScene": dict(scene="complexScene", W=80, H=80, spp=2, random_state=12345),
}


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def load_case(c):
    if isinstance(c["scene"], str):
        from raytracergpu_mastersproject_b200 import scenes
        sc = scenes.load_scene(c["scene"])
        ubo = SU.make_ubo(sc, max_depth=sc["max_depth"], random_state=c["random_state"], vfov=sc["vfov"])
    else:
        sc = SU.random_scene(**c["scene"])
        ubo = SU.make_ubo(sc, random_state=c["random_state"])
    return sc, ubo


def run_case(c):
    sc, ubo = load_case(c)
    b = O.build_bvh(sc["models"], sc["triangles"], sc["spheres"])
    r = O.raytrace(ubo, c["W"], c["H"], b["tris"], b["sphs"], sc["materials"], b["nodes"], c["spp"])
    return dict(nodes_sha=sha(b["nodes"]), morton_sha=sha(b["morton"]), enclosing=b["enclosing"].view(np.uint32).reshape(-1),
                image=r["image"], hit_prim=r["hit_prim"], rng=r["rng"],
                counters=np.array([r["counters"][k] for k in ("rays", "nodeVisits", "triTests", "sphTests", "matReads", "samples")], np.uint64))


if __name__ == "__main__":
    for name, c in CASES.items():
        out = run_case(c)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), nodes_sha=np.array(out["nodes_sha"]), morton_sha=np.array(out["morton_sha"]),
                            enclosing=out["enclosing"], image=out["image"], hit_prim=out["hit_prim"], rng=out["rng"], counters=out["counters"])
        print(name, out["nodes_sha"][:16], out["counters"].tolist())
