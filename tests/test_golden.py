"""Golden fixtures (tests/golden/*.npz, produced by tests/golden/make_golden.py from the oracle): CPU test pins the
oracle against them, GPU test pins the CUDA path against them."""
import hashlib
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden as G  # noqa: E402


def _load(name):
    return np.load(os.path.join(HERE, "golden", name + ".npz"))


@pytest.mark.parametrize("name", sorted(G.CASES))
def test_oracle_reproduces_golden(name):
    g = _load(name)
    out = G.run_case(G.CASES[name])
    assert out["nodes_sha"] == str(g["nodes_sha"]) and out["morton_sha"] == str(g["morton_sha"])
    assert np.array_equal(out["enclosing"], g["enclosing"])
    assert np.array_equal(out["image"].view(np.uint32), g["image"].view(np.uint32))
    assert np.array_equal(out["hit_prim"], g["hit_prim"]) and np.array_equal(out["rng"], g["rng"])
    assert np.array_equal(out["counters"], g["counters"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(G.CASES))
def test_cuda_reproduces_golden(device, name):
    from oracle import oracle as O
    from raytracergpu_mastersproject_b200 import Buffer, Raytracer, capi
    g = _load(name)
    c = G.CASES[name]
    sc, ubo = G.load_case(c)
    W, H = c["W"], c["H"]
    n = len(sc["triangles"]) + len(sc["spheres"])
    rt = Raytracer(device, W, H)
    rt.update_scene(sc["models"], sc["triangles"], sc["spheres"], sc["materials"])
    rt.build_bvh(ubo)
    assert hashlib.sha256(rt.nodes.read(O.NODE, 2 * n - 1).tobytes()).hexdigest() == str(g["nodes_sha"])
    assert hashlib.sha256(rt.morton1.read(O.MORTON, n).tobytes()).hexdigest() == str(g["morton_sha"])
    hp = Buffer(device, 4, W * H); rg = Buffer(device, 4, W * H)
    rt.clear_image(); rt.counters.zero()
    rt.raytrace(ubo, c["spp"], flags=capi.TRACE_COUNT, hit_prim=hp, rng_out=rg)
    device.wait_idle()
    assert np.array_equal(rt.read_image().view(np.uint32), g["image"].view(np.uint32))
    assert np.array_equal(hp.read(np.uint32).reshape(H, W), g["hit_prim"])
    assert np.array_equal(rg.read(np.uint32).reshape(H, W), g["rng"])
    cnt = rt.read_counters()
    assert [cnt[k] for k in capi.COUNTER_FIELDS] == g["counters"].tolist()
