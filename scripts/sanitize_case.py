"""One small frame over every production kernel (fused build, traversal hierarchy, radix sort, trace + tail + accumulate), for compute-sanitizer."""
import os
import sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import scene_util as SU                                        # noqa: E402
from raytracergpu_mastersproject_b200 import Device, Raytracer  # noqa: E402
from oracle import oracle as O                                  # noqa: E402  (checker only)

W, H, spp = 96, 64, 3
n_tris = int(sys.argv[1]) if len(sys.argv) > 1 else 3000        # <= 1800: the single-block sort; default: the tiled sort
sc = SU.random_scene(77, n_tris=n_tris, n_spheres=200, sort_morton=True)
ubo = SU.make_ubo(sc, random_state=9)
dev = Device(0)
rt = Raytracer(dev, W, H)
rt.update_scene(sc["models"], sc["triangles"], sc["spheres"], sc["materials"])
rt.build_bvh(ubo)
rt.clear_image(); rt.raytrace(ubo, spp); dev.wait_idle()
img = rt.read_image()
ref = O.build_bvh(sc["models"], sc["triangles"], sc["spheres"])
rr = O.raytrace(ubo, W, H, ref["tris"], ref["sphs"], sc["materials"], ref["nodes"], spp)
assert np.array_equal(img.view(np.uint32), rr["image"].view(np.uint32)), "frame differs from the oracle"
print(f"sanitize case ok: {len(sc['triangles']) + len(sc['spheres'])} primitives, {W}x{H}, {spp} spp, {dev.launch_count()} launches, bit-exact vs oracle")
