"""CPU: host-side logic of the drop-in API (C++ host mirror through librtb200_host.so): TransformComponent::mat4,
the GameObject -> device-array flatten order, the OBJ loader, the reference's named scenes and the synthetic generators."""
import os

import numpy as np
import pytest

from oracle import oracle as O
from raytracergpu_mastersproject_b200 import capi, scenes


def _mat4_restated(t, s, r):
    """numpy restatement of the reference's TransformComponent::mat4 (TransformComponent.cpp:9-45), binary32"""
    f = np.float32
    c3, s3 = f(np.cos(f(r[2]))), f(np.sin(f(r[2])))
    c2, s2 = f(np.cos(f(r[0]))), f(np.sin(f(r[0])))
    c1, s1 = f(np.cos(f(r[1]))), f(np.sin(f(r[1])))
    sx, sy, sz = f(s[0]), f(s[1]), f(s[2])
    return np.array([
        sx * (c1 * c3 + s1 * s2 * s3), sx * (c2 * s3), sx * (c1 * s2 * s3 - c3 * s1), 0,
        sy * (c3 * s1 * s2 - c1 * s3), sy * (c2 * c3), sy * (c1 * c3 * s2 + s1 * s3), 0,
        sz * (c2 * s1), sz * (-s2), sz * (c1 * c2), 0,
        t[0], t[1], t[2], 1], np.float32)


@pytest.mark.parametrize("seed", range(5))
def test_transform_mat4_matches_restatement(seed):
    rng = np.random.default_rng(seed)
    t, s, r = rng.uniform(-500, 500, 3), rng.uniform(0.1, 300, 3), rng.uniform(-3.2, 3.2, 3)
    got = scenes.transform_mat4(t, s, r)
    ref = _mat4_restated(np.float32(t), np.float32(s), np.float32(r))
    # libm float sin/cos vs numpy's: identical glibc here; allow 1 ulp of the products in case numpy vectorises differently
    assert np.allclose(got, ref, rtol=3e-7, atol=1e-6)
    assert got[3] == got[7] == got[11] == 0 and got[15] == 1


def test_complex_scene_flatten_order_and_parameters():
    s = scenes.load_scene("complexScene")                   # Scenes.cpp:337-398
    assert (s["rays_per_pixel"], s["max_depth"], s["vfov"]) == (8, 8, 40.0)
    # objects in insertion order: light quad, floor, back wall, monkey, sphere -> one Model + one Material each
    assert len(s["models"]) == 5 and len(s["materials"]) == 5
    assert len(s["triangles"]) == 2 + 2 + 2 + 968 and len(s["spheres"]) == 1
    t = s["triangles"]
    assert t["modelIndex"][:6].tolist() == [0, 0, 1, 1, 2, 2] and np.all(t["modelIndex"][6:] == 3)
    assert np.array_equal(t["materialIndex"], t["modelIndex"])
    assert s["spheres"]["modelIndex"][0] == 4 and s["spheres"]["radius"][0] == 40.0
    m = s["materials"]
    assert m["materialType"].tolist() == [0, 1, 1, 1, 1]
    assert np.allclose(m["albedo"][0, :3], 15.0) and np.allclose(m["albedo"][3, :3], [0.12, 0.15, 0.45])
    # light panel: quad scaled (65,1,50) at (275,549,300); floor quad spans [0,550]^2 at y = 0
    tw, _ = O.model_to_world(s["models"], s["triangles"], s["spheres"])
    light = np.concatenate([tw["v0"][:2, :3], tw["v1"][:2, :3], tw["v2"][:2, :3]])
    assert light[:, 0].min() == 210 and light[:, 0].max() == 340 and np.all(light[:, 1] == 549)
    floor = np.concatenate([tw["v0"][2:4, :3], tw["v1"][2:4, :3], tw["v2"][2:4, :3]])
    assert floor[:, 0].min() == 0 and floor[:, 0].max() == 550 and np.all(floor[:, 1] == 0)
    # padding bytes of the uploaded records are zero
    assert np.all(t["_pad"] == 0) and np.all(m["_pad"] == 0)


@pytest.mark.parametrize("name,counts", [
    ("simpleScene", (4, 6, 1, 4, 16, 8, 40.0)),
    ("cornellBoxScene", (9, 36, 1, 9, 128, 25, 40.0)),
    ("cornellMixedScene", (9, 34, 2, 9, 1, 100, 80.0)),
    ("randomSpheres", (10, 2, 9, 10, 1, 5, 90.0)),
])
def test_reference_scenes(name, counts):
    s = scenes.load_scene(name)
    got = (len(s["models"]), len(s["triangles"]), len(s["spheres"]), len(s["materials"]), s["rays_per_pixel"], s["max_depth"], s["vfov"])
    assert got == counts


def test_unknown_scene_raises():
    with pytest.raises(capi.RtbError):
        scenes.load_scene("noSuchScene")


def test_obj_loader(tmp_path):
    """host/VulkanWrapper/ObjLoader.cpp restates tinyobjloader's triangulation (the reference calls LoadObj with triangulate = true,
    RTModel.cpp:53): triangles as they are, quads along the shorter diagonal, larger polygons ear-clipped from corner 0.  The
    expected triangles below were derived by hand from those rules (the library is un-vendored: pinned by restatement)."""
    def tris_of(text, name="t.obj"):
        p = tmp_path / name
        p.write_text(text)
        return scenes.load_obj(str(p))[:, :, :2].tolist()
    # square quad: equal diagonals -> NOT "shorter 0-2" -> (0,1,3) (1,2,3); then a negative-index triangle with vt / vn
    t = tris_of("v 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0\nvt 0 0\nvn 0 0 1\n"
                "f 1/1/1 2/1/1 3/1/1 4/1/1\n"
                "f -4//1 -3//1 -1//1\n")
    assert t == [[[0, 0], [1, 0], [0, 1]], [[1, 0], [1, 1], [0, 1]], [[0, 0], [1, 0], [0, 1]]]
    # kite whose diagonal 0-2 (|.|^2 = 2) is shorter than 1-3 (18) -> (0,1,2) (0,2,3)
    t = tris_of("v 0 0 0\nv 3 0 0\nv 1 1 0\nv 0 3 0\nf 1 2 3 4\n")
    assert t == [[[0, 0], [3, 0], [1, 1]], [[0, 0], [1, 1], [0, 3]]]
    # ... and the mirror case: diagonal 1-3 shorter -> (0,1,3) (1,2,3)
    t = tris_of("v 0 0 0\nv 1 0 0\nv 3 3 0\nv 0 1 0\nf 1 2 3 4\n")
    assert t == [[[0, 0], [1, 0], [0, 1]], [[1, 0], [3, 3], [0, 1]]]
    # convex pentagon: every corner is an ear in turn -> a fan from corner 0
    P = [(0, 0), (2, 0), (3, 1.5), (1, 3), (-1, 1.5)]
    t = tris_of("".join(f"v {x} {y} 0\n" for x, y in P) + "f 1 2 3 4 5\n")
    f = lambda *ix: [[list(map(float, P[i])) for i in tri] for tri in ix]  # noqa: E731
    assert t == f((0, 1, 2), (0, 2, 3), (0, 3, 4))
    # concave pentagon, corner 1 reflex: corner 0's ear is rejected, clipping starts at corner 1 -> (1,2,3) (1,3,4) (0,1,4)
    P = [(0, 0), (1, 1), (2, 0), (2, 3), (0, 3)]
    t = tris_of("".join(f"v {x} {y} 0\n" for x, y in P) + "f 1 2 3 4 5\n")
    assert t == f((1, 2, 3), (1, 3, 4), (0, 1, 4))
    # the same polygon in the x = const plane (dominant-plane axes y, z) triangulates the same way
    t3 = scenes.load_obj(str(_write(tmp_path / "yz.obj", "".join(f"v 7 {x} {y}\n" for x, y in P) + "f 1 2 3 4 5\n")))[:, :, 1:].tolist()
    assert t3 == f((1, 2, 3), (1, 3, 4), (0, 1, 4))
    # faces with fewer than three corners are dropped (tinyobjloader warns), line continuation, vertex colours, o / g / s / usemtl ignored
    t = tris_of("o thing\ng part\ns off\nv 0 0 0 1 0 0\nv 1 0 0 0 1 0\nv 0 1 0 0 0 1\nusemtl none\nf 1 2\nf 1 2 \\\n 3\n")
    assert t == [[[0, 0], [1, 0], [0, 1]]]
    # an mtllib that exists is parsed, one that does not is only a warning: the reference never looks at OBJ materials
    (tmp_path / "m.mtl").write_text("newmtl red\nKd 1 0 0\nKe 0 0 0\nNs 10\nnewmtl glow\nKe 5 5 5\n")
    t = tris_of("mtllib m.mtl\nmtllib missing.mtl\nv 0 0 0\nv 1 0 0\nv 0 1 0\nusemtl glow\nf 1 2 3\n")
    assert t == [[[0, 0], [1, 0], [0, 1]]]
    with pytest.raises(capi.RtbError):
        scenes.load_obj(str(tmp_path / "missing.obj"))
    bad = tmp_path / "bad.obj"; bad.write_text("v 0 0 0\nf 1 2 3\n")
    with pytest.raises(capi.RtbError):
        scenes.load_obj(str(bad))


def _write(path, text):
    path.write_text(text)
    return path


def test_png_writer_round_trip(tmp_path):
    """host/utils/Png.hpp (the headless hosts' present-to-PNG): signature, chunk CRCs, zlib stream and pixels survive a decode,
    including an image whose raw data needs several stored deflate blocks (> 65535 bytes)."""
    rng = np.random.default_rng(5)
    for h, w in [(1, 1), (7, 5), (200, 150)]:
        img = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
        p = str(tmp_path / f"t{h}.png")
        scenes.write_png(p, img)
        assert np.array_equal(scenes.read_png_rgb(p), img[..., :3])
    try:
        from PIL import Image
        assert np.array_equal(np.asarray(Image.open(p).convert("RGB")), img[..., :3])     # an independent decoder agrees
    except ImportError:
        pass


def test_standin_assets():
    d = os.path.join(os.path.dirname(scenes.__file__), "host", "models")
    assert scenes.load_obj(os.path.join(d, "quad.obj")).shape[0] == 2
    assert scenes.load_obj(os.path.join(d, "cube.obj")).shape[0] == 12
    assert scenes.load_obj(os.path.join(d, "monkey.obj")).shape[0] == 968


@pytest.mark.parametrize("spec,first_sorted", [("meshRoom:40:1", 6), ("heightField:60:30:3:0", 2), ("heightField:48:48:4:80", 2)])
def test_generators_emit_reference_morton_order(spec, first_sorted):
    """D8: generated meshes are emitted in the reference's own Morton order (the LBVH keeps leaves in input order)."""
    s = scenes.load_scene(spec)
    tw, sw = O.model_to_world(s["models"], s["triangles"], s["spheres"])
    codes = O.morton_codes(tw, sw, O.enclosing_aabb(tw, sw))["code"][first_sorted:len(tw)]
    assert np.all(np.diff(codes.astype(np.int64)) >= 0)
    # no degenerate triangles (a zero-area triangle makes the reference's normalize() produce NaN hits)
    n = np.cross(tw["v1"][:, :3] - tw["v0"][:, :3], tw["v2"][:, :3] - tw["v0"][:, :3])
    assert np.all(np.linalg.norm(n, axis=1) > 0)


def test_sphere_field_generator():
    s = scenes.load_scene("sphereField:2000:2")
    assert len(s["spheres"]) == 2000 and len(s["triangles"]) == 4 and len(s["materials"]) == 2002
    types = s["materials"]["materialType"][2:]
    frac = [(types == k).mean() for k in (1, 2, 3)]
    assert abs(frac[0] - 0.70) < 0.05 and abs(frac[1] - 0.15) < 0.04 and abs(frac[2] - 0.15) < 0.04
    tw, sw = O.model_to_world(s["models"], s["triangles"], s["spheres"])
    codes = O.morton_codes(tw, sw, O.enclosing_aabb(tw, sw))["code"][4:]
    assert np.all(np.diff(codes.astype(np.int64)) >= 0)
    assert sw["center"][:, :3].min() >= 25 and sw["center"][:, :3].max() <= 525
    assert s["spheres"]["radius"].min() >= 1 and s["spheres"]["radius"].max() <= 4


def test_dielectric_heavy_height_field():
    s = scenes.load_scene("heightField:128:128:4:80")
    mt = s["materials"]["materialType"][s["triangles"]["materialIndex"]]
    assert 0.6 < (mt == 3).mean() < 0.95 and s["max_depth"] == 16 and s["rays_per_pixel"] == 1024


def test_generators_are_deterministic():
    a = scenes.load_scene("meshRoom:24:1"); b = scenes.load_scene("meshRoom:24:1")
    assert a["triangles"].tobytes() == b["triangles"].tobytes()
    c = scenes.load_scene("meshRoom:24:2")
    assert a["triangles"].tobytes() != c["triangles"].tobytes()
