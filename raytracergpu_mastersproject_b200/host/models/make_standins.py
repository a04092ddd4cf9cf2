"""Writes the stand-in OBJ assets.  The reference's scenes load models/quad.obj, cube.obj and monkey.obj
(Scenes.cpp:105,117-122,210-214,294-296,342-345) but every .obj is git-ignored there (/root/reference/.gitignore:77),
so the files are inferred from their usage:
  quad.obj   unit square in the XZ plane, corners (+-1, 0, +-1), 2 triangles
  cube.obj   [-1,1]^3, 12 triangles
  monkey.obj NOT Blender's Suzanne: a 968-triangle closed bumpy head-sized blob of roughly unit radius
             (Suzanne triangulates to 968 triangles), so config C1 matches the reference scene in structure only.
"""
import math
import os

HERE = os.path.dirname(os.path.abspath(__file__))


def write(name, verts, faces, comment):
    with open(os.path.join(HERE, name), "w") as f:
        f.write(f"# {comment}\n")
        for v in verts:
            f.write("v %.6f %.6f %.6f\n" % v)
        for a, b, c in faces:
            f.write(f"f {a} {b} {c}\n")


write("quad.obj", [(-1, 0, -1), (1, 0, -1), (1, 0, 1), (-1, 0, 1)], [(1, 2, 3), (1, 3, 4)], "stand-in quad: unit square in the XZ plane")

cv = [(x, y, z) for x in (-1, 1) for y in (-1, 1) for z in (-1, 1)]
cf = [(1, 2, 4), (1, 4, 3), (5, 8, 6), (5, 7, 8), (1, 5, 6), (1, 6, 2), (3, 4, 8), (3, 8, 7), (1, 3, 7), (1, 7, 5), (2, 6, 8), (2, 8, 4)]
write("cube.obj", cv, cf, "stand-in cube [-1,1]^3")

lon, lat = 22, 23          # 2 * lon * (lat - 1) = 968 triangles
verts, faces = [], []
for r in range(lat):
    phi = math.pi * (0.03 + 0.94 * r / (lat - 1))
    for s in range(lon):
        th = 2 * math.pi * s / lon
        bump = 1.0 + 0.18 * math.sin(3 * th) * math.sin(2 * phi) + 0.10 * math.cos(5 * th + 1.0) * math.sin(phi) ** 2
        ear = 0.35 * math.exp(-((th - 1.2) ** 2 + (phi - 1.0) ** 2) * 6) + 0.35 * math.exp(-((th - 1.94) ** 2 + (phi - 1.0) ** 2) * 6)
        rad = 0.8 * (bump + ear)
        verts.append((rad * math.sin(phi) * math.cos(th), rad * math.cos(phi), rad * math.sin(phi) * math.sin(th)))
for r in range(lat - 1):
    for s in range(lon):
        a = r * lon + s + 1; b = r * lon + (s + 1) % lon + 1
        c = (r + 1) * lon + (s + 1) % lon + 1; d = (r + 1) * lon + s + 1
        faces += [(a, b, c), (a, c, d)]
assert len(faces) == 968
write("monkey.obj", verts, faces, "stand-in for monkey.obj (NOT Suzanne): 968-triangle bumpy blob")
