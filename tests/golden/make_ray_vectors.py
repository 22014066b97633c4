"""Extracts the ray known-answer vectors of the reference's test/tstRay.cpp (intersects_box :17-175,
intersects_triangle :408-470) into tests/golden/ray_vectors.json.  Run once in the build container
(where /root/reference exists); the tests only read the committed JSON."""
import json
import os
import re

SRC = "/root/reference/test/tstRay.cpp"
txt = open(SRC).read()
lines = txt.splitlines()
num = r"[-+]?[0-9]*\.?[0-9]+(?:[eE][-+]?[0-9]+)?f?"
pat = re.compile(r"BOOST_TEST\((!?)intersects\(Ray\{\{(" + num + r"),\s*(" + num + r"),\s*(" + num + r")\},\s*\{(" + num +
                 r"),\s*(" + num + r"),\s*(" + num + r")\}\},\s*(\w+)\)\);")
shapes = {"unit_box": ("box", [0, 0, 0, 1, 1, 1]),
          "unit_triangle": ("triangle", [0, 0, 0, 1, 0, 0, 0, 1, 0]),
          "tilted_triangle": ("triangle", [0, 0, 0, 2, 0, 1, 0, 2, 1])}
out = []
for ln, line in enumerate(lines, 1):
    m = pat.search(line)
    if not m:
        continue
    neg = m.group(1) == "!"
    vals = [float(x.rstrip("f")) for x in m.groups()[1:7]]
    shape = m.group(8)
    if shape not in shapes:
        continue
    kind, geom = shapes[shape]
    out.append({"line": ln, "ray": vals, "kind": kind, "geom": geom, "hit": not neg})
json.dump(out, open(os.path.join(os.path.dirname(__file__), "ray_vectors.json"), "w"), indent=0)
print(len(out), "vectors;", sum(1 for o in out if o["kind"] == "box"), "box,", sum(1 for o in out if o["kind"] == "triangle"), "triangle")
