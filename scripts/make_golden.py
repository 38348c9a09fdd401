#!/usr/bin/env python
"""Generate tests/golden/reference_goldens.json from the reference checkout.

Run in the build container only (it reads /root/reference, which does not exist on
the GPU box).  It extracts, without copying any reference source:

  * the reference's test inputs   test/test_{0,1,2}.cdl, grids/rect3030.res.cdl  (values only)
  * the 5 integration goldens     test/<case>/ref_partition_{mask,metadata}_3.cdl
      - parsed values (pid map, boxes, neighbour tables, dimension lengths)
      - sha256 of the exact ncdump text, used to check the CDL emitter byte for byte
  * the 9 bounding-box known-answer tests of test/test_zoltan_partitioner_{0,1,2}.cpp,
    transcribed below with their file:line.
"""
import hashlib
import json
import os
import re
import sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden",
                   "reference_goldens.json")


def read(path):
    with open(os.path.join(REF, path)) as f:
        return f.read()


def sha(text):
    return hashlib.sha256(text.encode()).hexdigest()


def parse_dims(text):
    dims = {}
    m = re.search(r"dimensions:\n(.*?)\n(variables:|\n|group:)", text, re.S)
    for line in m.group(1).splitlines():
        mm = re.match(r"\s*(\w+) = (\w+) ;(?: // \((\d+) currently\))?", line)
        if mm:
            dims[mm.group(1)] = int(mm.group(3)) if mm.group(2) == "UNLIMITED" else int(mm.group(2))
    return dims


def parse_data(text):
    """all `name = v, v, ... ;` statements after any `data:` marker"""
    out = {}
    for body in re.split(r"\bdata:\n", text)[1:]:
        for mm in re.finditer(r"^\s*(\w+) =\s*([-0-9,\s]+?);", body, re.S | re.M):
            out[mm.group(1)] = [int(v) for v in mm.group(2).replace("\n", " ").split(",")]
    return out


def parse_input(path, var):
    text = read(path)
    dims = parse_dims(text)
    mm = re.search(r"^\s*%s =\s*([-0-9,.\s]+?);" % var, text, re.S | re.M)
    vals = [int(float(v)) for v in mm.group(1).replace("\n", " ").split(",")]
    return text, dims, vals


def main():
    g = {"_generated_by": "scripts/make_golden.py from the reference checkout (read-only)",
         "inputs": {}, "box_kats": [], "integration": {}}

    for name, var, xd, yd in (("test_0", "mask", "x", "y"), ("test_1", "mask", "x", "y"),
                              ("test_2", "land_mask", "m", "n")):
        text, dims, vals = parse_input("test/%s.cdl" % name, var)
        title = re.search(r':title = "(.*)"', text).group(1)
        g["inputs"][name] = {"source": "test/%s.cdl" % name, "nx": dims[xd], "ny": dims[yd],
                             "xdim": xd, "ydim": yd, "order": "yx", "mask_name": var,
                             "title": title, "mask": vals, "cdl_sha256": sha(text)}
    # rect3030: `double mask(x, y)` inside `group: data`; needs `-o xy` (Grid.cpp:110-113).  The
    # raw values are kept in FILE order; Grid indexes them x-fastest (quirk Q7).
    text = read("grids/rect3030.res.cdl")
    mm = re.search(r"^\s*mask =\s*([-0-9,.\s]+?);", text, re.S | re.M)
    vals = [int(float(v)) for v in mm.group(1).replace("\n", " ").split(",")]
    assert len(vals) == 900
    g["inputs"]["rect3030"] = {"source": "grids/rect3030.res.cdl", "nx": 30, "ny": 30, "xdim": "x",
                               "ydim": "y", "order": "xy", "mask_name": "mask", "group": "data",
                               "mask": vals, "n_ocean": sum(1 for v in vals if v > 0)}

    # bounding-box known-answer tests: [x0, y0, ext_x, ext_y] per rank
    K = g["box_kats"]
    K.append({"input": "test_0", "P": 1, "boxes": [[0, 0, 6, 4]],
              "cite": "test/test_zoltan_partitioner_0.cpp:14-37"})
    K.append({"input": "test_0", "P": 2, "boxes": [[0, 0, 3, 4], [3, 0, 3, 4]],
              "cite": "test/test_zoltan_partitioner_0.cpp:39-66"})
    K.append({"input": "test_1", "P": 1, "boxes": [[0, 0, 6, 4]],
              "cite": "test/test_zoltan_partitioner_1.cpp:14-37"})
    K.append({"input": "test_1", "P": 2, "boxes": [[0, 0, 3, 4], [3, 0, 3, 4]],
              "cite": "test/test_zoltan_partitioner_1.cpp:39-66"})
    K.append({"input": "test_2", "P": 1, "boxes": [[0, 0, 6, 4]],
              "cite": "test/test_zoltan_partitioner_2.cpp:14-37"})
    K.append({"input": "test_2", "P": 2, "boxes": [[0, 0, 3, 4], [3, 0, 3, 4]],
              "cite": "test/test_zoltan_partitioner_2.cpp:39-66"})
    K.append({"input": "test_2", "P": 3, "boxes": [[0, 0, 2, 4], [2, 0, 2, 4], [4, 0, 2, 4]],
              "cite": "test/test_zoltan_partitioner_2.cpp:68-98"})
    K.append({"input": "test_2", "P": 4,
              "boxes": [[0, 0, 1, 4], [1, 0, 2, 4], [3, 0, 1, 4], [4, 0, 2, 4]],
              "cite": "test/test_zoltan_partitioner_2.cpp:100-139"})
    # Grid known answers (test/test_grid_{0,1,2}.cpp): extents / object counts on 1-2 ranks
    g["grid_kats"] = [
        {"input": "test_0", "P": 1, "rank": 0, "num_objects": 24, "num_nonzero_objects": 0,
         "cite": "test/test_grid_0.cpp:11-29"},
        {"input": "test_1", "P": 1, "rank": 0, "num_objects": 24, "num_nonzero_objects": 24,
         "cite": "test/test_grid_1.cpp:11-29"},
        {"input": "test_2", "P": 1, "rank": 0, "num_objects": 24, "num_nonzero_objects": 12,
         "cite": "test/test_grid_2.cpp:11-30"},
        {"input": "test_2", "P": 2, "rank": 0, "num_objects": 12, "num_nonzero_objects": 6,
         "cite": "test/test_grid_2.cpp:32-51"},
        {"input": "test_2", "P": 2, "rank": 1, "num_objects": 12, "num_nonzero_objects": 6,
         "cite": "test/test_grid_2.cpp:32-51"},
    ]

    cases = {"test_1": ("test_1", 0, 0), "test_2": ("test_2", 0, 0), "test_1_px": ("test_1", 1, 0),
             "test_1_py": ("test_1", 0, 1), "test_1_px_py": ("test_1", 1, 1)}
    for case, (inp, px, py) in cases.items():
        mtext = read("test/%s/ref_partition_mask_3.cdl" % case)
        dtext = read("test/%s/ref_partition_metadata_3.cdl" % case)
        data = parse_data(dtext)
        g["integration"][case] = {
            "input": inp, "P": 3, "px": px, "py": py,
            "cite": "test/integration-test.sh:6-35, test/%s/" % case,
            "pid": parse_data(mtext)["pid"],
            "dims": parse_dims(dtext),
            "metadata": data,
            "mask_cdl_sha256": sha(mtext),
            "metadata_cdl_sha256": sha(dtext),
        }
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    with open(OUT, "w") as f:
        json.dump(g, f, indent=1, sort_keys=True)
        f.write("\n")
    print("wrote", os.path.normpath(OUT))


if __name__ == "__main__":
    main()
