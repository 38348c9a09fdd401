#!/usr/bin/env python
"""Reads the part maps off the reference's own pictures (TEST INFRASTRUCTURE; run where /root/reference exists).

    python scripts/make_golden_png.py      -> tests/golden/rect3030_png_parts.json

img/partition_2.png and img/partition_4.png of the reference show `grids/rect3030.res.cdl` decomposed into 2 and 4
parts by the reference (Zoltan) itself.  They are the only reference-held evidence of Zoltan's behaviour on an
irregular coastline and with more than one cut level per direction: for every cell of the 30 x 30 grid the colour at
the cell centre is mapped back through the picture's colour bar to a value in [-1, P - 1] (-1 = land).  The plots are
shaded (colours blend between cell centres next to land) and the last column is clipped by the plot frame, so a test
only trusts cells whose value is close to an integer (tests/test_oracle_golden.py).  The pictures are drawn with the
file's first dimension horizontally, i.e. TRANSPOSED with respect to the x-fastest mask the decomposer sees for
`-o xy`, and partition_4.png numbers the parts with 1 and 2 exchanged (stale w.r.t. REMAP=0).
"""
import json
import os

import numpy as np
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/img"
# plot frame in pixels: x = -0.5 at 117, 29.5 at 923; y = 29.5 at 72, -0.5 at 473; colour bar row 567, x 240..800
X0, X1, Y0, Y1, BAR_Y, BAR_X0, BAR_X1 = 117, 923, 72, 473, 567, 240, 800


def read(name, P):
    im = np.asarray(Image.open(os.path.join(REF, name)).convert("RGB")).astype(int)
    bar = im[BAR_Y, BAR_X0:BAR_X1 + 1]
    vals = np.linspace(-1, P - 1, bar.shape[0])
    out = np.zeros((30, 30))
    for j in range(30):
        for i in range(30):
            px = int(round(X0 + (i + 0.5) * (X1 - X0) / 30.0))
            py = int(round(Y0 + (29.5 - j) * (Y1 - Y0) / 30.0))
            c = im[py - 1:py + 2, px - 1:px + 2].reshape(-1, 3).mean(axis=0)
            out[j, i] = vals[np.abs(bar - c).sum(axis=1).argmin()]
    return out


def main():
    out = {"_generated_by": "scripts/make_golden_png.py from img/partition_{2,4}.png of the reference checkout (read-only)",
           "layout": "value[j][i]: picture row j (vertical axis, upwards), column i; the decomposer's pid[y][x] for -o xy is "
                     "value[x][y]; column i = 29 is clipped by the plot frame",
           "swap_1_2_in_partition_4": True}
    for name, P in (("partition_2.png", 2), ("partition_4.png", 4)):
        out[name] = {"parts": P, "value": [[round(float(v), 3) for v in row] for row in read(name, P)]}
    with open(os.path.join(ROOT, "tests", "golden", "rect3030_png_parts.json"), "w") as f:
        json.dump(out, f)
        f.write("\n")


if __name__ == "__main__":
    main()
