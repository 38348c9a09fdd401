"""Shared by the CPU-emulation and the GPU test of the halo exchange (ddc_halo_exchange_f64): what every framed tile
must hold afterwards, derived from the boxes alone -- the ghost cell of part p beyond an edge holds the id of the part
whose box contains the cell across that edge (wrapping when the direction is periodic), everything else is untouched."""
import numpy as np

UNTOUCHED = -7.0


def initial_tiles(boxes, offsets):
    buf = np.full(int(offsets[-1]), UNTOUCHED, dtype=np.float64)
    for p, (x0, y0, ex, ey) in enumerate(boxes):
        t = buf[offsets[p]:offsets[p + 1]].reshape(ey + 2, ex + 2)
        t[1:-1, 1:-1] = p
    return buf


def expected_tiles(boxes, offsets, nx, ny, px, py, periodic):
    owner = np.full((ny, nx), -1, dtype=np.int64)
    for p, (x0, y0, ex, ey) in enumerate(boxes):
        owner[y0:y0 + ey, x0:x0 + ex] = p
    assert (owner >= 0).all(), "boxes do not tile the domain"
    buf = initial_tiles(boxes, offsets)

    def at(x, y):
        if x < 0 or x >= nx:
            if not (periodic and px):
                return UNTOUCHED
            x %= nx
        if y < 0 or y >= ny:
            if not (periodic and py):
                return UNTOUCHED
            y %= ny
        return float(owner[y, x])

    for p, (x0, y0, ex, ey) in enumerate(boxes):
        t = buf[offsets[p]:offsets[p + 1]].reshape(ey + 2, ex + 2)
        for j in range(ey):
            t[j + 1, 0] = at(x0 - 1, y0 + j)
            t[j + 1, ex + 1] = at(x0 + ex, y0 + j)
        for i in range(ex):
            t[0, i + 1] = at(x0 + i, y0 - 1)
            t[ey + 1, i + 1] = at(x0 + i, y0 + ey)
    return buf
