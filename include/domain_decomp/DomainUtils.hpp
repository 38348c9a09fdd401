// DomainUtils.hpp -- box geometry shared by the host API and (restated) by the neighbour kernel.
// API-compatible with the reference's DomainUtils.hpp:13-58.
#pragma once

#include <array>

#include "domain_decomp_export.hpp"

// Edge numbering is part of the metadata wire format (left, right, bottom, top).
enum Edge { LEFT, RIGHT, BOTTOM, TOP, N_EDGE };
static constexpr std::array<Edge, N_EDGE> edges = { LEFT, RIGHT, BOTTOM, TOP };

struct Point {
    int x, y;
};

// Half-open box: p1 = (x0, y0) is the first cell, p2 = (x0 + width, y0 + height) one past the last.
struct LIB_EXPORT Domain {
    Point p1, p2;
    int get_width() const;
    int get_height() const;
};

// Length of the shared 1-D interval of two boxes along the axis parallel to `edge`
// (x for TOP/BOTTOM, y for LEFT/RIGHT); 0 when they only touch at a corner or not at all.
// It does not test adjacency.  Same contract as the reference (DomainUtils.cpp:15-35),
// including exit(EXIT_FAILURE) on an invalid edge.  (Exported here; the reference keeps it hidden.)
LIB_EXPORT int domain_overlap(const Domain d1, const Domain d2, const Edge edge);
