// DomainUtils.cpp -- host-side box geometry (contract of the reference's DomainUtils.cpp:12-35).
#include "DomainUtils.hpp"

#include <algorithm>
#include <cstdlib>
#include <iostream>

int Domain::get_width() const { return p2.x - p1.x; }
int Domain::get_height() const { return p2.y - p1.y; }

namespace {
// closed-interval intersection test followed by the half-open length: touching intervals give 0
inline int interval_overlap(int a1, int a2, int b1, int b2)
{
    if (a2 < b1 || a1 > b2)
        return 0;
    return std::min(a2, b2) - std::max(a1, b1);
}
} // namespace

int domain_overlap(const Domain d1, const Domain d2, const Edge edge)
{
    switch (edge) {
    case TOP:
    case BOTTOM:
        return interval_overlap(d1.p1.x, d1.p2.x, d2.p1.x, d2.p2.x);
    case LEFT:
    case RIGHT:
        return interval_overlap(d1.p1.y, d1.p2.y, d2.p1.y, d2.p2.y);
    default:
        std::cerr << "ERROR: edge must be LEFT, RIGHT, BOTTOM, TOP." << std::endl;
        exit(EXIT_FAILURE);
    }
}
