// C entry point around the REFERENCE's own domain_overlap (DomainUtils.cpp:15-35),
// compiled from the reference checkout where it lies (see Makefile target `ref`).
// Test infrastructure only: used to validate orc_domain_overlap in ddc_oracle.c.
#include "DomainUtils.hpp"

extern "C" int ref_domain_overlap(int ax1, int ay1, int ax2, int ay2, int bx1, int by1, int bx2,
    int by2, int edge)
{
    Domain a { { ax1, ay1 }, { ax2, ay2 } };
    Domain b { { bx1, by1 }, { bx2, by2 } };
    return domain_overlap(a, b, static_cast<Edge>(edge));
}
