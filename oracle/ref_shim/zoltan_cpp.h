/* zoltan_cpp.h -- TEST INFRASTRUCTURE.  The reference's ZoltanPartitioner.hpp holds a
 * std::unique_ptr<Zoltan>; Zoltan itself is not available, and ZoltanPartitioner.cpp is NOT compiled
 * by oracle/Makefile (only the Zoltan-free host path is: Grid.cpp, Partitioner.cpp, DomainUtils.cpp). */
#ifndef DDC_REF_SHIM_ZOLTAN_CPP_H
#define DDC_REF_SHIM_ZOLTAN_CPP_H
class Zoltan {
};
#endif
