// domain_decomp_export.hpp -- TEST INFRASTRUCTURE.  The reference generates this header with CMake
// (generate_export_header, CMakeLists.txt:80-84); for the oracle/_ref build the macro is empty.
#pragma once
#define LIB_EXPORT
