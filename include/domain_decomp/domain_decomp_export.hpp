#pragma once
// Symbol visibility of the host library (the reference generates this header with CMake's
// GenerateExportHeader, CMakeLists.txt:80-84).
#if defined(__GNUC__)
#define LIB_EXPORT __attribute__((visibility("default")))
#else
#define LIB_EXPORT
#endif
