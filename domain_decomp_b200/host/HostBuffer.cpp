// HostBuffer.cpp -- storage of ddc_host::IntBuffer (HostBuffer.hpp): page-locked through the C ABI when large.
#include "HostBuffer.hpp"

#include <cstdint>
#include <cstdlib>

#include "ddc.h"

namespace ddc_host {
namespace {
constexpr std::size_t HEADER = 64; // keeps the payload 64-byte aligned
constexpr std::uint64_t PINNED = 0x64646370696e6e64ull, HEAP = 0x6464636865617021ull;
}

void* buffer_alloc(std::size_t bytes)
{
    void* base = nullptr;
    std::uint64_t kind = HEAP;
    if (bytes >= (std::size_t)1 << 20 && ddc_host_alloc(&base, bytes + HEADER) == DDC_OK && base)
        kind = PINNED;
    else
        base = std::malloc(bytes + HEADER); // small, or no CUDA device: tools that only read / write files
    if (!base)
        return nullptr;
    *static_cast<std::uint64_t*>(base) = kind;
    return static_cast<char*>(base) + HEADER;
}

void buffer_free(void* p) noexcept
{
    if (!p)
        return;
    void* base = static_cast<char*>(p) - HEADER;
    if (*static_cast<std::uint64_t*>(base) == PINNED)
        ddc_host_free(base);
    else
        std::free(base);
}
} // namespace ddc_host
