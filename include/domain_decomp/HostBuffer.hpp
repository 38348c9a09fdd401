// HostBuffer.hpp -- host arrays that cross PCIe: the global land-sea mask of a Grid and the pid map of a
// Partitioner.  Large ones live in page-locked memory obtained through the C ABI (ddc_host_alloc), so that the
// CUDA partitioner copies them at link speed and asynchronously, straight from where the file reader put the
// mask and straight into what save_mask writes; and they are NOT value-initialised (a resize() of 4 GiB does
// not touch 4 GiB).  No counterpart in the reference: its ranks each hold one block of a host-only mask.
#pragma once

#include <cstddef>
#include <new>
#include <utility>
#include <vector>

#include "domain_decomp_export.hpp"

namespace ddc_host {
// page-locked when bytes >= 1 MiB and a CUDA device is usable, otherwise plain heap memory
LIB_EXPORT void* buffer_alloc(std::size_t bytes);
LIB_EXPORT void buffer_free(void* p) noexcept;

template <typename T>
struct BufferAllocator {
    using value_type = T;
    BufferAllocator() = default;
    template <class U>
    BufferAllocator(const BufferAllocator<U>&) noexcept
    {
    }
    T* allocate(std::size_t n)
    {
        void* p = buffer_alloc(n * sizeof(T));
        if (!p)
            throw std::bad_alloc();
        return static_cast<T*>(p);
    }
    void deallocate(T* p, std::size_t) noexcept { buffer_free(p); }
    // default-initialisation: resize(n) leaves the new elements (and their pages) untouched
    template <class U>
    void construct(U* p) noexcept
    {
        ::new (static_cast<void*>(p)) U;
    }
    template <class U, class... A>
    void construct(U* p, A&&... a)
    {
        ::new (static_cast<void*>(p)) U(std::forward<A>(a)...);
    }
    template <class U>
    bool operator==(const BufferAllocator<U>&) const noexcept
    {
        return true;
    }
    template <class U>
    bool operator!=(const BufferAllocator<U>&) const noexcept
    {
        return false;
    }
};
using IntBuffer = std::vector<int, BufferAllocator<int>>;
} // namespace ddc_host
