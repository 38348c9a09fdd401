// cuda_emu.h -- TEST INFRASTRUCTURE: launch interface of the host emulation of the CUDA execution model
// (cuda_emu.cpp).  Kernels are the product's own sources compiled with -DDDC_HOST_EMU (ddc_host_emu.h).
#pragma once
#ifndef DDC_HOST_EMU
#define DDC_HOST_EMU
#endif
#include <cstddef>
#include <functional>
#include <string>

#include "ddc_host_emu.h"

namespace cuda_emu {
struct Dim3 {
    unsigned x = 1, y = 1, z = 1;
    Dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1)
        : x(x_)
        , y(y_)
        , z(z_)
    {
    }
};
// kernel<<<grid, block, dyn_smem>>>(args...)  ==  launch(grid, block, dyn_smem, [&] { kernel(args...); })
// false: the launch was refused or deadlocked (last_error() says why)
bool launch(Dim3 grid, Dim3 block, size_t dyn_smem, const std::function<void()>& body);
const char* last_error();
// 0 (default): deterministic index order; otherwise blocks and threads are scheduled in seeded random orders
void set_schedule_seed(unsigned long long seed);
} // namespace cuda_emu
