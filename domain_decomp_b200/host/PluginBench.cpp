// PluginBench.cpp -- the reference-facing call, timed: Grid + Partitioner::Factory::create + partition(grid), the
// call sequence of the reference's main.cpp:84-94, behind one C entry point so that bench.py can measure the
// end-to-end number of the PLUGIN (host mask in the Grid -> boxes, neighbour tables and the pid map in the
// Partitioner's host memory; every host <-> device copy inside the timed call) and not only of the C ABI.
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <memory>
#include <string>

#include "CudaRcbPartitioner.hpp"
#include "Grid.hpp"
#include "Partitioner.hpp"

extern "C" {
// mask[ny][nx] (host).  Runs partition() 1 + reps times (the first one untimed: allocations) and returns the
// seconds of each timed call in seconds[reps].  expect_pid / expect_boxes (may be NULL): what the C ABI gave for
// the same mask -- *same is set to 1 when the plugin's pid map (and x0 y0 ex ey [4][parts]) are identical.
// grid_seconds: building the Grid (a copy of the mask into its page-locked buffer), not part of partition().
LIB_EXPORT int ddc_plugin_bench(const int* mask, int nx, int ny, int parts, int px, int py, int device, int gpus, int reps,
    double* seconds, double* grid_seconds, const int* expect_pid, const int* expect_boxes, int* same, char* err, int errlen)
{
    try {
        const auto g0 = std::chrono::steady_clock::now();
        std::unique_ptr<Grid> grid(Grid::create_from_mask(ddc_shim_comm(0, 1), mask, nx, ny, px != 0, py != 0));
        if (grid_seconds)
            *grid_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - g0).count();
        std::string a0 = "bench", a1 = "--device", a2 = std::to_string(device), a3 = "--gpus", a4 = std::to_string(gpus);
        char* argv[] = { &a0[0], &a1[0], &a2[0], &a3[0], &a4[0], nullptr };
        std::unique_ptr<Partitioner> part(Partitioner::Factory::create(ddc_shim_comm(0, 1), 5, argv, PartitionerType::Cuda_RCB));
        part->set_num_parts(parts);
        part->partition(*grid);
        for (int r = 0; r < reps; r++) {
            const auto t0 = std::chrono::steady_clock::now();
            part->partition(*grid);
            seconds[r] = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        }
        if (same) {
            *same = 1;
            if (expect_pid && std::memcmp(expect_pid, part->get_partition_ids().data(), sizeof(int) * (size_t)nx * ny) != 0)
                *same = 0;
            for (int p = 0; expect_boxes && p < parts; p++) {
                int b[4];
                part->get_bounding_box(p, b[0], b[1], b[2], b[3]);
                for (int i = 0; i < 4; i++)
                    if (b[i] != expect_boxes[(size_t)i * parts + p])
                        *same = 0;
            }
        }
        return 0;
    } catch (const std::exception& e) {
        if (err && errlen > 0)
            std::snprintf(err, (size_t)errlen, "%s", e.what());
        return -1;
    }
}
}
