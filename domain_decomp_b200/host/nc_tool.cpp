// nc_tool.cpp -- command-line access to the netCDF classic reader / writer and to Grid, so that the
// CPU test-suite can check them against an independent implementation (scipy.io.netcdf_file)
// without a GPU.
//
//   nc_tool grid <file> <xdim> <ydim> <yx|xy> <maskvar> [ranks N r]   Grid::create -> extents, counts, block mask
//   nc_tool dump <file>                                               header + values of every variable
//   nc_tool savemask <out.nc> <num_processes> <nx> <ny>               Partitioner::save_mask on ids from stdin
//   nc_tool write <out.nc> <version 0|1|2|5> <num_processes> <nx> <ny>   pid values (nx * ny ints) from stdin
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <stdexcept>
#include <string>
#include <vector>

#include "Grid.hpp"
#include "NcClassic.hpp"
#include "Partitioner.hpp"

// a Partitioner whose "partition" is handed in: exercises the base-class writers without a GPU
struct GivenPartitioner final : Partitioner {
    GivenPartitioner(int nparts, int nx, int ny, std::vector<int> pid)
        : Partitioner(ddc_shim_comm(0, 1))
    {
        _num_parts = nparts;
        _global_ext = { nx, ny };
        _pid_global.assign(pid.begin(), pid.end());
    }
    void partition(Grid&) override {}
};

static int usage()
{
    std::fprintf(stderr, "usage: nc_tool grid|dump|write ... (see nc_tool.cpp)\n");
    return 2;
}

int main(int argc, char** argv)
{
    try {
        if (argc < 3)
            return usage();
        const std::string cmd = argv[1];
        if (cmd == "grid") {
            if (argc != 7 && argc != 10)
                return usage();
            const std::string order = argv[5];
            const std::vector<int> ord = order == "xy" ? std::vector<int> { 0, 1 } : std::vector<int> { 1, 0 };
            const int ranks = argc == 10 ? std::atoi(argv[8]) : 1, rank = argc == 10 ? std::atoi(argv[9]) : 0;
            Grid* g = Grid::create(ddc_shim_comm(rank, ranks), argv[2], argv[3], argv[4], ord, argv[6]);
            int g0, g1, e0, e1;
            g->get_bounding_box(g0, g1, e0, e1);
            std::printf("extent %d %d\nobjects %d nonzero %d\nblock %d %d %d %d\nmask", g->get_global_ext()[0],
                g->get_global_ext()[1], g->get_num_objects(), g->get_num_nonzero_objects(), g0, g1, e0, e1);
            const int* m = g->get_land_mask();
            for (int i = 0; i < g->get_num_objects(); i++)
                std::printf(" %d", m[i]);
            std::printf("\n");
            delete g;
            return 0;
        }
        if (cmd == "dump") {
            const ddc_host::CdlFile f = ddc_host::read_netcdf_classic(argv[2]);
            for (const auto& d : f.root.dims)
                std::printf("dim %s %ld\n", d.first.c_str(), d.second);
            for (const auto& v : f.root.vars) {
                std::printf("var %s %s (", v.first.c_str(), v.second.type.c_str());
                for (size_t i = 0; i < v.second.dims.size(); i++)
                    std::printf("%s%s", i ? "," : "", v.second.dims[i].c_str());
                std::printf(")");
                for (double x : v.second.data)
                    std::printf(" %.17g", x);
                std::printf("\n");
            }
            return 0;
        }
        if (cmd == "write") {
            if (argc != 7)
                return usage();
            const int version = std::atoi(argv[3]), P = std::atoi(argv[4]), nx = std::atoi(argv[5]), ny = std::atoi(argv[6]);
            std::vector<int32_t> pid((size_t)nx * ny);
            for (auto& v : pid)
                if (!(std::cin >> v))
                    throw std::runtime_error("ERROR: not enough values on stdin");
            ddc_host::write_netcdf_classic(argv[2], { { "y", (uint64_t)ny }, { "x", (uint64_t)nx } },
                { { "num_processes", P } }, { { "pid", { 0, 1 }, pid.data() } }, version);
            return 0;
        }
        if (cmd == "savemask") { // Partitioner::save_mask on given ids: <out.nc> <num_processes> <nx> <ny>, ids on stdin
            if (argc != 6)
                return usage();
            const int P = std::atoi(argv[3]), nx = std::atoi(argv[4]), ny = std::atoi(argv[5]);
            std::vector<int> pid((size_t)nx * ny);
            for (auto& v : pid)
                if (!(std::cin >> v))
                    throw std::runtime_error("ERROR: not enough values on stdin");
            GivenPartitioner part(P, nx, ny, std::move(pid));
            part.save_mask(argv[2]);
            return 0;
        }
        return usage();
    } catch (const std::exception& e) {
        std::fprintf(stderr, "%s\n", e.what());
        return 1;
    }
}
