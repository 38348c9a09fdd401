// main.cpp -- the `decomp` command line tool (same options as the reference's main.cpp:45-56).
//
//   decomp -g grid.cdl [--parts N] [-x x] [-y y] [-o yx] [-m mask] [-i] [--px] [--py]
//
// The reference takes the number of parts from `mpirun -n P`; here one process drives the GPU and
// the number of parts is `--parts N` (default: the communicator size).  Outputs are
// partition_mask_<P> and partition_metadata_<P> (CDL text unless built with netCDF).
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <stdexcept>
#include <string>
#include <vector>

#include "CudaRcbPartitioner.hpp"
#include "Grid.hpp"
#include "Partitioner.hpp"

namespace {
struct Options {
    std::string grid, xdim = "x", ydim = "y", order = "yx", mask = "mask";
    bool ignore_mask = false, px = false, py = false, help = false, stats = false;
    int parts = -1;
};

void usage(const char* argv0)
{
    std::cout << "Usage: " << argv0 << " [options]\n"
              << "Options:\n"
                 "  -h [ --help ]             Display this help message\n"
                 "  -g [ --grid ] arg         NetCDF grid file\n"
                 "  -x [ --xdim ] arg (=x)    Name of x dimension in netCDF grid file\n"
                 "  -y [ --ydim ] arg (=y)    Name of y dimension in netCDF grid file\n"
                 "  -o [ --order ] arg (=yx)  Order of dimensions in netCDF grid file, e.g., 'yx'\n"
                 "                            or 'xy'\n"
                 "  -m [ --mask ] arg (=mask) Mask variable name in netCDF grid file\n"
                 "  -i [ --ignore-mask ]      Ignore mask in netCDF grid file\n"
                 "  --periodic-x [ --px ]     Periodicity in x-direction\n"
                 "  --periodic-y [ --py ]     Periodicity in y-direction\n"
                 "  -n [ --parts ] arg        Number of parts (the reference uses the MPI world size)\n"
                 "  --device arg (=0)         CUDA device\n"
                 "  --stats                   Print partitioning statistics\n";
}

// returns false (after printing an error) on a malformed command line
bool parse(int argc, char** argv, Options& o)
{
    auto need = [&](int& i, const std::string& name, std::string& out) {
        if (i + 1 >= argc) {
            std::cerr << "ERROR: the required argument for option '--" << name << "' is missing" << std::endl;
            return false;
        }
        out = argv[++i];
        return true;
    };
    for (int i = 1; i < argc; i++) {
        const std::string a = argv[i];
        std::string v;
        if (a == "-h" || a == "--help")
            o.help = true;
        else if (a == "-g" || a == "--grid") {
            if (!need(i, "grid", o.grid))
                return false;
        } else if (a == "-x" || a == "--xdim") {
            if (!need(i, "xdim", o.xdim))
                return false;
        } else if (a == "-y" || a == "--ydim") {
            if (!need(i, "ydim", o.ydim))
                return false;
        } else if (a == "-o" || a == "--order") {
            if (!need(i, "order", o.order))
                return false;
        } else if (a == "-m" || a == "--mask") {
            if (!need(i, "mask", o.mask))
                return false;
        } else if (a == "-i" || a == "--ignore-mask")
            o.ignore_mask = true;
        else if (a == "--px" || a == "--periodic-x")
            o.px = true;
        else if (a == "--py" || a == "--periodic-y")
            o.py = true;
        else if (a == "-n" || a == "--parts") {
            if (!need(i, "parts", v))
                return false;
            o.parts = std::atoi(v.c_str());
        } else if (a == "--device") {
            if (!need(i, "device", v))
                return false; // consumed again by the partitioner from argv
        } else if (a == "--stats")
            o.stats = true;
        else {
            std::cerr << "ERROR: unrecognised option '" << a << "'" << std::endl;
            return false;
        }
    }
    return true;
}
} // namespace

int main(int argc, char* argv[])
{
    MPI_Comm comm = MPI_COMM_WORLD;
    MPI_Init(&argc, &argv);

    Options opt;
    if (!parse(argc, argv, opt))
        return 1;
    if (opt.help) {
        usage(argv[0]);
        return 0;
    }
    if (opt.grid.empty()) {
        std::cerr << "ERROR: the option '--grid' is required but missing" << std::endl;
        return 1;
    }
    if (opt.order != "xy" && opt.order != "yx") {
        std::cerr << "ERROR: invalid option. [order] must be either 'xy' or 'yx'." << std::endl;
        return 1;
    }
    const std::vector<int> order = opt.order[0] == 'x' ? std::vector<int>({ 0, 1 }) : std::vector<int>({ 1, 0 });

    int rc = 0;
    Grid* grid = nullptr;
    Partitioner* partitioner = nullptr;
    try {
        grid = Grid::create(comm, opt.grid, opt.xdim, opt.ydim, order, opt.mask, opt.ignore_mask, opt.px, opt.py);
        partitioner = Partitioner::Factory::create(comm, argc, argv, PartitionerType::Cuda_RCB);
        if (opt.parts > 0)
            partitioner->set_num_parts(opt.parts);
        partitioner->partition(*grid);
        const int P = partitioner->get_num_parts();
        partitioner->save_mask("partition_mask_" + std::to_string(P) + ".nc");
        partitioner->save_metadata("partition_metadata_" + std::to_string(P) + ".nc");
        if (opt.stats) {
            const ddc_stats& s = static_cast<CudaRcbPartitioner*>(partitioner)->stats();
            const double ave = s.nparts ? (double)s.n_ocean / s.nparts : 0.0;
            std::cout << "Partitioning Statistics:\n"
                      << " Total weight of dots = " << s.n_ocean << "\n"
                      << " Weight on each part: ave = " << ave << ", max = " << s.load_max << ", min = " << s.load_min << "\n"
                      << " RCB levels: " << s.nlev << " (" << s.n_xlev << " cut x, " << s.n_ylev << " cut y), strips = " << s.nstrips << "\n"
                      << " Median find iterations (all cuts): " << s.median_iters << "\n"
                      << " changes = " << s.changes << ", imbalance = " << (ave > 0 ? s.load_max / ave : 1.0)
                      << ", edge cut = " << s.edge_cut << "\n";
        }
    } catch (const std::exception& e) {
        std::cerr << e.what() << std::endl;
        rc = 1;
    }
    delete grid;
    delete partitioner;
    MPI_Finalize();
    return rc;
}
