// main.cpp -- the `decomp` command line tool (same options as the reference's main.cpp:45-56).
//
//   decomp -g grid.cdl [--parts N] [--gpus G] [-x x] [-y y] [-o yx] [-m mask] [-i] [--px] [--py] [--stats]
//
// The reference takes the number of parts from `mpirun -n P`; here one process drives the GPU and
// the number of parts is `--parts N` (default: the communicator size).  Outputs are
// partition_mask_<P> and partition_metadata_<P> (CDL text unless built with netCDF).
#include <chrono>
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
                 "  --device arg (=0)         (first) CUDA device\n"
                 "  --gpus arg (=1)           Row-shard the mask over this many GPUs of the box (the reference:\n"
                 "                            mpirun -n)\n"
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
        } else if (a == "--device" || a == "--gpus") {
            if (!need(i, a.substr(2), v))
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
        else if (partitioner->get_num_parts() == 1)
            std::cerr << "WARNING: no --parts given and a single-process communicator: decomposing into 1 part "
                         "(the reference takes the part count from mpirun -n)" << std::endl;
        CudaRcbPartitioner* cuda = static_cast<CudaRcbPartitioner*>(partitioner);
        cuda->set_profile(opt.stats);
        const auto t0 = std::chrono::steady_clock::now();
        partitioner->partition(*grid);
        const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        const int P = partitioner->get_num_parts();
        partitioner->save_mask("partition_mask_" + std::to_string(P) + ".nc");
        partitioner->save_metadata("partition_metadata_" + std::to_string(P) + ".nc");
        if (opt.stats) {
            // the layout of Zoltan's DEBUG_LEVEL statistics that the reference prints (README.md:165-193), with
            // what the GPU pipeline has to say: host <-> device copies are inside "Partitioning total time"
            const ddc_stats& s = cuda->stats();
            const double ave = s.nparts ? (double)s.n_ocean / s.nparts : 0.0;
            const double imbal = ave > 0 ? s.load_max / ave : 1.0;
            static const char* stage[DDC_N_STAGES] = { "Mask scan", "x cuts", "Strip row counts", "y cuts", "Labelling",
                "Step end", "Neighbours (beside the labelling)", "Device total" };
            std::cout << "Partitioning total time: " << secs << " (secs)\n"
                      << "Partitioning Statistics:\n"
                      << " GPUs = " << cuda->num_gpus() << ", parts = " << s.nparts << ", grid = " << s.nx << " x " << s.ny << "\n"
                      << " Total weight of dots = " << s.n_ocean << "\n"
                      << " Weight on each part: ave = " << ave << ", max = " << s.load_max << ", min = " << s.load_min << "\n"
                      << " Maximum weight of single dot = 1\n"
                      << " RCB levels: " << s.nlev << " (" << s.n_xlev << " cut x, " << s.n_ylev << " cut y), strips = " << s.nstrips << "\n"
                      << " Median find iteration counts:\n"
                      << "     Total for all cuts: " << s.median_iters << "\n";
            for (int i = 0; i < DDC_N_STAGES; i++)
                std::cout << " " << stage[i] << " time (secs): " << s.stage_ms[i] * 1e-3 << "\n";
            std::cout << " Kernel launches: " << s.gpu_launches << ", changes = " << s.changes << ", edge cut = " << s.edge_cut << "\n"
                      << " STATS Runs 1  bal  CURRENT " << imbal << "  MAX " << imbal << "  MIN " << imbal << "  AVG " << imbal << "\n"
                      << " STATS DETAIL count:  min " << s.load_min << "  max " << s.load_max << "  avg " << ave << "  imbal " << imbal << "\n";
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
