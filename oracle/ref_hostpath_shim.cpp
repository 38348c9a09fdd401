// ref_hostpath_shim.cpp -- TEST INFRASTRUCTURE (never linked into the product).
//
// Runs the REFERENCE's own host path -- Grid.cpp, Partitioner.cpp and DomainUtils.cpp, compiled from
// the reference checkout where they lie (oracle/Makefile target `ref`, output oracle/_ref/) -- on P
// "MPI ranks" that are P threads of this process, with
//   * a miniature MPI (ref_shim/mpi.h): the collectives of Grid.cpp:142-146 and
//     Partitioner.cpp:85-86,190-205,378-388 over a generation barrier;
//   * an in-memory netCDF (ref_shim/netcdf.h): the files the reference reads and writes are tables
//     in this process (Grid.cpp:51-130, Partitioner.cpp:128-318);
//   * a Partitioner subclass that receives the part boxes and the pid map instead of computing them
//     with Zoltan (which is not available) and then does exactly what ZoltanPartitioner::partition
//     does around the Zoltan call (ZoltanPartitioner.cpp:96-121,172-219): P == 1 shortcut, box
//     clamp, discover_neighbours(), labelling of the rank's naive block.
// What this validates against REAL reference code: the naive block decomposition and ocean lists
// (SURVEY 8a a1, a2), neighbour discovery / halo sizes / halo starts incl. the periodic variants
// (a7-a10), the flattening and the layout of both output files (a11).  The RCB itself (a4) stays a
// restatement of Zoltan.
#include <algorithm>
#include <condition_variable>
#include <cstdio>
#include <cstring>
#include <functional>
#include <map>
#include <mutex>
#include <sstream>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include <mpi.h>
#include <netcdf.h>
#include <netcdf_par.h>

#include "Grid.hpp"
#include "Partitioner.hpp"
#include "ZoltanPartitioner.hpp"

// ------------------------------------------------------------------------------------------------
// miniature MPI: ranks are threads
// ------------------------------------------------------------------------------------------------
namespace {
struct World {
    int size = 1;
    std::mutex m;
    std::condition_variable cv;
    int arrived = 0;
    long generation = 0;
    std::vector<int> slots; // one contribution per rank (every collective here moves ints)
    std::vector<const int*> sendptr; // MPI_Allgatherv: every rank's send buffer
    void barrier()
    {
        std::unique_lock<std::mutex> lk(m);
        const long gen = generation;
        if (++arrived == size) {
            arrived = 0;
            generation++;
            cv.notify_all();
        } else
            cv.wait(lk, [&] { return generation != gen; });
    }
};
} // namespace
struct ref_shim_comm {
    World* world;
    int rank;
};

extern "C" {
int MPI_Comm_rank(MPI_Comm c, int* rank)
{
    *rank = c->rank;
    return MPI_SUCCESS;
}
int MPI_Comm_size(MPI_Comm c, int* size)
{
    *size = c->world->size;
    return MPI_SUCCESS;
}
// every rank deposits `count` ints, waits for the others, and `use` reads the table
static void exchange(MPI_Comm c, const int* send, int count, const std::function<void(const std::vector<int>&)>& use)
{
    World& w = *c->world;
    {
        std::lock_guard<std::mutex> lk(w.m);
        if ((int)w.slots.size() != w.size * count)
            w.slots.assign((size_t)w.size * count, 0);
    }
    w.barrier(); // the table has its size before anyone writes
    std::copy(send, send + count, w.slots.begin() + (size_t)c->rank * count);
    w.barrier();
    use(w.slots);
    w.barrier(); // nobody overwrites the table while another rank still reads it
}
int MPI_Allgather(const void* sendbuf, int sendcount, MPI_Datatype, void* recvbuf, int, MPI_Datatype, MPI_Comm c)
{
    exchange(c, static_cast<const int*>(sendbuf), sendcount,
        [&](const std::vector<int>& t) { std::copy(t.begin(), t.end(), static_cast<int*>(recvbuf)); });
    return MPI_SUCCESS;
}
int MPI_Allgatherv(const void* sendbuf, int, MPI_Datatype, void* recvbuf, const int* recvcounts, const int* displs,
    MPI_Datatype, MPI_Comm c)
{
    World& w = *c->world;
    {
        std::lock_guard<std::mutex> lk(w.m);
        w.sendptr.resize(w.size);
    }
    w.barrier();
    w.sendptr[c->rank] = static_cast<const int*>(sendbuf);
    w.barrier();
    for (int r = 0; r < w.size; r++)
        std::copy(w.sendptr[r], w.sendptr[r] + recvcounts[r], static_cast<int*>(recvbuf) + displs[r]);
    w.barrier(); // every rank has copied before any send buffer goes away
    return MPI_SUCCESS;
}
// rooted collectives (integration/reference_binding): every rank publishes its buffer pointer, the data is copied
// between the barriers
int MPI_Gatherv(const void* sendbuf, int, MPI_Datatype, void* recvbuf, const int* recvcounts, const int* displs,
    MPI_Datatype, int root, MPI_Comm c)
{
    World& w = *c->world;
    {
        std::lock_guard<std::mutex> lk(w.m);
        w.sendptr.resize(w.size);
    }
    w.barrier();
    w.sendptr[c->rank] = static_cast<const int*>(sendbuf);
    w.barrier();
    if (c->rank == root)
        for (int r = 0; r < w.size; r++)
            std::copy(w.sendptr[r], w.sendptr[r] + recvcounts[r], static_cast<int*>(recvbuf) + displs[r]);
    w.barrier();
    return MPI_SUCCESS;
}
int MPI_Scatterv(const void* sendbuf, const int* sendcounts, const int* displs, MPI_Datatype, void* recvbuf, int recvcount,
    MPI_Datatype, int root, MPI_Comm c)
{
    World& w = *c->world;
    {
        std::lock_guard<std::mutex> lk(w.m);
        w.sendptr.resize(w.size);
        w.slots.resize(2 * (size_t)w.size);
    }
    w.barrier();
    if (c->rank == root) {
        w.sendptr[0] = static_cast<const int*>(sendbuf);
        for (int r = 0; r < w.size; r++) {
            w.slots[2 * r] = sendcounts[r];
            w.slots[2 * r + 1] = displs[r];
        }
    }
    w.barrier();
    const int n = std::min(recvcount, w.slots[2 * c->rank]);
    std::copy(w.sendptr[0] + w.slots[2 * c->rank + 1], w.sendptr[0] + w.slots[2 * c->rank + 1] + n, static_cast<int*>(recvbuf));
    w.barrier();
    return MPI_SUCCESS;
}
int MPI_Bcast(void* buffer, int count, MPI_Datatype, int root, MPI_Comm c)
{
    World& w = *c->world;
    {
        std::lock_guard<std::mutex> lk(w.m);
        w.sendptr.resize(w.size);
    }
    w.barrier();
    if (c->rank == root)
        w.sendptr[0] = static_cast<const int*>(buffer);
    w.barrier();
    if (c->rank != root)
        std::copy(w.sendptr[0], w.sendptr[0] + count, static_cast<int*>(buffer));
    w.barrier();
    return MPI_SUCCESS;
}
int MPI_Allreduce(const void* sendbuf, void* recvbuf, int count, MPI_Datatype, MPI_Op, MPI_Comm c)
{
    exchange(c, static_cast<const int*>(sendbuf), count, [&](const std::vector<int>& t) {
        for (int k = 0; k < count; k++) {
            int s = 0;
            for (int r = 0; r < c->world->size; r++)
                s += t[(size_t)r * count + k];
            static_cast<int*>(recvbuf)[k] = s;
        }
    });
    return MPI_SUCCESS;
}
int MPI_Exscan(const void* sendbuf, void* recvbuf, int count, MPI_Datatype, MPI_Op, MPI_Comm c)
{
    exchange(c, static_cast<const int*>(sendbuf), count, [&](const std::vector<int>& t) {
        if (c->rank == 0)
            return; // MPI leaves rank 0's receive buffer untouched
        for (int k = 0; k < count; k++) {
            int s = 0;
            for (int r = 0; r < c->rank; r++)
                s += t[(size_t)r * count + k];
            static_cast<int*>(recvbuf)[k] = s;
        }
    });
    return MPI_SUCCESS;
}
int MPI_Error_string(int code, char* s, int* len)
{
    *len = std::snprintf(s, MPI_MAX_ERROR_STRING, "shim MPI error %d", code);
    return MPI_SUCCESS;
}
int MPI_Finalize(void) { return MPI_SUCCESS; }
}

#include "netcdf_mem.hpp" // the in-memory netCDF (shared with oracle/host_netcdf_shim.cpp)

extern "C" {
int nc_open_par(const char* path, int, MPI_Comm, MPI_Info, int* ncidp)
{
    std::lock_guard<std::mutex> lk(g_fs_mutex);
    const int f = find_file(path);
    if (f < 0)
        return NC_ENOENT;
    *ncidp = f << 8;
    return NC_NOERR;
}
int nc_create_par(const char* path, int, MPI_Comm, MPI_Info, int* ncidp)
{
    // collective in the reference: the first rank to arrive creates the file, the others join it
    std::lock_guard<std::mutex> lk(g_fs_mutex);
    int f = find_file(path);
    if (f < 0) {
        MemFile nf;
        nf.path = path;
        g_files.push_back(nf);
        f = (int)g_files.size() - 1;
    }
    *ncidp = f << 8;
    return NC_NOERR;
}
}

// ------------------------------------------------------------------------------------------------
// the Zoltan-free partitioner and the driver
// ------------------------------------------------------------------------------------------------
// never called: Partitioner::Factory::create (Partitioner.cpp:320-327) refers to it
ZoltanPartitioner* ZoltanPartitioner::create(MPI_Comm, int, char**) { return nullptr; }

namespace {
struct Given {
    int changes; // what Zoltan's LB_Partition would have reported
    const int* boxes; // [P][4] x0, y0, ext_x, ext_y of every part (what RCB_Box + ceil would give)
    const int* pid; // [NY][NX] owner of every cell, -1 on land
};

class GivenBoxesPartitioner final : public Partitioner {
public:
    GivenBoxesPartitioner(MPI_Comm comm, const Given& g)
        : Partitioner(comm)
        , _g(g)
    {
    }
    // the code around the Zoltan call of ZoltanPartitioner::partition, with Zoltan's answers handed in
    void partition(Grid& grid) override
    {
        _num_procs = grid.get_num_procs();
        _global_ext = grid.get_global_ext();
        grid.get_bounding_box(_global[0], _global[1], _local_ext[0], _local_ext[1]);
        _px = grid.get_px();
        _py = grid.get_py();
        const bool single = _total_num_procs == 1;
        for (int d = 0; d < 2; d++) {
            const bool from_rcb = !single && _g.changes == 1;
            _global_new[d] = from_rcb ? _g.boxes[4 * _rank + d] : _global[d];
            _local_ext_new[d] = from_rcb ? _g.boxes[4 * _rank + 2 + d] : _local_ext[d];
        }
        if (!single) {
            for (int d = 0; d < 2; d++) // "adapt to blocking"
                if (_global_new[d] + _local_ext_new[d] > _global_ext[d])
                    _local_ext_new[d] = _global_ext[d] - _global_new[d];
            discover_neighbours();
        }
        // owner of every cell of this rank's ORIGINAL block: own rank on ocean, then the exports
        const int n = grid.get_num_objects();
        const int* land = grid.get_land_mask();
        const bool masked = n != grid.get_num_nonzero_objects();
        _proc_id.assign(n, masked ? -1 : _rank);
        for (int i = 0; i < n; i++) {
            if (masked && !(land[i] > 0))
                continue;
            const int x = _global[0] + i % _local_ext[0], y = _global[1] + i / _local_ext[0];
            _proc_id[i] = single ? _rank : _g.pid[(size_t)y * _global_ext[0] + x];
        }
    }

private:
    Given _g;
};

std::string g_report;
std::string g_error;

} // namespace

namespace {
struct RunArgs {
    int P, nx, ny;
    const int* mask;
    const char *xdim, *ydim, *maskname;
    int order_xy, file_order_xy, data_group, ignore_mask, px, py;
};
// Grid::create + <partitioner>::partition + the getters + save_mask + save_metadata on P thread-ranks;
// make(comm) returns the partitioner of a rank (nullptr: only Grid is exercised).  The text report:
//   rank r block x0 y0 ex ey objects N nonzero M
//   rank r mask v v v ...            the rank's land-mask slab as Grid read it
//   rank r ids g g g ...             Grid's global ids of the ocean cells
//   rank r box x0 y0 ex ey           Partitioner::get_bounding_box
//   rank r nbr <edge> <periodic> id:halo:start ...
//   file mask|metadata / dim / att / var lines: the two output files as written
const char* run_ranks(const RunArgs& a, const std::function<Partitioner*(MPI_Comm)>& make)
{
    g_error.clear();
    g_report.clear();
    const int P = a.P;
    {
        std::lock_guard<std::mutex> lk(g_fs_mutex);
        g_files.clear();
        MemFile in;
        in.path = "grid.nc";
        in.dims = { { a.xdim, (size_t)a.nx }, { a.ydim, (size_t)a.ny } };
        if (a.data_group)
            in.groups.push_back("data");
        MemVar v;
        v.name = a.maskname;
        v.group = a.data_group ? 1 : 0;
        v.dimids = a.file_order_xy ? std::vector<int> { 0, 1 } : std::vector<int> { 1, 0 };
        v.data.assign(a.mask, a.mask + (size_t)a.nx * a.ny);
        v.written.assign(v.data.size(), 1);
        in.vars.push_back(v);
        g_files.push_back(in);
    }
    World world;
    world.size = P;
    std::vector<ref_shim_comm> comms(P);
    std::vector<std::string> part(P), err(P);
    auto body = [&](int r) {
        comms[r] = { &world, r };
        std::ostringstream os;
        try {
            const std::vector<int> order = a.order_xy ? std::vector<int> { 0, 1 } : std::vector<int> { 1, 0 };
            Grid* grid = Grid::create(&comms[r], "grid.nc", a.xdim, a.ydim, order, a.maskname, a.ignore_mask != 0,
                a.px != 0, a.py != 0);
            int b[4];
            grid->get_bounding_box(b[0], b[1], b[2], b[3]);
            os << "rank " << r << " block " << b[0] << " " << b[1] << " " << b[2] << " " << b[3] << " objects "
               << grid->get_num_objects() << " nonzero " << grid->get_num_nonzero_objects() << "\n";
            os << "rank " << r << " mask";
            if (!a.ignore_mask)
                for (int i = 0; i < grid->get_num_objects(); i++)
                    os << " " << grid->get_land_mask()[i];
            os << "\nrank " << r << " ids";
            for (int i = 0; i < grid->get_num_nonzero_objects(); i++)
                os << " " << grid->get_nonzero_object_ids()[i];
            os << "\n";
            if (Partitioner* p = make(&comms[r])) {
                p->partition(*grid);
                p->get_bounding_box(b[0], b[1], b[2], b[3]);
                os << "rank " << r << " box " << b[0] << " " << b[1] << " " << b[2] << " " << b[3] << "\n";
                for (int per = 0; per < 2; per++) {
                    std::vector<std::vector<int>> ids(4), halos(4), starts(4);
                    if (per)
                        p->get_neighbour_info_periodic(ids, halos, starts);
                    else
                        p->get_neighbour_info(ids, halos, starts);
                    for (int e = 0; e < 4; e++) {
                        os << "rank " << r << " nbr " << e << " " << per;
                        for (size_t k = 0; k < ids[e].size(); k++)
                            os << " " << ids[e][k] << ":" << halos[e][k] << ":" << starts[e][k];
                        os << "\n";
                    }
                }
                p->save_mask("partition_mask.nc");
                p->save_metadata("partition_metadata.nc");
                delete p;
            }
            delete grid;
        } catch (const std::exception& e) {
            err[r] = e.what();
        }
        part[r] = os.str();
    };
    // a rank that throws would leave the others waiting in a collective: Grid::create and the
    // writers throw on the same condition on every rank, so either all ranks throw or none does
    std::vector<std::thread> threads;
    for (int r = 0; r < P; r++)
        threads.emplace_back(body, r);
    for (auto& t : threads)
        t.join();
    for (int r = 0; r < P; r++)
        if (!err[r].empty()) {
            g_error = "rank " + std::to_string(r) + ": " + err[r];
            return nullptr;
        }
    std::ostringstream os;
    for (int r = 0; r < P; r++)
        os << part[r];
    {
        std::lock_guard<std::mutex> lk(g_fs_mutex);
        dump_file(os, "mask", "partition_mask.nc");
        dump_file(os, "metadata", "partition_metadata.nc");
    }
    g_report = os.str();
    return g_report.c_str();
}
} // namespace

#ifdef DDC_WITH_BINDING
#include "CudaRcbPartitioner.hpp" // integration/reference_binding: the binding a maintainer adds
#endif

extern "C" {
// The reference's host path with Zoltan's answers handed in (see the header of this file).
// order_xy: the caller's `-o xy`; file_order_xy: the mask variable is DECLARED (xdim, ydim) in the file;
// data_group: dims + variable live in group "data".  boxes == NULL: only Grid is exercised.
// Returns a text report owned by the library (valid until the next call), NULL on failure
// (ref_host_error() has the message).
__attribute__((visibility("default"))) const char* ref_host_run(int P, int nx, int ny, const int* mask, const char* xdim,
    const char* ydim, const char* maskname, int order_xy, int file_order_xy, int data_group, int ignore_mask, int px,
    int py, int changes, const int* boxes, const int* pid)
{
    const RunArgs a { P, nx, ny, mask, xdim, ydim, maskname, order_xy, file_order_xy, data_group, ignore_mask, px, py };
    const Given given { changes, boxes, pid };
    return run_ranks(a, [&](MPI_Comm comm) -> Partitioner* {
        return boxes ? new GivenBoxesPartitioner(comm, given) : nullptr;
    });
}
#ifdef DDC_WITH_BINDING
// The same run with the CUDA-backed partitioner of integration/reference_binding in place of Zoltan:
// reference Grid + binding over the C ABI + reference neighbour discovery + reference writers.
__attribute__((visibility("default"))) const char* ref_binding_run(int P, int nx, int ny, const int* mask, const char* xdim,
    const char* ydim, const char* maskname, int ignore_mask, int px, int py, int device)
{
    const RunArgs a { P, nx, ny, mask, xdim, ydim, maskname, 0, 0, 0, ignore_mask, px, py };
    std::string dev = std::to_string(device);
    char arg0[] = "decomp", arg1[] = "--device";
    char* argv[] = { arg0, arg1, &dev[0], nullptr };
    return run_ranks(a, [&](MPI_Comm comm) -> Partitioner* { return CudaRcbPartitioner::create(comm, 3, argv); });
}
#endif
__attribute__((visibility("default"))) const char* ref_host_error(void) { return g_error.c_str(); }
}
