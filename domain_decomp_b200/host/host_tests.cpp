// host_tests.cpp -- the reference's unit tests (test/test_grid_{0,1,2}.cpp,
// test/test_zoltan_partitioner_{0,1,2}.cpp) re-expressed against this library.
//
// The reference runs them under `mpirun -n 4` with MPI_TEST_CASE(name, N); here "rank r of N" is a
// shim communicator and the ranks are visited in a loop (the CUDA partitioner computes every part
// from the global mask, no communication is involved).  Needs a GPU.
//
//   host_tests <dir with test_0.cdl test_1.cdl test_2.cdl>
#include <cstdio>
#include <functional>
#include <stdexcept>
#include <string>
#include <vector>

#include "DomainUtils.hpp"
#include "Grid.hpp"
#include "Partitioner.hpp"

static int g_fail = 0, g_checks = 0;
#define REQUIRE(cond)                                                                              \
    do {                                                                                           \
        g_checks++;                                                                                \
        if (!(cond)) {                                                                             \
            g_fail++;                                                                              \
            std::printf("  FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond);                        \
        }                                                                                          \
    } while (0)

static std::string g_dir;
static int g_argc = 1;
static char g_arg0[] = "host_tests";
static char* g_argv[] = { g_arg0, nullptr };

static void for_ranks(int n, const std::function<void(MPI_Comm, int)>& body)
{
    for (int r = 0; r < n; r++)
        body(ddc_shim_comm(r, n), r);
}

static void grid_tests()
{
    // test_grid_0.cpp: all land
    for (int n : { 1, 2 })
        for_ranks(n, [&](MPI_Comm comm, int) {
            Grid* grid = Grid::create(comm, g_dir + "/test_0.cdl");
            REQUIRE(grid->get_global_ext()[0] == 6);
            REQUIRE(grid->get_global_ext()[1] == 4);
            REQUIRE(grid->get_num_nonzero_objects() == 0);
            REQUIRE(grid->get_num_objects() == 24 / n);
            const int* mask = grid->get_land_mask();
            for (int i = 0; i < grid->get_num_objects(); i++)
                REQUIRE(mask[i] == 0);
            delete grid;
        });
    // test_grid_1.cpp: no land
    for (int n : { 1, 2 })
        for_ranks(n, [&](MPI_Comm comm, int) {
            Grid* grid = Grid::create(comm, g_dir + "/test_1.cdl");
            REQUIRE(grid->get_global_ext()[0] == 6);
            REQUIRE(grid->get_global_ext()[1] == 4);
            REQUIRE(grid->get_num_objects() == 24 / n);
            REQUIRE(grid->get_num_nonzero_objects() == 24 / n);
            const int* mask = grid->get_land_mask();
            for (int i = 0; i < grid->get_num_objects(); i++)
                REQUIRE(mask[i] == 1);
            delete grid;
        });
    // test_grid_2.cpp: non-default dimension naming
    for (int n : { 1, 2 })
        for_ranks(n, [&](MPI_Comm comm, int) {
            Grid* grid = Grid::create(comm, g_dir + "/test_2.cdl", "m", "n", { 1, 0 }, "land_mask");
            REQUIRE(grid->get_global_ext()[0] == 6);
            REQUIRE(grid->get_global_ext()[1] == 4);
            REQUIRE(grid->get_num_objects() == 24 / n);
            REQUIRE(grid->get_num_nonzero_objects() == 12 / n);
            const int* mask = grid->get_land_mask();
            for (int i = 0; i < grid->get_num_objects(); i++)
                if ((i / grid->get_local_ext()[0]) % 2 == 0)
                    REQUIRE(mask[i] == 0);
            delete grid;
        });
    // wrong dimension order must throw (Grid.cpp:110-113)
    bool threw = false;
    try {
        Grid* grid = Grid::create(MPI_COMM_WORLD, g_dir + "/test_2.cdl", "m", "n", { 0, 1 }, "land_mask");
        delete grid;
    } catch (const std::runtime_error&) {
        threw = true;
    }
    REQUIRE(threw);
}

struct Box {
    int x0, y0, ex, ey;
};

static void expect_boxes(const std::string& file, const std::vector<std::string>& names, int n, const std::vector<Box>& want)
{
    for_ranks(n, [&](MPI_Comm comm, int r) {
        Grid* grid = names.empty() ? Grid::create(comm, g_dir + "/" + file)
                                   : Grid::create(comm, g_dir + "/" + file, names[0], names[1], { 1, 0 }, names[2]);
        Partitioner* partitioner = Partitioner::Factory::create(comm, g_argc, g_argv, PartitionerType::Zoltan_RCB);
        partitioner->partition(*grid);
        int global_0, global_1, local_ext_0, local_ext_1;
        partitioner->get_bounding_box(global_0, global_1, local_ext_0, local_ext_1);
        REQUIRE(global_0 == want[r].x0);
        REQUIRE(global_1 == want[r].y0);
        REQUIRE(local_ext_0 == want[r].ex);
        REQUIRE(local_ext_1 == want[r].ey);
        delete grid;
        delete partitioner;
    });
}

static void partitioner_tests()
{
    const std::vector<std::string> none, t2 = { "m", "n", "land_mask" };
    // test_zoltan_partitioner_0.cpp (all land), _1.cpp (no land)
    for (const char* f : { "test_0.cdl", "test_1.cdl" }) {
        expect_boxes(f, none, 1, { { 0, 0, 6, 4 } });
        expect_boxes(f, none, 2, { { 0, 0, 3, 4 }, { 3, 0, 3, 4 } });
    }
    // test_zoltan_partitioner_2.cpp (striped mask, non-default names)
    expect_boxes("test_2.cdl", t2, 1, { { 0, 0, 6, 4 } });
    expect_boxes("test_2.cdl", t2, 2, { { 0, 0, 3, 4 }, { 3, 0, 3, 4 } });
    expect_boxes("test_2.cdl", t2, 3, { { 0, 0, 2, 4 }, { 2, 0, 2, 4 }, { 4, 0, 2, 4 } });
    expect_boxes("test_2.cdl", t2, 4, { { 0, 0, 1, 4 }, { 1, 0, 2, 4 }, { 3, 0, 1, 4 }, { 4, 0, 2, 4 } });
    // invalid partitioner type
    bool threw = false;
    try {
        Partitioner::Factory::create(MPI_COMM_WORLD, g_argc, g_argv, static_cast<PartitionerType>(99));
    } catch (const std::runtime_error& e) {
        threw = std::string(e.what()) == "Invalid partitioner!";
    }
    REQUIRE(threw);
    // per-rank neighbour getters agree with the all-parts tables (test_1, 3 parts, periodic x + y)
    for_ranks(3, [&](MPI_Comm comm, int r) {
        Grid* grid = Grid::create(comm, g_dir + "/test_1.cdl", false, true, true);
        Partitioner* p = Partitioner::Factory::create(comm, g_argc, g_argv, PartitionerType::Cuda_RCB);
        p->partition(*grid);
        std::vector<std::vector<int>> a(N_EDGE), b(N_EDGE), c(N_EDGE), a2(N_EDGE), b2(N_EDGE), c2(N_EDGE);
        p->get_neighbour_info(a, b, c);
        p->get_neighbour_info(r, a2, b2, c2);
        REQUIRE(a == a2 && b == b2 && c == c2);
        std::vector<std::vector<int>> pa(N_EDGE), pb(N_EDGE), pc(N_EDGE), pa2(N_EDGE), pb2(N_EDGE), pc2(N_EDGE);
        p->get_neighbour_info_periodic(pa, pb, pc);
        p->get_neighbour_info_periodic(r, pa2, pb2, pc2);
        REQUIRE(pa == pa2 && pb == pb2 && pc == pc2);
        if (r == 2) { // ref_partition_metadata_3.cdl of test_1_px_py: part 2 is its own B/T periodic neighbour
            REQUIRE(pa[BOTTOM] == std::vector<int>({ 2 }));
            REQUIRE(pb[BOTTOM] == std::vector<int>({ 2 }));
            REQUIRE(pc[BOTTOM] == std::vector<int>({ 6 }));
            REQUIRE(a[LEFT] == std::vector<int>({ 0, 1 }));
        }
        delete grid;
        delete p;
    });
}

static void domain_tests()
{
    Domain a { { 0, 0 }, { 4, 2 } }, b { { 0, 2 }, { 4, 4 } }, c { { 4, 0 }, { 6, 4 } };
    REQUIRE(a.get_width() == 4 && a.get_height() == 2);
    REQUIRE(domain_overlap(a, b, TOP) == 4);
    REQUIRE(domain_overlap(a, c, RIGHT) == 2);
    REQUIRE(domain_overlap(a, c, TOP) == 0); // touching in x only
    REQUIRE(domain_overlap(b, c, LEFT) == 2);
}

int main(int argc, char** argv)
{
    if (argc < 2) {
        std::printf("usage: host_tests <fixture dir>\n");
        return 2;
    }
    g_dir = argv[1];
    MPI_Init(&argc, &argv);
    try {
        domain_tests();
        grid_tests();
        partitioner_tests();
    } catch (const std::exception& e) {
        std::printf("EXCEPTION: %s\n", e.what());
        g_fail++;
    }
    MPI_Finalize();
    std::printf("host_tests: %d checks, %d failed\n", g_checks, g_fail);
    return g_fail ? 1 : 0;
}
