"""netCDF classic reader / writer of the host library (domain_decomp_b200/host/NcClassic.cpp) against an
independent implementation (scipy.io.netcdf_file), and Grid::create on real binary netCDF grids --
the reference's own test inputs (test/test_{0,1,2}.cdl, compiled by `ncgen -b` to classic files in the
reference's build, test/CMakeLists.txt:37-59).  CPU only."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT

scipy_io = pytest.importorskip("scipy.io")
NC_TOOL = os.path.join(ROOT, "domain_decomp_b200", "nc_tool")


@pytest.fixture(scope="module", autouse=True)
def built():
    from domain_decomp_b200 import build
    build.build_all()
    assert os.path.exists(NC_TOOL)


def run(*args, stdin=None):
    out = subprocess.run([NC_TOOL, *map(str, args)], input=stdin, capture_output=True, text=True, timeout=120)
    return out.returncode, out.stdout, out.stderr


def write_grid(path, mask, xname, yname, var, order, version, dtype="i4"):
    """the reference's input layout: dims in CDL order, `var(ydim, xdim)` (or (xdim, ydim) for -o xy)"""
    ny, nx = mask.shape
    f = scipy_io.netcdf_file(path, "w", version=version)
    f.createDimension(xname, nx)
    f.createDimension(yname, ny)
    dims = (yname, xname) if order == "yx" else (xname, yname)
    v = f.createVariable(var, dtype, dims)
    v[:] = mask if order == "yx" else mask.reshape(nx, ny)  # raw reinterpretation (quirk Q7)
    f.createVariable("unrelated", "f8", (xname,))[:] = np.arange(nx) * 0.5
    f.title = b"classic test grid"
    f.close()


def parse_grid(out):
    lines = dict(l.split(" ", 1) for l in out.strip().splitlines())
    return (list(map(int, lines["extent"].split())), lines["objects"], list(map(int, lines["block"].split())),
            list(map(int, lines["mask"].split())))


@pytest.mark.parametrize("version", [1, 2])
@pytest.mark.parametrize("name", ["test_0", "test_1", "test_2"])
def test_grid_reads_reference_inputs_as_binary_netcdf(goldens, tmp_path, name, version):
    """test_grid_{0,1,2}.cpp's known answers, with the grid read from a classic netCDF file"""
    inp = goldens["inputs"][name]
    mask = np.asarray(inp["mask"], dtype=np.int32).reshape(inp["ny"], inp["nx"])
    path = str(tmp_path / (name + ".nc"))
    write_grid(path, mask, inp["xdim"], inp["ydim"], inp["mask_name"], "yx", version)
    rc, out, err = run("grid", path, inp["xdim"], inp["ydim"], "yx", inp["mask_name"])
    assert rc == 0, err
    ext, objects, block, m = parse_grid(out)
    assert ext == [6, 4] and block == [0, 0, 6, 4]
    assert objects == "24 nonzero %d" % int((mask > 0).sum())
    assert m == mask.ravel().tolist()
    # two ranks: the naive blocks of Grid.cpp:150-166 (test_grid_2.cpp:29-51)
    for rank in range(2):
        rc, out, err = run("grid", path, inp["xdim"], inp["ydim"], "yx", inp["mask_name"], "ranks", 2, rank)
        assert rc == 0, err
        ext, objects, block, m = parse_grid(out)
        assert block == [3 * rank, 0, 3, 4]
        assert m == mask[:, 3 * rank:3 * rank + 3].ravel().tolist()


@pytest.mark.parametrize("dtype,version", [("f8", 1), ("i2", 2), ("i1", 1), ("f4", 2)])
def test_mask_types_convert_like_nc_get_vara_int(tmp_path, dtype, version):
    rng = np.random.default_rng(3)
    mask = rng.integers(0, 3, size=(7, 9)).astype(np.int32)
    path = str(tmp_path / "g.nc")
    write_grid(path, mask.astype(dtype), "x", "y", "mask", "yx", version, dtype)
    rc, out, err = run("grid", path, "x", "y", "yx", "mask")
    assert rc == 0, err
    assert parse_grid(out)[3] == mask.ravel().tolist()


def test_xy_order_and_errors(tmp_path):
    mask = (np.arange(30).reshape(5, 6) % 3).astype(np.int32)
    path = str(tmp_path / "g.nc")
    write_grid(path, mask, "m", "n", "land_mask", "xy", 2)
    rc, out, err = run("grid", path, "m", "n", "xy", "land_mask")
    assert rc == 0, err
    assert parse_grid(out)[3] == mask.ravel().tolist()
    # the declared order must match (Grid.cpp:110-113)
    rc, out, err = run("grid", path, "m", "n", "yx", "land_mask")
    assert rc != 0 and "Dimension ordering provided does not match" in err
    rc, out, err = run("grid", path, "m", "n", "xy", "nosuchvar")
    assert rc != 0 and "not found" in err
    rc, out, err = run("grid", path, "q", "n", "xy", "land_mask")
    assert rc != 0 and "Invalid dimension" in err
    hdf = tmp_path / "g4.nc"
    hdf.write_bytes(b"\x89HDF\r\n\x1a\n" + b"\0" * 64)
    rc, out, err = run("grid", hdf, "x", "y", "yx", "mask")
    assert rc != 0 and "netCDF-4" in err
    short = tmp_path / "short.nc"
    short.write_bytes(open(path, "rb").read()[:100])
    rc, out, err = run("grid", short, "m", "n", "xy", "land_mask")
    assert rc != 0 and "ERROR" in err


def test_record_variables_and_dump(tmp_path):
    """record (unlimited-dimension) variables are interleaved per record in a classic file"""
    path = str(tmp_path / "rec.nc")
    f = scipy_io.netcdf_file(path, "w", version=1)
    f.createDimension("t", None)
    f.createDimension("x", 3)
    a = f.createVariable("a", "i4", ("t", "x"))
    b = f.createVariable("b", "f8", ("t",))
    c = f.createVariable("c", "i2", ("x",))
    c[:] = [7, 8, 9]
    for t in range(4):
        a[t] = [10 * t, 10 * t + 1, 10 * t + 2]
        b[t] = t + 0.25
    f.close()
    rc, out, err = run("dump", path)
    assert rc == 0, err
    lines = out.strip().splitlines()
    assert "dim t 4" in lines and "dim x 3" in lines
    got = {l.split()[1]: l for l in lines if l.startswith("var ")}
    assert got["a"].split()[2:4] == ["int", "(t,x)"]
    assert list(map(float, got["a"].split()[4:])) == [10 * t + k for t in range(4) for k in range(3)]
    assert list(map(float, got["b"].split()[4:])) == [t + 0.25 for t in range(4)]
    assert list(map(float, got["c"].split()[4:])) == [7, 8, 9]


@pytest.mark.parametrize("version", [0, 1, 2])
def test_written_mask_file_reads_back_with_scipy(tmp_path, version):
    """partition_mask_<P>.nc as save_mask writes it (Partitioner.cpp:128-166): dims y, x; int pid(y, x);
    global attribute num_processes"""
    rng = np.random.default_rng(5)
    nx, ny, P = 11, 6, 3
    pid = rng.integers(-1, P, size=(ny, nx)).astype(np.int32)
    path = str(tmp_path / "partition_mask_3.nc")
    rc, out, err = run("write", path, version, P, nx, ny, stdin=" ".join(map(str, pid.ravel())))
    assert rc == 0, err
    f = scipy_io.netcdf_file(path, "r", mmap=False)
    assert list(f.dimensions.items()) == [("y", ny), ("x", nx)]
    assert f.num_processes == P
    v = f.variables["pid"]
    assert v.dimensions == ("y", "x") and v.data.dtype.kind == "i" and v.data.dtype.itemsize == 4
    assert np.array_equal(v.data, pid)
    assert f.version_byte == (1 if version in (0, 1) else 2)
    f.close()


def test_cdf5_round_trip(tmp_path):
    """the 64-bit-data variant (needed for a pid map of 4 GiB or more) through the library's own reader"""
    pid = (np.arange(35, dtype=np.int32) - 3).reshape(5, 7)
    path = str(tmp_path / "m5.nc")
    rc, out, err = run("write", path, 5, 4, 7, 5, stdin=" ".join(map(str, pid.ravel())))
    assert rc == 0, err
    assert open(path, "rb").read(4) == b"CDF\x05"
    rc, out, err = run("dump", path)
    assert rc == 0, err
    lines = out.strip().splitlines()
    assert "dim x 7" in lines and "dim y 5" in lines
    row = [l for l in lines if l.startswith("var pid int (y,x)")][0]
    assert list(map(float, row.split()[4:])) == pid.ravel().astype(float).tolist()
    rc, out, err = run("grid", path, "x", "y", "yx", "pid")
    assert rc == 0, err
    assert parse_grid(out)[3] == pid.ravel().tolist()


def test_save_mask_writes_cdl_and_netcdf(goldens, tmp_path):
    """Partitioner::save_mask("partition_mask_3.nc") on the golden ids of test_1: the CDL text is the
    reference's ncdump output byte for byte, and the .nc file beside it holds the same thing"""
    import hashlib
    G = goldens["integration"]["test_1"]
    pid = np.asarray(G["pid"], dtype=np.int32).reshape(4, 6)
    path = str(tmp_path / "partition_mask_3.nc")
    rc, out, err = run("savemask", path, 3, 6, 4, stdin=" ".join(map(str, pid.ravel())))
    assert rc == 0, err
    text = open(str(tmp_path / "partition_mask_3.cdl")).read()
    assert hashlib.sha256(text.encode()).hexdigest() == G["mask_cdl_sha256"]
    f = scipy_io.netcdf_file(path, "r", mmap=False)
    assert list(f.dimensions.items()) == [("y", 4), ("x", 6)] and f.num_processes == 3
    assert np.array_equal(f.variables["pid"].data, pid)
    f.close()
