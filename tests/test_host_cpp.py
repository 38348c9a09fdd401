"""The C++ host layer (Grid / Partitioner / CudaRcbPartitioner / decomp) on top of the C ABI.

  * host_tests: the reference's doctest cases re-expressed against this library
  * decomp: the reference's integration test (test/integration-test.sh): run the CLI on the
    golden inputs into 3 parts and compare both output files, byte for byte, with the text
    `ncdump` printed for the reference's outputs (sha256 of the goldens).
"""
import hashlib
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT

PKG = os.path.join(ROOT, "domain_decomp_b200")


def write_input_cdl(path, name, inp):
    """a grid file as `ncdump` prints it (values only; built from the golden JSON)"""
    nx, ny = inp["nx"], inp["ny"]
    vals = np.asarray(inp["mask"]).reshape(ny, nx)
    rows = ",\n".join("  " + ", ".join(str(int(v)) for v in r) for r in vals)
    text = ("netcdf %s {\ndimensions:\n\t%s = %d ;\n\t%s = %d ;\nvariables:\n\tint %s(%s, %s) ;\n\n"
            "// global attributes:\n\t\t:title = \"%s\" ;\ndata:\n\n %s =\n%s ;\n}\n"
            % (name, inp["xdim"], nx, inp["ydim"], ny, inp["mask_name"], inp["ydim"], inp["xdim"],
               inp.get("title", ""), inp["mask_name"], rows))
    with open(path, "w") as f:
        f.write(text)
    return text


@pytest.fixture(scope="module")
def fixture_dir(tmp_path_factory, goldens):
    d = tmp_path_factory.mktemp("grids")
    for name in ("test_0", "test_1", "test_2"):
        write_input_cdl(os.path.join(d, name + ".cdl"), name, goldens["inputs"][name])
    return str(d)


def test_regenerated_inputs_equal_reference_files(goldens, tmp_path):
    """the grid files rebuilt from the golden JSON are byte-identical to the reference's test/*.cdl"""
    for name in ("test_0", "test_1", "test_2"):
        text = write_input_cdl(os.path.join(tmp_path, name + ".cdl"), name, goldens["inputs"][name])
        assert hashlib.sha256(text.encode()).hexdigest() == goldens["inputs"][name]["cdl_sha256"], name


@pytest.mark.gpu
def test_host_unit_tests(fixture_dir):
    exe = os.path.join(PKG, "host_tests")
    assert os.path.exists(exe), "build with python -m domain_decomp_b200.build"
    out = subprocess.run([exe, fixture_dir], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert " 0 failed" in out.stdout


CASES = {
    "test_1": ("test_1", ["-x", "x", "-y", "y", "-m", "mask", "-o", "yx"]),
    "test_2": ("test_2", ["-x", "m", "-y", "n", "-m", "land_mask", "-o", "yx"]),
    "test_1_px": ("test_1", ["-x", "x", "-y", "y", "-m", "mask", "-o", "yx", "--px"]),
    "test_1_py": ("test_1", ["-x", "x", "-y", "y", "-m", "mask", "-o", "yx", "--py"]),
    "test_1_px_py": ("test_1", ["-x", "x", "-y", "y", "-m", "mask", "-o", "yx", "--px", "--py"]),
}


@pytest.mark.gpu
@pytest.mark.parametrize("case", sorted(CASES))
def test_decomp_cli_integration(goldens, fixture_dir, tmp_path, case):
    inp, flags = CASES[case]
    exe = os.path.join(PKG, "decomp")
    cmd = [exe, "-g", os.path.join(fixture_dir, inp + ".cdl"), "--parts", "3"] + flags
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=tmp_path)
    assert out.returncode == 0, out.stdout + out.stderr
    G = goldens["integration"][case]
    for fname, key in (("partition_mask_3.cdl", "mask_cdl_sha256"), ("partition_metadata_3.cdl", "metadata_cdl_sha256")):
        text = open(os.path.join(tmp_path, fname)).read()
        assert hashlib.sha256(text.encode()).hexdigest() == G[key], "%s differs from the golden:\n%s" % (fname, text)


@pytest.mark.gpu
def test_decomp_cli_errors(fixture_dir, tmp_path):
    exe = os.path.join(PKG, "decomp")
    r = subprocess.run([exe], capture_output=True, text=True, cwd=tmp_path)
    assert r.returncode == 1 and "'--grid' is required" in r.stderr
    r = subprocess.run([exe, "-g", os.path.join(fixture_dir, "test_1.cdl"), "-o", "zz"], capture_output=True, text=True, cwd=tmp_path)
    assert r.returncode == 1 and "must be either 'xy' or 'yx'" in r.stderr
    r = subprocess.run([exe, "-g", os.path.join(fixture_dir, "test_2.cdl"), "--parts", "2"], capture_output=True, text=True, cwd=tmp_path)
    assert r.returncode == 1  # default names x / y / mask do not exist in test_2
    r = subprocess.run([exe, "-h"], capture_output=True, text=True, cwd=tmp_path)
    assert r.returncode == 0 and "--grid" in r.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("version", [1, 2])
def test_decomp_cli_on_binary_netcdf(goldens, tmp_path, version):
    """the same integration case with the grid in a real (classic-format) netCDF file -- what the
    reference's build makes of test/test_1.cdl with `ncgen -b` -- and the mask read back from the
    partition_mask_3.nc the CLI wrote"""
    scipy_io = pytest.importorskip("scipy.io")
    inp = goldens["inputs"]["test_1"]
    grid = str(tmp_path / "test_1.nc")
    f = scipy_io.netcdf_file(grid, "w", version=version)
    f.createDimension(inp["xdim"], inp["nx"])
    f.createDimension(inp["ydim"], inp["ny"])
    f.createVariable(inp["mask_name"], "i4", (inp["ydim"], inp["xdim"]))[:] = np.asarray(
        inp["mask"], dtype=np.int32).reshape(inp["ny"], inp["nx"])
    f.close()
    exe = os.path.join(PKG, "decomp")
    out = subprocess.run([exe, "-g", grid, "--parts", "3", "--px"], capture_output=True, text=True, timeout=300, cwd=tmp_path)
    assert out.returncode == 0, out.stdout + out.stderr
    G = goldens["integration"]["test_1_px"]
    for fname, key in (("partition_mask_3.cdl", "mask_cdl_sha256"), ("partition_metadata_3.cdl", "metadata_cdl_sha256")):
        text = open(os.path.join(tmp_path, fname)).read()
        assert hashlib.sha256(text.encode()).hexdigest() == G[key], "%s differs from the golden:\n%s" % (fname, text)
    m = scipy_io.netcdf_file(str(tmp_path / "partition_mask_3.nc"), "r", mmap=False)
    assert m.num_processes == 3 and list(m.dimensions.items()) == [("y", inp["ny"]), ("x", inp["nx"])]
    assert m.variables["pid"].data.ravel().tolist() == list(G["pid"])
    m.close()
