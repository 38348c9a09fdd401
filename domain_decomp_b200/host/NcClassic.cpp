// NcClassic.cpp -- see NcClassic.hpp.  Grammar followed (netCDF classic format specification):
//
//   netcdf_file = header data
//   header      = magic numrecs dim_list gatt_list var_list
//   magic       = 'C' 'D' 'F' VERSION                      VERSION = 1 | 2 | 5
//   dim_list    = ABSENT | NC_DIMENSION nelems [dim ...]   dim = name dim_length (0 = the record dimension)
//   att_list    = ABSENT | NC_ATTRIBUTE nelems [attr ...]  attr = name nc_type nelems [values ...] padding
//   var_list    = ABSENT | NC_VARIABLE nelems [var ...]    var = name nelems [dimid ...] vatt_list nc_type vsize begin
//   name        = nelems namestring padding                (padding to a multiple of 4 bytes)
//   ABSENT      = ZERO ZERO,  NC_DIMENSION = 10, NC_VARIABLE = 11, NC_ATTRIBUTE = 12   (tags are 32-bit)
//   nelems, dim_length, dimid, vsize, numrecs: 32-bit in CDF-1 / CDF-2, 64-bit in CDF-5
//   begin: 32-bit in CDF-1, 64-bit in CDF-2 / CDF-5;  everything big-endian
//   data: fixed-size variables at `begin`, each padded to 4 bytes; record variables interleaved per record
#include "NcClassic.hpp"

#include <cstring>
#include <fstream>
#include <limits>
#include <stdexcept>

namespace ddc_host {
namespace {

enum NcType { NC_BYTE = 1, NC_CHAR, NC_SHORT, NC_INT, NC_FLOAT, NC_DOUBLE, NC_UBYTE, NC_USHORT, NC_UINT, NC_INT64, NC_UINT64 };
constexpr uint32_t TAG_DIM = 10, TAG_VAR = 11, TAG_ATT = 12;

[[noreturn]] void bad(const std::string& path, const std::string& what)
{
    throw std::runtime_error("ERROR: NetCDF: " + what + " (" + path + ")");
}

size_t type_size(int t)
{
    switch (t) {
    case NC_BYTE:
    case NC_CHAR:
    case NC_UBYTE:
        return 1;
    case NC_SHORT:
    case NC_USHORT:
        return 2;
    case NC_INT:
    case NC_UINT:
    case NC_FLOAT:
        return 4;
    case NC_DOUBLE:
    case NC_INT64:
    case NC_UINT64:
        return 8;
    default:
        return 0;
    }
}
const char* type_name(int t)
{
    static const char* names[] = { "", "byte", "char", "short", "int", "float", "double", "ubyte", "ushort", "uint",
        "int64", "uint64" };
    return (t >= 1 && t <= 11) ? names[t] : "?";
}

class Reader {
public:
    Reader(const std::string& path)
        : _path(path)
        , _in(path, std::ios::binary)
    {
        if (!_in)
            bad(path, "No such file or directory");
        _in.seekg(0, std::ios::end);
        _size = (uint64_t)_in.tellg();
        _in.seekg(0);
    }
    void bytes(void* dst, size_t n)
    {
        _in.read(static_cast<char*>(dst), (std::streamsize)n);
        if ((size_t)_in.gcount() != n)
            bad(_path, "file is truncated");
    }
    uint32_t u32()
    {
        unsigned char b[4];
        bytes(b, 4);
        return ((uint32_t)b[0] << 24) | ((uint32_t)b[1] << 16) | ((uint32_t)b[2] << 8) | b[3];
    }
    uint64_t u64()
    {
        const uint64_t hi = u32();
        return (hi << 32) | u32();
    }
    uint64_t count() { return version == 5 ? u64() : u32(); } // nelems, lengths, ids, sizes
    uint64_t offset() { return version == 1 ? u32() : u64(); }
    std::string name()
    {
        const uint64_t n = count();
        if (n > 65536)
            bad(_path, "implausible name length: not a netCDF classic header");
        std::string s((size_t)n, '\0');
        bytes(&s[0], (size_t)n);
        skip((4 - n % 4) % 4);
        return s;
    }
    void skip(uint64_t n) { _in.seekg((std::streamoff)n, std::ios::cur); }
    void seek(uint64_t pos)
    {
        _in.clear();
        _in.seekg((std::streamoff)pos);
    }
    uint64_t size() const { return _size; }
    int version = 0;

private:
    std::string _path;
    std::ifstream _in;
    uint64_t _size = 0;
};

struct VarHeader {
    std::string name;
    std::vector<uint64_t> dimids;
    int type = 0;
    uint64_t vsize = 0, begin = 0;
    bool record = false;
};

void skip_att_list(Reader& r, const std::string& path)
{
    const uint32_t tag = r.u32();
    const uint64_t n = r.count();
    if (tag == 0 && n == 0)
        return;
    if (tag != TAG_ATT)
        bad(path, "malformed attribute list");
    for (uint64_t i = 0; i < n; i++) {
        r.name();
        const int t = (int)r.u32();
        const uint64_t ne = r.count();
        const size_t ts = type_size(t);
        if (!ts)
            bad(path, "unknown attribute type");
        const uint64_t nbytes = ne * ts;
        r.skip(nbytes + (4 - nbytes % 4) % 4);
    }
}

double element(const unsigned char* p, int t)
{
    auto be = [&](int n) {
        uint64_t v = 0;
        for (int i = 0; i < n; i++)
            v = (v << 8) | p[i];
        return v;
    };
    switch (t) {
    case NC_BYTE:
        return (double)(int8_t)p[0];
    case NC_CHAR:
    case NC_UBYTE:
        return (double)p[0];
    case NC_SHORT:
        return (double)(int16_t)be(2);
    case NC_USHORT:
        return (double)(uint16_t)be(2);
    case NC_INT:
        return (double)(int32_t)be(4);
    case NC_UINT:
        return (double)(uint32_t)be(4);
    case NC_FLOAT: {
        const uint32_t u = (uint32_t)be(4);
        float f;
        std::memcpy(&f, &u, 4);
        return (double)f;
    }
    case NC_DOUBLE: {
        const uint64_t u = be(8);
        double d;
        std::memcpy(&d, &u, 8);
        return d;
    }
    case NC_INT64:
        return (double)(int64_t)be(8);
    case NC_UINT64:
        return (double)be(8);
    }
    return 0.0;
}

// ---- writer helpers ----------------------------------------------------------------------------
struct Out {
    std::string buf;
    int version = 1;
    void u32(uint32_t v)
    {
        const char b[4] = { (char)(v >> 24), (char)(v >> 16), (char)(v >> 8), (char)v };
        buf.append(b, 4);
    }
    void u64(uint64_t v)
    {
        u32((uint32_t)(v >> 32));
        u32((uint32_t)v);
    }
    void count(uint64_t v)
    {
        if (version == 5)
            u64(v);
        else
            u32((uint32_t)v);
    }
    void offset(uint64_t v)
    {
        if (version == 1)
            u32((uint32_t)v);
        else
            u64(v);
    }
    void name(const std::string& s)
    {
        count(s.size());
        buf.append(s);
        buf.append((4 - s.size() % 4) % 4, '\0');
    }
};

std::string build_header(int version, const std::vector<NcDim>& dims, const std::vector<NcIntAttr>& atts,
    const std::vector<NcIntVar>& vars, const std::vector<uint64_t>& vsize, const std::vector<uint64_t>& begin)
{
    Out o;
    o.version = version;
    o.buf.append("CDF", 3);
    o.buf.push_back((char)version);
    o.count(0); // numrecs: no record variables
    auto list = [&](uint32_t tag, size_t n) {
        o.u32(n ? tag : 0);
        o.count(n);
    };
    list(TAG_DIM, dims.size());
    for (const NcDim& d : dims) {
        o.name(d.name);
        o.count(d.len);
    }
    list(TAG_ATT, atts.size());
    for (const NcIntAttr& a : atts) {
        o.name(a.name);
        o.u32(NC_INT);
        o.count(1);
        o.u32((uint32_t)a.value);
    }
    list(TAG_VAR, vars.size());
    for (size_t i = 0; i < vars.size(); i++) {
        o.name(vars[i].name);
        o.count(vars[i].dimids.size());
        for (int id : vars[i].dimids)
            o.count((uint64_t)id);
        list(TAG_ATT, 0);
        o.u32(NC_INT);
        o.count(vsize[i]);
        o.offset(begin[i]);
    }
    return o.buf;
}

} // namespace

FileKind sniff_file_kind(const std::string& path)
{
    std::ifstream in(path, std::ios::binary);
    if (!in)
        throw std::runtime_error("ERROR: NetCDF: No such file or directory (" + path + ")");
    unsigned char m[8] = { 0 };
    in.read(reinterpret_cast<char*>(m), 8);
    if (m[0] == 'C' && m[1] == 'D' && m[2] == 'F' && (m[3] == 1 || m[3] == 2 || m[3] == 5))
        return FileKind::NetcdfClassic;
    if (m[0] == 0x89 && m[1] == 'H' && m[2] == 'D' && m[3] == 'F')
        return FileKind::Hdf5;
    return FileKind::Text;
}

CdlFile read_netcdf_classic(const std::string& path, const std::string& only_var, bool as_int)
{
    Reader r(path);
    unsigned char magic[4];
    r.bytes(magic, 4);
    if (!(magic[0] == 'C' && magic[1] == 'D' && magic[2] == 'F' && (magic[3] == 1 || magic[3] == 2 || magic[3] == 5)))
        bad(path, "Unknown file format");
    r.version = magic[3];
    uint64_t numrecs = r.count();
    const bool streaming = r.version == 5 ? numrecs == ~0ull : numrecs == 0xffffffffull;

    CdlFile file;
    file.name = netcdf_name_of(path);
    std::vector<NcDim> dims;
    {
        const uint32_t tag = r.u32();
        const uint64_t n = r.count();
        if (!(tag == 0 && n == 0)) {
            if (tag != TAG_DIM)
                bad(path, "malformed dimension list");
            for (uint64_t i = 0; i < n; i++) {
                NcDim d;
                d.name = r.name();
                d.len = r.count();
                dims.push_back(d);
            }
        }
    }
    skip_att_list(r, path); // global attributes
    std::vector<VarHeader> vars;
    {
        const uint32_t tag = r.u32();
        const uint64_t n = r.count();
        if (!(tag == 0 && n == 0)) {
            if (tag != TAG_VAR)
                bad(path, "malformed variable list");
            for (uint64_t i = 0; i < n; i++) {
                VarHeader v;
                v.name = r.name();
                const uint64_t nd = r.count();
                if (nd > 1024)
                    bad(path, "implausible number of dimensions");
                for (uint64_t k = 0; k < nd; k++) {
                    const uint64_t id = r.count();
                    if (id >= dims.size())
                        bad(path, "Invalid dimension ID or name");
                    v.dimids.push_back(id);
                }
                skip_att_list(r, path);
                v.type = (int)r.u32();
                if (!type_size(v.type))
                    bad(path, "unknown variable type");
                v.vsize = r.count();
                v.begin = r.offset();
                v.record = !v.dimids.empty() && dims[v.dimids[0]].len == 0;
                vars.push_back(v);
            }
        }
    }
    // record size: the sum of the record variables' slabs; a single record variable is not padded
    uint64_t recsize = 0;
    size_t nrecvars = 0;
    for (const VarHeader& v : vars)
        if (v.record) {
            recsize += v.vsize;
            nrecvars++;
        }
    auto slab_elems = [&](const VarHeader& v) { // values per record (record var) or in total
        uint64_t n = 1;
        for (size_t k = v.record ? 1 : 0; k < v.dimids.size(); k++)
            n *= dims[v.dimids[k]].len;
        return n;
    };
    if (nrecvars == 1)
        for (const VarHeader& v : vars)
            if (v.record)
                recsize = slab_elems(v) * type_size(v.type);
    if (streaming) { // numrecs was never written back: derive it from the file size
        numrecs = 0;
        uint64_t first = ~0ull;
        for (const VarHeader& v : vars)
            if (v.record && v.begin < first)
                first = v.begin;
        if (nrecvars && recsize && r.size() > first)
            numrecs = (r.size() - first) / recsize;
    }
    for (const NcDim& d : dims)
        file.root.dims[d.name] = (long)(d.len == 0 ? numrecs : d.len);

    for (const VarHeader& v : vars) {
        CdlVar cv;
        cv.type = type_name(v.type);
        for (uint64_t id : v.dimids)
            cv.dims.push_back(dims[id].name);
        if (only_var.empty() || only_var == v.name) {
            const size_t ts = type_size(v.type);
            const uint64_t per = slab_elems(v), slabs = v.record ? numrecs : 1;
            if (per * slabs > (uint64_t)std::numeric_limits<int>::max())
                bad(path, "variable '" + v.name + "' has more values than an int can index");
            if (as_int)
                cv.idata.reserve((size_t)(per * slabs));
            else
                cv.data.reserve((size_t)(per * slabs));
            std::vector<unsigned char> buf;
            const size_t chunk = 1 << 20; // values per read
            for (uint64_t s = 0; s < slabs; s++) {
                r.seek(v.begin + s * recsize);
                for (uint64_t done = 0; done < per; done += chunk) {
                    const size_t n = (size_t)std::min<uint64_t>(chunk, per - done);
                    buf.resize(n * ts);
                    r.bytes(buf.data(), n * ts);
                    if (as_int && v.type == NC_INT) { // the common case: a byte swap
                        for (size_t i = 0; i < n; i++) {
                            const unsigned char* q = buf.data() + 4 * i;
                            cv.idata.push_back((int)(((uint32_t)q[0] << 24) | ((uint32_t)q[1] << 16) | ((uint32_t)q[2] << 8) | q[3]));
                        }
                    } else if (as_int) {
                        for (size_t i = 0; i < n; i++)
                            cv.idata.push_back((int)element(buf.data() + i * ts, v.type));
                    } else {
                        for (size_t i = 0; i < n; i++)
                            cv.data.push_back(element(buf.data() + i * ts, v.type));
                    }
                }
            }
            cv.has_data = true;
        }
        file.root.vars[v.name] = std::move(cv);
    }
    return file;
}

void write_netcdf_classic(const std::string& path, const std::vector<NcDim>& dims,
    const std::vector<NcIntAttr>& global_attrs, const std::vector<NcIntVar>& vars, int version)
{
    std::vector<uint64_t> vsize(vars.size()), begin(vars.size(), 0);
    uint64_t total = 0, largest = 0;
    for (size_t i = 0; i < vars.size(); i++) {
        uint64_t n = 1;
        for (int id : vars[i].dimids) {
            if (id < 0 || (size_t)id >= dims.size())
                bad(path, "Invalid dimension ID or name");
            n *= dims[id].len;
        }
        vsize[i] = n * 4; // NC_INT: already a multiple of 4
        total += vsize[i];
        largest = std::max(largest, vsize[i]);
    }
    if (version == 0) {
        const uint64_t slack = 1 << 20; // header
        version = largest >= (1ull << 32) - 4 ? 5 : (total + slack < (1ull << 31) ? 1 : 2);
    }
    if (version != 1 && version != 2 && version != 5)
        bad(path, "unsupported classic format version");
    if (version != 5 && largest >= (1ull << 32) - 4)
        bad(path, "a variable of 4 GiB or more needs the CDF-5 format");
    // the header's own size does not depend on the begin offsets: build once to measure, once for real
    const uint64_t hsize = build_header(version, dims, global_attrs, vars, vsize, begin).size();
    uint64_t pos = (hsize + 3) & ~(uint64_t)3;
    for (size_t i = 0; i < vars.size(); i++) {
        begin[i] = pos;
        pos += vsize[i];
    }
    if (version == 1 && pos >= (1ull << 31))
        bad(path, "file too large for the CDF-1 format");
    std::string header = build_header(version, dims, global_attrs, vars, vsize, begin);
    header.append(((hsize + 3) & ~(uint64_t)3) - hsize, '\0');

    std::ofstream out(path, std::ios::binary | std::ios::trunc);
    if (!out)
        throw std::runtime_error("ERROR: cannot write '" + path + "'");
    out.write(header.data(), (std::streamsize)header.size());
    std::vector<unsigned char> buf;
    for (size_t i = 0; i < vars.size(); i++) {
        const uint64_t n = vsize[i] / 4;
        const size_t chunk = 1 << 20;
        for (uint64_t done = 0; done < n; done += chunk) {
            const size_t m = (size_t)std::min<uint64_t>(chunk, n - done);
            buf.resize(m * 4);
            for (size_t k = 0; k < m; k++) {
                const uint32_t v = (uint32_t)vars[i].data[done + k];
                buf[4 * k] = (unsigned char)(v >> 24);
                buf[4 * k + 1] = (unsigned char)(v >> 16);
                buf[4 * k + 2] = (unsigned char)(v >> 8);
                buf[4 * k + 3] = (unsigned char)v;
            }
            out.write(reinterpret_cast<const char*>(buf.data()), (std::streamsize)buf.size());
        }
    }
    if (!out)
        throw std::runtime_error("ERROR: short write to '" + path + "'");
}

} // namespace ddc_host
