// netcdf_mem.hpp -- TEST INFRASTRUCTURE.  The in-memory netCDF behind oracle/ref_shim/netcdf.h: a file table in
// memory and the netCDF-C calls the reference's Grid.cpp / Partitioner.cpp make (only int data, only what is
// called).  Included by oracle/ref_hostpath_shim.cpp (the REFERENCE's sources run on it) and by
// oracle/host_netcdf_shim.cpp (THIS repository's host layer built with -DHAVE_NETCDF runs on it), so that the two
// writers can be compared call for call: same dimensions (a zero length makes a dimension UNLIMITED), groups,
// variables and values.  Not part of the product.
#pragma once
#include <cstring>
#include <mutex>
#include <sstream>
#include <string>
#include <vector>

#include "netcdf.h"

// ------------------------------------------------------------------------------------------------
// in-memory netCDF
// ------------------------------------------------------------------------------------------------
namespace {
struct MemDim {
    std::string name;
    size_t len;
};
struct MemVar {
    std::string name;
    int group;
    std::vector<int> dimids;
    std::vector<int> data;
    std::vector<char> written;
};
struct MemFile {
    std::string path;
    std::vector<MemDim> dims; // file-wide (the reference defines all dimensions in the root group)
    std::vector<std::string> groups { "" }; // 0 = root
    std::vector<MemVar> vars;
    std::vector<std::pair<std::string, int>> atts; // global int attributes
};
std::mutex g_fs_mutex;
std::vector<MemFile> g_files;

// ncid = file index * 256 + group index
MemFile* file_of(int ncid)
{
    const int f = ncid >> 8;
    return f >= 0 && f < (int)g_files.size() ? &g_files[f] : nullptr;
}
int group_of(int ncid) { return ncid & 255; }
int find_file(const std::string& path)
{
    for (size_t i = 0; i < g_files.size(); i++)
        if (g_files[i].path == path)
            return (int)i;
    return -1;
}
size_t var_size(const MemFile& f, const MemVar& v)
{
    size_t n = 1;
    for (int d : v.dimids)
        n *= f.dims[d].len;
    return n;
}
void ensure_storage(const MemFile& f, MemVar& v)
{
    const size_t n = var_size(f, v);
    if (v.data.size() != n) {
        v.data.assign(n, 0);
        v.written.assign(n, 0);
    }
}
} // namespace

extern "C" {
const char* nc_strerror(int err)
{
    switch (err) {
    case NC_NOERR:
        return "No error";
    case NC_EBADID:
        return "NetCDF: Not a valid ID";
    case NC_EBADDIM:
        return "NetCDF: Invalid dimension ID or name";
    case NC_ENOTVAR:
        return "NetCDF: Variable not found";
    case NC_EEDGE:
        return "NetCDF: Start+count exceeds dimension bound";
    case NC_ENOGRP:
        return "NetCDF: Bad group ID";
    case NC_ENOENT:
        return "No such file or directory";
    }
    return "NetCDF: unknown error";
}
int nc_close(int) { return NC_NOERR; }
int nc_enddef(int) { return NC_NOERR; }
int nc_var_par_access(int, int, int) { return NC_NOERR; }
int nc_inq_ncid(int ncid, const char* name, int* grp)
{
    std::lock_guard<std::mutex> lk(g_fs_mutex);
    MemFile* f = file_of(ncid);
    if (!f)
        return NC_EBADID;
    for (size_t g = 1; g < f->groups.size(); g++)
        if (f->groups[g] == name) {
            *grp = (ncid & ~255) | (int)g;
            return NC_NOERR;
        }
    return NC_ENOGRP;
}
int nc_inq_dimid(int ncid, const char* name, int* idp)
{
    std::lock_guard<std::mutex> lk(g_fs_mutex);
    MemFile* f = file_of(ncid);
    if (!f)
        return NC_EBADID;
    for (size_t d = 0; d < f->dims.size(); d++)
        if (f->dims[d].name == name) {
            *idp = (int)d;
            return NC_NOERR;
        }
    return NC_EBADDIM;
}
int nc_inq_dimlen(int ncid, int dimid, size_t* lenp)
{
    std::lock_guard<std::mutex> lk(g_fs_mutex);
    MemFile* f = file_of(ncid);
    if (!f || dimid < 0 || dimid >= (int)f->dims.size())
        return NC_EBADDIM;
    *lenp = f->dims[dimid].len;
    return NC_NOERR;
}
int nc_inq_dimname(int ncid, int dimid, char* name)
{
    std::lock_guard<std::mutex> lk(g_fs_mutex);
    MemFile* f = file_of(ncid);
    if (!f || dimid < 0 || dimid >= (int)f->dims.size())
        return NC_EBADDIM;
    std::strcpy(name, f->dims[dimid].name.c_str());
    return NC_NOERR;
}
int nc_inq_varid(int ncid, const char* name, int* varidp)
{
    std::lock_guard<std::mutex> lk(g_fs_mutex);
    MemFile* f = file_of(ncid);
    if (!f)
        return NC_EBADID;
    for (size_t v = 0; v < f->vars.size(); v++)
        if (f->vars[v].group == group_of(ncid) && f->vars[v].name == name) {
            *varidp = (int)v;
            return NC_NOERR;
        }
    return NC_ENOTVAR;
}
int nc_inq_vardimid(int ncid, int varid, int* dimidsp)
{
    std::lock_guard<std::mutex> lk(g_fs_mutex);
    MemFile* f = file_of(ncid);
    if (!f || varid < 0 || varid >= (int)f->vars.size())
        return NC_ENOTVAR;
    std::copy(f->vars[varid].dimids.begin(), f->vars[varid].dimids.end(), dimidsp);
    return NC_NOERR;
}
// row-major hyperslab walk shared by get / put
static int slab(MemFile* f, MemVar& v, const size_t* start, const size_t* count, int* out, const int* in)
{
    const size_t nd = v.dimids.size();
    size_t total = 1;
    for (size_t d = 0; d < nd; d++) {
        if (count[d] && start[d] + count[d] > f->dims[v.dimids[d]].len)
            return NC_EEDGE;
        total *= count[d];
    }
    std::vector<size_t> idx(nd, 0);
    for (size_t k = 0; k < total; k++) {
        size_t off = 0;
        for (size_t d = 0; d < nd; d++)
            off = off * f->dims[v.dimids[d]].len + start[d] + idx[d];
        if (out)
            out[k] = v.data[off];
        else {
            v.data[off] = in[k];
            v.written[off] = 1;
        }
        for (size_t d = nd; d-- > 0;) {
            if (++idx[d] < count[d])
                break;
            idx[d] = 0;
        }
    }
    return NC_NOERR;
}
int nc_get_vara_int(int ncid, int varid, const size_t* startp, const size_t* countp, int* ip)
{
    std::lock_guard<std::mutex> lk(g_fs_mutex);
    MemFile* f = file_of(ncid);
    if (!f || varid < 0 || varid >= (int)f->vars.size())
        return NC_ENOTVAR;
    return slab(f, f->vars[varid], startp, countp, ip, nullptr);
}
int nc_def_dim(int ncid, const char* name, size_t len, int* idp)
{
    std::lock_guard<std::mutex> lk(g_fs_mutex);
    MemFile* f = file_of(ncid);
    if (!f)
        return NC_EBADID;
    for (size_t d = 0; d < f->dims.size(); d++)
        if (f->dims[d].name == name) { // another rank defined it already (collective call)
            *idp = (int)d;
            return f->dims[d].len == len ? NC_NOERR : NC_EBADDIM;
        }
    f->dims.push_back({ name, len });
    *idp = (int)f->dims.size() - 1;
    return NC_NOERR;
}
int nc_def_grp(int parent, const char* name, int* new_ncid)
{
    std::lock_guard<std::mutex> lk(g_fs_mutex);
    MemFile* f = file_of(parent);
    if (!f)
        return NC_EBADID;
    size_t g = 1;
    for (; g < f->groups.size(); g++)
        if (f->groups[g] == name)
            break;
    if (g == f->groups.size())
        f->groups.push_back(name);
    *new_ncid = (parent & ~255) | (int)g;
    return NC_NOERR;
}
int nc_def_var(int ncid, const char* name, int, int ndims, const int* dimidsp, int* varidp)
{
    std::lock_guard<std::mutex> lk(g_fs_mutex);
    MemFile* f = file_of(ncid);
    if (!f)
        return NC_EBADID;
    for (size_t v = 0; v < f->vars.size(); v++)
        if (f->vars[v].group == group_of(ncid) && f->vars[v].name == name) {
            *varidp = (int)v;
            return NC_NOERR;
        }
    MemVar nv;
    nv.name = name;
    nv.group = group_of(ncid);
    nv.dimids.assign(dimidsp, dimidsp + ndims);
    f->vars.push_back(nv);
    *varidp = (int)f->vars.size() - 1;
    return NC_NOERR;
}
int nc_put_att_int(int ncid, int, const char* name, int, size_t, const int* op)
{
    std::lock_guard<std::mutex> lk(g_fs_mutex);
    MemFile* f = file_of(ncid);
    if (!f)
        return NC_EBADID;
    for (auto& a : f->atts)
        if (a.first == name) {
            a.second = *op;
            return NC_NOERR;
        }
    f->atts.push_back({ name, *op });
    return NC_NOERR;
}
int nc_put_vara_int(int ncid, int varid, const size_t* startp, const size_t* countp, const int* op)
{
    std::lock_guard<std::mutex> lk(g_fs_mutex);
    MemFile* f = file_of(ncid);
    if (!f || varid < 0 || varid >= (int)f->vars.size())
        return NC_ENOTVAR;
    ensure_storage(*f, f->vars[varid]);
    return slab(f, f->vars[varid], startp, countp, nullptr, op);
}
int nc_put_var1_int(int ncid, int varid, const size_t* indexp, const int* op)
{
    std::lock_guard<std::mutex> lk(g_fs_mutex);
    MemFile* f = file_of(ncid);
    if (!f || varid < 0 || varid >= (int)f->vars.size())
        return NC_ENOTVAR;
    MemVar& v = f->vars[varid];
    ensure_storage(*f, v);
    std::vector<size_t> one(v.dimids.size(), 1);
    return slab(f, v, indexp, one.data(), nullptr, op);
}
}


namespace {
void dump_file(std::ostringstream& os, const char* label, const std::string& path)
{
    const int fi = find_file(path);
    if (fi < 0)
        return;
    const MemFile& f = g_files[fi];
    os << "file " << label << "\n";
    for (const MemDim& d : f.dims)
        os << "dim " << d.name << " " << d.len << "\n";
    for (const auto& a : f.atts)
        os << "att " << a.first << " " << a.second << "\n";
    for (const MemVar& v : f.vars) {
        os << "var " << (f.groups[v.group].empty() ? "/" : f.groups[v.group]) << " " << v.name << " (";
        for (size_t k = 0; k < v.dimids.size(); k++)
            os << (k ? "," : "") << f.dims[v.dimids[k]].name;
        os << ")";
        size_t unwritten = 0;
        for (size_t k = 0; k < v.data.size(); k++) {
            os << " " << v.data[k];
            unwritten += v.written[k] ? 0 : 1;
        }
        os << "\n";
        if (unwritten || v.data.size() != var_size(f, v))
            os << "unwritten " << v.name << " " << unwritten + (var_size(f, v) - v.data.size()) << "\n";
    }
}
} // namespace
