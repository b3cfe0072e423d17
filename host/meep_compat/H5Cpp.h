// host/meep_compat/H5Cpp.h -- the slice of the HDF5 C++ API that bound_geom::save_field_times uses
// (reference src/disp.cpp:758-923), on this repository's own header-only HDF5 encoder (host/sj_hdf5.hpp).  Only for
// machines without libhdf5 (this image has none): where the real library exists, the reference keeps using it and this
// header is not on the include path.  Writes a real HDF5 file (superblock v0, symbol-table groups, contiguous datasets,
// f64 / u64 / compound-of-f64 types); NATIVE_FLOAT (STO_PREC_32 builds) is not supported.
#ifndef SJ_MEEP_COMPAT_H5CPP_H
#define SJ_MEEP_COMPAT_H5CPP_H

#include <cstddef>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../sj_hdf5.hpp"

typedef long long hid_t;
typedef int herr_t;
typedef unsigned long long hsize_t;
#define H5F_ACC_TRUNC 2u
#define HOFFSET(S, M) (offsetof(S, M))

namespace H5 {
struct compat_type { int kind; size_t size; std::vector<std::pair<std::string, unsigned> > members; };   // kind 1 f64, 3 u64, 4 compound
inline std::vector<compat_type> &compat_types() {
    static std::vector<compat_type> t;
    if (t.empty()) {
        t.resize(4);
        t[1].kind = 1; t[1].size = 8; t[2].kind = 2; t[2].size = 4; t[3].kind = 3; t[3].size = 8;
    }
    return t;
}
class DataType {
public:
    constexpr DataType(hid_t id = 0) : id(id) {}
    hid_t getId() const { return id; }
protected:
    hid_t id;
};
class PredType : public DataType {
public:
    constexpr PredType(hid_t id) : DataType(id) {}
    static const PredType NATIVE_DOUBLE, NATIVE_FLOAT, NATIVE_HSIZE;
};
class CompType : public DataType {
public:
    CompType(size_t size) : DataType(0) {
        compat_type c; c.kind = 4; c.size = size;
        compat_types().push_back(c);
        id = (hid_t)compat_types().size() - 1;
    }
};
class DataSpace {
public:
    DataSpace() {}
    DataSpace(int rank, const hsize_t *d) : dims(d, d + rank) {}
    std::vector<hsize_t> dims;
};
struct compat_file { std::string name; sj_h5::Writer w; bool open; };
class DataSet {
public:
    DataSet() : f(0), count(0) {}
    void write(const void *buf, const DataType &mem_type) {
        const compat_type &t = compat_types().at((size_t)mem_type.getId());
        sj_h5::bytes dt;
        if (t.kind == 1) dt = sj_h5::f64_type();
        else if (t.kind == 3) dt = sj_h5::u64_type();
        else if (t.kind == 4) dt = sj_h5::compound_type(t.members, (unsigned)t.size);
        else throw std::runtime_error("H5Cpp compat: unsupported datatype");
        f->w.dataset(path, dt, buf, count, t.size);
    }
    compat_file *f;
    std::string path;
    unsigned long long count;
};
class Group {
public:
    Group() : f(0) {}
    Group createGroup(const char *name) {
        if (!name) throw std::runtime_error("H5::Group::createGroup: NULL name");
        Group g; g.f = f; g.path = path.empty() ? std::string(name) : path + "/" + name;
        f->w.group(g.path);
        return g;
    }
    Group createGroup(const std::string &name) { return createGroup(name.c_str()); }
    DataSet createDataSet(const char *name, const DataType &, const DataSpace &space) {
        if (!name) throw std::runtime_error("H5::Group::createDataSet: NULL name");     // libhdf5 rejects it, H5Cpp throws
        DataSet d; d.f = f; d.path = path.empty() ? std::string(name) : path + "/" + name;
        d.count = 1;
        for (size_t i = 0; i < space.dims.size(); ++i) d.count *= space.dims[i];
        return d;
    }
    DataSet createDataSet(const std::string &name, const DataType &type, const DataSpace &space) { return createDataSet(name.c_str(), type, space); }
    compat_file *f;
    std::string path;
};
class H5File : public Group {
public:
    H5File(const char *name, unsigned) { f = new compat_file; f->name = name; f->open = true; }
    ~H5File() { close(); delete f; f = 0; }
    void close() {
        if (!f || !f->open) return;
        f->open = false;
        if (f->w.save(f->name.c_str())) throw std::runtime_error("H5Cpp compat: cannot write " + f->name);
    }
};
}  // namespace H5

inline herr_t H5Tinsert(hid_t compound, const char *name, size_t offset, hid_t) {
    H5::compat_type &t = H5::compat_types().at((size_t)compound);
    if (t.kind != 4) return -1;
    t.members.push_back(std::make_pair(std::string(name), (unsigned)offset));
    return 0;
}

#endif
