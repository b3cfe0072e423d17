/*
 * oracle/shim/H5Cpp.h -- TEST INFRASTRUCTURE.  A recorder with the shape of the HDF5 C++ API slice that the
 * reference's bound_geom::save_field_times uses (src/disp.cpp:758-923).  libhdf5 is absent from this image; this
 * header (our code) lets the reference's own writer run unmodified and records what it WOULD have put in the file:
 * on H5File::close() it writes <file>.manifest.json (every group and dataset in creation order: path, datatype
 * -- for compound types the member names, offsets and sizes the reference passed to H5Tinsert --, dimensions,
 * byte offset) and <file>.manifest.bin (the bytes handed to DataSet::write).  tests/test_ref_shim.py compares
 * that record with the file sim_juncs_b200/hdf5.py writes.
 */
#ifndef SJ_ORACLE_SHIM_H5CPP_H
#define SJ_ORACLE_SHIM_H5CPP_H

#include <cstddef>
#include <cstdio>
#include <string>
#include <vector>

typedef long long hid_t;
typedef int herr_t;
typedef unsigned long long hsize_t;
#define H5F_ACC_TRUNC 2u
#define HOFFSET(S, M) (offsetof(S, M))

namespace H5 {
struct shim_member { std::string name; size_t offset; hid_t type; };
struct shim_type { std::string kind; size_t size; std::vector<shim_member> members; };
shim_type &shim_type_of(hid_t id);
hid_t shim_new_compound(size_t size);

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
    CompType(size_t size) : DataType(shim_new_compound(size)) {}
};
class DataSpace {
public:
    DataSpace() {}
    DataSpace(int rank, const hsize_t *d) : dims(d, d + rank) {}
    std::vector<hsize_t> dims;
};
struct shim_file;
class DataSet {
public:
    DataSet() : f(0), index(0) {}
    void write(const void *buf, const DataType &mem_type);
    shim_file *f;
    size_t index;
};
class Group {
public:
    Group() : f(0) {}
    Group createGroup(const char *name);
    Group createGroup(const std::string &name) { return createGroup(name.c_str()); }
    DataSet createDataSet(const char *name, const DataType &type, const DataSpace &space);
    DataSet createDataSet(const std::string &name, const DataType &type, const DataSpace &space) { return createDataSet(name.c_str(), type, space); }
    shim_file *f;
    std::string path;
};
class H5File : public Group {
public:
    H5File(const char *name, unsigned flags);
    ~H5File();
    void close();
};
}  // namespace H5

herr_t H5Tinsert(hid_t compound, const char *name, size_t offset, hid_t member);

#endif
