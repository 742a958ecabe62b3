// h5lite — a self-contained writer for the subset of HDF5 the reference's output.h5 uses.
//
// The reference writes output.h5 through the HDF5 C++ API (src/simulator/report.cpp:9-158,
// src/Estimator.cpp:368-422,562-594); there is no libhdf5 in this image, so the
// host program carries its own writer.  It emits the classic, most widely readable layout —
// the one the reference's committed output files have: superblock v0, v1 object
// headers, symbol-table groups (B-tree v1 "TREE" + local heap "HEAP" + "SNOD" nodes),
// contiguous little-endian datasets (IEEE f64, u64, {r, i} compounds of f64), variable-length
// strings in one global heap collection ("GCOL"), v1 attribute messages.
//
// Usage: build the tree in memory, then write():
//   h5lite::File f;  auto& g = f.root.group("summary");  g.dataset_u64("Ncycle", 200);
//   g.dataset_f64("k", {n}, ptr);  g.attr_string("indexing", "[energy]");  f.write(path, err);
#ifndef MCB_H5LITE_H
#define MCB_H5LITE_H

#include <cstdint>
#include <deque>
#include <string>
#include <vector>

namespace h5lite {

enum class Type { F64, U64, VLEN_STRING, C128 };  // C128: complex numbers as the compound {r: f64, i: f64} (EigenHDF5 / h5py)

struct Attribute {
    std::string name;
    std::string value;  // variable-length string attribute, scalar (the only kind the reference writes)
};

struct Dataset {
    std::string name;
    Type type = Type::F64;
    std::vector<uint64_t> dims;        // empty = scalar
    std::vector<double> f64;           // C128: (re, im) pairs
    std::vector<uint64_t> u64;
    std::string str;                   // VLEN_STRING scalar
    std::vector<Attribute> attrs;
    void attr_string(const std::string& n, const std::string& v) { attrs.push_back({n, v}); }
    // filled by the writer
    uint64_t header_addr = 0, data_addr = 0;
};

struct Group {
    std::string name;
    std::deque<Group> groups;          // deque: references handed out stay valid
    std::deque<Dataset> datasets;
    std::vector<Attribute> attrs;

    Group& group(const std::string& n);
    Dataset& dataset_f64(const std::string& n, const std::vector<uint64_t>& dims, const double* data);
    Dataset& dataset_f64(const std::string& n, double scalar);
    Dataset& dataset_c128(const std::string& n, const std::vector<uint64_t>& dims, const double* re_im_pairs);
    Dataset& dataset_u64(const std::string& n, uint64_t scalar);
    Dataset& dataset_string(const std::string& n, const std::string& v);
    void attr_string(const std::string& n, const std::string& v) { attrs.push_back({n, v}); }
    // filled by the writer
    uint64_t header_addr = 0, btree_addr = 0, heap_addr = 0;
};

struct File {
    Group root;
    bool write(const std::string& path, std::string& error);
};

// Reader for what the TRMM post-processor needs (reference TRMM.cpp:20-24,48): one f64 dataset of the ROOT group of a
// classic-layout file (superblock v0/v1, v1 object headers, symbol-table groups, contiguous or compact layout) — the
// files h5lite writes and the ones the reference's libhdf5 wrote.  Anything else is an error, not a guess.
bool read_root_f64(const std::string& path, const std::string& name, std::vector<uint64_t>& dims, std::vector<double>& data,
                   std::string& error);

}  // namespace h5lite
#endif
