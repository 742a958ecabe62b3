// TEST INFRASTRUCTURE ONLY (oracle/): a text-dumping stand-in for the HDF5 C++
// API, used solely to compile the unmodified reference (ilhamv/MC-old) into
// oracle/_ref/ in an image that has no libhdf5.  It provides exactly the call
// shapes the reference uses in src/simulator/report.cpp:12-157 and
// src/Estimator.cpp:368-422,562-594.  Every dataset/attribute is written as one
// text line so tests can parse the reference's results:
//
//   D <path> <f64|u64|str> <rank> <dim0> ... : <v0> <v1> ...      (doubles %.17g)
//   A <path>@<name> str : <value>
//   G <path>
//
// This is NOT the product's HDF5 writer (see mc_old_b200/host/h5_writer.*).
#ifndef ORACLE_H5STUB_H5CPP_H
#define ORACLE_H5STUB_H5CPP_H

#include <cstdio>
#include <memory>
#include <string>
#include <vector>

typedef unsigned long long hsize_t;
typedef std::string H5std_string;
enum { H5F_ACC_TRUNC = 2 };
enum H5S_class_t { H5S_SCALAR = 0, H5S_SIMPLE = 1 };
#define H5T_VARIABLE ((size_t)(-1))

namespace H5 {

struct Sink {
    FILE* f;
    explicit Sink(const std::string& name) { f = std::fopen(name.c_str(), "w"); }
    ~Sink() { if (f) std::fclose(f); }
};

class DataType {
  public:
    int kind;  // 0 = f64, 1 = u64, 2 = string
    DataType(int k = 0) : kind(k) {}
};
class PredType : public DataType {
  public:
    PredType(int k) : DataType(k) {}
    static const PredType NATIVE_DOUBLE;
    static const PredType NATIVE_ULLONG;
};
class StrType : public DataType {
  public:
    StrType(int, size_t) : DataType(2) {}
};

class DataSpace {
  public:
    std::vector<hsize_t> dims;
    DataSpace() {}
    DataSpace(H5S_class_t) {}
    DataSpace(int rank, const hsize_t* d) : dims(d, d + rank) {}
    size_t count() const { size_t n = 1; for (auto d : dims) n *= d; return n; }
};

class Attribute {
  public:
    std::shared_ptr<Sink> sink; std::string path;
    void write(const StrType&, const std::string& v) {
        std::fprintf(sink->f, "A %s str : %s\n", path.c_str(), v.c_str());
    }
};

class DataSet {
  public:
    std::shared_ptr<Sink> sink; std::string path; DataSpace space;
    void header(const char* t) {
        std::fprintf(sink->f, "D %s %s %zu", path.c_str(), t, space.dims.size());
        for (auto d : space.dims) std::fprintf(sink->f, " %llu", d);
        std::fprintf(sink->f, " :");
    }
    void write(const void* buf, const DataType& t) {
        const size_t n = space.count();
        if (t.kind == 0) {
            header("f64");
            const double* p = static_cast<const double*>(buf);
            for (size_t i = 0; i < n; i++) std::fprintf(sink->f, " %.17g", p[i]);
        } else {
            header("u64");
            const unsigned long long* p = static_cast<const unsigned long long*>(buf);
            for (size_t i = 0; i < n; i++) std::fprintf(sink->f, " %llu", p[i]);
        }
        std::fprintf(sink->f, "\n");
    }
    void write(const std::string& v, const StrType&) {
        header("str");
        std::fprintf(sink->f, " %s\n", v.c_str());
    }
    Attribute createAttribute(const std::string& n, const StrType&, const DataSpace&) {
        Attribute a; a.sink = sink; a.path = path + "@" + n; return a;
    }
};

class Group {
  public:
    std::shared_ptr<Sink> sink; std::string path;
    static std::string join(const std::string& base, const std::string& n) {
        if (!n.empty() && n[0] == '/') return n;
        if (base.empty() || base == "/") return "/" + n;
        return base + "/" + n;
    }
    Group createGroup(const std::string& n) {
        Group g; g.sink = sink; g.path = join(path, n);
        std::fprintf(sink->f, "G %s\n", g.path.c_str());
        return g;
    }
    DataSet createDataSet(const std::string& n, const DataType&, const DataSpace& s) {
        DataSet d; d.sink = sink; d.path = join(path, n); d.space = s; return d;
    }
    Attribute createAttribute(const std::string& n, const StrType&, const DataSpace&) {
        Attribute a; a.sink = sink; a.path = path + "@" + n; return a;
    }
};

class H5File : public Group {
  public:
    H5File(const std::string& name, unsigned) { sink = std::make_shared<Sink>(name); path = "/"; }
};

}  // namespace H5

#endif
