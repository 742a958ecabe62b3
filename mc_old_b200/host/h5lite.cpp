// h5lite.cpp — see h5lite.h.  Format references are to the HDF5 File Format Specification 1.x ("classic" objects).
#include "h5lite.h"

#include <algorithm>
#include <cstdio>
#include <cstring>

namespace h5lite {

Group& Group::group(const std::string& n)
{
    for (auto& g : groups) if (g.name == n) return g;
    groups.emplace_back();
    groups.back().name = n;
    return groups.back();
}
Dataset& Group::dataset_f64(const std::string& n, const std::vector<uint64_t>& dims, const double* data)
{
    datasets.emplace_back();
    Dataset& d = datasets.back();
    d.name = n; d.type = Type::F64; d.dims = dims;
    uint64_t count = 1;
    for (uint64_t v : dims) count *= v;
    d.f64.assign(data, data + count);
    return d;
}
Dataset& Group::dataset_f64(const std::string& n, double scalar) { return dataset_f64(n, {}, &scalar); }
Dataset& Group::dataset_c128(const std::string& n, const std::vector<uint64_t>& dims, const double* re_im_pairs)
{
    datasets.emplace_back();
    Dataset& d = datasets.back();
    d.name = n; d.type = Type::C128; d.dims = dims;
    uint64_t count = 2;
    for (uint64_t v : dims) count *= v;
    d.f64.assign(re_im_pairs, re_im_pairs + count);
    return d;
}
Dataset& Group::dataset_u64(const std::string& n, uint64_t scalar)
{
    datasets.emplace_back();
    Dataset& d = datasets.back();
    d.name = n; d.type = Type::U64; d.u64.assign(1, scalar);
    return d;
}
Dataset& Group::dataset_string(const std::string& n, const std::string& v)
{
    datasets.emplace_back();
    Dataset& d = datasets.back();
    d.name = n; d.type = Type::VLEN_STRING; d.str = v;
    return d;
}

namespace {

constexpr uint64_t UNDEF = 0xffffffffffffffffull;
constexpr int LEAF_K = 4, INTERNAL_K = 16;           // superblock defaults of the library
constexpr uint64_t SNOD_SIZE = 8 + 2 * LEAF_K * 40;  // 328
constexpr uint64_t TREE_SIZE = 24 + (2 * INTERNAL_K + 1) * 8 + 2 * INTERNAL_K * 8;  // 544
inline uint64_t align8(uint64_t v) { return (v + 7) & ~7ull; }

struct Image {
    std::vector<uint8_t> b;
    uint64_t alloc(uint64_t size)
    {
        const uint64_t at = align8(b.size());
        b.resize(at + size, 0);
        return at;
    }
    void put(uint64_t at, const void* p, size_t n) { memcpy(b.data() + at, p, n); }
    void u8(uint64_t at, uint8_t v) { b[at] = v; }
    void u16(uint64_t at, uint16_t v) { put(at, &v, 2); }
    void u32(uint64_t at, uint32_t v) { put(at, &v, 4); }
    void u64(uint64_t at, uint64_t v) { put(at, &v, 8); }
};

// one global heap collection holds every variable-length string of the file
struct GlobalHeap {
    uint64_t addr = 0, size = 0;
    std::vector<std::string> objects;  // index i+1
    uint32_t add(const std::string& s) { objects.push_back(s); return (uint32_t)objects.size(); }
    uint64_t needed() const
    {
        uint64_t n = 16;
        for (auto& s : objects) n += 16 + align8(s.size());
        return std::max<uint64_t>(4096, align8(n + 16));
    }
};

struct Writer {
    Image img;
    GlobalHeap gh;
    std::string too_wide;  // first group with more members than one B-tree node addresses

    // ---- datatype messages (spec IV.A.2.d) ----
    static std::vector<uint8_t> datatype(Type t)
    {
        std::vector<uint8_t> m;
        auto p32 = [&](uint32_t v) { for (int i = 0; i < 4; i++) m.push_back((uint8_t)(v >> (8 * i))); };
        auto p16 = [&](uint16_t v) { m.push_back((uint8_t)v); m.push_back((uint8_t)(v >> 8)); };
        switch (t) {
        case Type::F64:  // class 1 (floating point) version 1; little-endian, implied-msb mantissa, sign bit 63
            m = {0x11, 0x20, 0x3f, 0x00};
            p32(8);
            p16(0); p16(64); m.push_back(52); m.push_back(11); m.push_back(0); m.push_back(52); p32(1023);
            break;
        case Type::U64:  // class 0 (fixed point) version 1; little-endian, unsigned
            m = {0x10, 0x00, 0x00, 0x00};
            p32(8);
            p16(0); p16(64);
            break;
        case Type::C128: {  // class 6 (compound) version 1, two members: "r" at 0, "i" at 8, both IEEE f64 (TRMM.cpp:75-78
                            // through EigenHDF5::save of complex matrices); the layout libhdf5 1.10 gives it
            m = {0x16, 0x02, 0x00, 0x00};
            p32(16);
            const std::vector<uint8_t> f64 = datatype(Type::F64);
            for (int k = 0; k < 2; k++) {
                m.push_back(k == 0 ? 'r' : 'i');
                for (int i = 0; i < 7; i++) m.push_back(0);   // name, null-terminated, padded to 8 bytes
                p32(8 * k);                                    // byte offset of the member
                for (int i = 0; i < 28; i++) m.push_back(0);  // dimensionality 0, reserved, permutation, reserved, 4 sizes
                m.insert(m.end(), f64.begin(), f64.end());
            }
            break;
        }
        case Type::VLEN_STRING:  // class 9 version 1; type = string (1), null-terminated, ASCII; base = 1-byte C string
            m = {0x19, 0x01, 0x00, 0x00};
            p32(16);
            m.insert(m.end(), {0x13, 0x00, 0x00, 0x00});
            p32(1);
            break;
        }
        return m;
    }
    static std::vector<uint8_t> dataspace(const std::vector<uint64_t>& dims)
    {  // version 1: version, rank, flags, reserved, reserved(4), dims
        std::vector<uint8_t> m = {1, (uint8_t)dims.size(), 0, 0, 0, 0, 0, 0};
        for (uint64_t d : dims) for (int i = 0; i < 8; i++) m.push_back((uint8_t)(d >> (8 * i)));
        return m;
    }
    // element of a variable-length string in the file: length, collection address, object index
    std::vector<uint8_t> vlen_ref(const std::string& s)
    {
        const uint32_t idx = gh.add(s);
        std::vector<uint8_t> m(16, 0);
        const uint32_t len = (uint32_t)s.size();
        memcpy(m.data(), &len, 4);
        memcpy(m.data() + 4, &gh.addr, 8);
        memcpy(m.data() + 12, &idx, 4);
        return m;
    }
    // attribute message, version 1 (spec IV.A.2.m)
    std::vector<uint8_t> attribute(const Attribute& a)
    {
        const std::vector<uint8_t> dt = datatype(Type::VLEN_STRING), ds = dataspace({});
        const std::vector<uint8_t> data = vlen_ref(a.value);
        std::vector<uint8_t> m(8, 0);
        m[0] = 1;
        const uint16_t ns = (uint16_t)(a.name.size() + 1), dts = (uint16_t)dt.size(), dss = (uint16_t)ds.size();
        memcpy(m.data() + 2, &ns, 2); memcpy(m.data() + 4, &dts, 2); memcpy(m.data() + 6, &dss, 2);
        auto append_padded = [&](const uint8_t* p, size_t n) { m.insert(m.end(), p, p + n); m.resize(align8(m.size()), 0); };
        append_padded((const uint8_t*)a.name.c_str(), a.name.size() + 1);
        append_padded(dt.data(), dt.size());
        append_padded(ds.data(), ds.size());
        m.insert(m.end(), data.begin(), data.end());
        return m;
    }

    // version-1 object header: prefix (16 bytes) + messages (8-byte header each, data padded to 8)
    uint64_t object_header(const std::vector<std::pair<uint16_t, std::vector<uint8_t>>>& msgs)
    {
        uint64_t body = 0;
        for (auto& m : msgs) body += 8 + align8(m.second.size());
        const uint64_t at = img.alloc(16 + body);
        img.u8(at, 1);
        img.u16(at + 2, (uint16_t)msgs.size());
        img.u32(at + 4, 1);               // object reference count
        img.u32(at + 8, (uint32_t)body);  // header data size
        uint64_t p = at + 16;
        for (auto& m : msgs) {
            const uint64_t sz = align8(m.second.size());
            img.u16(p, m.first);
            img.u16(p + 2, (uint16_t)sz);
            img.u8(p + 4, 0);
            if (!m.second.empty()) img.put(p + 8, m.second.data(), m.second.size());
            p += 8 + sz;
        }
        return at;
    }

    void emit_dataset(Dataset& d)
    {
        uint64_t count = 1;
        for (uint64_t v : d.dims) count *= v;
        std::vector<uint8_t> raw;
        if (d.type == Type::F64) { raw.resize(count * 8); if (count) memcpy(raw.data(), d.f64.data(), count * 8); }
        else if (d.type == Type::C128) { raw.resize(count * 16); if (count) memcpy(raw.data(), d.f64.data(), count * 16); }
        else if (d.type == Type::U64) { raw.resize(count * 8); if (count) memcpy(raw.data(), d.u64.data(), count * 8); }
        else raw = vlen_ref(d.str);
        d.data_addr = raw.empty() ? UNDEF : img.alloc(raw.size());
        if (!raw.empty()) img.put(d.data_addr, raw.data(), raw.size());
        std::vector<std::pair<uint16_t, std::vector<uint8_t>>> msgs;
        msgs.push_back({0x0001, dataspace(d.dims)});
        msgs.push_back({0x0003, datatype(d.type)});
        msgs.push_back({0x0005, {2, 1, 2, 1, 0, 0, 0, 0}});  // fill value v2: early allocation, write if set, default value
        std::vector<uint8_t> layout(18, 0);                   // data layout v3, contiguous
        layout[0] = 3; layout[1] = 1;
        memcpy(layout.data() + 2, &d.data_addr, 8);
        const uint64_t nbytes = raw.size();
        memcpy(layout.data() + 10, &nbytes, 8);
        msgs.push_back({0x0008, layout});
        for (auto& a : d.attrs) msgs.push_back({0x000c, attribute(a)});
        d.header_addr = object_header(msgs);
    }

    void emit_group(Group& g)
    {
        struct Child { std::string name; bool is_group; size_t index; };
        std::vector<Child> kids;
        for (size_t i = 0; i < g.groups.size(); i++) kids.push_back({g.groups[i].name, true, i});
        for (size_t i = 0; i < g.datasets.size(); i++) kids.push_back({g.datasets[i].name, false, i});
        std::sort(kids.begin(), kids.end(), [](const Child& a, const Child& b) { return strcmp(a.name.c_str(), b.name.c_str()) < 0; });
        const size_t n = kids.size();
        const size_t n_snod = (n + 2 * LEAF_K - 1) / (2 * LEAF_K);
        // one leaf-level B-tree node holds 2 * INTERNAL_K symbol nodes = 256 links; a larger group needs an internal
        // level, which this writer does not emit: refuse instead of writing past the node
        if (n_snod > (size_t)(2 * INTERNAL_K)) {
            if (too_wide.empty()) too_wide = g.name.empty() ? "/" : g.name;
            return;
        }
        // local heap data segment: "" at offset 0, then the names
        std::vector<uint64_t> name_off(n);
        uint64_t heap_data = 8;
        for (size_t i = 0; i < n; i++) { name_off[i] = heap_data; heap_data += align8(kids[i].name.size() + 1); }
        // allocate this group's own structures, then the children (their addresses go into the symbol nodes)
        std::vector<std::pair<uint16_t, std::vector<uint8_t>>> msgs;
        msgs.push_back({0x0011, std::vector<uint8_t>(16, 0)});  // symbol table message, patched below
        for (auto& a : g.attrs) msgs.push_back({0x000c, attribute(a)});
        g.header_addr = object_header(msgs);
        g.btree_addr = img.alloc(TREE_SIZE);
        g.heap_addr = img.alloc(32 + heap_data);
        std::vector<uint64_t> snod(n_snod);
        for (size_t i = 0; i < n_snod; i++) snod[i] = img.alloc(SNOD_SIZE);
        img.u64(g.header_addr + 16 + 8, g.btree_addr);
        img.u64(g.header_addr + 16 + 16, g.heap_addr);
        for (auto& c : g.groups) emit_group(c);
        for (auto& d : g.datasets) emit_dataset(d);
        // local heap (spec III.D): signature, version, data segment size, free-list head (1 = none), data segment address
        img.put(g.heap_addr, "HEAP", 4);
        img.u64(g.heap_addr + 8, heap_data);
        img.u64(g.heap_addr + 16, 1);
        img.u64(g.heap_addr + 24, g.heap_addr + 32);
        for (size_t i = 0; i < n; i++) img.put(g.heap_addr + 32 + name_off[i], kids[i].name.c_str(), kids[i].name.size() + 1);
        // B-tree v1 node of a group (spec III.A.1): one leaf level, child i = symbol node i, key i+1 = last name in it
        img.put(g.btree_addr, "TREE", 4);
        img.u8(g.btree_addr + 4, 0);
        img.u8(g.btree_addr + 5, 0);
        img.u16(g.btree_addr + 6, (uint16_t)n_snod);
        img.u64(g.btree_addr + 8, UNDEF);
        img.u64(g.btree_addr + 16, UNDEF);
        img.u64(g.btree_addr + 24, 0);  // key 0: the empty string
        for (size_t s = 0; s < n_snod; s++) {
            const size_t first = s * 2 * LEAF_K, last = std::min(n, first + 2 * LEAF_K) - 1;
            img.u64(g.btree_addr + 24 + 8 + 16 * s, snod[s]);
            img.u64(g.btree_addr + 24 + 16 + 16 * s, name_off[last]);
            // symbol table node (spec III.C)
            img.put(snod[s], "SNOD", 4);
            img.u8(snod[s] + 4, 1);
            img.u16(snod[s] + 6, (uint16_t)(last - first + 1));
            for (size_t i = first; i <= last; i++) {
                const uint64_t e = snod[s] + 8 + 40 * (i - first);
                img.u64(e, name_off[i]);
                if (kids[i].is_group) {
                    const Group& c = g.groups[kids[i].index];
                    img.u64(e + 8, c.header_addr);
                    img.u32(e + 16, 1);  // cached: B-tree and heap addresses
                    img.u64(e + 24, c.btree_addr);
                    img.u64(e + 32, c.heap_addr);
                } else {
                    img.u64(e + 8, g.datasets[kids[i].index].header_addr);
                }
            }
        }
    }

    static void count_strings(const Group& g, GlobalHeap& h)
    {
        for (auto& a : g.attrs) h.add(a.value);
        for (auto& d : g.datasets) { if (d.type == Type::VLEN_STRING) h.add(d.str); for (auto& a : d.attrs) h.add(a.value); }
        for (auto& c : g.groups) count_strings(c, h);
    }

    void run(File& f)
    {
        img.alloc(96);  // superblock, filled last
        GlobalHeap probe;
        count_strings(f.root, probe);
        gh.size = probe.needed();
        gh.addr = img.alloc(gh.size);
        emit_group(f.root);
        // global heap collection (spec III.E)
        img.put(gh.addr, "GCOL", 4);
        img.u8(gh.addr + 4, 1);
        img.u64(gh.addr + 8, gh.size);
        uint64_t p = gh.addr + 16;
        for (size_t i = 0; i < gh.objects.size(); i++) {
            const std::string& s = gh.objects[i];
            img.u16(p, (uint16_t)(i + 1));
            img.u16(p + 2, 1);  // reference count
            img.u64(p + 8, s.size());
            if (!s.empty()) img.put(p + 16, s.data(), s.size());
            p += 16 + align8(s.size());
        }
        const uint64_t left = gh.addr + gh.size - p;  // object 0 = the free space, its size includes its own header
        if (left >= 16) img.u64(p + 8, left);
        // superblock version 0 (spec II.A)
        const uint64_t eof = align8(img.b.size());
        img.b.resize(eof, 0);
        static const uint8_t sig[8] = {0x89, 'H', 'D', 'F', '\r', '\n', 0x1a, '\n'};
        img.put(0, sig, 8);
        img.u8(13, 8); img.u8(14, 8);  // size of offsets, size of lengths
        img.u16(16, LEAF_K); img.u16(18, INTERNAL_K);
        img.u64(24, 0);      // base address
        img.u64(32, UNDEF);  // free-space info
        img.u64(40, eof);
        img.u64(48, UNDEF);  // driver info
        // root group symbol table entry
        img.u64(56, 0);
        img.u64(64, f.root.header_addr);
        img.u32(72, 1);
        img.u64(80, f.root.btree_addr);
        img.u64(88, f.root.heap_addr);
    }
};

}  // namespace

bool File::write(const std::string& path, std::string& error)
{
    Writer w;
    w.run(*this);
    if (!w.too_wide.empty()) { error = "group " + w.too_wide + " has more than 256 members (not supported by the built-in HDF5 writer)"; return false; }
    FILE* fp = fopen(path.c_str(), "wb");
    if (!fp) { error = "cannot open " + path + " for writing"; return false; }
    const size_t n = fwrite(w.img.b.data(), 1, w.img.b.size(), fp);
    const bool ok = n == w.img.b.size() && fclose(fp) == 0;
    if (!ok) error = "short write to " + path;
    return ok;
}

// ---------------------------------------------------------------------------------------------
// reader (see h5lite.h)
// ---------------------------------------------------------------------------------------------
namespace {

struct Reader {
    std::vector<uint8_t> b;
    std::string err;

    bool fail(const std::string& m) { if (err.empty()) err = m; return false; }
    bool in(uint64_t at, uint64_t n) const { return at <= b.size() && n <= b.size() - at; }
    uint64_t u(uint64_t at, int n)
    {
        if (!in(at, (uint64_t)n)) { fail("read past the end of the file"); return 0; }
        uint64_t v = 0;
        for (int i = 0; i < n; i++) v |= (uint64_t)b[at + i] << (8 * i);
        return v;
    }
    bool magic(uint64_t at, const char* m) { return in(at, 4) && !memcmp(b.data() + at, m, 4); }

    struct Msg { uint16_t type; uint64_t at, size; };
    // the messages of a version-1 object header, continuation blocks included
    bool messages(uint64_t addr, std::vector<Msg>& out)
    {
        if (u(addr, 1) != 1) return fail("object header version " + std::to_string(u(addr, 1)) + " (only version 1 is read)");
        const uint64_t n_msg = u(addr + 2, 2);
        std::vector<std::pair<uint64_t, uint64_t>> blocks{{addr + 16, u(addr + 8, 4)}};
        for (size_t k = 0; k < blocks.size() && err.empty(); k++) {
            uint64_t p = blocks[k].first;
            const uint64_t end = p + blocks[k].second;
            if (!in(p, blocks[k].second)) return fail("object header block outside the file");
            while (p + 8 <= end && out.size() < n_msg) {
                const Msg m{(uint16_t)u(p, 2), p + 8, u(p + 2, 2)};
                if (m.at + m.size > end) return fail("object header message runs past its block");
                if (m.type == 0x0010) blocks.push_back({u(m.at, 8), u(m.at + 8, 8)});
                out.push_back(m);
                p += 8 + m.size;
            }
        }
        return err.empty();
    }
    // object header address of `name` in the group whose B-tree / local heap are given; 0 if absent
    uint64_t lookup(uint64_t btree, uint64_t heap, const std::string& name, int depth = 0)
    {
        if (depth > 16 || !magic(btree, "TREE") || u(btree + 4, 1) != 0) { fail("bad group B-tree node"); return 0; }
        if (!magic(heap, "HEAP")) { fail("bad local heap"); return 0; }
        const uint64_t level = u(btree + 5, 1), used = u(btree + 6, 2);
        const uint64_t heap_size = u(heap + 8, 8), heap_data = u(heap + 24, 8);
        for (uint64_t i = 0; i < used && err.empty(); i++) {
            const uint64_t child = u(btree + 24 + 8 + 16 * i, 8);
            if (level > 0) { const uint64_t h = lookup(child, heap, name, depth + 1); if (h) return h; continue; }
            if (!magic(child, "SNOD")) { fail("bad symbol table node"); return 0; }
            const uint64_t n_sym = u(child + 6, 2);
            for (uint64_t s = 0; s < n_sym; s++) {
                const uint64_t name_off = u(child + 8 + 40 * s, 8), header = u(child + 8 + 40 * s + 8, 8);
                if (name_off >= heap_size || !in(heap_data + name_off, 1)) { fail("symbol name outside the heap"); return 0; }
                const char* c = reinterpret_cast<const char*>(b.data() + heap_data + name_off);
                const size_t max_len = (size_t)std::min<uint64_t>(heap_size - name_off, b.size() - (heap_data + name_off));
                if (strnlen(c, max_len) == name.size() && !memcmp(c, name.data(), name.size())) return header;
            }
        }
        return 0;
    }
};

}  // namespace

bool read_root_f64(const std::string& path, const std::string& name, std::vector<uint64_t>& dims, std::vector<double>& data,
                   std::string& error)
{
    Reader r;
    FILE* fp = fopen(path.c_str(), "rb");
    if (!fp) { error = "cannot open " + path; return false; }
    uint8_t buf[1 << 16];
    size_t n;
    while ((n = fread(buf, 1, sizeof(buf), fp)) > 0) r.b.insert(r.b.end(), buf, buf + n);
    fclose(fp);
    auto bad = [&](const std::string& m) { error = path + ": " + (r.err.empty() ? m : r.err); return false; };
    static const uint8_t SIG[8] = {0x89, 'H', 'D', 'F', '\r', '\n', 0x1a, '\n'};
    if (r.b.size() < 96 || memcmp(r.b.data(), SIG, 8)) return bad("not an HDF5 file");
    const uint64_t ver = r.b[8];
    if (ver > 1) return bad("superblock version " + std::to_string(ver) + " (only the classic versions 0 and 1 are read)");
    if (r.b[13] != 8 || r.b[14] != 8) return bad("offsets / lengths are not 8 bytes");
    const uint64_t root_entry = (ver == 0 ? 24 : 28) + 32;  // the root group's symbol table entry
    const uint64_t root_header = r.u(root_entry + 8, 8);
    uint64_t btree = 0, heap = 0;
    if (r.u(root_entry + 16, 4) == 1) { btree = r.u(root_entry + 24, 8); heap = r.u(root_entry + 32, 8); }
    else {
        std::vector<Reader::Msg> msgs;
        if (!r.messages(root_header, msgs)) return bad("");
        for (const auto& m : msgs) if (m.type == 0x0011) { btree = r.u(m.at, 8); heap = r.u(m.at + 8, 8); }
        if (!btree) return bad("the root group has no symbol table");
    }
    const uint64_t header = r.lookup(btree, heap, name);
    if (!r.err.empty()) return bad("");
    if (!header) return bad("no dataset \"" + name + "\" in the root group");
    std::vector<Reader::Msg> msgs;
    if (!r.messages(header, msgs)) return bad("");
    bool have_space = false, have_type = false, have_layout = false;
    uint64_t addr = 0, nbytes = 0;
    dims.clear();
    for (const auto& m : msgs) {
        if (m.type == 0x0001) {  // dataspace, version 1 or 2
            const uint64_t v = r.u(m.at, 1), rank = r.u(m.at + 1, 1);
            if (v != 1 && v != 2) return bad("dataspace version " + std::to_string(v));
            if (v == 2 && r.u(m.at + 3, 1) == 2) return bad("dataset \"" + name + "\" has a null dataspace");
            const uint64_t p = m.at + (v == 1 ? 8 : 4);
            for (uint64_t i = 0; i < rank; i++) dims.push_back(r.u(p + 8 * i, 8));
            have_space = true;
        } else if (m.type == 0x0003) {  // datatype: little-endian IEEE f64 and nothing else
            const uint64_t cls = r.u(m.at, 1) & 0x0f, bits0 = r.u(m.at + 1, 1), size = r.u(m.at + 4, 4);
            if (cls != 1 || (bits0 & 1) || size != 8 || r.u(m.at + 10, 2) != 64 || r.u(m.at + 12, 1) != 52 || r.u(m.at + 13, 1) != 11 ||
                r.u(m.at + 15, 1) != 52 || r.u(m.at + 16, 4) != 1023)
                return bad("dataset \"" + name + "\" is not little-endian IEEE f64");
            have_type = true;
        } else if (m.type == 0x0008) {  // data layout version 3: compact (0) or contiguous (1)
            if (r.u(m.at, 1) != 3) return bad("data layout version " + std::to_string(r.u(m.at, 1)));
            const uint64_t cls = r.u(m.at + 1, 1);
            if (cls == 0) { nbytes = r.u(m.at + 2, 2); addr = m.at + 4; }
            else if (cls == 1) { addr = r.u(m.at + 2, 8); nbytes = r.u(m.at + 10, 8); }
            else return bad("dataset \"" + name + "\" is chunked (not read)");
            have_layout = true;
        } else if (m.type == 0x000B) return bad("dataset \"" + name + "\" is filtered (not read)");
    }
    if (!r.err.empty()) return bad("");
    if (!have_space || !have_type || !have_layout) return bad("dataset \"" + name + "\" lacks a dataspace, datatype or layout message");
    uint64_t count = 1;
    for (uint64_t d : dims) { if (d && count > (1ull << 40) / d) return bad("dataset too large"); count *= d; }
    if (nbytes != count * 8) return bad("dataset \"" + name + "\": " + std::to_string(nbytes) + " bytes stored for " + std::to_string(count) + " elements");
    if (count && (addr == UNDEF || !r.in(addr, nbytes))) return bad("dataset \"" + name + "\" has no storage in the file");
    data.resize(count);
    if (count) memcpy(data.data(), r.b.data() + addr, nbytes);
    return true;
}

}  // namespace h5lite
