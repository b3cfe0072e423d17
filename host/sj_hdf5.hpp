// sj_hdf5.hpp -- a header-only HDF5 writer for the subset field_samples.h5 needs (src/disp.cpp:758-923 writes it
// through H5Cpp; libhdf5 does not exist in this image).  Same encoding as sim_juncs_b200/hdf5.py, which documents
// and tests the byte layout against a file written by the real library:
//   superblock v0 | v1 object headers | symbol-table groups (local heap + one v1 B-tree node + SNOD nodes)
//   dataspace v1 | datatype v1 (u64, f64, compounds of f64) | fill value | contiguous layout v3
// Usage: sj_h5::Writer w; w.dataset_f64("info/time_bounds", p, 3); w.compound("c/p/time", {"Re","Im"}, ...); w.save(path).
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <utility>
#include <vector>

namespace sj_h5 {

typedef std::string bytes;
static const uint64_t UNDEF = ~0ull;

inline void put(bytes &b, uint64_t v, int n) { for (int i = 0; i < n; ++i) b.push_back((char)((v >> (8 * i)) & 255)); }
inline void pad8(bytes &b) { while (b.size() % 8) b.push_back('\0'); }

inline bytes f64_type() {
    bytes t; put(t, 0x11, 1); put(t, 0x20, 1); put(t, 63, 1); put(t, 0, 1); put(t, 8, 4);
    put(t, 0, 2); put(t, 64, 2); put(t, 52, 1); put(t, 11, 1); put(t, 0, 1); put(t, 52, 1); put(t, 1023, 4);
    return t;
}
inline bytes u64_type() { bytes t; put(t, 0x10, 1); put(t, 0, 3); put(t, 8, 4); put(t, 0, 2); put(t, 64, 2); return t; }
// compound of doubles: (member name, byte offset) pairs, total size in bytes (version-1 member layout)
inline bytes compound_type(const std::vector<std::pair<std::string, unsigned> > &members, unsigned size) {
    bytes t; put(t, 0x16, 1); put(t, members.size(), 2); put(t, 0, 1); put(t, size, 4);
    for (size_t i = 0; i < members.size(); ++i) {
        bytes nm = members[i].first; nm.push_back('\0'); pad8(nm);
        t += nm; put(t, members[i].second, 4); put(t, 0, 28);     // rank 0, permutation, reserved, four dimension sizes
        t += f64_type();
    }
    return t;
}

struct Node {
    bool is_group;
    std::map<std::string, Node> children;      // std::map iterates in byte order of the names, the B-tree's order
    bytes dtype, raw; uint64_t count;
    std::vector<uint64_t> dims;                // empty: rank 1 with `count` elements
    Node() : is_group(true), count(0) {}
};

class Writer {
public:
    Node root;
    Node &group(const std::string &path) {
        Node *g = &root; size_t a = 0;
        while (a < path.size()) {
            size_t b = path.find('/', a); if (b == std::string::npos) b = path.size();
            if (b > a) g = &g->children[path.substr(a, b - a)];
            a = b + 1;
        }
        return *g;
    }
    void dataset(const std::string &path, const bytes &dtype, const void *data, uint64_t count, size_t item) {
        const size_t cut = path.rfind('/');
        Node &g = cut == std::string::npos ? root : group(path.substr(0, cut));
        Node &d = g.children[cut == std::string::npos ? path : path.substr(cut + 1)];
        d.is_group = false; d.dtype = dtype; d.count = count; d.raw.assign((const char *)data, count * item);
    }
    void dataset_f64(const std::string &path, const double *p, uint64_t n) { dataset(path, f64_type(), p, n, 8); }
    // rank-N array of doubles, C order (the whole-grid dumps eps-*.h5 / ex-*.h5)
    void dataset_f64_nd(const std::string &path, const double *p, const std::vector<uint64_t> &dims) {
        uint64_t n = 1;
        for (size_t i = 0; i < dims.size(); ++i) n *= dims[i];
        dataset(path, f64_type(), p, n, 8);
        const size_t cut = path.rfind('/');
        Node &g = cut == std::string::npos ? root : group(path.substr(0, cut));
        g.children[cut == std::string::npos ? path : path.substr(cut + 1)].dims = dims;
    }
    void dataset_u64(const std::string &path, const uint64_t *p, uint64_t n) { dataset(path, u64_type(), p, n, 8); }

    int save(const char *path) {
        leaf_k = 4;
        const size_t most = max_entries(root);
        if ((most + 63) / 64 > leaf_k) leaf_k = (unsigned)((most + 63) / 64);
        buf.assign(96, '\0');
        uint64_t bt, hp; const uint64_t hdr = emit_group(root, bt, hp);
        pad8(buf);
        bytes sb("\x89HDF\r\n\x1a\n", 8);
        put(sb, 0, 5); put(sb, 8, 1); put(sb, 8, 1); put(sb, 0, 1); put(sb, leaf_k, 2); put(sb, 16, 2); put(sb, 0, 4);
        put(sb, 0, 8); put(sb, UNDEF, 8); put(sb, buf.size(), 8); put(sb, UNDEF, 8);
        put(sb, 0, 8); put(sb, hdr, 8); put(sb, 1, 4); put(sb, 0, 4); put(sb, bt, 8); put(sb, hp, 8);
        buf.replace(0, 96, sb);
        FILE *fp = fopen(path, "wb");
        if (!fp) return -1;
        const size_t n = fwrite(buf.data(), 1, buf.size(), fp);
        fclose(fp);
        return n == buf.size() ? 0 : -1;
    }

private:
    bytes buf; unsigned leaf_k;
    static size_t max_entries(const Node &g) {
        size_t m = g.children.size();
        for (std::map<std::string, Node>::const_iterator it = g.children.begin(); it != g.children.end(); ++it)
            if (it->second.is_group) { const size_t c = max_entries(it->second); if (c > m) m = c; }
        return m;
    }
    uint64_t alloc(const bytes &blob) { pad8(buf); const uint64_t a = buf.size(); buf += blob; return a; }
    static bytes message(unsigned type, bytes body, unsigned flags) {
        pad8(body); bytes m; put(m, type, 2); put(m, body.size(), 2); put(m, flags, 1); put(m, 0, 3); return m + body;
    }
    uint64_t object_header(const bytes &msgs, unsigned n) {
        bytes h; put(h, 1, 1); put(h, 0, 1); put(h, n, 2); put(h, 1, 4); put(h, msgs.size(), 4); put(h, 0, 4);
        return alloc(h + msgs);
    }
    uint64_t emit_dataset(const Node &d) {
        const uint64_t addr = d.raw.empty() ? UNDEF : alloc(d.raw);
        bytes space; put(space, 1, 1); put(space, d.dims.empty() ? 1 : d.dims.size(), 1); put(space, 0, 6);
        if (d.dims.empty()) put(space, d.count, 8);
        else for (size_t i = 0; i < d.dims.size(); ++i) put(space, d.dims[i], 8);
        bytes fill; put(fill, 1, 1); put(fill, 2, 1); put(fill, 2, 1); put(fill, 1, 1); put(fill, 0, 4);
        bytes lay; put(lay, 3, 1); put(lay, 1, 1); put(lay, addr, 8); put(lay, d.raw.size(), 8);
        return object_header(message(1, space, 0) + message(3, d.dtype, 1) + message(5, fill, 1) + message(8, lay, 0), 4);
    }
    uint64_t emit_group(const Node &g, uint64_t &bt_addr, uint64_t &heap_addr) {
        struct Ent { uint64_t hdr, bt, hp, name_off; bool grp; };
        std::vector<Ent> ents;
        bytes heap(8, '\0');
        for (std::map<std::string, Node>::const_iterator it = g.children.begin(); it != g.children.end(); ++it) {
            Ent e; e.grp = it->second.is_group; e.bt = e.hp = 0;
            e.hdr = e.grp ? emit_group(it->second, e.bt, e.hp) : emit_dataset(it->second);
            e.name_off = heap.size();
            bytes nm = it->first; nm.push_back('\0'); pad8(nm); heap += nm;
            ents.push_back(e);
        }
        const uint64_t free_off = heap.size();
        put(heap, 1, 8); put(heap, 16, 8);                  // one free block closes the segment (H5HL_FREE_NULL = 1)
        pad8(buf); heap_addr = buf.size();
        bytes hh("HEAP", 4); put(hh, 0, 4); put(hh, heap.size(), 8); put(hh, free_off, 8); put(hh, heap_addr + 32, 8);
        alloc(hh + heap);
        const size_t per = 2 * leaf_k;
        std::vector<uint64_t> nodes, last;
        for (size_t s = 0; s < ents.size(); s += per) {
            const size_t e = s + per < ents.size() ? s + per : ents.size();
            bytes n("SNOD", 4); put(n, 1, 1); put(n, 0, 1); put(n, e - s, 2);
            for (size_t q = s; q < e; ++q) {
                put(n, ents[q].name_off, 8); put(n, ents[q].hdr, 8); put(n, ents[q].grp ? 1 : 0, 4); put(n, 0, 4);
                put(n, ents[q].grp ? ents[q].bt : 0, 8); put(n, ents[q].grp ? ents[q].hp : 0, 8);
            }
            n.resize(8 + per * 40, '\0');
            nodes.push_back(alloc(n)); last.push_back(ents[e - 1].name_off);
        }
        bytes t("TREE", 4); put(t, 0, 1); put(t, 0, 1); put(t, nodes.size(), 2); put(t, UNDEF, 8); put(t, UNDEF, 8); put(t, 0, 8);
        for (size_t q = 0; q < nodes.size(); ++q) { put(t, nodes[q], 8); put(t, last[q], 8); }
        t.resize(24 + (4 * 16 + 1) * 8, '\0');
        bt_addr = alloc(t);
        bytes stab; put(stab, bt_addr, 8); put(stab, heap_addr, 8);
        return object_header(message(0x11, stab, 0), 1);
    }
};

}  // namespace sj_h5
