// libclane_b200.so -- fast ingest of the reference's on-disk edge list (host only).
//
// Reference call site replaced: /root/reference/clane/graph.py:73-81
//     lines = io.read().strip().split("\n");  src_id, dst_id = line.split("\t");  vertex_ids.index(id)
// i.e. text mode (universal newlines: "\r\n" and "\r" read as "\n"), surrounding whitespace of the whole file
// stripped, one edge per line, exactly one TAB per line, ids resolved to their FIRST position in V.  The
// reference's list.index makes this O(E*N); here: one hash map over V, the file cut at line starts into one
// piece per thread, each piece parsed independently.  Errors are reported for the first offending line in file
// order, as the reference's loop would hit them.
#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstring>
#include <new>
#include <memory>
#include <string>
#include <string_view>
#include <thread>
#include <unordered_map>
#include <vector>

#include "clane_b200.h"

struct clane_edge_file {
    std::vector<int64_t> src, dst;
};

namespace {

inline bool py_space(unsigned char c) {   // str.strip() on ASCII input
    return c == ' ' || (c >= '\t' && c <= '\r') || (c >= 0x1c && c <= 0x1f);
}

// a position is a line start iff it follows a line terminator ("\n", "\r\n" or a lone "\r")
inline bool line_start(const char* b, const char* p) {
    if (p == b) return true;
    return p[-1] == '\n' || (p[-1] == '\r' && *p != '\n');
}

// end of the line starting at p (exclusive) and the start of the next one
inline void line_extent(const char* p, const char* end, const char** stop, const char** next) {
    const char* q = p;
    while (q < end && *q != '\n' && *q != '\r') ++q;
    *stop = q;
    if (q < end) q += (*q == '\r' && q + 1 < end && q[1] == '\n') ? 2 : 1;
    *next = q;
}

struct Piece {
    const char* lo;
    const char* hi;      // lines starting in [lo, hi)
    int64_t lines = 0;   // pass 1
    int64_t first = 0;   // index of its first line
};

struct Failure {
    int64_t line = INT64_MAX;
    int code = CLANE_OK;
    std::string text;
};


// job(c) for every piece c, one thread each; an exception inside a worker (bad_alloc from a Failure's string) is
// flagged and re-raised on the caller after the join, and a thread that cannot be created runs on the caller
template <class F>
void run_pieces(int npieces, F&& job) {
    std::atomic<int> failed{0};
    std::vector<std::thread> th;
    auto guarded = [&](int c) { try { job(c); } catch (...) { failed.store(1); } };
    for (int c = 0; c < npieces; ++c) {
        try { th.emplace_back(guarded, c); } catch (...) { guarded(c); }
    }
    for (auto& t : th) t.join();
    if (failed.load()) throw std::bad_alloc();
}

}  // namespace

extern "C" {

int clane_edges_open(const char* v_ids, int64_t v_bytes, int64_t n_vertices, const char* e_path, int32_t threads,
                     clane_edge_file** out, int64_t* e_raw, int64_t* err_line, char* err_text, int32_t err_cap) {
    if (!out || !e_raw || !e_path || n_vertices < 0 || v_bytes < 0 || (v_bytes > 0 && !v_ids)) return CLANE_EINVAL;
    *out = nullptr;
    auto fail_text = [&](const std::string& t) {
        if (err_text && err_cap > 0) {
            const size_t k = std::min<size_t>(t.size(), (size_t)err_cap - 1);
            memcpy(err_text, t.data(), k);
            err_text[k] = 0;
        }
    };
    try {
        // ---- V: n_vertices ids joined by "\n" (parsed by the caller with the reference's own expression) ----
        std::unordered_map<std::string_view, int64_t> first;
        first.reserve((size_t)n_vertices * 2 + 16);
        {
            const char* p = v_ids;
            const char* end = v_ids + v_bytes;
            for (int64_t i = 0; i < n_vertices; ++i) {
                const char* q = static_cast<const char*>(memchr(p, '\n', (size_t)(end - p)));
                if (!q) q = end;
                first.emplace(std::string_view(p, (size_t)(q - p)), i);   // keeps the first position of a repeated id
                p = q < end ? q + 1 : end;
                if (p == end && i + 1 < n_vertices && !(q < end)) return CLANE_EINVAL;   // fewer ids than announced
            }
        }
        // ---- E: whole file in memory ----
        FILE* f = fopen(e_path, "rb");
        if (!f) { fail_text(e_path); return CLANE_ENOENT; }
        std::vector<char> buf;
        {
            fseek(f, 0, SEEK_END);
            const long sz = ftell(f);
            fseek(f, 0, SEEK_SET);
            buf.resize((size_t)std::max<long>(sz, 0) + 1);
            const size_t got = sz > 0 ? fread(buf.data(), 1, (size_t)sz, f) : 0;
            fclose(f);
            buf.resize(got + 1);
            buf[got] = 0;   // sentinel: line_start / line_extent may look one byte ahead
        }
        const char* b = buf.data();
        const char* e = b + buf.size() - 1;
        while (b < e && py_space((unsigned char)*b)) ++b;
        while (e > b && py_space((unsigned char)e[-1])) --e;
        // "".split("\n") == [""]: an empty file is one empty line (which then fails to unpack, as upstream)
        const int nthreads = std::max(1, std::min<int>(threads > 0 ? threads : (int)std::thread::hardware_concurrency(), 64));
        const int64_t bytes = e - b;
        const int npieces = (int)std::max<int64_t>(1, std::min<int64_t>(nthreads, bytes / (1 << 16)));
        std::vector<Piece> pieces((size_t)npieces);
        for (int c = 0; c < npieces; ++c) {
            const char* lo = b + bytes * c / npieces;
            if (c > 0) while (lo < e && !line_start(b, lo)) ++lo;
            pieces[(size_t)c].lo = lo;
        }
        for (int c = 0; c < npieces; ++c) pieces[(size_t)c].hi = c + 1 < npieces ? pieces[(size_t)c + 1].lo : e;
        // the text after the last terminator is a line too, even when empty: "a\tb\n" was stripped, but "a\tb\n\nc\td"
        // has an empty middle line.  A piece whose range is empty contributes nothing, except piece 0 of an empty file.
        auto for_lines = [&](const Piece& pc, auto&& fn) {
            const char* p = pc.lo;
            if (pc.lo == pc.hi) {
                if (&pc == &pieces[0] && bytes == 0) fn(p, p);
                return;
            }
            while (p < pc.hi) {
                const char *stop, *next;
                line_extent(p, e, &stop, &next);
                fn(p, stop);
                if (next == e && stop < e && &pc == &pieces.back()) { fn(e, e); break; }   // terminator at the very end (not after strip)
                p = next;
            }
        };
        run_pieces(npieces, [&](int c) {   // pass 1: count lines
            int64_t k = 0;
            for_lines(pieces[(size_t)c], [&](const char*, const char*) { ++k; });
            pieces[(size_t)c].lines = k;
        });
        int64_t total = 0;
        for (auto& pc : pieces) { pc.first = total; total += pc.lines; }
        std::unique_ptr<clane_edge_file> ef(new clane_edge_file());
        ef->src.resize((size_t)total);
        ef->dst.resize((size_t)total);
        std::vector<Failure> fails((size_t)npieces);
        run_pieces(npieces, [&](int c) {   // pass 2: parse
                {
                    int64_t k = pieces[(size_t)c].first;
                    Failure& fl = fails[(size_t)c];
                    for_lines(pieces[(size_t)c], [&](const char* p, const char* stop) {
                        const int64_t line = k++;
                        if (fl.code != CLANE_OK) return;
                        const char* tab = static_cast<const char*>(memchr(p, '\t', (size_t)(stop - p)));
                        if (!tab) { fl = Failure{line, CLANE_EPARSE, "not enough values to unpack (expected 2, got 1)"}; return; }
                        if (memchr(tab + 1, '\t', (size_t)(stop - tab - 1))) {
                            fl = Failure{line, CLANE_EPARSE, "too many values to unpack (expected 2)"};
                            return;
                        }
                        const std::string_view a(p, (size_t)(tab - p)), d(tab + 1, (size_t)(stop - tab - 1));
                        auto ia = first.find(a);
                        if (ia == first.end()) { fl = Failure{line, CLANE_EUNKNOWNID, std::string(a)}; return; }
                        auto id = first.find(d);
                        if (id == first.end()) { fl = Failure{line, CLANE_EUNKNOWNID, std::string(d)}; return; }
                        ef->src[(size_t)line] = ia->second;
                        ef->dst[(size_t)line] = id->second;
                    });
                }
        });
        const Failure* worst = nullptr;
        for (const Failure& fl : fails)
            if (fl.code != CLANE_OK && (!worst || fl.line < worst->line)) worst = &fl;
        if (worst) {
            if (err_line) *err_line = worst->line;
            fail_text(worst->text);
            return worst->code;
        }
        *e_raw = total;
        *out = ef.release();
        return CLANE_OK;
    } catch (...) {
        return CLANE_ENOMEM;
    }
}

int clane_edges_read(const clane_edge_file* f, int64_t* h_src, int64_t* h_dst) {
    if (!f || (!f->src.empty() && (!h_src || !h_dst))) return CLANE_EINVAL;
    if (!f->src.empty()) {
        memcpy(h_src, f->src.data(), f->src.size() * sizeof(int64_t));
        memcpy(h_dst, f->dst.data(), f->dst.size() * sizeof(int64_t));
    }
    return CLANE_OK;
}

int clane_edges_close(clane_edge_file* f) {
    delete f;
    return CLANE_OK;
}

}  // extern "C"
