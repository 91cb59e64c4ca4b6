// N2: the wire format the second half of RECON reads.
//   save_embed (GAT/main.py:406-413) writes  json.dump({idx: row.tolist()}, f, indent=4, cls=CustomEncoder)
//   and the consumer does json.load + gat_embeddings["<idx>"] (train.py:100-130, utils/context_utils.py:444-504).
// For the C2 table (2M x 200 floats) that is ~10 GB of text produced by a pure-Python encoder; this file emits the
// byte-identical text from a host fp32 buffer with all cores (rows formatted in parallel, written in order), plus a
// binary side-car (header + raw fp32 rows) for consumers that can mmap.
//
// Byte-identical means Python's float repr of the float32 value widened to double (ndarray.tolist()):
//   shortest digit string that round-trips the double (std::to_chars == David Gay's mode 0), fixed notation when
//   -4 < decpt <= 16 with ".0" for integral values, else d[.ddd]e[+-]XX with at least two exponent digits
//   (CPython Python/pystrtod.c format_float_short, type 'r'); NaN / Infinity / -Infinity as json emits them.
#include "../../include/spkbgat.h"

#include <algorithm>
#include <charconv>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <atomic>
#include <thread>
#include <vector>
#include <condition_variable>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif
#include <functional>
#include <mutex>

namespace {
// Minimal persistent worker pool for the host-side packers (one job at a time; callers are serialised by the mutex).
class PackPool {
public:
    ~PackPool() {
        { std::lock_guard<std::mutex> lk(m_); stop_ = true; ++gen_; }
        go_.notify_all();
        for (auto& t : th_) if (t.joinable()) t.join();
    }
    void run(int n, const std::function<void(int)>& f) {
        std::lock_guard<std::mutex> serial(run_m_);
        std::unique_lock<std::mutex> lk(m_);
        while ((int)th_.size() < n) { const int id = (int)th_.size(); th_.emplace_back([this, id] { loop(id); }); }
        job_ = &f; active_ = n; pending_ = n; ++gen_;
        go_.notify_all();
        done_.wait(lk, [this] { return pending_ == 0; });
        job_ = nullptr;
    }
private:
    void loop(int id) {
        uint64_t seen = 0;
        std::unique_lock<std::mutex> lk(m_);
        for (;;) {
            go_.wait(lk, [&] { return gen_ != seen; });
            seen = gen_;
            if (stop_) return;
            if (id >= active_) continue;
            const std::function<void(int)>* f = job_;
            lk.unlock();
            (*f)(id);
            lk.lock();
            if (--pending_ == 0) done_.notify_all();
        }
    }
    std::vector<std::thread> th_;
    std::mutex m_, run_m_;
    std::condition_variable go_, done_;
    const std::function<void(int)>* job_ = nullptr;
    uint64_t gen_ = 0;
    int active_ = 0, pending_ = 0;
    bool stop_ = false;
};
PackPool& pack_pool() { static PackPool* p = new PackPool(); return *p; }   // leaked on purpose: no join at process exit
}  // namespace


namespace spk {
void set_error(const char* fmt, ...);
}

namespace {

// appends repr(float(v)) to out
inline void append_pyfloat(std::string& out, float f) {
    const double v = (double)f;
    if (std::isnan(v)) { out += "NaN"; return; }
    if (std::isinf(v)) { out += v > 0 ? "Infinity" : "-Infinity"; return; }
    if (v == 0.0) { out += std::signbit(v) ? "-0.0" : "0.0"; return; }
    char buf[40];
    const auto res = std::to_chars(buf, buf + sizeof(buf), v, std::chars_format::scientific);
    // buf = [-]d[.ddd]e[+-]XX
    const char* p = buf;
    if (*p == '-') { out += '-'; ++p; }
    char digits[24];
    int nd = 0;
    digits[nd++] = *p++;
    if (*p == '.') {
        ++p;
        while (*p != 'e') digits[nd++] = *p++;
    }
    ++p;                                      // 'e'
    int esign = 1;
    if (*p == '-') { esign = -1; ++p; } else if (*p == '+') { ++p; }
    int ex = 0;
    while (p < res.ptr) ex = ex * 10 + (*p++ - '0');
    const int decpt = esign * ex + 1;         // value = 0.d1d2... x 10^decpt
    if (decpt <= -4 || decpt > 16) {          // exponent notation
        out += digits[0];
        if (nd > 1) { out += '.'; out.append(digits + 1, nd - 1); }
        int e10 = decpt - 1;
        out += 'e';
        out += e10 < 0 ? '-' : '+';
        e10 = e10 < 0 ? -e10 : e10;
        char eb[8];
        int n = 0;
        do { eb[n++] = (char)('0' + e10 % 10); e10 /= 10; } while (e10);
        if (n < 2) eb[n++] = '0';
        while (n) out += eb[--n];
    } else if (decpt <= 0) {
        out += "0.";
        out.append((size_t)(-decpt), '0');
        out.append(digits, nd);
    } else if (decpt >= nd) {
        out.append(digits, nd);
        out.append((size_t)(decpt - nd), '0');
        out += ".0";
    } else {
        out.append(digits, decpt);
        out += '.';
        out.append(digits + decpt, nd - decpt);
    }
}

// rows [r0, r1) of the dict body:  "<idx>": [\n        v,\n        v\n    ]  joined by ",\n    "
void format_rows(const float* data, int64_t ld, int64_t width, int64_t r0, int64_t r1, int64_t n_rows, std::string& out) {
    out.clear();
    out.reserve((size_t)((r1 - r0) * (width * 30 + 32)));
    char kb[32];
    for (int64_t r = r0; r < r1; ++r) {
        out += "\n    \"";
        const int kn = snprintf(kb, sizeof(kb), "%lld", (long long)r);
        out.append(kb, kn);
        out += "\": [";
        if (width == 0) {
            out += "]";
        } else {
            const float* row = data + r * ld;
            for (int64_t c = 0; c < width; ++c) {
                out += "\n        ";
                append_pyfloat(out, row[c]);
                if (c + 1 < width) out += ',';
            }
            out += "\n    ]";
        }
        if (r + 1 < n_rows) out += ',';
    }
}

}  // namespace

extern "C" {

int spk_export_json(const float* host, int64_t rows, int64_t width, int64_t ld, const char* path, int32_t n_threads) {
    if (rows < 0 || width < 0 || (rows > 0 && width > 0 && host == nullptr) || ld < width || path == nullptr) {
        spk::set_error("export_json: bad arguments (rows=%lld width=%lld ld=%lld)", (long long)rows, (long long)width, (long long)ld);
        return 2;
    }
    FILE* f = fopen(path, "wb");
    if (!f) { spk::set_error("export_json: cannot open %s", path); return 3; }
    int rc = 0;
    if (rows == 0) {
        if (fputs("{}", f) < 0) rc = 4;
    } else {
        int nt = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
        nt = std::max(1, std::min(nt, 256));
        const int64_t block = std::max<int64_t>(1, (int64_t)(4 << 20) / std::max<int64_t>(1, width * 26));   // ~4 MB of text per task
        const int64_t wave = block * nt;
        std::vector<std::string> bufs((size_t)nt);
        if (fputc('{', f) == EOF) rc = 4;
        for (int64_t w0 = 0; w0 < rows && rc == 0; w0 += wave) {
            std::vector<std::thread> th;
            int used = 0;
            for (int t = 0; t < nt; ++t) {
                const int64_t r0 = w0 + t * block, r1 = std::min(rows, r0 + block);
                if (r0 >= rows) break;
                ++used;
                if (nt == 1) format_rows(host, ld, width, r0, r1, rows, bufs[t]);
                else th.emplace_back(format_rows, host, ld, width, r0, r1, rows, std::ref(bufs[(size_t)t]));
            }
            for (auto& x : th) x.join();
            for (int t = 0; t < used && rc == 0; ++t)
                if (fwrite(bufs[t].data(), 1, bufs[t].size(), f) != bufs[t].size()) rc = 4;
        }
        if (rc == 0 && fputs("\n}", f) < 0) rc = 4;
    }
    if (fclose(f) != 0 && rc == 0) rc = 4;
    if (rc) spk::set_error("export_json: write to %s failed", path);
    return rc;
}

/* side-car: 64-byte header {magic "SPKEMB01", int64 rows, int64 width, int64 dtype(0 = fp32 LE), zero pad} + rows*width fp32 */
int spk_export_bin(const float* host, int64_t rows, int64_t width, int64_t ld, const char* path) {
    if (rows < 0 || width < 0 || (rows > 0 && width > 0 && host == nullptr) || ld < width || path == nullptr) {
        spk::set_error("export_bin: bad arguments");
        return 2;
    }
    FILE* f = fopen(path, "wb");
    if (!f) { spk::set_error("export_bin: cannot open %s", path); return 3; }
    char hdr[64];
    memset(hdr, 0, sizeof(hdr));
    memcpy(hdr, "SPKEMB01", 8);
    const int64_t meta[3] = {rows, width, 0};
    memcpy(hdr + 8, meta, sizeof(meta));
    int rc = fwrite(hdr, 1, sizeof(hdr), f) == sizeof(hdr) ? 0 : 4;
    if (rc == 0 && rows > 0 && width > 0) {
        if (ld == width) {
            if (fwrite(host, sizeof(float), (size_t)(rows * width), f) != (size_t)(rows * width)) rc = 4;
        } else {
            for (int64_t r = 0; r < rows && rc == 0; ++r)
                if (fwrite(host + r * ld, sizeof(float), (size_t)width, f) != (size_t)width) rc = 4;
        }
    }
    if (fclose(f) != 0 && rc == 0) rc = 4;
    if (rc) spk::set_error("export_bin: write to %s failed", path);
    return rc;
}

}  // extern "C"

// ---- reader of the same text (what train.py:103-104 does with json.load, minus the Python object per float) ----------
// Works on any JSON object of the shape {"<int>": [numbers...], ...} (whitespace-insensitive): the file is mapped, cut
// into byte ranges at key boundaries (only keys carry quotes) and parsed by all host threads with std::from_chars;
// row "<i>" lands in out[i, :]. Values are narrowed to fp32, which is exact for text written from fp32 tables.
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

namespace {

struct Mapped {
    const char* p = nullptr; size_t n = 0; int fd = -1;
    bool open_file(const char* path) {
        fd = ::open(path, O_RDONLY);
        if (fd < 0) return false;
        struct stat st;
        if (fstat(fd, &st) != 0) { ::close(fd); fd = -1; return false; }
        n = (size_t)st.st_size;
        if (n == 0) { p = ""; return true; }
        void* m = mmap(nullptr, n, PROT_READ, MAP_PRIVATE, fd, 0);
        if (m == MAP_FAILED) { ::close(fd); fd = -1; return false; }
        p = (const char*)m;
        return true;
    }
    ~Mapped() { if (p && n) munmap((void*)p, n); if (fd >= 0) ::close(fd); }
};

inline const char* skip_ws(const char* s, const char* e) {
    while (s < e && (*s == ' ' || *s == '\n' || *s == '\t' || *s == '\r')) ++s;
    return s;
}

// parses one number (JSON number, NaN, Infinity, -Infinity); returns nullptr on error
inline const char* parse_number(const char* s, const char* e, float* out) {
    if (s < e && *s == 'N') { if (e - s >= 3 && !memcmp(s, "NaN", 3)) { *out = NAN; return s + 3; } return nullptr; }
    if (s < e && *s == 'I') { if (e - s >= 8 && !memcmp(s, "Infinity", 8)) { *out = INFINITY; return s + 8; } return nullptr; }
    if (e - s >= 9 && !memcmp(s, "-Infinity", 9)) { *out = -INFINITY; return s + 9; }
    double v;
    const auto r = std::from_chars(s, e, v);
    if (r.ec == std::errc::result_out_of_range) {           // denormal-range or huge text: let strtod decide
        char* end = nullptr;
        std::string tmp(s, (size_t)std::min<ptrdiff_t>(e - s, 64));
        v = strtod(tmp.c_str(), &end);
        if (end == tmp.c_str()) return nullptr;
        *out = (float)v;
        return s + (end - tmp.c_str());
    }
    if (r.ec != std::errc()) return nullptr;
    *out = (float)v;
    return r.ptr;
}

// parses rows whose key's opening quote lies in [s, stop); returns rows parsed or -1
long parse_rows(const char* s, const char* stop, const char* e, float* out, int64_t rows, int64_t width,
                int64_t ld, std::atomic<uint8_t>* seen) {
    long done = 0;
    while (true) {
        s = (const char*)memchr(s, '"', (size_t)(e - s));
        if (!s || s >= stop) return done;
        ++s;
        int64_t idx = 0;
        const char* k0 = s;
        while (s < e && *s >= '0' && *s <= '9' && s - k0 < 19) idx = idx * 10 + (*s++ - '0');   // <= 18 digits: no overflow
        if (s - k0 >= 19) return -1;
        if (s == k0 || s >= e || *s != '"') {
            // not "<digits>": this was the CLOSING quote of a key whose opening quote belongs to the previous range
            // (a range may start inside a key); a closing quote is followed by [ws] ':' -- anything else is an error
            const char* c = skip_ws(k0, e);
            if (s == k0 && c < e && *c == ':') continue;
            return -1;
        }
        if (idx < 0 || idx >= rows) return -1;
        if (seen[idx].exchange(1, std::memory_order_relaxed)) return -1;      // duplicate key: json.load would return fewer rows
        s = skip_ws(s + 1, e);
        if (s >= e || *s != ':') return -1;
        s = skip_ws(s + 1, e);
        if (s >= e || *s != '[') return -1;
        s = skip_ws(s + 1, e);
        float* row = out + idx * ld;
        int64_t c = 0;
        if (s < e && *s == ']') { ++s; }
        else {
            while (true) {
                float v;
                s = parse_number(s, e, &v);
                if (!s) return -1;
                if (c < width) row[c] = v;
                ++c;
                s = skip_ws(s, e);
                if (s < e && *s == ',') { s = skip_ws(s + 1, e); continue; }
                if (s < e && *s == ']') { ++s; break; }
                return -1;
            }
        }
        if (c != width) return -1;
        ++done;
    }
}

}  // namespace

extern "C" {

/* rows = number of keys, width = length of the first row's list; 0 on success */
int spk_import_json_shape(const char* path, int64_t* rows, int64_t* width) {
    Mapped f;
    if (!path || !rows || !width || !f.open_file(path)) { spk::set_error("import_json: cannot open %s", path ? path : "(null)"); return 3; }
    const char* s = f.p; const char* e = f.p + f.n;
    int64_t quotes = 0;
    for (const char* q = s; (q = (const char*)memchr(q, '"', (size_t)(e - q))) != nullptr; ++q) ++quotes;
    if (quotes % 2) { spk::set_error("import_json: unbalanced quotes in %s", path); return 5; }
    *rows = quotes / 2;
    *width = 0;
    if (*rows) {
        const char* b = (const char*)memchr(s, '[', (size_t)(e - s));
        if (!b) { spk::set_error("import_json: no list in %s", path); return 5; }
        b = skip_ws(b + 1, e);
        int64_t w = 0;
        if (b < e && *b != ']') {
            w = 1;
            for (; b < e && *b != ']'; ++b) if (*b == ',') ++w;
        }
        *width = w;
    }
    return 0;
}

int spk_import_json(const char* path, float* out, int64_t rows, int64_t width, int64_t ld, int32_t n_threads) {
    Mapped f;
    if (!path || !f.open_file(path)) { spk::set_error("import_json: cannot open %s", path ? path : "(null)"); return 3; }
    if (rows == 0) return 0;
    if (!out || ld < width) { spk::set_error("import_json: bad output buffer"); return 2; }
    int nt = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
    nt = std::max(1, std::min(nt, 256));
    if (f.n < (size_t)(1 << 20)) nt = 1;
    const char* b = f.p; const char* e = f.p + f.n;
    std::vector<long> res((size_t)nt, 0);
    std::vector<std::atomic<uint8_t>> seen((size_t)rows);                      // every key exactly once (total == rows below)
    for (auto& x : seen) x.store(0, std::memory_order_relaxed);
    std::vector<std::thread> th;
    const size_t chunk = (f.n + nt - 1) / nt;
    for (int t = 0; t < nt; ++t) {
        const char* s = b + std::min(f.n, (size_t)t * chunk);
        const char* stop = b + std::min(f.n, (size_t)(t + 1) * chunk);
        if (nt == 1) res[0] = parse_rows(s, stop, e, out, rows, width, ld, seen.data());
        else th.emplace_back([&, t, s, stop]() { res[(size_t)t] = parse_rows(s, stop, e, out, rows, width, ld, seen.data()); });
    }
    for (auto& x : th) x.join();
    long total = 0;
    for (long r : res) { if (r < 0) { spk::set_error("import_json: %s is not an {\"i\": [numbers]} table of width %lld", path, (long long)width); return 5; } total += r; }
    if (total != rows) { spk::set_error("import_json: parsed %ld of %lld rows of %s", total, (long long)rows, path); return 5; }
    return 0;
}


/* Host staging of the edge list (HOST pointers, no GPU work): int64 index column (stride in elements, e.g. one column
 * of the [E2,4] 2-hop rows) -> int32, range-checked against [lo, hi), by n_threads host threads (0 = all cores). The int64
 * tensors of the reference API carry 32 bits of information per element; packing them into pinned memory halves the
 * bytes that cross PCIe. Returns 0, or 5 if a value is out of range (IndexError in the reference). */
int spk_pack_index_host(const int64_t* src, int64_t n, int64_t stride, int64_t lo, int64_t hi, int32_t* dst, int32_t n_threads) {
    if (n < 0 || stride < 1 || (n > 0 && (!src || !dst))) { spk::set_error("pack_index_host: bad arguments"); return 1; }
    int nt = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
    nt = std::max(1, std::min(nt, 64));
    if (n < (1 << 18)) nt = 1;
    std::atomic<int> bad{0};
    // The destination is a pinned staging buffer that the copy engine reads next and the cores never read back: written
    // with non-temporal stores (no read-for-ownership of the destination lines: a third less memory traffic per element).
    const uint64_t span = (uint64_t)(hi - lo);
    auto work = [&](int64_t b, int64_t e) {
        int local_bad = 0;
        int64_t i = b;
#if defined(__SSE2__)
        for (; i < e && (reinterpret_cast<uintptr_t>(dst + i) & 15); ++i) {
            const int64_t v = src[i * stride];
            local_bad |= (uint64_t)(v - lo) >= span;
            dst[i] = (int32_t)v;
        }
        for (; i + 4 <= e; i += 4) {
            const int64_t v0 = src[i * stride], v1 = src[(i + 1) * stride], v2 = src[(i + 2) * stride], v3 = src[(i + 3) * stride];
            local_bad |= ((uint64_t)(v0 - lo) >= span) | ((uint64_t)(v1 - lo) >= span) | ((uint64_t)(v2 - lo) >= span) |
                         ((uint64_t)(v3 - lo) >= span);
            _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i), _mm_set_epi32((int32_t)v3, (int32_t)v2, (int32_t)v1, (int32_t)v0));
        }
#endif
        for (; i < e; ++i) {
            const int64_t v = src[i * stride];
            local_bad |= (uint64_t)(v - lo) >= span;
            dst[i] = (int32_t)v;
        }
#if defined(__SSE2__)
        _mm_sfence();
#endif
        if (local_bad) bad.store(1, std::memory_order_relaxed);
    };
    if (nt == 1) work(0, n);
    else {
        // persistent workers: the stager calls this once per 16 MB chunk, and spawning up to 64 threads per call cost more
        // than packing the chunk
        const int64_t chunk = (n + nt - 1) / nt;
        pack_pool().run(nt, [&](int t) {
            const int64_t b = std::min(n, (int64_t)t * chunk), e = std::min(n, b + chunk);
            if (b < e) work(b, e);
        });
    }
    if (bad.load()) { spk::set_error("pack_index_host: index outside [%lld, %lld)", (long long)lo, (long long)hi); return 5; }
    return 0;
}

}  // extern "C"
