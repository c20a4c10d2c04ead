// libfz_rasterio.so -- raster file I/O at the two ends of the zonal path (include/flair_zonal_rasterio.h).
//
// Host code only (g++ -O3 -pthread -lz): a TIFF / BigTIFF / GeoTIFF block reader and writer whose unit of work is one
// block (tile or strip) per task, dealt to all host cores.  The reference reads one rasterio window per tile on one core
// (flair_zonal_detection/dataset.py:89-117) and writes LZW windows as tiles arrive (inference.py:343-352); here the
// raster is decoded once into the page-locked array the GPU upload starts from and the class raster is encoded once from
// the array the read-back filled, both block-parallel.
//
// The LZW codec follows the TIFF 6.0 specification (MSB-first codes, 9..12 bits, ClearCode 256, EndOfInformation 257,
// the "early change" of the code width) with libtiff's choices where the specification leaves room (when the table
// is reset, how the stream ends); tests/test_rasterio.py pins it against libtiff itself through Pillow in both directions.
#include "flair_zonal_rasterio.h"

#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace {

thread_local std::string g_error;

int fail(int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_error = buf;
    return code;
}

// ------------------------------------------------------------------------------------------------ block-parallel loop
// fn(i, worker) -> empty string, or an error message (the first one wins; remaining tasks are skipped)
template <class F>
std::string parallel_for(int64_t n, int threads, F&& fn) {
    int t = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
    if (t < 1) t = 1;
    if ((int64_t)t > n) t = (int)std::max<int64_t>(n, 1);
    std::atomic<int64_t> next{0};
    std::atomic<bool> failed{false};
    std::mutex mu;
    std::string first;
    auto work = [&](int worker) {
        for (;;) {
            int64_t i = next.fetch_add(1);
            if (i >= n || failed.load()) return;
            std::string e;
            try {
                e = fn(i, worker);
            } catch (const std::bad_alloc&) {             // an exception must not leave a worker thread
                e = "out of memory";
            } catch (const std::exception& ex) {
                e = ex.what();
            }
            if (!e.empty()) {
                std::lock_guard<std::mutex> lock(mu);
                if (first.empty()) first = e;
                failed.store(true);
                return;
            }
        }
    };
    if (t == 1) {
        work(0);
    } else {
        std::vector<std::thread> pool;
        for (int k = 0; k < t; ++k) pool.emplace_back(work, k);
        for (auto& th : pool) th.join();
    }
    return first;
}

// ------------------------------------------------------------------------------------------------ LZW (TIFF flavour)
constexpr int LZW_CLEAR = 256, LZW_EOI = 257, LZW_FIRST = 258, LZW_BITS_MIN = 9, LZW_BITS_MAX = 12;
constexpr int LZW_CODE_MAX = (1 << LZW_BITS_MAX) - 1;    // 4095

int64_t lzw_bound(int64_t n) { return 2 * n + 64; }

int64_t lzw_encode(const uint8_t* src, int64_t n, uint8_t* dst, int64_t cap) {
    // dictionary: child[code][byte] -> code, a direct 4096 x 256 table (2 MB per thread, allocated once).  One dependent load
    // per input byte instead of a hash probe; only the <= 3837 entries made since the last ClearCode are non-zero, and a reset
    // zeroes exactly those (their positions are remembered), never the whole table.
    struct Dict {
        std::vector<uint16_t> child;
        std::vector<uint32_t> touched;
        Dict() : child((size_t)4096 * 256, 0) { touched.reserve(4096); }
        void reset() {
            for (uint32_t t : touched) child[t] = 0;
            touched.clear();
        }
    };
    static thread_local std::unique_ptr<Dict> dict_holder;
    if (!dict_holder) dict_holder.reset(new Dict());
    Dict& dict = *dict_holder;
    dict.reset();                                        // a previous call may have ended anywhere
    uint16_t* const child = dict.child.data();
    uint64_t acc = 0;
    int nacc = 0;
    int64_t o = 0;
    bool overflow = false;
    auto put = [&](int code, int nbits) {
        acc = (acc << nbits) | (uint64_t)code;
        nacc += nbits;
        while (nacc >= 8) {
            if (o >= cap) { overflow = true; nacc -= 8; continue; }
            dst[o++] = (uint8_t)(acc >> (nacc - 8));
            nacc -= 8;
        }
    };
    int nbits = LZW_BITS_MIN, maxcode = (1 << LZW_BITS_MIN) - 1, free_ent = LZW_FIRST;
    put(LZW_CLEAR, nbits);
    if (n > 0) {
        uint32_t ent = src[0];
        for (int64_t i = 1; i < n; ++i) {
            const uint32_t slot = ent << 8 | src[i];
            const uint32_t next = child[slot];
            if (next) { ent = next; continue; }
            put((int)ent, nbits);
            child[slot] = (uint16_t)free_ent++;
            dict.touched.push_back(slot);
            ent = src[i];
            if (free_ent == LZW_CODE_MAX - 1) {                                     // table full: start over
                dict.reset();
                free_ent = LZW_FIRST;
                put(LZW_CLEAR, nbits);
                nbits = LZW_BITS_MIN;
                maxcode = (1 << LZW_BITS_MIN) - 1;
            } else if (free_ent > maxcode) {
                ++nbits;
                maxcode = (1 << nbits) - 1;
            }
        }
        put((int)ent, nbits);
        ++free_ent;                                     // the decoder adds an entry for this code too
        if (free_ent == LZW_CODE_MAX - 1) {
            put(LZW_CLEAR, nbits);
            nbits = LZW_BITS_MIN;
        } else if (free_ent > maxcode) {
            ++nbits;
        }
    }
    put(LZW_EOI, nbits);
    if (nacc > 0) put(0, 8 - nacc);
    return overflow ? -1 : o;
}

// Every string of the table is a run of bytes that was already written: the entry made while decoding a code is "the previous
// string + the first byte of this one", and those lie next to each other in the output.  So an entry is just (where in the
// output it starts, how long it is) and decoding a code is one forward copy.
int64_t lzw_decode(const uint8_t* src, int64_t n, uint8_t* dst, int64_t cap) {
    int64_t start[4096];
    int32_t length[4096];
    uint64_t acc = 0;
    int nacc = 0;
    int64_t ip = 0, o = 0;
    int nbits = LZW_BITS_MIN, free_ent = LZW_FIRST;
    int64_t old_start = -1;
    int32_t old_len = 0;
    auto next_code = [&]() -> int {
        if (nacc < nbits) {
            if (ip + 8 <= n) {                          // refill with as many whole bytes as fit (6..8)
                uint64_t w;
                memcpy(&w, src + ip, 8);
                w = __builtin_bswap64(w);
                const int take = (64 - nacc) >> 3;
                acc = take == 8 ? w : (acc << (take * 8)) | (w >> (64 - take * 8));
                ip += take;
                nacc += take * 8;
            } else {
                while (nacc < nbits) {
                    if (ip >= n) return LZW_EOI;        // a stream cut short ends like libtiff's does
                    acc = (acc << 8) | src[ip++];
                    nacc += 8;
                }
            }
        }
        int code = (int)((acc >> (nacc - nbits)) & ((1u << nbits) - 1));
        nacc -= nbits;
        return code;
    };
    while (o < cap) {
        int code = next_code();
        if (code == LZW_EOI) break;
        if (code == LZW_CLEAR) {
            nbits = LZW_BITS_MIN;
            free_ent = LZW_FIRST;
            do { code = next_code(); } while (code == LZW_CLEAR);
            if (code == LZW_EOI) break;
            if (code > 255) return -2;
            old_start = o; old_len = 1;
            dst[o++] = (uint8_t)code;
            continue;
        }
        if (old_start < 0) return -2;                   // data must start with ClearCode
        if (code > free_ent || (code == free_ent && free_ent >= 4096)) return -2;
        // the new entry: previous string + first byte of this one = the bytes from old_start on, one longer
        if (free_ent < 4096) { start[free_ent] = old_start; length[free_ent] = old_len + 1; }
        const int64_t here = o;
        int32_t len;
        if (code < 256) {
            dst[o++] = (uint8_t)code;
            len = 1;
        } else {
            const int64_t from = start[code];
            len = length[code];
            const int64_t take = std::min<int64_t>(len, cap - o);
            if (from + len <= o) memcpy(dst + o, dst + from, (size_t)take);
            else for (int64_t k = 0; k < take; ++k) dst[o + k] = dst[from + k];      // code == free_ent: overlaps its own tail
            o += take;
        }
        old_start = here; old_len = len;
        if (free_ent < 4096) ++free_ent;
        if (free_ent > (1 << nbits) - 2 && nbits < LZW_BITS_MAX) ++nbits;            // early change
    }
    return o;
}

// ------------------------------------------------------------------------------------------------ TIFF directory
struct Dir {
    bool big = false, be = false;
    int64_t width = 0, height = 0;
    int spp = 1, bits = 8, fmt = 1, comp = 1, pred = 1, planar = 1, subfile = 0;
    bool tiled = false;
    int64_t bw = 0, bh = 0;
    std::vector<uint64_t> offsets, counts;
    std::vector<double> scale, tie, matrix;
    std::vector<uint16_t> geokeys;
    uint64_t next = 0;
};

struct File {
    int fd = -1;
    int64_t size = 0;
    ~File() { if (fd >= 0) close(fd); }
    bool read_at(uint64_t off, void* dst, size_t n) const {
        uint8_t* p = (uint8_t*)dst;
        while (n > 0) {
            ssize_t r = pread(fd, p, n, (off_t)off);
            if (r <= 0) return false;
            p += r; off += (uint64_t)r; n -= (size_t)r;
        }
        return true;
    }
};

inline uint64_t get_uint(const uint8_t* p, int n, bool be) {
    uint64_t v = 0;
    if (be) for (int i = 0; i < n; ++i) v = (v << 8) | p[i];
    else for (int i = n - 1; i >= 0; --i) v = (v << 8) | p[i];
    return v;
}

int type_size(int t) {
    switch (t) {
        case 1: case 2: case 6: case 7: return 1;
        case 3: case 8: return 2;
        case 4: case 9: case 11: case 13: return 4;
        case 5: case 10: case 12: case 16: case 17: case 18: return 8;
        default: return 0;
    }
}

// values of one directory entry as doubles (exact for every integer a TIFF offset can hold below 2^53)
bool entry_values(const File& f, const Dir& d, int type, uint64_t count, const uint8_t* value_field, std::vector<double>& out,
                  std::vector<uint64_t>* as_uint = nullptr) {
    const int ts = type_size(type);
    if (ts == 0 || type == 5 || type == 10) return false;
    const uint64_t bytes = (uint64_t)ts * count;
    const uint64_t inline_cap = d.big ? 8 : 4;
    std::vector<uint8_t> buf;
    const uint8_t* p = value_field;
    if (bytes > inline_cap) {
        if (bytes > (uint64_t)1 << 31) return false;
        buf.resize(bytes);
        uint64_t off = get_uint(value_field, d.big ? 8 : 4, d.be);
        if (!f.read_at(off, buf.data(), bytes)) return false;
        p = buf.data();
    }
    out.resize(count);
    if (as_uint) as_uint->resize(count);
    for (uint64_t i = 0; i < count; ++i) {
        const uint8_t* q = p + i * ts;
        double v;
        uint64_t u = 0;
        if (type == 12) { u = get_uint(q, 8, d.be); memcpy(&v, &u, 8); }
        else if (type == 11) { uint32_t w = (uint32_t)get_uint(q, 4, d.be); float x; memcpy(&x, &w, 4); v = x; }
        else if (type == 6) v = (int8_t)q[0];
        else if (type == 8) v = (int16_t)get_uint(q, 2, d.be);
        else if (type == 9) v = (int32_t)get_uint(q, 4, d.be);
        else if (type == 17) v = (double)(int64_t)get_uint(q, 8, d.be);
        else { u = get_uint(q, ts, d.be); v = (double)u; }
        out[i] = v;
        if (as_uint) (*as_uint)[i] = u;
    }
    return true;
}

int open_file(const char* path, File& f) {
    f.fd = open(path, O_RDONLY);
    if (f.fd < 0) return fail(-2, "%s: cannot open", path);
    struct stat st;
    if (fstat(f.fd, &st) != 0) return fail(-2, "%s: cannot stat", path);
    f.size = st.st_size;
    return 0;
}

// reads the directory at `off` (0: the first one)
int read_dir(const char* path, const File& f, uint64_t off, Dir& d) {
    uint8_t head[16];
    if (!f.read_at(0, head, 8)) return fail(-3, "%s: not a TIFF file (too short)", path);
    if (head[0] == 'I' && head[1] == 'I') d.be = false;
    else if (head[0] == 'M' && head[1] == 'M') d.be = true;
    else return fail(-3, "%s: not a TIFF file", path);
    const uint64_t magic = get_uint(head + 2, 2, d.be);
    if (magic == 42) d.big = false;
    else if (magic == 43) d.big = true;
    else return fail(-3, "%s: not a TIFF file (magic %d)", path, (int)magic);
    if (off == 0) {
        if (d.big) {
            if (!f.read_at(0, head, 16)) return fail(-3, "%s: truncated BigTIFF header", path);
            off = get_uint(head + 8, 8, d.be);
        } else {
            off = get_uint(head + 4, 4, d.be);
        }
    }
    uint8_t cnt[8];
    if (!f.read_at(off, cnt, d.big ? 8 : 2)) return fail(-3, "%s: directory outside the file", path);
    const uint64_t n = get_uint(cnt, d.big ? 8 : 2, d.be);
    if (n == 0 || n > 4096) return fail(-3, "%s: implausible directory (%llu entries)", path, (unsigned long long)n);
    const int esz = d.big ? 20 : 12;
    std::vector<uint8_t> raw(n * esz + 8);
    if (!f.read_at(off + (d.big ? 8 : 2), raw.data(), n * esz + (d.big ? 8 : 4))) return fail(-3, "%s: truncated directory", path);
    d.next = get_uint(raw.data() + n * esz, d.big ? 8 : 4, d.be);
    int64_t rows_per_strip = -1, tile_w = 0, tile_h = 0;
    std::vector<uint64_t> strip_off, strip_cnt, tile_off, tile_cnt;
    std::vector<double> v;
    for (uint64_t i = 0; i < n; ++i) {
        const uint8_t* e = raw.data() + i * esz;
        const int tag = (int)get_uint(e, 2, d.be), type = (int)get_uint(e + 2, 2, d.be);
        const uint64_t count = get_uint(e + 4, d.big ? 8 : 4, d.be);
        const uint8_t* val = e + (d.big ? 12 : 8);
        std::vector<uint64_t> u;
        switch (tag) {
            case 254: if (entry_values(f, d, type, count, val, v) && count) d.subfile = (int)v[0]; break;
            case 256: if (entry_values(f, d, type, count, val, v) && count) d.width = (int64_t)v[0]; break;
            case 257: if (entry_values(f, d, type, count, val, v) && count) d.height = (int64_t)v[0]; break;
            case 258:
                if (entry_values(f, d, type, count, val, v) && count) {
                    d.bits = (int)v[0];
                    for (double b : v) if ((int)b != d.bits) return fail(-4, "%s: bands of different depths", path);
                }
                break;
            case 259: if (entry_values(f, d, type, count, val, v) && count) d.comp = (int)v[0]; break;
            case 273: entry_values(f, d, type, count, val, v, &strip_off); break;
            case 277: if (entry_values(f, d, type, count, val, v) && count) d.spp = (int)v[0]; break;
            case 278: if (entry_values(f, d, type, count, val, v) && count) rows_per_strip = (int64_t)std::min(v[0], 4294967295.0); break;
            case 279: entry_values(f, d, type, count, val, v, &strip_cnt); break;
            case 284: if (entry_values(f, d, type, count, val, v) && count) d.planar = (int)v[0]; break;
            case 317: if (entry_values(f, d, type, count, val, v) && count) d.pred = (int)v[0]; break;
            case 322: if (entry_values(f, d, type, count, val, v) && count) tile_w = (int64_t)v[0]; break;
            case 323: if (entry_values(f, d, type, count, val, v) && count) tile_h = (int64_t)v[0]; break;
            case 324: entry_values(f, d, type, count, val, v, &tile_off); break;
            case 325: entry_values(f, d, type, count, val, v, &tile_cnt); break;
            case 339:
                if (entry_values(f, d, type, count, val, v) && count) d.fmt = (int)v[0];
                break;
            case 33550: entry_values(f, d, type, count, val, d.scale); break;
            case 33922: entry_values(f, d, type, count, val, d.tie); break;
            case 34264: entry_values(f, d, type, count, val, d.matrix); break;
            case 34735:
                if (entry_values(f, d, type, count, val, v)) {
                    d.geokeys.resize(v.size());
                    for (size_t k = 0; k < v.size(); ++k) d.geokeys[k] = (uint16_t)v[k];
                }
                break;
            default: break;
        }
    }
    if (d.width <= 0 || d.height <= 0) return fail(-3, "%s: no image size in the directory", path);
    if (tile_w > 0 && tile_h > 0) {
        d.tiled = true; d.bw = tile_w; d.bh = tile_h;
        d.offsets.swap(tile_off); d.counts.swap(tile_cnt);
    } else {
        d.tiled = false; d.bw = d.width;
        d.bh = rows_per_strip <= 0 || rows_per_strip > d.height ? d.height : rows_per_strip;
        d.offsets.swap(strip_off); d.counts.swap(strip_cnt);
    }
    if (d.comp == 32946) d.comp = 8;                    // the old Deflate code
    const int64_t nbx = (d.width + d.bw - 1) / d.bw, nby = (d.height + d.bh - 1) / d.bh;
    const int64_t planes = d.planar == 2 ? d.spp : 1;
    if ((int64_t)d.offsets.size() < nbx * nby * planes)
        return fail(-3, "%s: %lld block offsets for %lld blocks", path, (long long)d.offsets.size(), (long long)(nbx * nby * planes));
    if (d.counts.size() < d.offsets.size()) {
        if (d.comp != 1) return fail(-3, "%s: compressed blocks without byte counts", path);
        d.counts.assign(d.offsets.size(), 0);           // tolerated for raw data: the size follows from the geometry
        const uint64_t sppb = d.planar == 2 ? 1 : d.spp;
        for (auto& c : d.counts) c = (uint64_t)d.bw * d.bh * sppb * (d.bits / 8);
    }
    return 0;
}

int find_level(const char* path, const File& f, int level, Dir& d, int* n_overviews) {
    int rc = read_dir(path, f, 0, d);
    if (rc) return rc;
    // overviews: following directories flagged reduced-resolution (bit 0) and not a mask (bit 2)
    std::vector<uint64_t> ovr;
    uint64_t next = d.next;
    int guard = 0;
    while (next != 0 && guard++ < 64) {
        Dir o;
        if (read_dir(path, f, next, o)) break;
        if ((o.subfile & 1) && !(o.subfile & 4)) ovr.push_back(next);
        next = o.next;
    }
    if (n_overviews) *n_overviews = (int)ovr.size();
    if (level == 0) return 0;
    if (level < 0 || level > (int)ovr.size()) return fail(-5, "%s: no overview level %d (%d present)", path, level, (int)ovr.size());
    d = Dir();
    return read_dir(path, f, ovr[level - 1], d);
}

int check_supported(const char* path, const Dir& d) {
    if (d.bits != 8 && d.bits != 16 && d.bits != 32)
        return fail(-4, "%s: %d-bit samples are not supported (8, 16, 32)", path, d.bits);
    if (d.comp != 1 && d.comp != 5 && d.comp != 8)
        return fail(-4, "%s: TIFF compression %d is not supported (none, LZW, Deflate)", path, d.comp);
    if (d.pred != 1 && d.pred != 2 && d.pred != 3)
        return fail(-4, "%s: predictor %d is not supported (1, 2, 3)", path, d.pred);
    if (d.pred == 3 && (d.fmt != 3 || d.bits != 32))
        return fail(-4, "%s: the floating-point predictor is decoded for 32-bit float samples only", path);
    if (d.planar != 1 && d.planar != 2) return fail(-4, "%s: PlanarConfiguration %d", path, d.planar);
    if (d.spp < 1 || d.spp > 4096) return fail(-4, "%s: %d samples per pixel", path, d.spp);
    if (d.bw < 1 || d.bh < 1 || d.width > (1ll << 31) || d.height > (1ll << 31))
        return fail(-3, "%s: implausible geometry (%lld x %lld, blocks %lld x %lld)", path, (long long)d.width, (long long)d.height,
                    (long long)d.bw, (long long)d.bh);
    const double block_bytes = (double)d.bw * (double)d.bh * (d.planar == 2 ? 1 : d.spp) * (d.bits / 8);
    if (block_bytes > 1073741824.0)
        return fail(-3, "%s: a %lld x %lld block of %.0f bytes is beyond what this reader decodes in one piece (1 GiB)", path,
                    (long long)d.bw, (long long)d.bh, block_bytes);
    return 0;
}

struct Georef { bool ok = false; double left = 0, top = 0, rx = 0, ry = 0; int epsg = 0; bool geographic = false; };

Georef georef_of(const Dir& d) {
    Georef g;
    if (d.scale.size() >= 2 && d.tie.size() >= 6) {
        g.rx = d.scale[0]; g.ry = d.scale[1];
        g.left = d.tie[3] - d.tie[0] * g.rx;
        g.top = d.tie[4] + d.tie[1] * g.ry;
        g.ok = true;
    } else if (d.matrix.size() >= 16 && d.matrix[1] == 0.0 && d.matrix[4] == 0.0) {
        g.rx = d.matrix[0]; g.ry = -d.matrix[5]; g.left = d.matrix[3]; g.top = d.matrix[7];
        g.ok = true;
    }
    bool point = false;
    int proj = 0, geog = 0, model = 0;
    for (size_t i = 4; i + 3 < d.geokeys.size(); i += 4) {
        if (d.geokeys[i + 1] != 0) continue;            // value stored in another tag
        const int key = d.geokeys[i], value = d.geokeys[i + 3];
        if (key == 1024) model = value;
        else if (key == 1025) point = value == 2;
        else if (key == 2048) geog = value;
        else if (key == 3072) proj = value;
    }
    if (proj > 0 && proj < 32767) { g.epsg = proj; g.geographic = false; }
    else if (geog > 0 && geog < 32767) { g.epsg = geog; g.geographic = true; }
    if (model == 2 && proj == 0) g.geographic = true;
    if (g.ok && point) { g.left -= 0.5 * g.rx; g.top += 0.5 * g.ry; }   // PixelIsPoint: the tie point is a pixel CENTRE
    return g;
}

// ------------------------------------------------------------------------------------------------ block decode
inline void swap_samples(uint8_t* p, int64_t n, int bps) {
    if (bps == 2) for (int64_t i = 0; i < n; ++i) std::swap(p[2 * i], p[2 * i + 1]);
    else if (bps == 4) for (int64_t i = 0; i < n; ++i) { std::swap(p[4 * i], p[4 * i + 3]); std::swap(p[4 * i + 1], p[4 * i + 2]); }
}

template <class T>
void undo_hdiff(uint8_t* row, int64_t pixels, int stride) {
    T* r = (T*)row;
    for (int64_t i = stride; i < pixels * stride; ++i) r[i] = (T)(r[i] + r[i - stride]);
}

// decodes block `idx` of `rows` rows into buf (rows x bw x sppb samples)
std::string decode_block(const File& f, const Dir& d, int64_t idx, int64_t rows, std::vector<uint8_t>& comp, uint8_t* buf) {
    const int bps = d.bits / 8;
    const int sppb = d.planar == 2 ? 1 : d.spp;
    const int64_t want = rows * d.bw * sppb * bps;
    const uint64_t off = d.offsets[idx], cnt = d.counts[idx];
    if (off == 0 || cnt == 0) { memset(buf, 0, want); return ""; }                  // sparse file: a block never written
    if (off + cnt > (uint64_t)f.size) return "block " + std::to_string(idx) + " lies outside the file";
    if (d.comp == 1) {
        const uint64_t take = std::min<uint64_t>(cnt, want);
        if (!f.read_at(off, buf, take)) return "read error in block " + std::to_string(idx);
        if ((int64_t)take < want) memset(buf + take, 0, want - take);
    } else {
        comp.resize(cnt);
        if (!f.read_at(off, comp.data(), cnt)) return "read error in block " + std::to_string(idx);
        int64_t got;
        if (d.comp == 5) {
            got = lzw_decode(comp.data(), (int64_t)cnt, buf, want);
            if (got < 0) return "corrupt LZW data in block " + std::to_string(idx);
        } else {
            uLongf len = (uLongf)want;
            int z = uncompress(buf, &len, comp.data(), (uLong)cnt);
            if (z != Z_OK && z != Z_BUF_ERROR) return "corrupt Deflate data in block " + std::to_string(idx);
            got = (int64_t)len;
        }
        if (got < want) memset(buf + got, 0, want - got);
    }
    if (d.pred == 3 && d.comp != 1) {
        // floating-point predictor (TIFF Technical Note 3; elevation rasters written with GDAL's PREDICTOR=3): every row holds
        // the samples' bytes as planes, most significant byte first, differenced byte-wise along the whole row.  Undo the
        // differences, then gather each sample's bytes (into native little-endian order) -- the row's own byte order does not
        // matter, so the endian swap below is skipped.
        const int64_t row_bytes = d.bw * sppb * bps, wc = d.bw * sppb;
        std::vector<uint8_t> tmp((size_t)row_bytes);
        for (int64_t r = 0; r < rows; ++r) {
            uint8_t* row = buf + r * row_bytes;
            for (int64_t i = sppb; i < row_bytes; ++i) row[i] = (uint8_t)(row[i] + row[i - sppb]);
            memcpy(tmp.data(), row, (size_t)row_bytes);
            for (int64_t k = 0; k < wc; ++k)
                for (int b = 0; b < bps; ++b) row[bps * k + b] = tmp[(size_t)((bps - 1 - b) * wc + k)];
        }
        return "";
    }
    if (d.be && bps > 1) swap_samples(buf, rows * d.bw * sppb, bps);
    if (d.pred == 2 && d.comp != 1) {                   // libtiff applies the predictor inside the LZW / Deflate codecs only
        for (int64_t r = 0; r < rows; ++r) {
            uint8_t* row = buf + r * d.bw * sppb * bps;
            if (bps == 1) undo_hdiff<uint8_t>(row, d.bw, sppb);
            else if (bps == 2) undo_hdiff<uint16_t>(row, d.bw, sppb);
            else undo_hdiff<uint32_t>(row, d.bw, sppb);
        }
    }
    return "";
}

int read_window_impl(const char* path, const File& f, const Dir& d, int64_t row0, int64_t col0, int64_t win_h, int64_t win_w,
                     const int32_t* bands, int32_t n_bands, uint8_t* dst, int64_t band_stride, int64_t row_stride, int threads) {
    const int bps = d.bits / 8;
    std::vector<int> sel;
    if (bands == nullptr) { for (int b = 0; b < d.spp; ++b) sel.push_back(b); }
    else for (int j = 0; j < n_bands; ++j) {
        if (bands[j] < 1 || bands[j] > d.spp) return fail(-1, "%s: band index %d out of range 1..%d", path, bands[j], d.spp);
        sel.push_back(bands[j] - 1);
    }
    const int nsel = (int)sel.size();
    if (win_h < 0 || win_w < 0) return fail(-1, "negative window size");
    if (win_h == 0 || win_w == 0 || nsel == 0) return 0;
    if (row_stride == 0) row_stride = win_w * bps;
    if (band_stride == 0) band_stride = row_stride * win_h;
    if (row_stride < win_w * bps) return fail(-1, "row stride %lld < %lld bytes of a window row", (long long)row_stride, (long long)(win_w * bps));
    // the part of the window inside the raster
    const int64_t r_lo = std::max<int64_t>(row0, 0), r_hi = std::min<int64_t>(row0 + win_h, d.height);
    const int64_t c_lo = std::max<int64_t>(col0, 0), c_hi = std::min<int64_t>(col0 + win_w, d.width);
    const bool inside = r_lo == row0 && r_hi == row0 + win_h && c_lo == col0 && c_hi == col0 + win_w;
    if (!inside)                                         // boundless read: fill_value = 0 (dataset.py:113-114)
        for (int j = 0; j < nsel; ++j)
            for (int64_t r = 0; r < win_h; ++r) memset(dst + j * band_stride + r * row_stride, 0, win_w * bps);
    if (r_lo >= r_hi || c_lo >= c_hi) return 0;
    const int64_t nbx = (d.width + d.bw - 1) / d.bw, nby = (d.height + d.bh - 1) / d.bh;
    const int64_t bx0 = c_lo / d.bw, bx1 = (c_hi - 1) / d.bw, by0 = r_lo / d.bh, by1 = (r_hi - 1) / d.bh;
    struct Task { int64_t plane, by, bx; };
    std::vector<Task> tasks;
    std::vector<int> planes;
    if (d.planar == 2) { planes = sel; std::sort(planes.begin(), planes.end()); planes.erase(std::unique(planes.begin(), planes.end()), planes.end()); }
    else planes.push_back(0);
    for (int p : planes)
        for (int64_t by = by0; by <= by1; ++by)
            for (int64_t bx = bx0; bx <= bx1; ++bx) tasks.push_back({p, by, bx});
    const int sppb = d.planar == 2 ? 1 : d.spp;
    const int64_t block_bytes = d.bw * d.bh * sppb * bps;
    int t = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
    t = (int)std::max<int64_t>(1, std::min<int64_t>(t, (int64_t)tasks.size()));
    struct Scratch { std::vector<uint8_t> comp, raw; };
    std::vector<Scratch> scratch(t);
    std::string err = parallel_for((int64_t)tasks.size(), t, [&](int64_t i, int worker) -> std::string {
        Scratch& s = scratch[worker];
        if ((int64_t)s.raw.size() < block_bytes) s.raw.resize(block_bytes);
        const Task& k = tasks[i];
        const int64_t rows = d.tiled ? d.bh : std::min<int64_t>(d.bh, d.height - k.by * d.bh);
        const int64_t idx = k.plane * nbx * nby + k.by * nbx + k.bx;
        std::string e = decode_block(f, d, idx, rows, s.comp, s.raw.data());
        if (!e.empty()) return e;
        // the intersection of this block with the in-raster part of the window
        const int64_t br0 = k.by * d.bh, bc0 = k.bx * d.bw;
        const int64_t rr0 = std::max(r_lo, br0), rr1 = std::min(r_hi, br0 + rows);
        const int64_t cc0 = std::max(c_lo, bc0), cc1 = std::min(c_hi, bc0 + d.bw);
        const int64_t ncol = cc1 - cc0;
        for (int j = 0; j < nsel; ++j) {
            int comp_idx;
            if (d.planar == 2) { if (sel[j] != k.plane) continue; comp_idx = 0; }
            else comp_idx = sel[j];
            for (int64_t r = rr0; r < rr1; ++r) {
                const uint8_t* srow = s.raw.data() + ((r - br0) * d.bw + (cc0 - bc0)) * sppb * bps + comp_idx * bps;
                uint8_t* drow = dst + j * band_stride + (r - row0) * row_stride + (cc0 - col0) * bps;
                if (sppb == 1) memcpy(drow, srow, ncol * bps);
                else if (bps == 1) for (int64_t c = 0; c < ncol; ++c) drow[c] = srow[c * sppb];
                else if (bps == 2) for (int64_t c = 0; c < ncol; ++c) memcpy(drow + 2 * c, srow + 2 * c * sppb, 2);
                else for (int64_t c = 0; c < ncol; ++c) memcpy(drow + 4 * c, srow + 4 * c * sppb, 4);
            }
        }
        return "";
    });
    if (!err.empty()) return fail(-6, "%s: %s", path, err.c_str());
    return 0;
}

// ------------------------------------------------------------------------------------------------ writer
struct Level {
    int64_t width = 0, height = 0;
    const uint8_t* data = nullptr;                       // [count] planes
    int64_t band_stride = 0, row_stride = 0;
    std::vector<uint8_t> own;                            // overview pixels (dense)
    int64_t nbx = 0, nby = 0;
    std::vector<std::vector<uint8_t>> blocks;            // compressed, in TIFF block order
    std::vector<uint64_t> offsets;
    uint64_t ifd_pos = 0;
};

void put_uint(std::vector<uint8_t>& v, uint64_t x, int n) { for (int i = 0; i < n; ++i) v.push_back((uint8_t)(x >> (8 * i))); }

struct Entry { int tag, type; uint64_t count; std::vector<uint8_t> data; };

Entry ent_shorts(int tag, const std::vector<uint16_t>& x) { Entry e{tag, 3, x.size(), {}}; for (auto s : x) put_uint(e.data, s, 2); return e; }
Entry ent_long(int tag, uint64_t x) { Entry e{tag, 4, 1, {}}; put_uint(e.data, x, 4); return e; }
Entry ent_doubles(int tag, const std::vector<double>& x) {
    Entry e{tag, 12, x.size(), {}};
    for (double d : x) { uint64_t u; memcpy(&u, &d, 8); put_uint(e.data, u, 8); }
    return e;
}

void make_overview(const Level& src, Level& dst, int count, int bps, int resampling, int threads) {
    dst.width = (src.width + 1) / 2;
    dst.height = (src.height + 1) / 2;
    dst.row_stride = dst.width * bps;
    dst.band_stride = dst.row_stride * dst.height;
    dst.own.resize((size_t)dst.band_stride * count);
    dst.data = dst.own.data();
    const double rx = (double)src.width / (double)dst.width, ry = (double)src.height / (double)dst.height;
    parallel_for(dst.height * count, threads, [&](int64_t i, int) -> std::string {
        const int64_t b = i / dst.height, y = i % dst.height;
        uint8_t* out = dst.own.data() + b * dst.band_stride + y * dst.row_stride;
        const uint8_t* plane = src.data + b * src.band_stride;
        if (resampling == FZIO_OVR_MODE && bps == 1) {
            const int64_t y0 = (int64_t)std::floor(y * ry), y1 = std::min<int64_t>(src.height, (int64_t)std::ceil((y + 1) * ry));
            for (int64_t x = 0; x < dst.width; ++x) {
                const int64_t x0 = (int64_t)std::floor(x * rx), x1 = std::min<int64_t>(src.width, (int64_t)std::ceil((x + 1) * rx));
                uint8_t vals[16]; int nv = 0;
                for (int64_t yy = y0; yy < y1 && nv < 16; ++yy)
                    for (int64_t xx = x0; xx < x1 && nv < 16; ++xx) vals[nv++] = plane[yy * src.row_stride + xx];
                int best = 256, best_n = 0;
                for (int a = 0; a < nv; ++a) {
                    int c = 0;
                    for (int q = 0; q < nv; ++q) c += vals[q] == vals[a];
                    if (c > best_n || (c == best_n && vals[a] < best)) { best = vals[a]; best_n = c; }
                }
                out[x] = (uint8_t)best;
            }
        } else {
            const int64_t sy = std::min<int64_t>(src.height - 1, (int64_t)(0.5 + y * ry));
            const uint8_t* srow = plane + sy * src.row_stride;
            for (int64_t x = 0; x < dst.width; ++x) {
                const int64_t sx = std::min<int64_t>(src.width - 1, (int64_t)(0.5 + x * rx));
                memcpy(out + x * bps, srow + sx * bps, bps);
            }
        }
        return "";
    });
}

int write_impl(const char* path, const uint8_t* data, int count, int64_t height, int64_t width, int64_t band_stride,
               int64_t row_stride, const fzio_write_opts& o_in) {
    fzio_write_opts o = o_in;
    if (o.block == 0) o.block = 512;
    if (o.compression == 0) o.compression = FZIO_COMP_LZW;
    if (o.predictor == 0) o.predictor = 1;
    if (o.deflate_level == 0) o.deflate_level = 6;
    if (o.sample_format == 0) o.sample_format = FZIO_FMT_UINT;
    if (o.bits == 0) o.bits = 8;
    if (count < 1 || height < 1 || width < 1) return fail(-1, "empty raster (%d x %lld x %lld)", count, (long long)height, (long long)width);
    if (o.block < 16 || o.block % 16) return fail(-1, "block=%d: TIFF tiles are multiples of 16", o.block);
    if (o.block > 8192) return fail(-1, "block=%d: tiles above 8192 x 8192 are refused", o.block);
    if (o.overview_resampling != FZIO_OVR_NEAREST && o.overview_resampling != FZIO_OVR_MODE)
        return fail(-1, "overview_resampling=%d (0 nearest, 1 mode)", o.overview_resampling);
    if (o.overview_resampling == FZIO_OVR_MODE && o.bits != 0 && o.bits != 8 && o.overviews != 0)
        return fail(-1, "mode resampling of overviews is written for 8-bit (class) rasters only");
    if (o.compression != 1 && o.compression != 5 && o.compression != 8) return fail(-1, "compression=%d (1 none, 5 LZW, 8 Deflate)", o.compression);
    if (o.bits != 8 && o.bits != 16 && o.bits != 32) return fail(-1, "bits=%d (8, 16, 32)", o.bits);
    if (o.predictor != 1 && o.predictor != 2) return fail(-1, "predictor=%d (1, 2)", o.predictor);
    if (o.predictor == 2 && o.bits != 8) return fail(-1, "predictor 2 is written for 8-bit samples only");
    if (o.predictor == 2 && o.compression == 1) return fail(-1, "predictor 2 needs LZW or Deflate (libtiff ignores it on raw data)");
    if (o.deflate_level < 1 || o.deflate_level > 9) return fail(-1, "deflate_level=%d", o.deflate_level);
    if (o.cog && count > 1) o.pixel_interleave = 1;     // the ghost area declares one block sequence per resolution
    const int bps = o.bits / 8;
    if (row_stride == 0) row_stride = width * bps;
    if (band_stride == 0) band_stride = row_stride * height;
    const bool chunky = o.pixel_interleave && count > 1;
    const int sppb = chunky ? count : 1;
    const int planes = chunky ? 1 : count;

    // resolutions
    std::vector<std::unique_ptr<Level>> levels;
    levels.emplace_back(new Level());
    levels[0]->width = width; levels[0]->height = height; levels[0]->data = data;
    levels[0]->band_stride = band_stride; levels[0]->row_stride = row_stride;
    int n_ovr = o.overviews;
    if (n_ovr < 0) {
        n_ovr = 0;
        int64_t w = width, h = height;
        while (w > o.block || h > o.block) { w = (w + 1) / 2; h = (h + 1) / 2; ++n_ovr; }
    }
    for (int k = 0; k < n_ovr; ++k) {
        const Level& prev = *levels.back();
        if (prev.width <= 1 && prev.height <= 1) break;
        std::unique_ptr<Level> next(new Level());
        make_overview(prev, *next, count, bps, o.overview_resampling, o.threads);
        levels.push_back(std::move(next));
    }

    // compress every block of every level, one task per block
    struct Task { int level; int64_t plane, by, bx; };
    std::vector<Task> tasks;
    for (size_t l = 0; l < levels.size(); ++l) {
        Level& L = *levels[l];
        L.nbx = (L.width + o.block - 1) / o.block;
        L.nby = (L.height + o.block - 1) / o.block;
        L.blocks.resize((size_t)(L.nbx * L.nby * planes));
        for (int p = 0; p < planes; ++p)
            for (int64_t by = 0; by < L.nby; ++by)
                for (int64_t bx = 0; bx < L.nbx; ++bx) tasks.push_back({(int)l, p, by, bx});
    }
    const int64_t raw_bytes = (int64_t)o.block * o.block * sppb * bps;
    int wt = o.threads > 0 ? o.threads : (int)std::thread::hardware_concurrency();
    wt = (int)std::max<int64_t>(1, std::min<int64_t>(wt, (int64_t)tasks.size()));
    std::vector<std::vector<uint8_t>> scratch(wt);
    std::string err = parallel_for((int64_t)tasks.size(), wt, [&](int64_t i, int worker) -> std::string {
        std::vector<uint8_t>& raw = scratch[worker];
        if ((int64_t)raw.size() < raw_bytes) raw.resize(raw_bytes);
        const Task& k = tasks[i];
        Level& L = *levels[k.level];
        const int64_t r0 = k.by * o.block, c0 = k.bx * o.block;
        const int64_t rows = std::min<int64_t>(o.block, L.height - r0), cols = std::min<int64_t>(o.block, L.width - c0);
        if (rows < o.block || cols < o.block) memset(raw.data(), 0, raw_bytes);
        for (int64_t r = 0; r < rows; ++r) {
            uint8_t* out = raw.data() + r * o.block * sppb * bps;
            if (!chunky) {
                memcpy(out, L.data + k.plane * L.band_stride + (r0 + r) * L.row_stride + c0 * bps, cols * bps);
            } else {
                for (int b = 0; b < count; ++b) {
                    const uint8_t* in = L.data + b * L.band_stride + (r0 + r) * L.row_stride + c0 * bps;
                    if (bps == 1) for (int64_t c = 0; c < cols; ++c) out[c * count + b] = in[c];
                    else for (int64_t c = 0; c < cols; ++c) memcpy(out + (c * count + b) * bps, in + c * bps, bps);
                }
            }
        }
        if (o.predictor == 2)
            for (int64_t r = 0; r < o.block; ++r) {
                uint8_t* row = raw.data() + r * o.block * sppb;
                for (int64_t c = (int64_t)o.block * sppb - 1; c >= sppb; --c) row[c] = (uint8_t)(row[c] - row[c - sppb]);
            }
        std::vector<uint8_t>& out = L.blocks[(size_t)(k.plane * L.nbx * L.nby + k.by * L.nbx + k.bx)];
        if (o.compression == 1) {
            out.assign(raw.data(), raw.data() + raw_bytes);
        } else if (o.compression == 5) {
            out.resize((size_t)lzw_bound(raw_bytes));
            int64_t n = lzw_encode(raw.data(), raw_bytes, out.data(), (int64_t)out.size());
            if (n < 0) return "LZW output overflow";
            out.resize((size_t)n);
            out.shrink_to_fit();
        } else {
            uLongf cap = compressBound((uLong)raw_bytes);
            out.resize(cap);
            if (compress2(out.data(), &cap, raw.data(), (uLong)raw_bytes, o.deflate_level) != Z_OK) return "zlib compress2 failed";
            out.resize(cap);
            out.shrink_to_fit();
        }
        return "";
    });
    if (!err.empty()) return fail(-6, "%s: %s", path, err.c_str());

    // size decides classic / BigTIFF
    uint64_t payload = 0, n_blocks = 0;
    for (auto& L : levels) for (auto& b : L->blocks) { payload += b.size() + (o.cog ? 8 : 0); ++n_blocks; }
    bool big = o.bigtiff > 0;
    const uint64_t estimate = payload + n_blocks * 16 + levels.size() * 4096 + 4096;
    if (o.bigtiff == 0 && estimate >= 4000000000ull) big = true;
    if (o.bigtiff < 0 && estimate >= 4294967295ull) return fail(-7, "%s: %llu bytes do not fit a classic TIFF", path, (unsigned long long)estimate);

    // directories: header [+ ghost area], then per level the IFD followed by its out-of-line values
    std::vector<uint8_t> head;
    head.push_back('I'); head.push_back('I');
    if (big) { put_uint(head, 43, 2); put_uint(head, 8, 2); put_uint(head, 0, 2); put_uint(head, 0, 8); }
    else { put_uint(head, 42, 2); put_uint(head, 0, 4); }
    if (o.cog) {
        const std::string rest = "LAYOUT=IFDS_BEFORE_DATA\nBLOCK_ORDER=ROW_MAJOR\nBLOCK_LEADER=SIZE_AS_UINT4\n"
                                 "BLOCK_TRAILER=LAST_4_BYTES_REPEATED\nKNOWN_INCOMPATIBLE_EDITION=NO\n ";
        char first[64];
        snprintf(first, sizeof first, "GDAL_STRUCTURAL_METADATA_SIZE=%06d bytes\n", (int)rest.size());
        for (const char* p = first; *p; ++p) head.push_back((uint8_t)*p);
        for (char c : rest) head.push_back((uint8_t)c);
        if (head.size() % 2) head.push_back(0);
    }
    const int off_size = big ? 8 : 4, inline_cap = big ? 8 : 4;
    auto build_ifd = [&](size_t l, uint64_t pos, uint64_t next_ifd, bool final_pass) -> std::vector<uint8_t> {
        Level& L = *levels[l];
        std::vector<Entry> ents;
        if (l > 0) ents.push_back(ent_long(254, 1));
        ents.push_back(ent_long(256, (uint64_t)L.width));
        ents.push_back(ent_long(257, (uint64_t)L.height));
        ents.push_back(ent_shorts(258, std::vector<uint16_t>(count, (uint16_t)o.bits)));
        ents.push_back(ent_shorts(259, {(uint16_t)o.compression}));
        const bool rgb = chunky && o.bits == 8 && (count == 3 || count == 4);
        ents.push_back(ent_shorts(262, {(uint16_t)(rgb ? 2 : 1)}));
        ents.push_back(ent_shorts(277, {(uint16_t)count}));
        ents.push_back(ent_shorts(284, {(uint16_t)(chunky || count == 1 ? 1 : 2)}));
        if (o.predictor == 2) ents.push_back(ent_shorts(317, {2}));
        ents.push_back(ent_long(322, (uint64_t)o.block));
        ents.push_back(ent_long(323, (uint64_t)o.block));
        {
            Entry e{324, big ? 16 : 4, L.blocks.size(), {}};
            for (size_t b = 0; b < L.blocks.size(); ++b) put_uint(e.data, final_pass ? L.offsets[b] : 0, off_size);
            ents.push_back(std::move(e));
            Entry c{325, 4, L.blocks.size(), {}};
            for (auto& b : L.blocks) put_uint(c.data, b.size(), 4);
            ents.push_back(std::move(c));
        }
        const int extra = rgb ? count - 3 : count - 1;
        if (extra > 0) ents.push_back(ent_shorts(338, std::vector<uint16_t>(extra, 0)));
        ents.push_back(ent_shorts(339, std::vector<uint16_t>(count, (uint16_t)o.sample_format)));
        if (l == 0 && o.has_georef) {
            ents.push_back(ent_doubles(33550, {o.res, o.res, 0.0}));
            ents.push_back(ent_doubles(33922, {0.0, 0.0, 0.0, o.left, o.top, 0.0}));
            std::vector<uint16_t> keys = {1, 1, 0, 0, 1024, 0, 1, (uint16_t)(o.geographic ? 2 : 1), 1025, 0, 1, 1};
            if (o.epsg > 0) { keys.push_back(o.geographic ? 2048 : 3072); keys.push_back(0); keys.push_back(1); keys.push_back((uint16_t)o.epsg); }
            keys[3] = (uint16_t)((keys.size() - 4) / 4);
            ents.push_back(ent_shorts(34735, keys));
        }
        std::sort(ents.begin(), ents.end(), [](const Entry& a, const Entry& b) { return a.tag < b.tag; });
        std::vector<uint8_t> ifd, extra_data;
        const uint64_t ifd_bytes = (big ? 8 : 2) + ents.size() * (big ? 20 : 12) + off_size;
        put_uint(ifd, ents.size(), big ? 8 : 2);
        for (auto& e : ents) {
            put_uint(ifd, (uint64_t)e.tag, 2);
            put_uint(ifd, (uint64_t)e.type, 2);
            put_uint(ifd, e.count, big ? 8 : 4);
            if ((int)e.data.size() <= inline_cap) {
                for (int b = 0; b < inline_cap; ++b) ifd.push_back(b < (int)e.data.size() ? e.data[b] : 0);
            } else {
                put_uint(ifd, pos + ifd_bytes + extra_data.size(), off_size);
                extra_data.insert(extra_data.end(), e.data.begin(), e.data.end());
                if (extra_data.size() % 2) extra_data.push_back(0);
            }
        }
        put_uint(ifd, next_ifd, off_size);
        ifd.insert(ifd.end(), extra_data.begin(), extra_data.end());
        return ifd;
    };
    // pass 1: sizes -> positions
    uint64_t pos = head.size();
    std::vector<uint64_t> ifd_size(levels.size());
    for (size_t l = 0; l < levels.size(); ++l) {
        levels[l]->ifd_pos = pos;
        ifd_size[l] = build_ifd(l, pos, 0, false).size();
        pos += ifd_size[l];
        if (pos % 2) ++pos;
    }
    // data: COG puts the smallest overview first, the full resolution last
    std::vector<size_t> order;
    if (o.cog) for (size_t l = levels.size(); l-- > 0;) order.push_back(l);
    else for (size_t l = 0; l < levels.size(); ++l) order.push_back(l);
    const uint64_t data_start = pos;
    for (size_t l : order) {
        Level& L = *levels[l];
        L.offsets.resize(L.blocks.size());
        for (size_t b = 0; b < L.blocks.size(); ++b) {
            if (o.cog) pos += 4;
            L.offsets[b] = pos;
            pos += L.blocks[b].size();
            if (o.cog) pos += 4;
        }
    }
    if (!big && pos >= 4294967295ull) return fail(-7, "%s: %llu bytes do not fit a classic TIFF", path, (unsigned long long)pos);
    // pass 2: write
    FILE* fp = fopen(path, "wb");
    if (!fp) return fail(-2, "%s: cannot create", path);
    std::vector<char> iobuf(8 << 20);
    setvbuf(fp, iobuf.data(), _IOFBF, iobuf.size());
    {   // patch the first-IFD offset into the header
        uint64_t first = levels[0]->ifd_pos;
        for (int b = 0; b < off_size; ++b) head[(big ? 8 : 4) + b] = (uint8_t)(first >> (8 * b));
    }
    bool ok = fwrite(head.data(), 1, head.size(), fp) == head.size();
    uint64_t at = head.size();
    for (size_t l = 0; l < levels.size() && ok; ++l) {
        const uint64_t next = l + 1 < levels.size() ? levels[l + 1]->ifd_pos : 0;
        std::vector<uint8_t> ifd = build_ifd(l, levels[l]->ifd_pos, next, true);
        ok = ok && fwrite(ifd.data(), 1, ifd.size(), fp) == ifd.size();
        at += ifd.size();
        if (at % 2) { ok = ok && fputc(0, fp) != EOF; ++at; }
    }
    if (ok && at != data_start) { fclose(fp); return fail(-8, "%s: internal layout error (%llu != %llu)", path, (unsigned long long)at, (unsigned long long)data_start); }
    for (size_t l : order) {
        Level& L = *levels[l];
        for (size_t b = 0; b < L.blocks.size() && ok; ++b) {
            const std::vector<uint8_t>& blk = L.blocks[b];
            if (o.cog) {
                uint8_t lead[4];
                for (int q = 0; q < 4; ++q) lead[q] = (uint8_t)((uint64_t)blk.size() >> (8 * q));
                ok = ok && fwrite(lead, 1, 4, fp) == 4;
            }
            ok = ok && fwrite(blk.data(), 1, blk.size(), fp) == blk.size();
            if (o.cog) {
                uint8_t trail[4] = {0, 0, 0, 0};
                for (int q = 0; q < 4; ++q) if (blk.size() >= 4) trail[q] = blk[blk.size() - 4 + q];
                ok = ok && fwrite(trail, 1, 4, fp) == 4;
            }
        }
    }
    ok = (fclose(fp) == 0) && ok;
    if (!ok) return fail(-2, "%s: write error", path);
    return 0;
}

}  // namespace

// ================================================================================================ C ABI
extern "C" {

int fzio_abi_version(void) { return FZIO_ABI_VERSION; }
const char* fzio_last_error(void) { return g_error.c_str(); }

#define FZIO_GUARD_BEGIN try {
#define FZIO_GUARD_END                                                         \
    }                                                                          \
    catch (const std::bad_alloc&) { return fail(-9, "out of memory"); }        \
    catch (const std::exception& ex) { return fail(-9, "%s", ex.what()); }

int fzio_tiff_info(const char* path, int level, fzio_info* out) {
    FZIO_GUARD_BEGIN
    if (!path || !out) return fail(-1, "fzio_tiff_info: null argument");
    File f;
    int rc = open_file(path, f);
    if (rc) return rc;
    Dir d;
    int n_ovr = 0;
    rc = find_level(path, f, level, d, &n_ovr);
    if (rc) return rc;
    memset(out, 0, sizeof *out);
    out->width = d.width; out->height = d.height; out->count = d.spp; out->bits = d.bits; out->sample_format = d.fmt;
    out->compression = d.comp; out->predictor = d.pred; out->planar = d.planar; out->tiled = d.tiled ? 1 : 0;
    out->block_w = (int32_t)d.bw; out->block_h = (int32_t)d.bh; out->bigtiff = d.big ? 1 : 0; out->big_endian = d.be ? 1 : 0;
    out->overviews = n_ovr;
    Georef g = georef_of(d);
    out->has_georef = g.ok ? 1 : 0; out->epsg = g.epsg; out->geographic = g.geographic ? 1 : 0;
    out->left = g.left; out->top = g.top; out->res_x = g.rx; out->res_y = g.ry;
    return 0;
    FZIO_GUARD_END
}

int fzio_read_window(const char* path, int level, int64_t row0, int64_t col0, int64_t win_h, int64_t win_w,
                     const int32_t* bands, int32_t n_bands, void* dst, int64_t dst_band_stride, int64_t dst_row_stride,
                     int32_t threads) {
    FZIO_GUARD_BEGIN
    if (!path || (!dst && win_h > 0 && win_w > 0)) return fail(-1, "fzio_read_window: null argument");
    File f;
    int rc = open_file(path, f);
    if (rc) return rc;
    Dir d;
    rc = find_level(path, f, level, d, nullptr);
    if (rc) return rc;
    rc = check_supported(path, d);
    if (rc) return rc;
    return read_window_impl(path, f, d, row0, col0, win_h, win_w, bands, n_bands, (uint8_t*)dst, dst_band_stride, dst_row_stride, threads);
    FZIO_GUARD_END
}

int fzio_write_geotiff(const char* path, const void* data, int32_t count, int64_t height, int64_t width,
                       int64_t band_stride, int64_t row_stride, const fzio_write_opts* opts) {
    FZIO_GUARD_BEGIN
    if (!path || !data) return fail(-1, "fzio_write_geotiff: null argument");
    fzio_write_opts o;
    memset(&o, 0, sizeof o);
    if (opts) o = *opts;
    return write_impl(path, (const uint8_t*)data, count, height, width, band_stride, row_stride, o);
    FZIO_GUARD_END
}

int fzio_convert_to_cog(const char* src_path, const char* dst_path, int32_t threads) {
    FZIO_GUARD_BEGIN
    if (!src_path || !dst_path) return fail(-1, "fzio_convert_to_cog: null argument");
    File f;
    int rc = open_file(src_path, f);
    if (rc) return rc;
    Dir d;
    rc = find_level(src_path, f, 0, d, nullptr);
    if (rc) return rc;
    rc = check_supported(src_path, d);
    if (rc) return rc;
    const int bps = d.bits / 8;
    std::vector<uint8_t> pixels((size_t)d.spp * d.height * d.width * bps);
    rc = read_window_impl(src_path, f, d, 0, 0, d.height, d.width, nullptr, 0, pixels.data(), 0, 0, threads);
    if (rc) return rc;
    Georef g = georef_of(d);
    fzio_write_opts o;
    memset(&o, 0, sizeof o);
    o.block = 512; o.compression = FZIO_COMP_LZW; o.overviews = -1; o.overview_resampling = FZIO_OVR_NEAREST; o.cog = 1;
    o.threads = threads; o.bits = d.bits; o.sample_format = d.fmt;
    if (g.ok) {
        if (std::fabs(g.rx - g.ry) > 1e-9 * std::max(g.rx, g.ry)) return fail(-4, "%s: non-square pixels (%g x %g)", src_path, g.rx, g.ry);
        o.has_georef = 1; o.left = g.left; o.top = g.top; o.res = g.rx; o.epsg = g.epsg; o.geographic = g.geographic ? 1 : 0;
    }
    return write_impl(dst_path, pixels.data(), d.spp, d.height, d.width, 0, 0, o);
    FZIO_GUARD_END
}

int64_t fzio_lzw_bound(int64_t n) { return lzw_bound(n); }
int64_t fzio_lzw_encode(const uint8_t* src, int64_t n, uint8_t* dst, int64_t dst_cap) {
    if (n < 0 || (!src && n > 0) || !dst) return fail(-1, "fzio_lzw_encode: bad argument");
    int64_t r = lzw_encode(src, n, dst, dst_cap);
    if (r < 0) return fail(-1, "fzio_lzw_encode: output buffer too small (%lld for %lld bytes)", (long long)dst_cap, (long long)n);
    return r;
}
int64_t fzio_lzw_decode(const uint8_t* src, int64_t n, uint8_t* dst, int64_t dst_cap) {
    if (n < 0 || (!src && n > 0) || (!dst && dst_cap > 0)) return fail(-1, "fzio_lzw_decode: bad argument");
    int64_t r = lzw_decode(src, n, dst, dst_cap);
    if (r < 0) return fail(-2, "fzio_lzw_decode: corrupt stream");
    return r;
}

}  // extern "C"
