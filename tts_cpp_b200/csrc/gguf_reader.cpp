// gguf_reader.cpp -- minimal GGUF v2/v3 reader (host side of the drop-in boundary).
//
// Replaces, for the hot path, what runner_from_file does with llama_mmap + gguf_init_from_file + the tensor iterator
// (reference src/models/loaders.cpp:34-95, ggml-patches/llama-mmap.cpp, ggml-patches/ggml-iterator.h): mmap the file, walk the
// metadata and tensor directory, and stream every (name, type, ne, data) to Kokoro::assign -- the analogue of
// tts_generation_runner::assign_weight -- followed by prepare() (prepare_post_load).  Written from the GGUF format
// specification; no ggml code is linked.
#include "kokoro.h"
#include "dac.h"
#include "orpheus.h"
#include "parler.h"
#include "dia.h"
#include "t5.h"

#include <exception>
#include <functional>

#include <cstdio>
#include <cstring>
#include <fcntl.h>
#include <string>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

namespace b2 {
namespace {

// every length read from the file is checked against the bytes that are left (by subtraction: a huge value cannot wrap a pointer past the check)
struct Cursor {
    const uint8_t * p; const uint8_t * end; bool ok = true;
    size_t left() const { return (size_t) (end - p); }
    template <class T> T rd() {
        T v{};
        if (!ok || sizeof(T) > left()) { ok = false; return v; }
        memcpy(&v, p, sizeof(T)); p += sizeof(T);
        return v;
    }
    bool skip(size_t n) { if (!ok || n > left()) { ok = false; return false; } p += n; return true; }
    std::string str() {
        uint64_t n = rd<uint64_t>();
        if (!ok || n > (uint64_t) left()) { ok = false; return std::string(); }
        std::string s((const char *) p, (size_t) n); p += n;
        return s;
    }
};

size_t scalar_size(uint32_t t) {
    switch (t) {
        case 0: case 1: case 7: return 1;
        case 2: case 3: return 2;
        case 4: case 5: case 6: return 4;
        case 10: case 11: case 12: return 8;
        default: return 0;
    }
}

// skips (or captures) one metadata value of GGUF type `t`
bool read_value(Cursor & c, uint32_t t, uint32_t * u32_out, std::string * str_out, std::vector<std::string> * arr_out) {
    if (t == 8) { std::string s = c.str(); if (str_out) *str_out = s; return c.ok; }
    if (t == 9) {
        uint32_t et = c.rd<uint32_t>(); uint64_t n = c.rd<uint64_t>();
        if (et != 8) {                                            // scalars: one bounds check for the whole array
            const size_t sz = scalar_size(et);
            if (!sz || n > (uint64_t) c.left() / sz) { c.ok = false; return false; }
            return c.skip((size_t) n * sz);
        }
        if (n > (uint64_t) c.left() / 8) { c.ok = false; return false; }      // every string costs at least its 8-byte length
        for (uint64_t i = 0; i < n && c.ok; i++) { std::string s = c.str(); if (arr_out && c.ok) arr_out->push_back(s); }
        return c.ok;
    }
    size_t sz = scalar_size(t);
    if (!sz) return false;
    if ((t == 4 || t == 6) && u32_out) { *u32_out = c.rd<uint32_t>(); return c.ok; }   // f32 values are captured as their bit pattern
    return c.skip(sz);
}

}  // namespace

namespace {

// walks the file: u32 metadata into `kv`, every F32/F16 tensor whose name starts with `prefix` to `on_tensor`; `arch` (if given) must
// match general.architecture
int read_gguf(const char * path, const char * prefix, const char * arch_required, std::map<std::string, uint32_t> & kv,
              const std::function<int(const char *, int, int, const int64_t *, const void *, size_t)> & on_tensor) {
    int fd = open(path, O_RDONLY);
    if (fd < 0) { set_error("cannot open '%s'", path); return 1; }
    struct stat st;
    if (fstat(fd, &st) != 0 || st.st_size < 24) { close(fd); set_error("%s: cannot stat, or too short to be a GGUF file", path); return 1; }
    void * map = mmap(nullptr, (size_t) st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
    close(fd);
    if (map == MAP_FAILED) { set_error("mmap of '%s' failed", path); return 1; }
    const uint8_t * base = (const uint8_t *) map;
    Cursor c{base, base + st.st_size};
    int rc = 1;
    try {
    do {
        if (c.rd<uint32_t>() != 0x46554747u) { set_error("%s: not a GGUF file", path); break; }
        uint32_t ver = c.rd<uint32_t>();
        if (ver < 2 || ver > 3) { set_error("%s: unsupported GGUF version %u", path, ver); break; }
        uint64_t n_tensors = c.rd<uint64_t>(), n_kv = c.rd<uint64_t>();
        // counts from an untrusted header: a key / value record is at least 12 bytes, a tensor record at least 24
        if (!c.ok || n_kv > (uint64_t) c.left() / 12 || n_tensors > (uint64_t) c.left() / 24) { set_error("%s: corrupt GGUF header (counts exceed the file size)", path); break; }
        uint32_t alignment = 32;
        std::string arch;
        bool bad = false;
        for (uint64_t i = 0; i < n_kv && !bad; i++) {
            std::string key = c.str();
            uint32_t t = c.rd<uint32_t>();
            uint32_t u = 0; bool is_u32 = (t == 4); std::string s;
            const bool is_f32 = (t == 6);
            if (!read_value(c, t, &u, &s, nullptr)) { bad = true; break; }
            if (is_u32) { kv[key] = u; if (key == "general.alignment") alignment = u; }
            if (alignment == 0 || alignment > (1u << 20) || (alignment & (alignment - 1))) { bad = true; break; }
            if (is_f32) kv[key + "#f32"] = u;   // e.g. dia.cfg_scale: the bit pattern of the float under a suffixed key
            if (key == "general.architecture") arch = s;
        }
        if (bad || !c.ok) { set_error("%s: corrupt GGUF metadata", path); break; }
        if (arch_required && arch != arch_required) { set_error("%s: general.architecture is '%s', this loader handles '%s'", path, arch.c_str(), arch_required); break; }
        struct TI { std::string name; int nd; int64_t ne[4]; uint32_t type; uint64_t off; };
        std::vector<TI> tis((size_t) n_tensors);
        for (auto & t : tis) {
            t.name = c.str(); t.nd = (int) c.rd<uint32_t>();
            if (t.nd > 4) { bad = true; break; }
            for (int d = 0; d < 4; d++) t.ne[d] = 1;
            for (int d = 0; d < t.nd; d++) t.ne[d] = (int64_t) c.rd<uint64_t>();
            t.type = c.rd<uint32_t>(); t.off = c.rd<uint64_t>();
            if (!c.ok) break;
        }
        if (bad || !c.ok) { set_error("%s: corrupt GGUF tensor directory", path); break; }
        size_t data0 = (size_t) (c.p - base);
        data0 = (data0 + alignment - 1) / alignment * alignment;
        for (auto & t : tis) {
            if (t.name.rfind(prefix, 0) != 0) continue;
            const size_t fsize = (size_t) st.st_size;
            if (data0 > fsize || t.off > fsize - data0) { set_error("%s: tensor '%s' starts past the end of the file", path, t.name.c_str()); bad = true; break; }
            const size_t avail = fsize - data0 - (size_t) t.off;
            int64_t n = 1;                                        // element count with overflow / sign checks: no tensor has more elements than the file has bytes x 2 (Q4_0)
            for (int d = 0; d < t.nd && !bad; d++) { if (t.ne[d] <= 0 || t.ne[d] > (int64_t) (2 * fsize) || n > (int64_t) (2 * fsize) / t.ne[d]) bad = true; else n *= t.ne[d]; }
            if (bad) { set_error("%s: tensor '%s' has an impossible shape", path, t.name.c_str()); break; }
            // F32, F16, or 32-value blocks: Q4_0 (18 bytes), Q5_0 (22), Q8_0 (34) -- ggml-common.h; the model decides whether it accepts a quantised tensor
            const size_t blk = t.type == 2 ? 18 : t.type == 6 ? 22 : t.type == 8 ? 34 : 0;
            size_t nbytes = t.type == 0 ? (size_t) n * 4 : t.type == 1 ? (size_t) n * 2 : 0;
            if (blk) { if (t.ne[0] % 32) { set_error("%s: quantised tensor '%s' has a row length that is not a multiple of 32", path, t.name.c_str()); bad = true; break; } nbytes = (size_t) n / 32 * blk; }
            if (!nbytes) { set_error("%s: tensor '%s' has ggml type %u; F32, F16, Q4_0, Q5_0 and Q8_0 are supported", path, t.name.c_str(), t.type); bad = true; break; }
            if (nbytes > avail) { set_error("%s: tensor '%s' runs past the end of the file", path, t.name.c_str()); bad = true; break; }
            if (on_tensor(t.name.c_str(), (int) t.type, t.nd, t.ne, base + data0 + t.off, nbytes)) { bad = true; break; }
        }
        if (bad) break;
        rc = 0;
    } while (false);
    } catch (const std::exception & e) { set_error("%s: %s while reading the file", path, e.what()); rc = 1; }      // (bad_alloc on a hostile header: never across the extern "C" boundary)
    munmap(map, (size_t) st.st_size);
    return rc;
}

}  // namespace

int load_gguf_into(Kokoro * m, const char * path) {
    if (read_gguf(path, "kokoro.", "kokoro", m->kv, [m](const char * n, int ty, int nd, const int64_t * ne, const void * d, size_t nb) { return m->assign(n, ty, nd, ne, d, nb); })) return 1;
    return m->prepare();
}

// the codec decoder's tensors live under "audio_encoder." in Parler / Dia GGUFs (reference src/decoder/dac_model.h:40-44)
int load_gguf_into(Dac * m, const char * path) {
    if (read_gguf(path, "audio_encoder.", nullptr, m->kv, [m](const char * n, int ty, int nd, const int64_t * ne, const void * d, size_t nb) { return m->assign(n, ty, nd, ne, d, nb); })) return 1;
    return m->prepare();
}

// the SNAC decoder's tensors live under "snac." in Orpheus GGUFs (reference src/decoder/snac_model.h:41-46, src/models/orpheus/model.cpp:440-441)
int load_gguf_into(Snac * m, const char * path) {
    if (read_gguf(path, "snac.", nullptr, m->kv, [m](const char * n, int ty, int nd, const int64_t * ne, const void * d, size_t nb) { return m->assign(n, ty, nd, ne, d, nb); })) return 1;
    return m->prepare();
}

int load_gguf_into(Orpheus * m, const char * path) {
    if (read_gguf(path, "orpheus.", "orpheus", m->kv, [m](const char * n, int ty, int nd, const int64_t * ne, const void * d, size_t nb) { return m->assign(n, ty, nd, ne, d, nb); })) return 1;
    return m->prepare();
}

// Parler's decoder tensors live under "decoder." (reference src/models/parler/model.cpp:3-28 assign_to_decoder); its DAC under "audio_encoder."
int load_gguf_into(Parler * m, const char * path) {
    if (read_gguf(path, "decoder.", "parler-tts", m->kv, [m](const char * n, int ty, int nd, const int64_t * ne, const void * d, size_t nb) { return m->assign(n, ty, nd, ne, d, nb); })) return 1;
    return m->prepare();
}

// Dia's encoder + decoder tensors live under "dia." (reference src/models/dia/model.cpp:3-132); its DAC under "audio_encoder."
int load_gguf_into(Dia * m, const char * path) {
    if (read_gguf(path, "dia.", "dia", m->kv, [m](const char * n, int ty, int nd, const int64_t * ne, const void * d, size_t nb) { return m->assign(n, ty, nd, ne, d, nb); })) return 1;
    return m->prepare();
}

// the T5 conditional-prompt encoder is its own GGUF (--text-encoder-path; reference src/models/parler/t5/model.cpp:373-400), tensors under "t5encoder."
int load_gguf_into(T5 * m, const char * path) {
    if (read_gguf(path, "t5encoder.", nullptr, m->kv, [m](const char * n, int ty, int nd, const int64_t * ne, const void * d, size_t nb) { return m->assign(n, ty, nd, ne, d, nb); })) return 1;
    return m->prepare();
}

}  // namespace b2
