// safetensors.cpp -- see safetensors.h.
#include "safetensors.h"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cmath>
#include <cstring>

namespace pk {
namespace {

// A just-big-enough JSON reader for the safetensors header.
struct JsonCur {
    const char *p, *e;
    bool fail = false;
    int depth = 0;                 // nesting of skip_value (bounded: a hostile header must not overflow the stack)
    void ws() {
        while (p < e && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r')) ++p;
    }
    bool eat(char c) {
        ws();
        if (p < e && *p == c) {
            ++p;
            return true;
        }
        return false;
    }
    std::string str() {
        ws();
        std::string s;
        if (p >= e || *p != '"') {
            fail = true;
            return s;
        }
        ++p;
        while (p < e && *p != '"') {
            if (*p == '\\' && p + 1 < e) {
                ++p;
                switch (*p) {
                case 'n': s += '\n'; break;
                case 't': s += '\t'; break;
                case 'u':
                    s += '?';
                    if (e - p < 5) { fail = true; return s; }      // \uXXXX must lie inside the header
                    p += 4;
                    break;
                default: s += *p;
                }
                ++p;
            } else {
                s += *p++;
            }
        }
        if (p >= e) fail = true;
        ++p;
        return s;
    }
    // JSON number inside [p, e): copied to a NUL-terminated buffer first (the mmap is not terminated)
    double num() {
        ws();
        char buf[64];
        size_t n = 0;
        while (p + n < e && n < sizeof(buf) - 1 && (strchr("+-.eE", p[n]) || (p[n] >= '0' && p[n] <= '9'))) {
            buf[n] = p[n];
            ++n;
        }
        buf[n] = 0;
        char *end = nullptr;
        double v = strtod(buf, &end);
        if (n == 0 || end != buf + n) fail = true;
        p += n;
        return v;
    }
    // non-negative integer (shape entries, data offsets): digits only, overflow-checked
    uint64_t uint() {
        ws();
        uint64_t v = 0;
        const char *s0 = p;
        while (p < e && *p >= '0' && *p <= '9') {
            const uint64_t d = (uint64_t)(*p - '0');
            if (v > (UINT64_MAX - d) / 10) { fail = true; return 0; }
            v = v * 10 + d;
            ++p;
        }
        if (p == s0) fail = true;          // (a '-' or any other character: not a valid size)
        return v;
    }
    void skip_value() {
        struct Depth { int &d; Depth(int &x) : d(x) { ++d; } ~Depth() { --d; } } guard(depth);
        if (depth > 64) { fail = true; return; }
        ws();
        if (p >= e) { fail = true; return; }
        if (*p == '"') { str(); return; }
        if (*p == '{') {
            ++p;
            if (eat('}')) return;
            do {
                str();
                if (!eat(':')) { fail = true; return; }
                skip_value();
            } while (!fail && eat(','));
            if (!eat('}')) fail = true;
            return;
        }
        if (*p == '[') {
            ++p;
            if (eat(']')) return;
            do skip_value(); while (!fail && eat(','));
            if (!eat(']')) fail = true;
            return;
        }
        if (e - p >= 4 && !strncmp(p, "true", 4)) { p += 4; return; }
        if (e - p >= 5 && !strncmp(p, "false", 5)) { p += 5; return; }
        if (e - p >= 4 && !strncmp(p, "null", 4)) { p += 4; return; }
        num();
    }
};

float half_to_float(uint16_t h) {
    uint32_t sign = (h >> 15) & 1, exp = (h >> 10) & 0x1f, man = h & 0x3ff, f;
    if (exp == 0) {
        if (man == 0) {
            f = sign << 31;
        } else {
            exp = 127 - 15 + 1;
            while (!(man & 0x400)) { man <<= 1; --exp; }
            man &= 0x3ff;
            f = (sign << 31) | (exp << 23) | (man << 13);
        }
    } else if (exp == 31) {
        f = (sign << 31) | 0x7f800000u | (man << 13);
    } else {
        f = (sign << 31) | ((exp + 127 - 15) << 23) | (man << 13);
    }
    float out;
    memcpy(&out, &f, 4);
    return out;
}

}  // namespace

SafeTensors::~SafeTensors() {
    if (map_) munmap(const_cast<uint8_t *>(map_), map_len_);
}

bool SafeTensors::open(const std::string &path, std::string &err) {
    int fd = ::open(path.c_str(), O_RDONLY);
    if (fd < 0) {
        err = "cannot open weights file: " + path;
        return false;
    }
    struct stat stt;
    if (fstat(fd, &stt) != 0 || stt.st_size < 8) {
        ::close(fd);
        err = "weights file too small: " + path;
        return false;
    }
    map_len_ = (size_t)stt.st_size;
    void *m = mmap(nullptr, map_len_, PROT_READ, MAP_PRIVATE, fd, 0);
    ::close(fd);
    if (m == MAP_FAILED) {
        map_ = nullptr;
        err = "mmap failed: " + path;
        return false;
    }
    map_ = static_cast<const uint8_t *>(m);
    uint64_t hlen;
    memcpy(&hlen, map_, 8);
    if (hlen > map_len_ - 8) {
        err = "safetensors: header length exceeds file size";
        return false;
    }
    data_base_ = 8 + (size_t)hlen;
    JsonCur c{reinterpret_cast<const char *>(map_ + 8), reinterpret_cast<const char *>(map_ + 8 + hlen)};
    if (!c.eat('{')) {
        err = "safetensors: header is not a JSON object";
        return false;
    }
    if (!c.eat('}')) {
        do {
            std::string name = c.str();
            if (!c.eat(':')) { c.fail = true; break; }
            if (name == "__metadata__") {
                c.skip_value();
                continue;
            }
            StTensor t;
            if (!c.eat('{')) { c.fail = true; break; }
            do {
                std::string key = c.str();
                if (!c.eat(':')) { c.fail = true; break; }
                if (key == "dtype") {
                    t.dtype = c.str();
                } else if (key == "shape") {
                    if (!c.eat('[')) { c.fail = true; break; }
                    if (!c.eat(']')) {
                        do {
                            const uint64_t dim = c.uint();
                            if (dim > (uint64_t)INT64_MAX) c.fail = true;
                            t.shape.push_back((int64_t)dim);
                        } while (!c.fail && c.eat(','));
                        if (!c.eat(']')) c.fail = true;
                    }
                } else if (key == "data_offsets") {
                    if (!c.eat('[')) { c.fail = true; break; }
                    t.begin = c.uint();
                    if (!c.eat(',')) c.fail = true;
                    t.end = c.uint();
                    if (!c.eat(']')) c.fail = true;
                } else {
                    c.skip_value();
                }
            } while (!c.fail && c.eat(','));
            if (!c.eat('}')) c.fail = true;
            if (c.fail) break;
            if (t.end < t.begin || t.end > map_len_ - data_base_) {     // (no wrap-around: data_base_ <= map_len_)
                err = "safetensors: tensor '" + name + "' data out of range";
                return false;
            }
            index_[name] = t;
        } while (!c.fail && c.eat(','));
    }
    if (c.fail) {
        err = "safetensors: malformed header JSON";
        return false;
    }
    return true;
}

const StTensor *SafeTensors::find(const std::string &name) const {
    auto it = index_.find(name);
    return it == index_.end() ? nullptr : &it->second;
}

bool SafeTensors::read_f32(const std::string &name, std::vector<float> &out, int64_t expect_numel,
                           std::string &err) const {
    const StTensor *t = find(name);
    if (!t) {
        err = "missing tensor '" + name + "'";
        return false;
    }
    const int64_t n = t->numel();
    if (n < 0) {
        err = "tensor '" + name + "': shape overflows";
        return false;
    }
    if (expect_numel >= 0 && n != expect_numel) {
        err = "tensor '" + name + "' has " + std::to_string(n) + " elements, expected " + std::to_string(expect_numel);
        return false;
    }
    const uint8_t *src = map_ + data_base_ + t->begin;
    const uint64_t bytes = t->end - t->begin;
    out.resize((size_t)n);
    if (t->dtype == "F32" && bytes == (uint64_t)n * 4) {
        memcpy(out.data(), src, bytes);
    } else if (t->dtype == "F16" && bytes == (uint64_t)n * 2) {
        for (int64_t i = 0; i < n; ++i) {
            uint16_t h;
            memcpy(&h, src + 2 * i, 2);
            out[i] = half_to_float(h);
        }
    } else if (t->dtype == "BF16" && bytes == (uint64_t)n * 2) {
        for (int64_t i = 0; i < n; ++i) {
            uint16_t h;
            memcpy(&h, src + 2 * i, 2);
            uint32_t f = (uint32_t)h << 16;
            memcpy(&out[i], &f, 4);
        }
    } else if (t->dtype == "F64" && bytes == (uint64_t)n * 8) {
        for (int64_t i = 0; i < n; ++i) {
            double d;
            memcpy(&d, src + 8 * i, 8);
            out[i] = (float)d;
        }
    } else {
        err = "tensor '" + name + "': unsupported dtype " + t->dtype + " or size mismatch";
        return false;
    }
    return true;
}

}  // namespace pk
