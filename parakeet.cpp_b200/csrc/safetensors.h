// safetensors.h -- minimal safetensors reader (host).
//
// Replaces axiom::io::safetensors::load (reference third_party/axiom/src/io/io_safetensors.cpp
// :123-160): 8-byte little-endian header length, a JSON object name -> {dtype, shape,
// data_offsets}, "__metadata__" skipped, raw little-endian tensor data after the header.
// Dtypes F32 / F16 / BF16 / F64 are converted to fp32 on load (the engine computes in fp32
// master weights); integer tensors (e.g. BatchNorm's num_batches_tracked, I64) are listed
// but not converted.
#pragma once
#include <cstdint>
#include <map>
#include <string>
#include <vector>

namespace pk {

struct StTensor {
    std::string dtype;
    std::vector<int64_t> shape;
    uint64_t begin = 0, end = 0;   // byte offsets into the data section
    int64_t numel() const {        // -1 on overflow
        int64_t n = 1;
        for (int64_t d : shape) {
            if (d < 0 || (d != 0 && n > INT64_MAX / d)) return -1;
            n *= d;
        }
        return n;
    }
};

class SafeTensors {
  public:
    // Returns false and fills err on failure.
    bool open(const std::string &path, std::string &err);
    ~SafeTensors();
    bool has(const std::string &name) const { return index_.count(name) != 0; }
    const StTensor *find(const std::string &name) const;
    // Converts to fp32; checks numel == expect_numel when expect_numel >= 0.
    bool read_f32(const std::string &name, std::vector<float> &out, int64_t expect_numel, std::string &err) const;
    size_t size() const { return index_.size(); }

  private:
    std::map<std::string, StTensor> index_;
    const uint8_t *map_ = nullptr;
    size_t map_len_ = 0;
    size_t data_base_ = 0;
};

}  // namespace pk
