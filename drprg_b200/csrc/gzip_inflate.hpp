// Parallel inflate of a single-stream gzip file (gzip_inflate.cpp).
#pragma once
#include <cstddef>
#include <cstdint>

namespace drprg {

// Inflates the gzip file image gz[0..n) on up to `threads` host threads into a malloc'ed, NUL-terminated buffer.
// Returns false — nothing allocated — when the image is not a single-member gzip stream this decoder handles, is too small
// to be worth cutting, or fails its CRC-32 / length check: the caller then uses zlib's sequential reader.
bool parallel_gunzip(const uint8_t* gz, size_t n, uint32_t threads, char** out, size_t* out_n);

}  // namespace drprg
