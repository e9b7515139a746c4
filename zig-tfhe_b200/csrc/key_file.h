// Flat cloud-key file: header layout and the host-side map/verify helpers shared by key_file.cu and capi.cu.
// The byte-level format is specified in include/tfhe_b200.h ("flat cloud-key file").
#pragma once
#include <cstddef>
#include <cstdint>

#include "../../include/tfhe_b200.h"

namespace tfhe_b200_keyfile {

constexpr uint32_t kVersion = 1;
constexpr uint32_t kAlign = 4096;
constexpr uint32_t kFlagHasKsk = 1;

struct Header {                 // little endian, 104 bytes, zero padded to kAlign in the file
    char magic[8];              // "TFHEB2CK"
    uint32_t version;           // kVersion
    uint32_t header_bytes;      // kAlign
    tfhe_b200_params params;    // n, N, L, bgbit, basebit, iks_t (6 x int32)
    uint32_t decomposition_offset;
    uint32_t flags;             // bit 0: key-switching key present (CloudKey.newNoKsk files have none, src/key.zig:80-100)
    uint64_t bsk_offset, bsk_bytes;
    uint64_t ksk_offset, ksk_bytes;
    uint64_t bsk_checksum, ksk_checksum;
    uint64_t header_checksum;   // checksum() of all bytes before this field
};
static_assert(sizeof(Header) == 104, "header layout is part of the file format");

struct View {                   // a mapped, validated file
    const void *base = nullptr;
    size_t map_bytes = 0;
    Header header{};
    const double *bsk = nullptr;
    const uint32_t *ksk = nullptr;   // nullptr when the file has none
};

// The one parameter validator (tfhe_b200_create and the key-file calls): N = 1024; n + 1 padded to a multiple of 4 must fit
// the key-switch kernel's 320 threads x 4 columns (n <= 1279; the reference's largest set has n = 1160); 1 + basebit * t <= 32
// keeps K2's precision offset 2^(31 - basebit * t) defined (src/trgsw.zig:483).
inline bool params_supported(const tfhe_b200_params &p) {
    return p.N == 1024 && p.n >= 1 && p.n <= 1279 && p.L >= 1 && p.L <= 4 && p.bgbit >= 1 && p.L * p.bgbit <= 32 && p.basebit >= 1 &&
           p.basebit <= 8 && p.iks_t >= 1 && 1 + p.basebit * p.iks_t <= 32;
}

uint64_t checksum(const void *data, size_t bytes);
uint64_t bsk_bytes(const tfhe_b200_params &p);
uint64_t ksk_bytes(const tfhe_b200_params &p);
int write(const char *path, const tfhe_b200_params &p, const double *bsk, const uint32_t *ksk, uint32_t offset, char *err, size_t cap);
int map(const char *path, bool verify_payload, View &v, char *err, size_t cap);
void unmap(View &v);

}  // namespace tfhe_b200_keyfile
