// Flat cloud-key file ("defines the missing on-disk key format"): key.CloudKey (src/key.zig:61-65) has no
// serialised form in the reference, and its generation takes ~30 s per test run (src/key.zig:240).  The file is the reference's
// in-memory arrays verbatim behind one 4 KiB header, sections 4 KiB aligned, so that a load is mmap + cudaMemcpy and a Zig /
// C++ / Python host can also map it directly as CloudKey.bootstrapping_key / key_switching_key.  Host code only.
#include <cerrno>
#include <cstdio>
#include <cstring>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include "key_file.h"

namespace tfhe_b200_keyfile {

namespace {

constexpr char kMagic[8] = {'T', 'F', 'H', 'E', 'B', '2', 'C', 'K'};
constexpr uint64_t kFnvBasis = 0xcbf29ce484222325ull, kFnvPrime = 0x100000001b3ull;

uint64_t round_up(uint64_t x, uint64_t a) { return (x + a - 1) / a * a; }

bool params_ok(const tfhe_b200_params &p) { return params_supported(p); }

uint64_t header_sum(const Header &h) { return checksum(&h, offsetof(Header, header_checksum)); }

int fill(char *err, size_t cap, int code, const char *fmt, const char *path, const char *what = "") {
    if (err && cap) snprintf(err, cap, fmt, path ? path : "(null)", what);
    return code;
}

}  // namespace

// Four interleaved FNV-1a-64 lanes over little-endian 64-bit words (lane = word index mod 4), folded lane 0..3 into one more
// FNV-1a pass together with the byte length; trailing bytes (< 8) are zero-extended into one last word.  Word-wise so that
// 150 MB of key checks in tens of milliseconds; documented in include/tfhe_b200.h so other hosts can produce the file.
uint64_t checksum(const void *data, size_t bytes) {
    const unsigned char *p = static_cast<const unsigned char *>(data);
    uint64_t lane[4] = {kFnvBasis, kFnvBasis + 1, kFnvBasis + 2, kFnvBasis + 3};
    const size_t words = bytes / 8;
    for (size_t i = 0; i < words; i++) {
        uint64_t w;
        memcpy(&w, p + 8 * i, 8);
        lane[i & 3] = (lane[i & 3] ^ w) * kFnvPrime;
    }
    if (bytes % 8) {
        uint64_t w = 0;
        memcpy(&w, p + 8 * words, bytes % 8);
        lane[words & 3] = (lane[words & 3] ^ w) * kFnvPrime;
    }
    uint64_t h = kFnvBasis;
    for (uint64_t l : lane) h = (h ^ l) * kFnvPrime;
    return (h ^ (uint64_t)bytes) * kFnvPrime;
}

uint64_t bsk_bytes(const tfhe_b200_params &p) { return (uint64_t)p.n * 2 * p.L * 2 * p.N * sizeof(double); }
uint64_t ksk_bytes(const tfhe_b200_params &p) { return (uint64_t)p.N * p.iks_t * (1ull << p.basebit) * (p.n + 1) * sizeof(uint32_t); }

int write(const char *path, const tfhe_b200_params &p, const double *bsk, const uint32_t *ksk, uint32_t offset, char *err, size_t cap) {
    if (!path || !bsk) return fill(err, cap, TFHE_B200_ERR_INVALID, "%s: null argument%s", path);
    if (!params_ok(p)) return fill(err, cap, TFHE_B200_ERR_INVALID, "%s: unsupported parameter set%s", path);
    Header h;
    memset(&h, 0, sizeof h);
    memcpy(h.magic, kMagic, 8);
    h.version = kVersion;
    h.header_bytes = kAlign;
    h.params = p;
    h.decomposition_offset = offset;
    h.flags = ksk ? kFlagHasKsk : 0;
    h.bsk_offset = kAlign;
    h.bsk_bytes = bsk_bytes(p);
    h.ksk_offset = round_up(h.bsk_offset + h.bsk_bytes, kAlign);
    h.ksk_bytes = ksk ? ksk_bytes(p) : 0;
    h.bsk_checksum = checksum(bsk, h.bsk_bytes);
    h.ksk_checksum = ksk ? checksum(ksk, h.ksk_bytes) : 0;
    h.header_checksum = header_sum(h);

    // write to a sibling temporary and rename: a reader never sees a torn key
    char tmp[4096];
    if (snprintf(tmp, sizeof tmp, "%s.tmp.%ld", path, (long)getpid()) >= (int)sizeof tmp)
        return fill(err, cap, TFHE_B200_ERR_INVALID, "%s: path too long%s", path);
    FILE *f = fopen(tmp, "wb");
    if (!f) return fill(err, cap, TFHE_B200_ERR_IO, "%s: cannot create (%s)", tmp, strerror(errno));
    static const char zeros[kAlign] = {0};
    bool ok = fwrite(&h, sizeof h, 1, f) == 1 && fwrite(zeros, kAlign - sizeof h, 1, f) == 1 && fwrite(bsk, 1, h.bsk_bytes, f) == h.bsk_bytes;
    if (ok && ksk) {
        const uint64_t pad = h.ksk_offset - (h.bsk_offset + h.bsk_bytes);
        ok = (pad == 0 || fwrite(zeros, 1, pad, f) == pad) && fwrite(ksk, 1, h.ksk_bytes, f) == h.ksk_bytes;
    }
    ok = (fflush(f) == 0) && ok;
    ok = (fclose(f) == 0) && ok;
    if (!ok || rename(tmp, path) != 0) {
        const int e = errno;
        unlink(tmp);
        return fill(err, cap, TFHE_B200_ERR_IO, "%s: write failed (%s)", path, strerror(e));
    }
    return TFHE_B200_OK;
}

void unmap(View &v) {
    if (v.base) munmap(const_cast<void *>(v.base), v.map_bytes);
    v = View{};
}

int map(const char *path, bool verify_payload, View &v, char *err, size_t cap) {
    v = View{};
    if (!path) return fill(err, cap, TFHE_B200_ERR_INVALID, "%s: null path%s", path);
    const int fd = open(path, O_RDONLY | O_CLOEXEC);
    if (fd < 0) return fill(err, cap, TFHE_B200_ERR_IO, "%s: cannot open (%s)", path, strerror(errno));
    struct stat st;
    if (fstat(fd, &st) != 0 || st.st_size < (off_t)sizeof(Header)) {
        close(fd);
        return fill(err, cap, TFHE_B200_ERR_IO, "%s: too short to be a cloud-key file%s", path);
    }
    void *base = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
    close(fd);
    if (base == MAP_FAILED) return fill(err, cap, TFHE_B200_ERR_IO, "%s: mmap failed (%s)", path, strerror(errno));
    v.base = base;
    v.map_bytes = (size_t)st.st_size;
    memcpy(&v.header, base, sizeof(Header));
    const Header &h = v.header;
    const char *bad = nullptr;
    if (memcmp(h.magic, kMagic, 8) != 0) bad = "bad magic";
    else if (h.version != kVersion) bad = "unknown version";
    else if (h.header_checksum != header_sum(h)) bad = "header checksum mismatch";
    else if (!params_ok(h.params)) bad = "unsupported parameter set";
    else if (h.bsk_bytes != bsk_bytes(h.params) || h.bsk_offset % kAlign || h.bsk_offset < sizeof(Header) ||
             h.bsk_offset + h.bsk_bytes > v.map_bytes)
        bad = "bootstrapping-key section out of bounds (truncated file?)";
    else if ((h.flags & kFlagHasKsk) && (h.ksk_bytes != ksk_bytes(h.params) || h.ksk_offset % kAlign ||
                                         h.ksk_offset < h.bsk_offset + h.bsk_bytes || h.ksk_offset + h.ksk_bytes > v.map_bytes))
        bad = "key-switching-key section out of bounds (truncated file?)";
    if (bad) {
        unmap(v);
        return fill(err, cap, TFHE_B200_ERR_INVALID, "%s: %s", path, bad);
    }
    v.bsk = reinterpret_cast<const double *>(static_cast<const char *>(base) + h.bsk_offset);
    v.ksk = (h.flags & kFlagHasKsk) ? reinterpret_cast<const uint32_t *>(static_cast<const char *>(base) + h.ksk_offset) : nullptr;
    if (verify_payload) {
        if (checksum(v.bsk, h.bsk_bytes) != h.bsk_checksum) bad = "bootstrapping-key checksum mismatch";
        else if (v.ksk && checksum(v.ksk, h.ksk_bytes) != h.ksk_checksum) bad = "key-switching-key checksum mismatch";
        if (bad) {
            unmap(v);
            return fill(err, cap, TFHE_B200_ERR_INVALID, "%s: %s", path, bad);
        }
    }
    return TFHE_B200_OK;
}

}  // namespace tfhe_b200_keyfile

namespace kf = tfhe_b200_keyfile;

namespace {
thread_local char g_err[512] = "";
}

extern "C" {

const char *tfhe_b200_key_file_last_error(void) { return g_err; }

int tfhe_b200_key_file_write(const char *path, const tfhe_b200_params *params, const double *bsk, const uint32_t *ksk,
                             uint32_t decomposition_offset) {
    g_err[0] = 0;
    if (!params) return snprintf(g_err, sizeof g_err, "null params"), TFHE_B200_ERR_INVALID;
    return kf::write(path, *params, bsk, ksk, decomposition_offset, g_err, sizeof g_err);
}

int tfhe_b200_key_file_info(const char *path, tfhe_b200_params *params, uint32_t *decomposition_offset, uint64_t *bsk_bytes,
                            uint64_t *ksk_bytes) {
    g_err[0] = 0;
    kf::View v;
    if (int r = kf::map(path, false, v, g_err, sizeof g_err)) return r;
    if (params) *params = v.header.params;
    if (decomposition_offset) *decomposition_offset = v.header.decomposition_offset;
    if (bsk_bytes) *bsk_bytes = v.header.bsk_bytes;
    if (ksk_bytes) *ksk_bytes = v.header.ksk_bytes;
    kf::unmap(v);
    return TFHE_B200_OK;
}

int tfhe_b200_key_file_read(const char *path, double *bsk, uint32_t *ksk) {
    g_err[0] = 0;
    kf::View v;
    if (int r = kf::map(path, true, v, g_err, sizeof g_err)) return r;
    int rc = TFHE_B200_OK;
    if (bsk) memcpy(bsk, v.bsk, v.header.bsk_bytes);
    if (ksk) {
        if (v.ksk) memcpy(ksk, v.ksk, v.header.ksk_bytes);
        else rc = (snprintf(g_err, sizeof g_err, "%s: file holds no key-switching key", path), TFHE_B200_ERR_NO_KEY);
    }
    kf::unmap(v);
    return rc;
}

}  // extern "C"
