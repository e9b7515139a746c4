// tfhe_b200.hpp -- C++ mirror of the reference's operator interface for the gate-bootstrapping path,
// over the C ABI of include/tfhe_b200.h.  (The reference's host language, Zig, has no toolchain in the
// build image; this header is the compiled-language host side, and zig/cuda.zig + zig/gpu.zig are the Zig one.)
//
//   reference                                                  here
//   gates.Gates{bootstrap}            src/gates.zig:25-151      tfhe_b200::Gates
//   gates.batchNand ... batchXnor     src/gates.zig:244-295     tfhe_b200::batchNand ... batchXnor
//   VanillaBootstrap                  src/bootstrap/vanilla.zig tfhe_b200::GpuBootstrap
//   key.CloudKey                      src/key.zig:61-65         tfhe_b200::CloudKey (views, not owners)
//   errors: Zig error unions                                    tfhe_b200::Error (std::runtime_error + status code)
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/tfhe_b200.h"

namespace tfhe_b200 {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string &m) : std::runtime_error("tfhe_b200: " + m), code(c) {}
};
struct NotImplemented : Error { using Error::Error; };   // Zig: error.NotImplemented

// utils.Ciphertext = tlwe.TLWELv0 (src/utils.zig:25, src/tlwe.zig:11-13): n+1 torus words, body last
using Ciphertext = std::vector<uint32_t>;

// params.SecurityParams (src/params.zig:36-67)
inline tfhe_b200_params security_128_bit() { return {700, 1024, 3, 6, 2, 9}; }   // params.zig:350-375
inline tfhe_b200_params security_110_bit() { return {630, 1024, 3, 6, 2, 8}; }   // params.zig:98-123
inline tfhe_b200_params security_80_bit() { return {550, 1024, 3, 6, 2, 7}; }    // params.zig:70-95

// key.CloudKey (src/key.zig:61-65) as non-owning views in the reference's layouts
struct CloudKey {
    uint32_t decomposition_offset;
    const double *bootstrapping_key;        // [n][2L][2][N]
    const uint32_t *key_switching_key;      // [N*t*base][n+1] or nullptr (CloudKey.newNoKsk, key.zig:80-100)
};

// CloudKey <-> flat file; host only
inline void saveCloudKey(const std::string &path, const tfhe_b200_params &p, const CloudKey &ck) {
    if (int rc = tfhe_b200_key_file_write(path.c_str(), &p, ck.bootstrapping_key, ck.key_switching_key, ck.decomposition_offset))
        throw Error(rc, tfhe_b200_key_file_last_error());
}

class GpuBootstrap {
public:
    GpuBootstrap(const tfhe_b200_params &p, const CloudKey &ck, const std::vector<int> &devices = {0}) : words_(p.n + 1) {
        int rc = tfhe_b200_create(&p, devices.data(), (int)devices.size(), &ctx_);
        if (rc) throw Error(rc, "no sm_100 CUDA device / bad parameters");
        check(tfhe_b200_load_key(ctx_, ck.bootstrapping_key, ck.key_switching_key, (size_t)words_ * 4, ck.decomposition_offset));
    }
    // key.CloudKey.new on the device (key.zig:70-77, 148-212): the evaluation keys are generated where they are used, from the
    // secret key the caller holds; nothing but the 6.9 KB secret key crosses the bus
    GpuBootstrap(const tfhe_b200_params &p, const uint32_t *key_lv0, const uint32_t *key_lv1, uint64_t seed, double ksk_alpha,
                 double bsk_alpha, const std::vector<int> &devices = {0})
        : words_(p.n + 1) {
        int rc = tfhe_b200_create(&p, devices.data(), (int)devices.size(), &ctx_);
        if (rc) throw Error(rc, "no sm_100 CUDA device / bad parameters");
        check(tfhe_b200_keygen(ctx_, key_lv0, key_lv1, seed, ksk_alpha, bsk_alpha, nullptr, nullptr));
    }
    // from a flat cloud-key file (include/tfhe_b200.h "flat cloud-key file"): mmap + checksum + upload, no 30 s key generation
    // per run (key.zig:240)
    GpuBootstrap(const tfhe_b200_params &p, const std::string &key_file, const std::vector<int> &devices = {0}) : words_(p.n + 1) {
        int rc = tfhe_b200_create(&p, devices.data(), (int)devices.size(), &ctx_);
        if (rc) throw Error(rc, "no sm_100 CUDA device / bad parameters");
        check(tfhe_b200_load_key_file(ctx_, key_file.c_str()));
    }
    GpuBootstrap(const GpuBootstrap &) = delete;
    GpuBootstrap &operator=(const GpuBootstrap &) = delete;
    ~GpuBootstrap() { tfhe_b200_destroy(ctx_); }

    // VanillaBootstrap.bootstrap (vanilla.zig:38-52)
    Ciphertext bootstrap(const Ciphertext &c) const {
        Ciphertext out(words_);
        check(tfhe_b200_bootstrap_batch(ctx_, c.data(), out.data(), 1, nullptr, 0));
        return out;
    }
    // VanillaBootstrap.bootstrapWithoutKeySwitch (vanilla.zig:58-69)
    Ciphertext bootstrapWithoutKeySwitch(const Ciphertext &c) const {
        Ciphertext out(words_);
        check(tfhe_b200_bootstrap_no_keyswitch_batch(ctx_, c.data(), out.data(), 1));
        return out;
    }
    // the bootstrapLut src/lut.zig:42 documents: `table[x]` = torus value to return for message x (Encoder.encode(f(x)),
    // lut/encoder.zig:66-73); the LookupTable (lut/generator.zig:150-191) is built on the device
    Ciphertext bootstrapLut(const Ciphertext &c, const std::vector<uint32_t> &table) const {
        Ciphertext out(words_);
        check(tfhe_b200_lut_bootstrap_batch(ctx_, c.data(), out.data(), 1, table.data(), (int)table.size(), 0));
        return out;
    }
    // several functions from ONE blind rotation: tables[f][x], f < tables.size() (a power of two), returns one ciphertext per function
    std::vector<Ciphertext> bootstrapLutMany(const Ciphertext &c, const std::vector<std::vector<uint32_t>> &tables) const {
        const size_t k = tables.size(), m = k ? tables[0].size() : 0;
        std::vector<uint32_t> flat;
        for (const auto &t : tables) flat.insert(flat.end(), t.begin(), t.end());
        std::vector<uint32_t> out(k * (size_t)words_);
        check(tfhe_b200_lut_bootstrap_many_batch(ctx_, c.data(), out.data(), 1, flat.data(), (int)k, (int)m));
        std::vector<Ciphertext> res;
        for (size_t f = 0; f < k; f++) res.emplace_back(out.begin() + f * words_, out.begin() + (f + 1) * words_);
        return res;
    }
    const char *name() const { return "b200"; }
    tfhe_b200_ctx *ctx() const { return ctx_; }
    int words() const { return words_; }
    void check(int rc) const {
        if (rc == TFHE_B200_ERR_NOT_IMPLEMENTED) throw NotImplemented(rc, tfhe_b200_last_error(ctx_));
        if (rc) throw Error(rc, tfhe_b200_last_error(ctx_));
    }

private:
    tfhe_b200_ctx *ctx_ = nullptr;
    int words_;
};

// gates.Gates (gates.zig:25-151)
class Gates {
public:
    explicit Gates(const GpuBootstrap &b) : b_(b) {}
    const char *bootstrapStrategy() const { return b_.name(); }
    Ciphertext nandGate(const Ciphertext &a, const Ciphertext &b) const { return one(TFHE_B200_NAND, a, b); }
    Ciphertext orGate(const Ciphertext &a, const Ciphertext &b) const { return one(TFHE_B200_OR, a, b); }
    Ciphertext andGate(const Ciphertext &a, const Ciphertext &b) const { return one(TFHE_B200_AND, a, b); }
    Ciphertext xorGate(const Ciphertext &a, const Ciphertext &b) const { return one(TFHE_B200_XOR, a, b); }
    Ciphertext xnorGate(const Ciphertext &a, const Ciphertext &b) const { return one(TFHE_B200_XNOR, a, b); }
    Ciphertext norGate(const Ciphertext &a, const Ciphertext &b) const { return one(TFHE_B200_NOR, a, b); }
    Ciphertext andNyGate(const Ciphertext &a, const Ciphertext &b) const { return one(TFHE_B200_ANDNY, a, b); }
    Ciphertext andYnGate(const Ciphertext &a, const Ciphertext &b) const { return one(TFHE_B200_ANDYN, a, b); }
    Ciphertext orNyGate(const Ciphertext &a, const Ciphertext &b) const { return one(TFHE_B200_ORNY, a, b); }
    Ciphertext orYnGate(const Ciphertext &a, const Ciphertext &b) const { return one(TFHE_B200_ORYN, a, b); }
    // gates.zig:131-134, no bootstrap
    Ciphertext notGate(const Ciphertext &a) const {
        Ciphertext out(a.size());
        for (size_t i = 0; i < a.size(); i++) out[i] = 0u - a[i];
        return out;
    }
    Ciphertext copy(const Ciphertext &a) const { return a; }
    // gates.zig:144-151 (false = 1 - 2^29, kept)
    Ciphertext constant(bool v) const {
        Ciphertext out(b_.words(), 0u);
        out.back() = v ? 0x20000000u : 1u - 0x20000000u;
        return out;
    }
    // gates.zig:124-129
    Ciphertext muxNaive(const Ciphertext &a, const Ciphertext &b, const Ciphertext &c) const {
        return orGate(andGate(a, b), andGate(notGate(a), c));
    }
    // batch form: a, b, out are [count][n+1] contiguous
    void batch(int op, const uint32_t *a, const uint32_t *b, uint32_t *out, size_t count) const {
        b_.check(tfhe_b200_gate_batch(b_.ctx(), op, a, b, out, count));
    }

private:
    Ciphertext one(int op, const Ciphertext &a, const Ciphertext &b) const {
        Ciphertext out(a.size());
        batch(op, a.data(), b.data(), out.data(), 1);
        return out;
    }
    const GpuBootstrap &b_;
};

// gates.batch* (gates.zig:244-295): inputs as pairs, results by value
inline std::vector<Ciphertext> batchGate(const GpuBootstrap &bs, int op, const std::vector<std::pair<Ciphertext, Ciphertext>> &in) {
    const size_t w = bs.words(), count = in.size();
    std::vector<uint32_t> a(count * w), b(count * w), o(count * w);
    for (size_t i = 0; i < count; i++) {
        std::copy(in[i].first.begin(), in[i].first.end(), a.begin() + i * w);
        std::copy(in[i].second.begin(), in[i].second.end(), b.begin() + i * w);
    }
    bs.check(tfhe_b200_gate_batch(bs.ctx(), op, a.data(), b.data(), o.data(), count));
    std::vector<Ciphertext> out(count);
    for (size_t i = 0; i < count; i++) out[i].assign(o.begin() + i * w, o.begin() + (i + 1) * w);
    return out;
}
inline auto batchNand(const GpuBootstrap &b, const std::vector<std::pair<Ciphertext, Ciphertext>> &in) { return batchGate(b, TFHE_B200_NAND, in); }
inline auto batchAnd(const GpuBootstrap &b, const std::vector<std::pair<Ciphertext, Ciphertext>> &in) { return batchGate(b, TFHE_B200_AND, in); }
inline auto batchOr(const GpuBootstrap &b, const std::vector<std::pair<Ciphertext, Ciphertext>> &in) { return batchGate(b, TFHE_B200_OR, in); }
inline auto batchXor(const GpuBootstrap &b, const std::vector<std::pair<Ciphertext, Ciphertext>> &in) { return batchGate(b, TFHE_B200_XOR, in); }
inline auto batchNor(const GpuBootstrap &b, const std::vector<std::pair<Ciphertext, Ciphertext>> &in) { return batchGate(b, TFHE_B200_NOR, in); }
inline auto batchXnor(const GpuBootstrap &b, const std::vector<std::pair<Ciphertext, Ciphertext>> &in) { return batchGate(b, TFHE_B200_XNOR, in); }

// A gate netlist compiled for the device (tfhe_b200_circuit_*): the batched form of a chain of Gates.* calls.
// Wire ids: 0..n_inputs-1 inputs, n_inputs + g = output of gate g; `w | kNot` = Gates.notGate of wire w (free).
class Circuit {
public:
    static constexpr uint32_t kNot = TFHE_B200_WIRE_NOT;
    static constexpr uint32_t kTrue = TFHE_B200_WIRE_TRUE, kFalse = TFHE_B200_WIRE_FALSE;   // Gates.constant wires (gates.zig:144-151)
    Circuit(const GpuBootstrap &b, const std::vector<tfhe_b200_gate_node> &gates, size_t n_inputs, const std::vector<uint32_t> &outputs)
        : b_(b), n_inputs_(n_inputs), n_outputs_(outputs.size()) {
        b_.check(tfhe_b200_circuit_create(b_.ctx(), gates.data(), gates.size(), n_inputs, outputs.data(), outputs.size(), &c_));
        tfhe_b200_circuit_info(c_, &levels_, &width_, &gates_);
    }
    ~Circuit() { tfhe_b200_circuit_destroy(c_); }
    Circuit(const Circuit &) = delete;
    Circuit &operator=(const Circuit &) = delete;
    size_t levels() const { return levels_; }
    size_t gates() const { return gates_; }
    // inputs [n_inputs][instances][n+1] -> outputs [n_outputs][instances][n+1]
    std::vector<uint32_t> run(const std::vector<uint32_t> &inputs, size_t instances) const {
        std::vector<uint32_t> out(n_outputs_ * instances * b_.words());
        b_.check(tfhe_b200_circuit_run(b_.ctx(), c_, inputs.data(), out.data(), instances));
        return out;
    }
    // examples/add_two_numbers.zig:24-73 (fullAdder chained over `width` bits): inputs a_0.., b_0.., cin; outputs sum_0.., carry
    static Circuit rippleCarryAdder(const GpuBootstrap &b, int width) {
        std::vector<tfhe_b200_gate_node> g;
        std::vector<uint32_t> outs;
        const uint32_t n_in = 2 * width + 1;
        uint32_t carry = 2 * width;
        for (int i = 0; i < width; i++) {
            const uint32_t a = i, bb = width + i, k = n_in + (uint32_t)g.size();
            g.push_back({TFHE_B200_XOR, a, bb});
            g.push_back({TFHE_B200_AND, a, bb});
            g.push_back({TFHE_B200_AND, k, carry});
            g.push_back({TFHE_B200_XOR, k, carry});
            g.push_back({TFHE_B200_OR, k + 1, k + 2});
            outs.push_back(k + 3);
            carry = k + 4;
        }
        outs.push_back(carry);
        return Circuit(b, g, n_in, outs);
    }
    Circuit(Circuit &&o) noexcept : b_(o.b_), c_(o.c_), n_inputs_(o.n_inputs_), n_outputs_(o.n_outputs_), levels_(o.levels_), width_(o.width_), gates_(o.gates_) { o.c_ = nullptr; }

private:
    const GpuBootstrap &b_;
    tfhe_b200_circuit *c_ = nullptr;
    size_t n_inputs_, n_outputs_, levels_ = 0, width_ = 0, gates_ = 0;
};

}  // namespace tfhe_b200
