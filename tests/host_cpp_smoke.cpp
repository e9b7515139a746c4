// C++ host-mirror smoke test (built and run by tests/test_gpu_host_cpp.py on the GPU box): reads a key +
// two ciphertext batches dumped by the python test, evaluates them through tfhe_b200::Gates / batchNand /
// GpuBootstrap and writes the results back for comparison with the oracle.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../zig-tfhe_b200/host/tfhe_b200.hpp"

template <class T>
static std::vector<T> slurp(const char *path) {
    FILE *f = std::fopen(path, "rb");
    if (!f) { std::perror(path); std::exit(2); }
    std::fseek(f, 0, SEEK_END);
    long sz = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    std::vector<T> v(sz / sizeof(T));
    if (std::fread(v.data(), 1, sz, f) != (size_t)sz) std::exit(3);
    std::fclose(f);
    return v;
}

int main(int argc, char **argv) {
    if (argc < 2) return 1;
    std::string dir = argv[1];
    auto bsk = slurp<double>((dir + "/bsk.bin").c_str());
    auto ksk = slurp<uint32_t>((dir + "/ksk.bin").c_str());
    auto a = slurp<uint32_t>((dir + "/a.bin").c_str());
    auto b = slurp<uint32_t>((dir + "/b.bin").c_str());
    const auto p = tfhe_b200::security_128_bit();
    const size_t w = p.n + 1, count = a.size() / w;
    try {
        tfhe_b200::GpuBootstrap bs(p, tfhe_b200::CloudKey{0x82080000u, bsk.data(), ksk.data()});
        tfhe_b200::Gates gates(bs);
        std::vector<std::pair<tfhe_b200::Ciphertext, tfhe_b200::Ciphertext>> in;
        for (size_t i = 0; i < count; i++)
            in.push_back({tfhe_b200::Ciphertext(a.begin() + i * w, a.begin() + (i + 1) * w),
                          tfhe_b200::Ciphertext(b.begin() + i * w, b.begin() + (i + 1) * w)});
        auto nand = tfhe_b200::batchNand(bs, in);
        auto mux = gates.muxNaive(in[0].first, in[0].second, in[1].first);
        auto boot = bs.bootstrap(in[0].first);
        // 1-bit full adder (examples/add_two_numbers.zig:24-39) as a circuit over all `count` instances: a, b, cin = NOT a
        auto adder = tfhe_b200::Circuit::rippleCarryAdder(bs, 1);
        std::vector<uint32_t> wires;
        wires.insert(wires.end(), a.begin(), a.end());
        wires.insert(wires.end(), b.begin(), b.end());
        for (size_t i = 0; i < a.size(); i++) wires.push_back(0u - a[i]);
        auto sum_carry = adder.run(wires, count);
        FILE *f = std::fopen((dir + "/out.bin").c_str(), "wb");
        for (auto &c : nand) std::fwrite(c.data(), 4, w, f);
        std::fwrite(mux.data(), 4, w, f);
        std::fwrite(boot.data(), 4, w, f);
        std::fwrite(sum_carry.data(), 4, sum_carry.size(), f);
        std::fclose(f);
        if (adder.levels() != 3 || adder.gates() != 5) return 5;
        // CloudKey -> flat file -> a second evaluator built from the file alone: same ciphertext bits
        const std::string key_path = dir + "/cloud.key";
        tfhe_b200::saveCloudKey(key_path, p, tfhe_b200::CloudKey{0x82080000u, bsk.data(), ksk.data()});
        tfhe_b200::GpuBootstrap from_file(p, key_path);
        if (from_file.bootstrap(in[0].first) != boot) return 6;
        // two functions from one blind rotation (tfhe_b200_lut_bootstrap_many_batch) and the documented bootstrapLut, through the C++ mirror
        const std::vector<uint32_t> ident = {0x00000000u, 0x20000000u, 0x40000000u, 0x60000000u}, shifted = {0x20000000u, 0x40000000u, 0x60000000u, 0x00000000u};
        auto many = bs.bootstrapLutMany(in[0].first, {ident, shifted});
        if (many.size() != 2 || many[0].size() != w || many[1].size() != w || many[0] == many[1]) return 7;
        if (bs.bootstrapLut(in[0].first, ident).size() != w) return 8;
        std::printf("ok strategy=%s count=%zu\n", gates.bootstrapStrategy(), count);
    } catch (const tfhe_b200::Error &e) {
        std::printf("error %d: %s\n", e.code, e.what());
        return 4;
    }
    return 0;
}
