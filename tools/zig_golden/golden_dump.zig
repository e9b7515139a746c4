//! golden_dump.zig -- dumps keys, inputs and outputs of the UNMODIFIED zig-tfhe hot path as raw little-endian arrays,
//! so that the CPU oracle and the CUDA path of tfhe-b200 can be pinned against bits produced by the real reference.
//!
//! The reference seeds its PRNGs from the clock (src/utils.zig:16-22); that does not matter here: whatever key and
//! ciphertexts this run draws are written out, and oracle / GPU must reproduce the outputs from those files.
//!
//! Usage (Zig 0.15.1, from a checkout of thedonutfactory/zig-tfhe; see tools/zig_golden/README.md):
//!   cp <tfhe-b200>/tools/zig_golden/golden_dump.zig examples/golden_dump.zig
//!   zig build-exe -O ReleaseFast --dep main -Mroot=examples/golden_dump.zig -Mmain=src/main.zig -lc -lm
//!   ./golden_dump <tfhe-b200>/tests/golden/zig
//! NOTE: written against Zig 0.15.1 but never compiled in the authoring environment (no Zig toolchain there).
//!
//! Files written (all little endian; n, N, L, t, base = params.implementation.*):
//!   manifest.txt      "n N L bgbit basebit iks_t decomposition_offset count_gates count_rotations" (decimal, one line)
//!   bsk.bin           f64 [n][2L][2 (a, b)][N]      CloudKey.bootstrapping_key           (src/key.zig:61-65, 182-212)
//!   ksk.bin           u32 [N*t*base][n+1]           CloudKey.key_switching_key           (src/key.zig:148-172)
//!   secret.bin        u32 [n] then u32 [N]          SecretKey.key_lv0, key_lv1           (src/key.zig:34-58)
//!   gate_ops.bin      i32 [G]                       opcode per gate (tfhe_b200_gate numbering)
//!   gate_a.bin, gate_b.bin, gate_out.bin   u32 [G][n+1]   Gates.{nand,and,or,xor,nor,xnor}Gate over the truth table (src/gates.zig:48-92)
//!   gate_bits.bin     u8  [G][3]                    plaintext a, b, decrypted result
//!   rot_in.bin        u32 [R][n+1]                  inputs of trgsw.blindRotate            (src/trgsw.zig:290-333)
//!   rot_trlwe.bin     u32 [R][2][N]                 its outputs
//!   rot_lv1.bin       u32 [R][N+1]                  trlwe.sampleExtractIndex(., 0)         (src/trlwe.zig:146-162)
//!   rot_lv0.bin       u32 [R][n+1]                  trgsw.identityKeySwitching             (src/trgsw.zig:471-502)
const std = @import("std");
const tfhe = @import("main");
const params = tfhe.params;
const key = tfhe.key;
const utils = tfhe.utils;
const trgsw = tfhe.trgsw;
const trlwe = tfhe.trlwe;
const gates = tfhe.gates;

const n = params.implementation.tlwe_lv0.N;
const N = params.implementation.trgsw_lv1.N;
const L = params.implementation.trgsw_lv1.L;

fn writeFile(dir: std.fs.Dir, name: []const u8, bytes: []const u8) !void {
    var f = try dir.createFile(name, .{});
    defer f.close();
    try f.writeAll(bytes);
}

const Op = enum(i32) { nand = 0, @"or" = 1, @"and" = 2, xor = 3, xnor = 4, nor = 5 };

pub fn main() !void {
    const allocator = std.heap.page_allocator;
    var args = try std.process.argsWithAllocator(allocator);
    defer args.deinit();
    _ = args.next();
    const out_path = args.next() orelse "golden_zig";
    try std.fs.cwd().makePath(out_path);
    var dir = try std.fs.cwd().openDir(out_path, .{});
    defer dir.close();

    const secret_key = key.SecretKey.new();
    const cloud_key = try key.CloudKey.new(allocator, &secret_key);

    // ---- keys, field by field (Zig structs have no guaranteed layout)
    {
        const bsk = try allocator.alloc(f64, n * 2 * L * 2 * N);
        defer allocator.free(bsk);
        for (cloud_key.bootstrapping_key.items, 0..) |*row, i| {
            for (&row.trlwe_fft, 0..) |*t, r| {
                const base = ((i * 2 * L + r) * 2) * N;
                @memcpy(bsk[base .. base + N], &t.a);
                @memcpy(bsk[base + N .. base + 2 * N], &t.b);
            }
        }
        try writeFile(dir, "bsk.bin", std.mem.sliceAsBytes(bsk));
        const rows = cloud_key.key_switching_key.items.len;
        const ksk = try allocator.alloc(u32, rows * (n + 1));
        defer allocator.free(ksk);
        for (cloud_key.key_switching_key.items, 0..) |*row, i| @memcpy(ksk[i * (n + 1) .. (i + 1) * (n + 1)], &row.p);
        try writeFile(dir, "ksk.bin", std.mem.sliceAsBytes(ksk));
        var sec: [n + N]u32 = undefined;
        @memcpy(sec[0..n], &secret_key.key_lv0);
        @memcpy(sec[n .. n + N], &secret_key.key_lv1);
        try writeFile(dir, "secret.bin", std.mem.sliceAsBytes(sec[0..]));
    }

    // ---- gates over their truth tables (src/gates.zig:374-511), three passes with fresh encryptions
    const g = gates.Gates.new();
    const ops = [_]Op{ .nand, .@"and", .@"or", .xor, .nor, .xnor };
    const passes = 3;
    const G = ops.len * 4 * passes;
    const ga = try allocator.alloc(u32, G * (n + 1));
    const gb = try allocator.alloc(u32, G * (n + 1));
    const go = try allocator.alloc(u32, G * (n + 1));
    var gops: [G]i32 = undefined;
    var gbits: [G * 3]u8 = undefined;
    var idx: usize = 0;
    for (0..passes) |_| {
        for (ops) |op| {
            for (0..4) |row| {
                const a = (row & 2) != 0;
                const b = (row & 1) != 0;
                const ct_a = try utils.Ciphertext.encryptBool(a, params.implementation.tlwe_lv0.ALPHA, &secret_key.key_lv0);
                const ct_b = try utils.Ciphertext.encryptBool(b, params.implementation.tlwe_lv0.ALPHA, &secret_key.key_lv0);
                const res = switch (op) {
                    .nand => try g.nandGate(&ct_a, &ct_b, &cloud_key),
                    .@"and" => try g.andGate(&ct_a, &ct_b, &cloud_key),
                    .@"or" => try g.orGate(&ct_a, &ct_b, &cloud_key),
                    .xor => try g.xorGate(&ct_a, &ct_b, &cloud_key),
                    .nor => try g.norGate(&ct_a, &ct_b, &cloud_key),
                    .xnor => try g.xnorGate(&ct_a, &ct_b, &cloud_key),
                };
                @memcpy(ga[idx * (n + 1) .. (idx + 1) * (n + 1)], &ct_a.p);
                @memcpy(gb[idx * (n + 1) .. (idx + 1) * (n + 1)], &ct_b.p);
                @memcpy(go[idx * (n + 1) .. (idx + 1) * (n + 1)], &res.p);
                gops[idx] = @intFromEnum(op);
                gbits[3 * idx] = @intFromBool(a);
                gbits[3 * idx + 1] = @intFromBool(b);
                gbits[3 * idx + 2] = @intFromBool(res.decryptBool(&secret_key.key_lv0));
                idx += 1;
            }
        }
    }
    try writeFile(dir, "gate_ops.bin", std.mem.sliceAsBytes(gops[0..]));
    try writeFile(dir, "gate_a.bin", std.mem.sliceAsBytes(ga));
    try writeFile(dir, "gate_b.bin", std.mem.sliceAsBytes(gb));
    try writeFile(dir, "gate_out.bin", std.mem.sliceAsBytes(go));
    try writeFile(dir, "gate_bits.bin", gbits[0..]);

    // ---- blind rotation -> sample extraction -> key switch, stage by stage (src/trgsw.zig:694-755)
    const R = 8;
    const rin = try allocator.alloc(u32, R * (n + 1));
    const rtr = try allocator.alloc(u32, R * 2 * N);
    const rl1 = try allocator.alloc(u32, R * (N + 1));
    const rl0 = try allocator.alloc(u32, R * (n + 1));
    for (0..R) |i| {
        const ct = try utils.Ciphertext.encryptBool((i & 1) != 0, params.implementation.tlwe_lv0.ALPHA, &secret_key.key_lv0);
        const tr = try trgsw.blindRotate(&ct, &cloud_key);
        const lv1 = trlwe.sampleExtractIndex(&tr, 0);
        const lv0 = trgsw.identityKeySwitching(&lv1, &cloud_key.key_switching_key);
        @memcpy(rin[i * (n + 1) .. (i + 1) * (n + 1)], &ct.p);
        @memcpy(rtr[i * 2 * N .. i * 2 * N + N], &tr.a);
        @memcpy(rtr[i * 2 * N + N .. (i + 1) * 2 * N], &tr.b);
        @memcpy(rl1[i * (N + 1) .. (i + 1) * (N + 1)], &lv1.p);
        @memcpy(rl0[i * (n + 1) .. (i + 1) * (n + 1)], &lv0.p);
    }
    try writeFile(dir, "rot_in.bin", std.mem.sliceAsBytes(rin));
    try writeFile(dir, "rot_trlwe.bin", std.mem.sliceAsBytes(rtr));
    try writeFile(dir, "rot_lv1.bin", std.mem.sliceAsBytes(rl1));
    try writeFile(dir, "rot_lv0.bin", std.mem.sliceAsBytes(rl0));

    var buf: [256]u8 = undefined;
    const line = try std.fmt.bufPrint(&buf, "{d} {d} {d} {d} {d} {d} {d} {d} {d}\n", .{
        n,                                       N,                                      L,
        params.implementation.trgsw_lv1.BGBIT,   params.implementation.trgsw_lv1.BASEBIT, params.implementation.trgsw_lv1.IKS_T,
        cloud_key.decomposition_offset,          G,                                      R,
    });
    try writeFile(dir, "manifest.txt", line);
    std.debug.print("wrote {s}: {d} gates, {d} rotations\n", .{ out_path, G, R });
}
