//! gpu.zig -- B200 bootstrap strategy + batch gates for zig-tfhe (goes to src/bootstrap/gpu.zig).
//!
//! Fills the slots the reference leaves open:
//!   * a bootstrap strategy next to VanillaBootstrap (src/bootstrap/vanilla.zig:25-75, trait in src/bootstrap.zig:30-47)
//!   * gates.batchNand/And/Or/Xor/Nor/Xnor, which return error.NotImplemented today (src/gates.zig:244-295)
//! NOTE: uncompiled in the authoring environment (no Zig toolchain); see INTEGRATION.md.
const std = @import("std");
const params = @import("../params.zig");
const utils = @import("../utils.zig");
const key = @import("../key.zig");
const tlwe = @import("../tlwe.zig");
const lut = @import("../lut.zig");
const cuda = @import("../cuda.zig");

const n = params.implementation.tlwe_lv0.N;
const N = params.implementation.trgsw_lv1.N;
const L = params.implementation.trgsw_lv1.L;
const CT_WORDS = n + 1; // TLWELv0.p, src/tlwe.zig:11-13

pub const GpuBootstrap = struct {
    ctx: *cuda.Ctx,

    const Self = @This();

    /// Create the device context for the compile-time parameter set (src/params.zig:386-416) and upload the
    /// cloud key.  Zig structs have no guaranteed layout, so the key is staged field by field into packed buffers.
    pub fn init(allocator: std.mem.Allocator, cloud_key: *const key.CloudKey, device_ids: []const c_int) !Self {
        const p = cuda.Params{
            .n = @intCast(n),
            .L = @intCast(L),
            .bgbit = @intCast(params.implementation.trgsw_lv1.BGBIT),
            .basebit = @intCast(params.implementation.trgsw_lv1.BASEBIT),
            .iks_t = @intCast(params.implementation.trgsw_lv1.IKS_T),
        };
        var ctx: ?*cuda.Ctx = null;
        try cuda.check(cuda.tfhe_b200_create(&p, device_ids.ptr, @intCast(device_ids.len), &ctx));
        errdefer cuda.tfhe_b200_destroy(ctx);

        // bootstrapping key: [n][2L][2 (a,b)][N] f64  (src/key.zig:61-65, src/trgsw.zig:75-76, src/trlwe.zig:104-106)
        const bsk = try allocator.alloc(f64, n * 2 * L * 2 * N);
        defer allocator.free(bsk);
        for (cloud_key.bootstrapping_key.items, 0..) |*row, i| {
            for (row.trlwe_fft, 0..) |*t, r| {
                const base = ((i * 2 * L + r) * 2) * N;
                @memcpy(bsk[base .. base + N], &t.a);
                @memcpy(bsk[base + N .. base + 2 * N], &t.b);
            }
        }
        // key-switching key: [N*t*base][n+1] u32 (src/key.zig:148-172); k = 0 rows are never read
        const rows = cloud_key.key_switching_key.items.len;
        const ksk = try allocator.alloc(u32, rows * CT_WORDS);
        defer allocator.free(ksk);
        for (cloud_key.key_switching_key.items, 0..) |*row, i| {
            @memcpy(ksk[i * CT_WORDS .. (i + 1) * CT_WORDS], &row.p);
        }
        try cuda.check(cuda.tfhe_b200_load_key(ctx.?, bsk.ptr, ksk.ptr, CT_WORDS * @sizeOf(u32), cloud_key.decomposition_offset));
        return Self{ .ctx = ctx.? };
    }

    pub fn deinit(self: *Self) void {
        cuda.tfhe_b200_destroy(self.ctx);
    }

    /// same contract as VanillaBootstrap.bootstrap (src/bootstrap/vanilla.zig:38-52)
    pub fn bootstrap(self: *const Self, ctxt: *const utils.Ciphertext, cloud_key: *const key.CloudKey) !utils.Ciphertext {
        _ = cloud_key; // already resident on the device
        var out = utils.Ciphertext.new();
        try cuda.check(cuda.tfhe_b200_bootstrap_batch(self.ctx, &ctxt.p, &out.p, 1, null, 0));
        return out;
    }

    /// src/bootstrap/vanilla.zig:58-69
    pub fn bootstrapWithoutKeySwitch(self: *const Self, ctxt: *const utils.Ciphertext, cloud_key: *const key.CloudKey) !utils.Ciphertext {
        _ = cloud_key;
        var out = utils.Ciphertext.new();
        try cuda.check(cuda.tfhe_b200_bootstrap_no_keyswitch_batch(self.ctx, &ctxt.p, &out.p, 1));
        return out;
    }

    pub fn name(self: *const Self) []const u8 {
        _ = self;
        return "b200";
    }

    /// The `bootstrapLut` src/lut.zig:42 documents and the reference never defines: programmable bootstrap of one
    /// ciphertext with a lookup table from lut.Generator (src/lut/generator.zig:85-135): blindRotateWithTestvec
    /// (src/trgsw.zig:336-400) -> sampleExtractIndex(., 0) -> identityKeySwitching, all on the device.
    pub fn bootstrapLut(self: *const Self, ctxt: *const utils.Ciphertext, table: *const lut.LookupTable) !utils.Ciphertext {
        var tv: [2 * N]u32 = undefined; // TRLWELv1 packed: a then b (src/trlwe.zig:15-17)
        @memcpy(tv[0..N], &table.poly.a);
        @memcpy(tv[N .. 2 * N], &table.poly.b);
        var out = utils.Ciphertext.new();
        try cuda.check(cuda.tfhe_b200_bootstrap_batch(self.ctx, &ctxt.p, &out.p, 1, &tv, 0));
        return out;
    }

    /// `count` programmable bootstraps in one call.  `tables`: torus values Encoder.encode(f(x)), x < message_modulus
    /// (src/lut/encoder.zig:66-73), one table shared by all items (tables.len == message_modulus) or one per item
    /// (tables.len == count * message_modulus); the test vectors are built on the device (generator.zig:150-191).
    /// Caller-freed result slice, like gates.batch*.
    pub fn batchLut(
        self: *const Self,
        allocator: std.mem.Allocator,
        inputs: []const utils.Ciphertext,
        tables: []const params.Torus,
        message_modulus: usize,
    ) ![]utils.Ciphertext {
        const count = inputs.len;
        std.debug.assert(tables.len == message_modulus or tables.len == count * message_modulus);
        const in_words = try allocator.alloc(u32, count * CT_WORDS);
        defer allocator.free(in_words);
        const out_words = try allocator.alloc(u32, count * CT_WORDS);
        defer allocator.free(out_words);
        for (inputs, 0..) |*c, i| @memcpy(in_words[i * CT_WORDS .. (i + 1) * CT_WORDS], &c.p);
        const per_item: c_int = if (tables.len == message_modulus) 0 else 1;
        try cuda.check(cuda.tfhe_b200_lut_bootstrap_batch(self.ctx, in_words.ptr, out_words.ptr, count, tables.ptr, @intCast(message_modulus), per_item));
        const result = try allocator.alloc(utils.Ciphertext, count);
        for (result, 0..) |*r, i| {
            r.* = utils.Ciphertext.new();
            @memcpy(&r.p, out_words[i * CT_WORDS .. (i + 1) * CT_WORDS]);
        }
        return result;
    }

    /// Body for gates.batchNand/And/Or/Xor/Nor/Xnor (src/gates.zig:244-295): same signature shape
    /// (inputs as pairs, caller-freed result slice from the given allocator).
    pub fn batchGate(
        self: *const Self,
        allocator: std.mem.Allocator,
        op: cuda.Gate,
        inputs: []const struct { utils.Ciphertext, utils.Ciphertext },
    ) ![]utils.Ciphertext {
        const count = inputs.len;
        const a = try allocator.alloc(u32, count * CT_WORDS);
        defer allocator.free(a);
        const b = try allocator.alloc(u32, count * CT_WORDS);
        defer allocator.free(b);
        const o = try allocator.alloc(u32, count * CT_WORDS);
        defer allocator.free(o);
        for (inputs, 0..) |pair, i| {
            @memcpy(a[i * CT_WORDS .. (i + 1) * CT_WORDS], &pair[0].p);
            @memcpy(b[i * CT_WORDS .. (i + 1) * CT_WORDS], &pair[1].p);
        }
        try cuda.check(cuda.tfhe_b200_gate_batch(self.ctx, @intFromEnum(op), a.ptr, b.ptr, o.ptr, count));
        const result = try allocator.alloc(utils.Ciphertext, count);
        for (result, 0..) |*r, i| {
            r.* = utils.Ciphertext.new();
            @memcpy(&r.p, o[i * CT_WORDS .. (i + 1) * CT_WORDS]);
        }
        return result;
    }

    /// `count` independent W-bit additions in one call: the batched form of examples/add_two_numbers.zig:56-73 (`add`),
    /// which chains fullAdder (:24-39) gate by gate.  a, b: [W][count] ciphertexts (LSB first), cin: [count];
    /// returns [W + 1][count] (sum bits, then the carry), caller-freed.
    pub fn batchAdd(
        self: *const Self,
        allocator: std.mem.Allocator,
        comptime W: usize,
        a: []const utils.Ciphertext,
        b: []const utils.Ciphertext,
        cin: []const utils.Ciphertext,
    ) ![]utils.Ciphertext {
        const count = cin.len;
        std.debug.assert(a.len == W * count and b.len == W * count);
        var nodes: [5 * W]cuda.GateNode = undefined;
        var outs: [W + 1]u32 = undefined;
        const n_in: u32 = 2 * W + 1;
        var carry: u32 = 2 * W;
        for (0..W) |i| {
            const g: u32 = n_in + @as(u32, @intCast(5 * i));
            nodes[5 * i + 0] = .{ .op = @intFromEnum(cuda.Gate.xor), .a = @intCast(i), .b = @intCast(W + i) };
            nodes[5 * i + 1] = .{ .op = @intFromEnum(cuda.Gate.@"and"), .a = @intCast(i), .b = @intCast(W + i) };
            nodes[5 * i + 2] = .{ .op = @intFromEnum(cuda.Gate.@"and"), .a = g, .b = carry };
            nodes[5 * i + 3] = .{ .op = @intFromEnum(cuda.Gate.xor), .a = g, .b = carry };
            nodes[5 * i + 4] = .{ .op = @intFromEnum(cuda.Gate.@"or"), .a = g + 1, .b = g + 2 };
            outs[i] = g + 3;
            carry = g + 4;
        }
        outs[W] = carry;
        var circuit: ?*cuda.Circuit = null;
        try cuda.check(cuda.tfhe_b200_circuit_create(self.ctx, &nodes, nodes.len, n_in, &outs, outs.len, &circuit));
        defer cuda.tfhe_b200_circuit_destroy(circuit);
        const in_words = try allocator.alloc(u32, n_in * count * CT_WORDS);
        defer allocator.free(in_words);
        for (a, 0..) |*c, i| @memcpy(in_words[i * CT_WORDS .. (i + 1) * CT_WORDS], &c.p);
        for (b, 0..) |*c, i| @memcpy(in_words[(W * count + i) * CT_WORDS .. (W * count + i + 1) * CT_WORDS], &c.p);
        for (cin, 0..) |*c, i| @memcpy(in_words[(2 * W * count + i) * CT_WORDS .. (2 * W * count + i + 1) * CT_WORDS], &c.p);
        const out_words = try allocator.alloc(u32, (W + 1) * count * CT_WORDS);
        defer allocator.free(out_words);
        try cuda.check(cuda.tfhe_b200_circuit_run(self.ctx, circuit.?, in_words.ptr, out_words.ptr, count));
        const result = try allocator.alloc(utils.Ciphertext, (W + 1) * count);
        for (result, 0..) |*r, i| {
            r.* = utils.Ciphertext.new();
            @memcpy(&r.p, out_words[i * CT_WORDS .. (i + 1) * CT_WORDS]);
        }
        return result;
    }
};

/// Process-wide device context for the free functions of src/gates.zig (batchNand ... batchXnor take only the cloud key):
/// created on first use for the cloud key it is first called with, on every visible GPU unless TFHE_B200_DEVICES=<count>
/// says otherwise; a different cloud key replaces the keys on the device.  Guarded by a mutex because the reference lets any
/// thread call gates with a shared *const CloudKey (src/fft.zig:983-992 keeps its only hidden state per thread); the device
/// context itself is not thread-safe, so batch calls through the singleton are serialised.
var singleton: ?GpuBootstrap = null;
var singleton_key: ?*const key.CloudKey = null;
var singleton_mutex: std.Thread.Mutex = .{};

pub fn gpu_singleton(cloud_key: *const key.CloudKey) !*GpuBootstrap {
    singleton_mutex.lock();
    defer singleton_mutex.unlock();
    if (singleton != null and singleton_key == cloud_key) return &singleton.?;
    if (singleton) |*old| {
        old.deinit();
        singleton = null;
    }
    var ids: [16]c_int = undefined;
    var count: usize = 1;
    if (std.posix.getenv("TFHE_B200_DEVICES")) |v| {
        count = @min(ids.len, @max(1, std.fmt.parseInt(usize, v, 10) catch 1));
    }
    for (0..count) |i| ids[i] = @intCast(i);
    singleton = try GpuBootstrap.init(std.heap.page_allocator, cloud_key, ids[0..count]);
    singleton_key = cloud_key;
    return &singleton.?;
}

/// serialises a batch call through the singleton (see above)
pub fn batchGateShared(op: cuda.Gate, inputs: []const struct { utils.Ciphertext, utils.Ciphertext }, cloud_key: *const key.CloudKey) ![]utils.Ciphertext {
    const g = try gpu_singleton(cloud_key);
    singleton_mutex.lock();
    defer singleton_mutex.unlock();
    // results come from page_allocator, as the reference's parallel module returns them (src/parallel/thread_pool.zig:55)
    return g.batchGate(std.heap.page_allocator, op, inputs);
}
