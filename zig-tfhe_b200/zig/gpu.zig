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
};
