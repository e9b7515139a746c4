//! cuda.zig -- `extern fn` view of include/tfhe_b200.h (libtfhe_b200.so, CUDA sm_100a).
//!
//! Drop this file into zig-tfhe's `src/` and link the library from build.zig (see INTEGRATION.md).
//! NOTE: written against Zig 0.15.1 but NOT compiled in the authoring environment (no Zig toolchain
//! there); the ABI it describes is exercised through the identical ctypes binding and the C++ header.
const std = @import("std");

pub const Ctx = opaque {};

/// runtime mirror of params.SecurityParams (src/params.zig:36-67); extern = C layout
pub const Params = extern struct {
    n: i32,
    N: i32 = 1024,
    L: i32,
    bgbit: i32,
    basebit: i32,
    iks_t: i32,
};

pub const Status = enum(c_int) { ok = 0, invalid = 1, no_device = 2, cuda = 3, no_key = 4, not_implemented = 5, _ };

pub const Gate = enum(c_int) { nand = 0, @"or" = 1, @"and" = 2, xor = 3, xnor = 4, nor = 5, andny = 6, andyn = 7, orny = 8, oryn = 9 };

pub const Error = error{ InvalidArgument, NoDevice, CudaFailure, NoKey, NotImplemented, KeyFileIo };

pub fn check(rc: c_int) Error!void {
    return switch (rc) {
        0 => {},
        1 => Error.InvalidArgument,
        2 => Error.NoDevice,
        4 => Error.NoKey,
        5 => Error.NotImplemented,
        6 => Error.KeyFileIo,
        else => Error.CudaFailure,
    };
}

pub extern fn tfhe_b200_create(params: *const Params, device_ids: ?[*]const c_int, n_dev: c_int, out: *?*Ctx) c_int;
pub extern fn tfhe_b200_destroy(ctx: ?*Ctx) void;
pub extern fn tfhe_b200_last_error(ctx: ?*const Ctx) [*:0]const u8;
pub extern fn tfhe_b200_num_devices(ctx: ?*const Ctx) c_int;
pub extern fn tfhe_b200_load_key(ctx: *Ctx, bsk: [*]const f64, ksk: ?[*]const u32, ksk_row_stride_bytes: usize, decomposition_offset: u32) c_int;
/// key.CloudKey.new on the device from the host-held secret key (key_lv0: [n]u32, key_lv1: [N]u32); bsk_out / ksk_out optional
pub extern fn tfhe_b200_keygen(ctx: *Ctx, key_lv0: [*]const u32, key_lv1: [*]const u32, seed: u64, ksk_alpha: f64, bsk_alpha: f64, bsk_out: ?[*]f64, ksk_out: ?[*]u32) c_int;
/// flat cloud-key file (format in include/tfhe_b200.h); the key_file_* calls are host only
pub extern fn tfhe_b200_key_file_write(path: [*:0]const u8, params: *const Params, bsk: [*]const f64, ksk: ?[*]const u32, decomposition_offset: u32) c_int;
pub extern fn tfhe_b200_key_file_info(path: [*:0]const u8, params: ?*Params, decomposition_offset: ?*u32, bsk_bytes: ?*u64, ksk_bytes: ?*u64) c_int;
pub extern fn tfhe_b200_key_file_read(path: [*:0]const u8, bsk: ?[*]f64, ksk: ?[*]u32) c_int;
pub extern fn tfhe_b200_key_file_last_error() [*:0]const u8;
pub extern fn tfhe_b200_load_key_file(ctx: *Ctx, path: [*:0]const u8) c_int;
pub extern fn tfhe_b200_decomposition_offset(ctx: *const Ctx) u32;
pub extern fn tfhe_b200_set_mode(ctx: *Ctx, mode: c_int) c_int;
pub extern fn tfhe_b200_gate_batch(ctx: *Ctx, op: c_int, a: [*]const u32, b: [*]const u32, out: [*]u32, count: usize) c_int;
pub extern fn tfhe_b200_gate_batch_ops(ctx: *Ctx, ops: [*]const i32, a: [*]const u32, b: [*]const u32, out: [*]u32, count: usize) c_int;
pub extern fn tfhe_b200_bootstrap_batch(ctx: *Ctx, in: [*]const u32, out: [*]u32, count: usize, testvec: ?[*]const u32, tv_per_item: c_int) c_int;
/// programmable bootstrap from function tables [count][message_modulus] (or one shared table): lut.Generator runs on the device
pub extern fn tfhe_b200_lut_bootstrap_batch(ctx: *Ctx, in: [*]const u32, out: [*]u32, count: usize, tables: [*]const u32, message_modulus: c_int, per_item: c_int) c_int;
pub extern fn tfhe_b200_lut_generate(ctx: *Ctx, table: [*]const u32, message_modulus: c_int, testvec_out: [*]u32) c_int;
pub extern fn tfhe_b200_bootstrap_no_keyswitch_batch(ctx: *Ctx, in: [*]const u32, out: [*]u32, count: usize) c_int;
pub extern fn tfhe_b200_blind_rotate_batch(ctx: *Ctx, in: [*]const u32, trlwe_out: [*]u32, count: usize, testvec: ?[*]const u32, tv_per_item: c_int) c_int;
pub extern fn tfhe_b200_keyswitch_batch(ctx: *Ctx, lv1: [*]const u32, lv0: [*]u32, count: usize) c_int;
/// proxy_reenc.ProxyReencryptionKey / reencryptTLWELv0 (src/proxy_reenc.zig:123-306), batched
pub extern fn tfhe_b200_load_reencryption_key(ctx: *Ctx, key: [*]const u32, basebit: c_int, t: c_int) c_int;
pub extern fn tfhe_b200_reencrypt_batch(ctx: *Ctx, in: [*]const u32, out: [*]u32, count: usize) c_int;
pub extern fn tfhe_b200_not_batch(ctx: *Ctx, a: [*]const u32, out: [*]u32, count: usize) c_int;
pub extern fn tfhe_b200_sync(ctx: *Ctx) c_int;
/// tuning keys are listed in include/tfhe_b200.h ("kct", "twt", "ks_tc", "host_copy_threads", "circuit_lanes", ...)
pub extern fn tfhe_b200_set_tuning(ctx: *Ctx, key: [*:0]const u8, value: c_int) c_int;

/// gate circuits (include/tfhe_b200.h): a netlist levelised once, every level one batched launch pair over
/// (gates of the level) x (instances) -- the batched form of examples/add_two_numbers.zig:24-73
pub const Circuit = opaque {};
pub const wire_not: u32 = 0x80000000; // Gates.notGate of the referenced wire, folded into the consumer
pub const wire_true: u32 = 0x7FFFFFFE; // Gates.constant(true), src/gates.zig:144-151
pub const wire_false: u32 = 0x7FFFFFFD; // Gates.constant(false) = (0, 1 - 2^29)
pub const GateNode = extern struct { op: i32, a: u32, b: u32 };
pub extern fn tfhe_b200_circuit_create(ctx: *Ctx, gates: [*]const GateNode, n_gates: usize, n_inputs: usize, outputs: [*]const u32, n_outputs: usize, out: *?*Circuit) c_int;
pub extern fn tfhe_b200_circuit_destroy(circuit: ?*Circuit) void;
/// host-only validation + levelisation (no device needed)
pub extern fn tfhe_b200_circuit_plan(gates: [*]const GateNode, n_gates: usize, n_inputs: usize, outputs: [*]const u32, n_outputs: usize, n_levels: ?*usize, max_level_width: ?*usize, gate_level: ?[*]u32) c_int;
pub extern fn tfhe_b200_circuit_info(circuit: *const Circuit, n_levels: ?*usize, max_level_width: ?*usize, n_gates: ?*usize) c_int;
pub extern fn tfhe_b200_circuit_run(ctx: *Ctx, circuit: *Circuit, inputs: [*]const u32, outputs: [*]u32, instances: usize) c_int;
