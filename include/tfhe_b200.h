/*
 * tfhe_b200.h -- C ABI of the B200-native gate-bootstrapping path for zig-tfhe.
 *
 * The reference (thedonutfactory/zig-tfhe, pure Zig) has no FFI layer; its seams for this path
 * are Zig-level.  Each entry point below names the reference interface it
 * replaces (file:line under /root/reference/) -- these are exactly the functions a Zig
 * `extern fn` block binds (see INTEGRATION.md and zig-tfhe_b200/zig/cuda.zig).
 *
 * Conventions
 *   - plain pointers and sizes only; all multi-byte data little-endian, row-major.
 *   - host entry points take HOST pointers (pageable or pinned) owned by the caller and are
 *     synchronous: when they return, `out` is filled.
 *   - `_device` entry points take DEVICE pointers valid on the context's device `dev`, enqueue
 *     on that device's stream (tfhe_b200_stream) and return without synchronising.
 *   - return value: 0 = TFHE_B200_OK, otherwise a tfhe_b200_status; tfhe_b200_last_error()
 *     gives the message.  There is NO CPU fallback: without a usable CUDA device every call fails.
 *   - a context is not thread-safe (one per calling thread, or lock externally), like the
 *     reference's per-thread FFT plan (src/fft.zig:983-992).
 *
 * Ciphertext layouts (identical to the reference's in-memory arrays)
 *   TLWELv0  : uint32_t[n+1], mask a[0..n), body b at index n          (src/tlwe.zig:11-31)
 *   TLWELv1  : uint32_t[N+1]                                           (src/tlwe.zig:243-262)
 *   TRLWELv1 : uint32_t[2][N]  = a[N] then b[N]                        (src/trlwe.zig:15-17)
 *   CloudKey.bootstrapping_key : double[n][2L][2 (a,b)][N], each polynomial the reference's
 *       ifft1024 spectrum, re[0..512) then im[0..512), reference bin order
 *                                                                      (src/key.zig:61-65, src/trgsw.zig:75-91, src/trlwe.zig:104-132)
 *   CloudKey.key_switching_key : uint32_t[N*t*base][n+1], row = base*t*i + base*j + k; rows with
 *       k == 0 are never read                                          (src/key.zig:148-172, src/trgsw.zig:491)
 */
#ifndef TFHE_B200_H
#define TFHE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct tfhe_b200_ctx tfhe_b200_ctx;

/* runtime mirror of params.SecurityParams / TrgswParams (src/params.zig:36-67);
 * N must be 1024 (true for all 11 sets, src/params.zig:85-365). */
typedef struct {
    int32_t n;       /* tlwe_lv0.n   */
    int32_t N;       /* trgsw_lv1.n  */
    int32_t L;       /* trgsw_lv1.l  (1..3) */
    int32_t bgbit;   /* trgsw_lv1.bgbit */
    int32_t basebit; /* trgsw_lv1.basebit */
    int32_t iks_t;   /* trgsw_lv1.iks_t */
} tfhe_b200_params;

typedef enum {
    TFHE_B200_OK = 0,
    TFHE_B200_ERR_INVALID = 1,     /* bad argument / unsupported parameter set */
    TFHE_B200_ERR_NO_DEVICE = 2,   /* no CUDA device, or device is not sm_100 */
    TFHE_B200_ERR_CUDA = 3,        /* CUDA runtime error (see last_error) */
    TFHE_B200_ERR_NO_KEY = 4,      /* hot-path call before load_key */
    TFHE_B200_ERR_NOT_IMPLEMENTED = 5,
    TFHE_B200_ERR_IO = 6           /* cloud-key file cannot be opened / read / written */
} tfhe_b200_status;

/* gate opcodes: the ten bootstrapped two-input gates of Gates (src/gates.zig:48-121) */
typedef enum {
    TFHE_B200_NAND = 0,  /* gates.zig:48  */
    TFHE_B200_OR = 1,    /* gates.zig:57  */
    TFHE_B200_AND = 2,   /* gates.zig:64  */
    TFHE_B200_XOR = 3,   /* gates.zig:71  (a + 2b + 1/4) */
    TFHE_B200_XNOR = 4,  /* gates.zig:78  (a - 2b - 1/4, reference semantics kept) */
    TFHE_B200_NOR = 5,   /* gates.zig:85  */
    TFHE_B200_ANDNY = 6, /* gates.zig:94  */
    TFHE_B200_ANDYN = 7, /* gates.zig:102 */
    TFHE_B200_ORNY = 8,  /* gates.zig:109 */
    TFHE_B200_ORYN = 9   /* gates.zig:117 */
} tfhe_b200_gate;

/* arithmetic mode of the blind rotation */
typedef enum {
    TFHE_B200_MODE_FAST = 0,  /* radix-8 FMA transform; bit-identical to the reference on L=3/BGBIT=6 sets */
    TFHE_B200_MODE_EXACT = 1  /* replays the reference's radix-2 / recurrence-twiddle / no-FMA DAG (fft.zig:582-619) */
} tfhe_b200_mode;

/* ---- lifetime ------------------------------------------------------------------------ */
/* Replaces nothing in the reference (it has no device); one context owns n_dev devices and
 * replicates the keys on each (the CPU thread pool of src/parallel/thread_pool.zig:39-83 is
 * what this sharding stands in for).  device_ids == NULL -> devices 0..n_dev-1. */
int tfhe_b200_create(const tfhe_b200_params *params, const int *device_ids, int n_dev, tfhe_b200_ctx **out);
void tfhe_b200_destroy(tfhe_b200_ctx *ctx);
const char *tfhe_b200_last_error(const tfhe_b200_ctx *ctx);
int tfhe_b200_num_devices(const tfhe_b200_ctx *ctx);
const char *tfhe_b200_version(void);

/* ---- keys: CloudKey (src/key.zig:61-77) --------------------------------------------- */
/* Uploads + re-lays-out the cloud key on every device of the context.
 * ksk may be NULL (CloudKey.newNoKsk, src/key.zig:80-100): key-switching calls then fail.
 * ksk_row_stride_bytes: distance between consecutive KSK rows (>= (n+1)*4). */
int tfhe_b200_load_key(tfhe_b200_ctx *ctx, const double *bsk, const uint32_t *ksk,
                       size_t ksk_row_stride_bytes, uint32_t decomposition_offset);
/* Same, from DEVICE buffers already resident on device `dev` (e.g. after an NCCL broadcast);
 * layouts as above, KSK rows packed ((n+1)*4 bytes). */
int tfhe_b200_load_key_device(tfhe_b200_ctx *ctx, int dev, const double *d_bsk, const uint32_t *d_ksk,
                              uint32_t decomposition_offset);
/* key.CloudKey.new (src/key.zig:70-77): genKeySwitchingKey + genBootstrappingKey (src/key.zig:148-212) ON THE DEVICE, from
 * the secret key the caller holds (key_lv0: uint32_t[n], key_lv1: uint32_t[N], entries 0/1 as in key.SecretKey,
 * src/key.zig:23-58).  ksk_alpha / bsk_alpha = params.KSK_ALPHA / BSK_ALPHA (src/params.zig:419-422).
 * Randomness: Philox4x32-10 keyed by `seed` (the reference's own seeding is a clock, src/utils.zig:16-22), so every
 * device of the context generates the identical key and a key is reproducible from its seed.
 * SECURITY: every mask and noise word of the cloud key is a function of `seed`, so whoever learns the seed can strip the
 * noise and solve for both secret keys -- the seed is secret-key-equivalent and caps the key's security at 64 bits.  Draw it
 * from the same entropy source as the secret key, never reuse or publish it; fixed seeds are for tests and benchmarks only.
 * (The reference's clock-seeded xoshiro is not cryptographic either; production keys want a >= 128-bit seed.)  The keys are written
 * straight into the device layouts; bsk_out / ksk_out (either may be NULL) receive the reference layouts described
 * at the top of this file, e.g. to hand the same CloudKey to the CPU implementation or to serialise it. */
int tfhe_b200_keygen(tfhe_b200_ctx *ctx, const uint32_t *key_lv0, const uint32_t *key_lv1, uint64_t seed, double ksk_alpha,
                     double bsk_alpha, double *bsk_out, uint32_t *ksk_out);
/* ---- flat cloud-key file ----------------------------------------------------------------
 * key.CloudKey (src/key.zig:61-65) has no serialised form in the reference and every test run regenerates it (~30 s,
 * src/key.zig:240).  File format, little endian, version 1:
 *   0    char[8]  "TFHEB2CK"          8   u32 version (1)           12  u32 header_bytes (4096)
 *   16   i32[6]   n, N, L, bgbit, basebit, iks_t (tfhe_b200_params)
 *   40   u32 decomposition_offset     44  u32 flags (bit 0: key-switching key present)
 *   48   u64 bsk_offset (4096), u64 bsk_bytes, u64 ksk_offset (4096-aligned), u64 ksk_bytes
 *   80   u64 bsk_checksum, u64 ksk_checksum       96  u64 header_checksum (over bytes 0..96)
 *   then zero padding; the two sections are CloudKey.bootstrapping_key and CloudKey.key_switching_key exactly as laid
 *   out at the top of this header (packed KSK rows), so a host can mmap the file and use the sections in place.
 *   checksum(data): four FNV-1a-64 lanes (basis 0xcbf29ce484222325 + lane, prime 0x100000001b3) over the little-endian
 *   64-bit words of data, word i into lane i mod 4 (a trailing partial word zero-extended); then h = basis; for each lane
 *   h = (h ^ lane) * prime; result (h ^ byte_length) * prime.
 * To save a generated key: tfhe_b200_keygen(..., bsk_out, ksk_out) then tfhe_b200_key_file_write.
 * The three key_file_* calls are host-only (no context, no device); errors: TFHE_B200_ERR_IO / _INVALID, text in
 * tfhe_b200_key_file_last_error() (thread local).  write() goes through a temporary + rename. */
int tfhe_b200_key_file_write(const char *path, const tfhe_b200_params *params, const double *bsk, const uint32_t *ksk /* or NULL */,
                             uint32_t decomposition_offset);
/* header only (no payload checksum pass); any out pointer may be NULL; ksk_bytes == 0: file has no key-switching key */
int tfhe_b200_key_file_info(const char *path, tfhe_b200_params *params, uint32_t *decomposition_offset, uint64_t *bsk_bytes,
                            uint64_t *ksk_bytes);
/* verifies both checksums, then copies the sections out (either pointer may be NULL) */
int tfhe_b200_key_file_read(const char *path, double *bsk, uint32_t *ksk);
const char *tfhe_b200_key_file_last_error(void);
/* tfhe_b200_load_key straight from a file: mmap, verify header + checksums, parameter set must equal the context's,
 * upload to every device of the context */
int tfhe_b200_load_key_file(tfhe_b200_ctx *ctx, const char *path);

/* CloudKey.decomposition_offset of the loaded / generated key (key.genDecompositionOffset, src/key.zig:121-131) */
uint32_t tfhe_b200_decomposition_offset(const tfhe_b200_ctx *ctx);
int tfhe_b200_set_mode(tfhe_b200_ctx *ctx, int mode);

/* ---- hot path, host buffers ---------------------------------------------------------- */
/* gates.batchNand/And/Or/Xor/Nor/Xnor (src/gates.zig:244-295, placeholders returning
 * error.NotImplemented) and the scalar Gates.*Gate (src/gates.zig:48-121): out[i] = gate(a[i], b[i]).
 * a, b, out: [B][n+1].  The batch is split contiguously over the context's devices. */
int tfhe_b200_gate_batch(tfhe_b200_ctx *ctx, int op, const uint32_t *a, const uint32_t *b, uint32_t *out, size_t B);
/* same with one opcode per item (mixed AND/XOR batches, one circuit level) */
int tfhe_b200_gate_batch_ops(tfhe_b200_ctx *ctx, const int32_t *ops, const uint32_t *a, const uint32_t *b,
                             uint32_t *out, size_t B);
/* VanillaBootstrap.bootstrap (src/bootstrap/vanilla.zig:38-52) over a batch; in/out [B][n+1].
 * testvec == NULL -> CloudKey.blind_rotate_testvec (src/key.zig:134-145); else a TRLWE [2][N]
 * shared by the batch (tv_per_item == 0) or one per item [B][2][N] (tv_per_item != 0):
 * trgsw.blindRotateWithTestvec (src/trgsw.zig:336-400) + sampleExtractIndex(.,0) + identityKeySwitching,
 * i.e. the programmable (LUT) bootstrap the reference documents in src/lut.zig:42. */
int tfhe_b200_bootstrap_batch(tfhe_b200_ctx *ctx, const uint32_t *in, uint32_t *out, size_t B,
                              const uint32_t *testvec, int tv_per_item);
/* Programmable bootstrap from FUNCTION TABLES: lut.Generator.generateLookupTableFull (src/lut/generator.zig:150-191)
 * runs on the device, then the `bootstrapLut` src/lut.zig:42 documents.  tables: [B][message_modulus] (per_item != 0)
 * or [message_modulus] torus values, entry x = the output for message x (Encoder.encode(f(x)), src/lut/encoder.zig:66-73).
 * Per-item tables cost message_modulus words of host-to-device traffic per item instead of an 8 KiB test vector. */
/* SEVERAL functions from ONE blind rotation (the reference has no such call).  tables:
 * [n_functions][message_modulus] torus values as above, n_functions a power of two with n_functions * 2 * message_modulus <= N.
 * The test vector interleaves the functions' lookup tables (position n_functions * i + f = table f at position
 * n_functions * i), the modulus switch of blindRotate (src/trgsw.zig:297,312) rounds to multiples of n_functions, so the
 * accumulator holds function f at every index congruent to f, and output f is sampleExtractIndex(., f)
 * (src/trlwe.zig:146-162) followed by identityKeySwitching.  out: [n_functions][B][n+1].  Cost: one blind rotation plus
 * n_functions key switches per item; the coarser switch adds up to n_functions / 2 positions of rounding error, to be kept
 * well inside the N / (2 * message_modulus) half-slot. */
int tfhe_b200_lut_bootstrap_many_batch(tfhe_b200_ctx *ctx, const uint32_t *in, uint32_t *out, size_t B, const uint32_t *tables,
                                       int n_functions, int message_modulus);
int tfhe_b200_lut_bootstrap_batch(tfhe_b200_ctx *ctx, const uint32_t *in, uint32_t *out, size_t B, const uint32_t *tables,
                                  int message_modulus, int per_item);
/* parity tap: the device-built LookupTable.poly [2][N] of one table (src/lut/lookup_table.zig:16-20) */
int tfhe_b200_lut_generate(tfhe_b200_ctx *ctx, const uint32_t *table, int message_modulus, uint32_t *testvec_out);
/* VanillaBootstrap.bootstrapWithoutKeySwitch (src/bootstrap/vanilla.zig:58-69): blind rotation +
 * sampleExtractIndex2(.,0) (src/trlwe.zig:165-180); out [B][n+1]. */
int tfhe_b200_bootstrap_no_keyswitch_batch(tfhe_b200_ctx *ctx, const uint32_t *in, uint32_t *out, size_t B);
/* trgsw.batchBlindRotate / blindRotate / blindRotateWithTestvec (src/trgsw.zig:290-436):
 * parity tap, trlwe_out [B][2][N]. */
int tfhe_b200_blind_rotate_batch(tfhe_b200_ctx *ctx, const uint32_t *in, uint32_t *trlwe_out, size_t B,
                                 const uint32_t *testvec, int tv_per_item);
/* trgsw.identityKeySwitching (src/trgsw.zig:471-502) over a batch: lv1 [B][N+1] -> lv0 [B][n+1]. */
int tfhe_b200_keyswitch_batch(tfhe_b200_ctx *ctx, const uint32_t *lv1, uint32_t *lv0, size_t B);
/* trlwe.sampleExtractIndex(., 0) (src/trlwe.zig:146-162) fused after the blind rotation:
 * in [B][n+1] -> lv1 [B][N+1] (parity tap between blind rotation and key switch). */
int tfhe_b200_blind_rotate_extract_batch(tfhe_b200_ctx *ctx, const uint32_t *in, uint32_t *lv1_out, size_t B);
/* proxy_reenc.ProxyReencryptionKey + reencryptTLWELv0 (src/proxy_reenc.zig:123-306) over a batch: the same digit-gather-
 * subtract as the key switch with source dimension n.  key: uint32_t[n*t*base][n+1], row = base*t*i + base*j + k
 * (k = 0 rows ignored); in/out [B][n+1]. */
int tfhe_b200_load_reencryption_key(tfhe_b200_ctx *ctx, const uint32_t *key, int basebit, int t);
int tfhe_b200_reencrypt_batch(tfhe_b200_ctx *ctx, const uint32_t *in, uint32_t *out, size_t B);
/* Gates.notGate / copy (src/gates.zig:131-141): no bootstrap, out = -a. */
int tfhe_b200_not_batch(tfhe_b200_ctx *ctx, const uint32_t *a, uint32_t *out, size_t B);

/* ---- gate circuits: level-batched, device-resident ------------------------------------- */
/* The reference evaluates a circuit (examples/add_two_numbers.zig:24-73: fullAdder, ripple-carry add) one
 * Gates.* call after another (src/gates.zig:48-121).  Here a netlist is levelised once and every dependency
 * level runs as ONE batched launch pair over (gates of the level) x (instances); wires never leave the device
 * between levels and the level sequence is replayed as a CUDA graph.  Instances are independent, so a
 * multi-device context shards them with no cross-device traffic.
 * Wire ids: 0 .. n_inputs-1 are the circuit inputs, n_inputs + g is the output of gate g; gates are given in
 * topological order.  TFHE_B200_WIRE_NOT on a wire reference is Gates.notGate (src/gates.zig:131-133) of that wire:
 * free, folded into the consuming gate's linear part (or applied on the way out for a circuit output). */
#define TFHE_B200_WIRE_NOT 0x80000000u
/* Gates.constant(true) / Gates.constant(false) (src/gates.zig:144-151) as wires: trivial ciphertexts (mask 0, body 2^29 resp.
 * 1 - 2^29 = 0xE0000001, the reference's value) that the executor materialises on the device; usable wherever a wire id is,
 * also under TFHE_B200_WIRE_NOT and as a circuit output.  Gates.copy (src/gates.zig:138-141) is the wire itself. */
#define TFHE_B200_WIRE_TRUE 0x7FFFFFFEu
#define TFHE_B200_WIRE_FALSE 0x7FFFFFFDu
typedef struct {
    int32_t op;  /* tfhe_b200_gate */
    uint32_t a;  /* wire id | TFHE_B200_WIRE_NOT */
    uint32_t b;
} tfhe_b200_gate_node;
typedef struct tfhe_b200_circuit tfhe_b200_circuit;
int tfhe_b200_circuit_create(tfhe_b200_ctx *ctx, const tfhe_b200_gate_node *gates, size_t n_gates, size_t n_inputs,
                             const uint32_t *outputs, size_t n_outputs, tfhe_b200_circuit **out);
void tfhe_b200_circuit_destroy(tfhe_b200_circuit *circuit);
/* The host-only half of circuit_create (no device, no context): validates the netlist and reports the number of dependency
 * levels, the widest level and (gate_level != NULL) the level of every gate, 1-based.  Same error codes as circuit_create. */
int tfhe_b200_circuit_plan(const tfhe_b200_gate_node *gates, size_t n_gates, size_t n_inputs, const uint32_t *outputs,
                           size_t n_outputs, size_t *n_levels, size_t *max_level_width, uint32_t *gate_level);
int tfhe_b200_circuit_info(const tfhe_b200_circuit *circuit, size_t *n_levels, size_t *max_level_width, size_t *n_gates);
/* inputs: [n_inputs][instances][n+1], outputs: [n_outputs][instances][n+1] (host buffers, wire-major). */
int tfhe_b200_circuit_run(tfhe_b200_ctx *ctx, tfhe_b200_circuit *circuit, const uint32_t *inputs, uint32_t *outputs,
                          size_t instances);

/* ---- hot path, device buffers (single device `dev` of the context, asynchronous) ------- */
int tfhe_b200_gate_batch_device(tfhe_b200_ctx *ctx, int dev, int op, const int32_t *d_ops, const uint32_t *d_a,
                                const uint32_t *d_b, uint32_t *d_out, size_t B);
int tfhe_b200_bootstrap_batch_device(tfhe_b200_ctx *ctx, int dev, const uint32_t *d_in, uint32_t *d_out, size_t B,
                                     const uint32_t *d_testvec, int tv_per_item);
int tfhe_b200_blind_rotate_batch_device(tfhe_b200_ctx *ctx, int dev, const uint32_t *d_in, uint32_t *d_trlwe_out,
                                        size_t B, const uint32_t *d_testvec, int tv_per_item);
int tfhe_b200_keyswitch_batch_device(tfhe_b200_ctx *ctx, int dev, const uint32_t *d_lv1, uint32_t *d_lv0, size_t B);
/* cudaStream_t of device `dev` (as void*), and a full synchronisation of all devices */
void *tfhe_b200_stream(tfhe_b200_ctx *ctx, int dev);
int tfhe_b200_sync(tfhe_b200_ctx *ctx);

/* ---- instrumentation ----------------------------------------------------------------- */
/* Largest |t - round(t)| seen in the inverse-transform rounding epilogue (src/fft.zig:416-424)
 * since the last reset, over all devices; tracked only while enabled (slower kernel variant).
 * Standing proof of the exactness margin claimed for the L=3/BGBIT=6 sets (alarm above 0.25). */
int tfhe_b200_track_margin(tfhe_b200_ctx *ctx, int enable);
double tfhe_b200_max_round_margin(tfhe_b200_ctx *ctx, int reset);
/* number of kernels this library launched since the context was created (bench `gpu_launches`) */
uint64_t tfhe_b200_launch_count(const tfhe_b200_ctx *ctx);
/* tuning knobs (tests/bench): "kct" ciphertexts per CTA of the blind-rotation kernel (0 = automatic: whole waves at 6 -- twiddles
 * in tensor memory -- plus a separately sized tail launch), "twt" tensor-memory twiddles (0 = where they win, i.e. at six per CTA;
 * 1 = also at 4 and 5; -1 = never, round-1 kernels), "ks_tc" key switch on the tensor cores (0 = from "ks_tc_min" = 192 ciphertexts
 * up on the BASEBIT = 2 sets, 1 = always there, -1 = never), "exact_kct" ciphertexts per CTA of the exact kernel (0 = automatic,
 * 1..4, 6), "exact_legacy" 1 = round-1 exact kernel, "host_copy_threads" memcpy threads that stage large PAGEABLE host buffers
 * through pinned memory (default 8, 0 = plain cudaMemcpyAsync), "host_pipeline" 1 = gate / bootstrap batches of >= 32,768 ciphertexts per
 * device run as four chunks whose copies (staging, from pageable memory) overlap the kernels (default; 2 = pageable callers only, 0 = never), "inject_fault" test hook: device value - 1 fails its next
 * host-batch shard, "use_tma" 0/1, "latency_mode" (1 = automatic: batches <= SMs/2 on two-CTA clusters, <= SMs
 * on one CTA per ciphertext; 2 = never the cluster kernel; 0 = throughput kernel only), "team" (2: two ciphertexts per
 * warp, a measured K1 variant kept for A/B runs), "max_chunk" ciphertexts per launch, "timing" 0/1 (record CUDA events
 * around K1/K2), "ks_tile" / "ks_vec" key-switch tile shape, "ks_fill" CTAs per SM the key switch splits its mask range for (0 = automatic),
 * "ks_rot" -1 = no staggered starting points, "circuit_graph" 0/1 (CUDA-graph replay of circuit levels),
 * "circuit_lanes" concurrent instance groups per device (read at circuit_create; default 4) */
int tfhe_b200_set_tuning(tfhe_b200_ctx *ctx, const char *key, int value);
/* with "timing" on: device time in ms of the last blind-rotation (which = 0) or key-switch (which = 1)
 * kernel enqueued on device `dev`, measured with CUDA events on the launching stream */
double tfhe_b200_last_kernel_ms(tfhe_b200_ctx *ctx, int dev, int which);
/* microbenchmarks used for the roofline denominators (profiles/): returns achieved rate */
double tfhe_b200_measure_fp64_tflops(tfhe_b200_ctx *ctx, int dev);

#ifdef __cplusplus
}
#endif
#endif
