// capi.cu -- the C-ABI boundary of libtfhe_b200 (include/tfhe_b200.h): contexts, key upload and
// re-layout, host<->device staging, batch sharding over the context's devices, kernel dispatch.
// There is no CPU fallback anywhere in this file: every hot-path entry point launches the CUDA
// kernels or fails with an error code.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/tfhe_b200.h"
#include "host_tables.h"
#include "kernels.cuh"
#include "key_file.h"

using namespace tfhe_b200;

namespace {

struct Buf {
    void *p = nullptr;
    size_t cap = 0;
};

struct Device {
    int id = -1;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    cplx *bsk = nullptr;
    uint32_t *ksk = nullptr;
    uint8_t *ksk_tc = nullptr;        // key-switching key as tensor-core operand tiles (keyswitch_tc.cu), BASEBIT = 2 sets
    uint32_t *reenc = nullptr;        // proxy re-encryption key, same device layout as ksk with N -> n
    cplx *tw2 = nullptr, *tw3 = nullptr;
    double *exact_tables = nullptr;   // make_exact_tables(), exact mode
    double *bsk_ref = nullptr;        // bootstrapping key in the reference layout (legacy exact kernel, key export)
    cplx *bsk_x = nullptr;            // bootstrapping key in the exact chunk layout (register-blocked exact kernel)
    cplx *exact_shared = nullptr;     // make_exact_shared_tables(): twist[512] in acc_pos order + 7 pass-A twiddles
    unsigned long long *margin_bits = nullptr;
    Buf a, b, out, lv1, ops, tv, trlwe, lut, ksdig;
    void *stage[2] = {nullptr, nullptr};            // host_copy: pinned staging buffers (allocated on first use)
    cudaEvent_t stage_ev[2] = {nullptr, nullptr};
    // pipelined host batches (run_host_device_pipelined): second set of batch buffers, a copy stream, hand-over events
    cudaStream_t copy_stream = nullptr;
    Buf a2, b2, out2, ops2;
    cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr};
    cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};   // K1 start, K1 end / K2 start, K2 end (timing mode)
    bool ev_valid = false;
    uint64_t launches = 0;             // kernels launched on this device by its host thread (summed by tfhe_b200_launch_count)
    bool has_key = false, has_ksk = false;   // this device holds a bootstrapping / key-switching key
};

}  // namespace

struct tfhe_b200_ctx {
    tfhe_b200_params prm{};
    std::vector<Device> devs;
    std::string err;
    std::mutex err_mu;                 // the per-device host threads of a multi-device batch may fail concurrently
    int mode = TFHE_B200_MODE_FAST;
    bool track_margin = false;
    bool has_key = false, has_ksk = false, has_reenc = false;
    int reenc_basebit = 0, reenc_t = 0;
    uint32_t offset = 0;
    int ksk_pitch = 0;
    BrTuning tune;
    uint64_t launches = 0;
    bool timing = false;
    int ks_tile = 0, ks_vec = 0, ks_fill = 0, ks_rot = 0;   // key-switch tuning overrides (0 = automatic)
    int inject_fault = 0;                 // test hook (tuning key "inject_fault"): device k = value - 1 fails its next host-batch shard
    int host_pipeline = 1;                // large host batches: overlap the copies / staging of chunk k + 1 with the kernels of chunk k (2: pageable callers only)
    int host_copy_threads = 8;            // large copies from / to PAGEABLE caller memory are staged through pinned buffers by this many
                                          // memcpy threads (0 = plain cudaMemcpyAsync from the caller's buffer)
    int ks_tc = 0;                        // tensor-core key switch: 0 = automatic (batches >= ks_tc_min), 1 = always, -1 = never
    int ks_tc_min = 192;
    size_t max_chunk = (size_t)1 << 18;   // ciphertexts per device per launch
    bool circuit_graph = true;            // replay a circuit's level sequence as one CUDA graph
    int exact_legacy = 0;                 // 1: round-1 exact kernel (one CTA per ciphertext, shared-memory butterflies)
    int exact_kct = 0;                    // ciphertexts per CTA of the register-blocked exact kernel (0 = automatic)
    bool exact_conjugate = true;          // host stage tables are conjugate-symmetric (else the legacy kernel is used)
    int circuit_lanes = 4;                // independent instance groups per device, each on its own stream (set before circuit_create)
    // Bumped by everything a captured circuit graph bakes into its kernel nodes: key buffers (load_key*, keygen), mode,
    // margin tracking, tuning.  A circuit re-captures when its graph was recorded under another epoch.
    uint64_t config_epoch = 1;
};

namespace {

int fail(tfhe_b200_ctx *c, int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (c) {
        std::lock_guard<std::mutex> lk(c->err_mu);
        c->err = buf;
    }
    return code;
}

 // a context has a key only when every one of its devices does (tfhe_b200_load_key_device loads one device at a time)
void refresh_key_flags(tfhe_b200_ctx *c) {
    bool k = !c->devs.empty(), s = !c->devs.empty();
    for (const Device &d : c->devs) { k = k && d.has_key; s = s && d.has_ksk; }
    c->has_key = k;
    c->has_ksk = s;
    c->config_epoch++;
}

#define CU(c, expr)                                                                                       \
    do {                                                                                                  \
        cudaError_t e__ = (expr);                                                                         \
        if (e__ != cudaSuccess)                                                                           \
            return fail(c, TFHE_B200_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)

int ensure(tfhe_b200_ctx *c, Buf &b, size_t bytes) {
    if (b.cap >= bytes) return 0;
    if (b.p) CU(c, cudaFree(b.p));
    b.p = nullptr;
    b.cap = 0;
    CU(c, cudaMalloc(&b.p, bytes));
    b.cap = bytes;
    return 0;
}

bool wide_round(const tfhe_b200_params &p) {
    // |coefficient| of the inverse transform < N * 2L * 2^(bgbit-1) * 2^31; magic rounding needs < 2^51
    double bits = 10 + 1 + (p.L > 1 ? 2 : 0) + (p.bgbit - 1) + 31;
    return bits >= 50;
}

// K2 on device buffers: tensor-core contraction for large batches on the BASEBIT = 2 sets, the scalar kernel otherwise.
// ks_digits: caller-owned digit scratch (circuit lanes, sized before graph capture) or nullptr -> the device's own.
int run_keyswitch(tfhe_b200_ctx *c, Device &d, const uint32_t *lv1, uint32_t *lv0, size_t B, uint64_t *ks_digits, uint64_t *launches) {
    if (B == 0) return 0;
    KsArgs K{lv1, lv0, d.ksk, (uint32_t)B, c->prm.n, c->prm.basebit, c->prm.iks_t, c->ksk_pitch, kN, c->ks_tile, c->ks_vec};
    K.fill = c->ks_fill; K.rot = c->ks_rot;
    const bool tc = d.ksk_tc && c->ks_tc >= 0 && (c->ks_tc > 0 || B >= (size_t)c->ks_tc_min);
    if (tc) {
        if (!ks_digits) {
            if (int r = ensure(c, d.ksdig, B * keyswitch_tc_digit_words(c->prm.basebit, c->prm.iks_t) * 8)) return r;
            ks_digits = (uint64_t *)d.ksdig.p;
        }
        CU(c, launch_keyswitch_tc(K, d.ksk_tc, ks_digits, d.stream, launches));
    } else {
        CU(c, launch_keyswitch(K, d.sm_count, d.stream, launches));
    }
    return 0;
}

// K1 (+K2) on device buffers of one device
struct LevelRef {            // one dependency level of a circuit: operands are wire rows of d_a (see BrArgs)
    const int32_t *ops;
    const uint32_t *a, *b;
    uint32_t inst;
};

int run_device(tfhe_b200_ctx *c, Device &d, int op, const int32_t *d_ops, const uint32_t *d_a, const uint32_t *d_b, uint32_t *d_lv0,
               uint32_t *d_lv1_out, uint32_t *d_trlwe, size_t B, const uint32_t *d_tv, int tv_per_item,
               const LevelRef *lvl = nullptr, bool concurrent = false, uint64_t *ks_digits = nullptr, int ms_shift = 0) {
    if (!c->has_key) return fail(c, TFHE_B200_ERR_NO_KEY, "no cloud key loaded");
    if (B == 0) return 0;
    CU(c, cudaSetDevice(d.id));
    uint32_t *lv1 = d_lv1_out;
    if (d_lv0 && !lv1) {
        if (int r = ensure(c, d.lv1, B * (size_t)(kN + 1) * 4)) return r;
        lv1 = (uint32_t *)d.lv1.p;
    }
    BrArgs A{};
    A.in_a = d_a; A.in_b = d_b; A.ops = d_ops; A.op = op;
    if (lvl) { A.lvl_ops = lvl->ops; A.lvl_a = lvl->a; A.lvl_b = lvl->b; A.inst = lvl->inst; }
    A.bsk = d.bsk; A.tw2 = d.tw2; A.tw3 = d.tw3;
    A.testvec = d_tv; A.tv_per_item = tv_per_item;
    A.out_lv1 = lv1; A.out_trlwe = d_trlwe;
    A.margin_bits = c->track_margin ? d.margin_bits : nullptr;
    A.B = (uint32_t)B; A.n = c->prm.n; A.L = c->prm.L; A.bgbit = c->prm.bgbit;
    A.offset = c->offset; A.wide_round = wide_round(c->prm) ? 1 : 0;
    A.ms_shift = ms_shift;
#ifdef TFHE_B200_DIAG
    A.diag = c->tune.diag;
#endif
    d.ev_valid = false;
    if (c->timing) CU(c, cudaEventRecord(d.ev[0], d.stream));
    if (c->mode == TFHE_B200_MODE_EXACT) {
        ExactArgs X{d.exact_tables, d.bsk_ref, d.bsk_x, d.exact_shared, (c->exact_legacy || !c->exact_conjugate) ? 1 : 0, c->exact_kct, d.sm_count};
        CU(c, launch_blind_rotate_exact(A, X, c->track_margin, d.stream, &d.launches));
    } else {
        BrTuning tune = c->tune;
        tune.sm_count = d.sm_count;
        tune.concurrent = concurrent ? 1 : 0;
        CU(c, launch_blind_rotate(A, tune, c->track_margin, d.stream, &d.launches));
    }
    if (c->timing) CU(c, cudaEventRecord(d.ev[1], d.stream));
    if (d_lv0) {
        if (!c->has_ksk) return fail(c, TFHE_B200_ERR_NO_KEY, "no key-switching key loaded");
        if (int r = run_keyswitch(c, d, lv1, d_lv0, B, ks_digits, &d.launches)) return r;
    }
    if (c->timing) {
        CU(c, cudaEventRecord(d.ev[2], d.stream));
        d.ev_valid = true;
    }
    return 0;
}

enum class Out { LV0, LV0_NOKS, LV1, TRLWE };

// One copy between a caller buffer and device memory, ordered on d.stream like a cudaMemcpyAsync there.
// Pageable host memory (what Zig's page_allocator hands out) goes through the driver's own staging at ~12 GB/s, blocks the
// issuing thread, and the driver serialises such copies even when several threads issue them on several streams (measured,
// profiles/r02_host_path.log: 1, 2, 4, 8 threads all give 96.0-96.4 k gates/s end to end against 101.2 k from pinned
// memory).  Page-locking the caller's buffers for the duration of the call is worse still: cudaHostRegister of the 552 MB
// of a 65,536-gate call costs ~390 ms (63.5 k gates/s).  So large pageable copies are staged here: `host_copy_threads`
// host threads memcpy 32 MiB pieces between the caller's buffer and two pinned staging buffers while the DMA engine
// moves the previous piece.  Pinned or registered caller memory keeps the single asynchronous copy.
constexpr size_t kStagePiece = (size_t)32 << 20;

void parallel_memcpy(void *dst, const void *src, size_t bytes, int T) {
    if (T <= 1 || bytes < ((size_t)4 << 20)) { memcpy(dst, src, bytes); return; }
    std::vector<std::thread> workers;
    const size_t slice = ((bytes / T) + 4095) & ~(size_t)4095;
    for (int k = 1; k < T; k++) {
        const size_t lo = std::min(bytes, slice * k), hi = std::min(bytes, slice * (k + 1));
        if (lo < hi) workers.emplace_back([=] { memcpy((char *)dst + lo, (const char *)src + lo, hi - lo); });
    }
    memcpy(dst, src, std::min(bytes, slice));
    for (auto &w : workers) w.join();
}

bool is_pageable(const void *host) {
    cudaPointerAttributes at{};
    if (cudaPointerGetAttributes(&at, host) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeUnregistered;
}

int host_copy(tfhe_b200_ctx *c, Device &d, void *dst, const void *src, size_t bytes, cudaMemcpyKind kind, cudaStream_t stream = nullptr) {
    if (!stream) stream = d.stream;
    const void *host = kind == cudaMemcpyHostToDevice ? src : dst;
    bool staged = c->host_copy_threads > 0 && bytes >= kStagePiece;
    if (staged) staged = is_pageable(host);
    if (!staged) {
        CU(c, cudaMemcpyAsync(dst, src, bytes, kind, stream));
        return 0;
    }
    for (int k = 0; k < 2; k++) {
        if (!d.stage[k]) CU(c, cudaHostAlloc(&d.stage[k], kStagePiece, cudaHostAllocPortable));
        if (!d.stage_ev[k]) CU(c, cudaEventCreateWithFlags(&d.stage_ev[k], cudaEventDisableTiming));
    }
    const int T = c->host_copy_threads;
    const size_t pieces = (bytes + kStagePiece - 1) / kStagePiece;
    if (kind == cudaMemcpyHostToDevice) {
        for (size_t p = 0; p < pieces; p++) {
            const size_t off = p * kStagePiece, nb = std::min(kStagePiece, bytes - off);
            const int k = (int)(p & 1);
            if (p >= 2) CU(c, cudaEventSynchronize(d.stage_ev[k]));          // the DMA that last read this staging buffer
            parallel_memcpy(d.stage[k], (const char *)src + off, nb, T);
            CU(c, cudaMemcpyAsync((char *)dst + off, d.stage[k], nb, cudaMemcpyHostToDevice, stream));
            CU(c, cudaEventRecord(d.stage_ev[k], stream));
        }
        CU(c, cudaEventSynchronize(d.stage_ev[(pieces - 1) & 1]));              // staging buffers are free again for the next call
        if (pieces > 1) CU(c, cudaEventSynchronize(d.stage_ev[(pieces - 2) & 1]));
    } else {
        for (size_t p = 0; p < pieces + 1; p++) {                               // DMA of piece p overlaps the memcpy of piece p - 1
            if (p < pieces) {
                const size_t off = p * kStagePiece, nb = std::min(kStagePiece, bytes - off);
                CU(c, cudaMemcpyAsync(d.stage[p & 1], (const char *)src + off, nb, cudaMemcpyDeviceToHost, stream));
                CU(c, cudaEventRecord(d.stage_ev[p & 1], stream));
            }
            if (p >= 1) {
                const size_t q = p - 1, off = q * kStagePiece, nb = std::min(kStagePiece, bytes - off);
                CU(c, cudaEventSynchronize(d.stage_ev[q & 1]));
                parallel_memcpy((char *)dst + off, d.stage[q & 1], nb, T);
            }
        }
    }
    return 0;
}

// host-buffer driver: shard contiguously over devices; every device is driven by its own host thread (chunk, stage,
// launch, copy back), the stand-in for the reference's CPU thread pool (src/parallel/thread_pool.zig:39-83).  Copies
// from pageable host memory block the issuing thread, so one thread per device is what lets the H2D / D2H traffic of
// all devices (each on its own PCIe link) and their kernels proceed concurrently.
// Gate / bootstrap batches of one device from PAGEABLE host memory, pipelined: the shard is cut into four chunks of whole CTA
// waves; while the kernels of chunk k run on the device's stream, this thread stages the inputs of chunk k + 1 through pinned
// memory on a copy stream and then the outputs of chunk k - 1.  Two sets of batch buffers; events hand a chunk from the copy
// stream to the compute stream and back.  Pinned callers take the same path (plain asynchronous copies on the copy stream): the
// chunks are whole waves, so K1 pays no extra tail, and three quarters of the 10 ms of copies per 65,536 gates disappear behind
// the kernels: 102.5 k -> 103.6 k gates/s end to end (tuning key "host_pipeline": 1 = always, 2 = pageable callers only, 0 = never).
int run_host_device_pipelined(tfhe_b200_ctx *c, Device &d, size_t lo, size_t hi, int op, const int32_t *ops, const uint32_t *a, const uint32_t *b,
                              uint32_t *out, const uint32_t *tv) {
    const size_t w0 = (size_t)c->prm.n + 1;
    const bool two_inputs = (op >= 0 || ops);
    CU(c, cudaSetDevice(d.id));
    if (!d.copy_stream) CU(c, cudaStreamCreateWithFlags(&d.copy_stream, cudaStreamNonBlocking));
    for (int k = 0; k < 2; k++) {
        if (!d.ev_in[k]) CU(c, cudaEventCreateWithFlags(&d.ev_in[k], cudaEventDisableTiming));
        if (!d.ev_done[k]) CU(c, cudaEventCreateWithFlags(&d.ev_done[k], cudaEventDisableTiming));
    }
    const uint32_t *d_tv = nullptr;
    if (tv) {
        if (int r = ensure(c, d.tv, (size_t)2 * kN * 4)) return r;
        CU(c, cudaMemcpyAsync(d.tv.p, tv, (size_t)2 * kN * 4, cudaMemcpyHostToDevice, d.stream));
        d_tv = (const uint32_t *)d.tv.p;
    }
    const size_t wave = (size_t)6 * (d.sm_count > 0 ? d.sm_count : 148);
    size_t chunk = (((hi - lo + 3) / 4 + wave - 1) / wave) * wave;
    chunk = std::min(chunk, c->max_chunk);
    Buf *A[2] = {&d.a, &d.a2}, *Bb[2] = {&d.b, &d.b2}, *O[2] = {&d.out, &d.out2}, *Ops[2] = {&d.ops, &d.ops2};
    size_t prev_off = 0, prev_nb = 0;
    int k = 0;
    for (size_t off = lo; off < hi; off += chunk, k++) {
        const size_t nb = std::min(chunk, hi - off);
        const int p = k & 1;
        if (int r = ensure(c, *A[p], nb * w0 * 4)) return r;
        if (int r = ensure(c, *O[p], nb * w0 * 4)) return r;
        if (int r = host_copy(c, d, A[p]->p, a + off * w0, nb * w0 * 4, cudaMemcpyHostToDevice, d.copy_stream)) return r;
        if (two_inputs) {
            if (int r = ensure(c, *Bb[p], nb * w0 * 4)) return r;
            if (int r = host_copy(c, d, Bb[p]->p, b + off * w0, nb * w0 * 4, cudaMemcpyHostToDevice, d.copy_stream)) return r;
        }
        const int32_t *d_ops = nullptr;
        if (ops) {
            if (int r = ensure(c, *Ops[p], nb * 4)) return r;
            CU(c, cudaMemcpyAsync(Ops[p]->p, ops + off, nb * 4, cudaMemcpyHostToDevice, d.copy_stream));
            d_ops = (const int32_t *)Ops[p]->p;
        }
        CU(c, cudaEventRecord(d.ev_in[p], d.copy_stream));
        CU(c, cudaStreamWaitEvent(d.stream, d.ev_in[p], 0));
        if (int r = run_device(c, d, ops ? 0 : op, d_ops, (uint32_t *)A[p]->p, two_inputs ? (uint32_t *)Bb[p]->p : nullptr, (uint32_t *)O[p]->p, nullptr, nullptr,
                               nb, d_tv, 0))
            return r;
        CU(c, cudaEventRecord(d.ev_done[p], d.stream));
        if (k >= 1) {   // results of the previous chunk; its buffers are written again only by chunk k + 1, enqueued behind this copy
            CU(c, cudaStreamWaitEvent(d.copy_stream, d.ev_done[p ^ 1], 0));
            if (int r = host_copy(c, d, out + prev_off * w0, O[p ^ 1]->p, prev_nb * w0 * 4, cudaMemcpyDeviceToHost, d.copy_stream)) return r;
        }
        prev_off = off;
        prev_nb = nb;
    }
    CU(c, cudaStreamWaitEvent(d.copy_stream, d.ev_done[(k - 1) & 1], 0));
    if (int r = host_copy(c, d, out + prev_off * w0, O[(k - 1) & 1]->p, prev_nb * w0 * 4, cudaMemcpyDeviceToHost, d.copy_stream)) return r;
    CU(c, cudaStreamSynchronize(d.copy_stream));
    CU(c, cudaStreamSynchronize(d.stream));
    return 0;
}

int run_host_device(tfhe_b200_ctx *c, Device &d, size_t lo, size_t hi, int op, const int32_t *ops, const uint32_t *a, const uint32_t *b,
                    void *out, Out kind, const uint32_t *tv, int tv_per_item, int lut_m) {
    const size_t w0 = (size_t)c->prm.n + 1, w1 = (size_t)kN + 1, wt = (size_t)2 * kN;
    const size_t wout = (kind == Out::LV0 || kind == Out::LV0_NOKS) ? w0 : (kind == Out::LV1) ? w1 : wt;
    const bool two_inputs = (op >= 0 || ops);
    if (c->inject_fault > 0 && &d == &c->devs[(size_t)(c->inject_fault - 1) % c->devs.size()]) {
        c->inject_fault = 0;       // one shot: the context must be usable again afterwards
        return fail(c, TFHE_B200_ERR_CUDA, "injected fault on device %d (test hook)", d.id);
    }
    if (c->host_pipeline != 0 && kind == Out::LV0 && !(tv && (tv_per_item || lut_m > 0)) && hi - lo >= 32768 && !c->timing && (c->host_pipeline == 1 || is_pageable(a + lo * w0)))
        return run_host_device_pipelined(c, d, lo, hi, op, ops, a, b, (uint32_t *)out, tv);
    for (size_t off = lo; off < hi; off += c->max_chunk) {
        const size_t nb = std::min(c->max_chunk, hi - off);
        CU(c, cudaSetDevice(d.id));
        if (int r = ensure(c, d.a, nb * w0 * 4)) return r;
        if (int r = host_copy(c, d, d.a.p, a + off * w0, nb * w0 * 4, cudaMemcpyHostToDevice)) return r;
        if (two_inputs) {
            if (int r = ensure(c, d.b, nb * w0 * 4)) return r;
            if (int r = host_copy(c, d, d.b.p, b + off * w0, nb * w0 * 4, cudaMemcpyHostToDevice)) return r;
        }
        const int32_t *d_ops = nullptr;
        if (ops) {
            if (int r = ensure(c, d.ops, nb * 4)) return r;
            CU(c, cudaMemcpyAsync(d.ops.p, ops + off, nb * 4, cudaMemcpyHostToDevice, d.stream));
            d_ops = (const int32_t *)d.ops.p;
        }
        const uint32_t *d_tv = nullptr;
        if (tv && lut_m > 0) {   // function tables in, test vectors built on the device
            const size_t cnt = tv_per_item ? nb : 1;
            if (int r = ensure(c, d.tv, cnt * wt * 4)) return r;
            if (int r = ensure(c, d.lut, cnt * lut_m * 4)) return r;
            CU(c, cudaMemcpyAsync(d.lut.p, tv + (tv_per_item ? off * lut_m : 0), cnt * lut_m * 4, cudaMemcpyHostToDevice, d.stream));
            CU(c, launch_build_testvec((const uint32_t *)d.lut.p, lut_m, (uint32_t *)d.tv.p, cnt, d.stream, &d.launches));
            d_tv = (const uint32_t *)d.tv.p;
        } else if (tv) {
            const size_t tvb = (tv_per_item ? nb : 1) * wt * 4;
            if (int r = ensure(c, d.tv, tvb)) return r;
            CU(c, cudaMemcpyAsync(d.tv.p, tv + (tv_per_item ? off * wt : 0), tvb, cudaMemcpyHostToDevice, d.stream));
            d_tv = (const uint32_t *)d.tv.p;
        }
        if (int r = ensure(c, d.out, nb * wout * 4)) return r;
        uint32_t *d_out = (uint32_t *)d.out.p;
        int r = 0;
        if (kind == Out::LV0)
            r = run_device(c, d, ops ? 0 : op, d_ops, (uint32_t *)d.a.p, (uint32_t *)d.b.p, d_out, nullptr, nullptr, nb, d_tv, tv_per_item);
        else if (kind == Out::LV1)
            r = run_device(c, d, ops ? 0 : op, d_ops, (uint32_t *)d.a.p, (uint32_t *)d.b.p, nullptr, d_out, nullptr, nb, d_tv, tv_per_item);
        else if (kind == Out::TRLWE)
            r = run_device(c, d, ops ? 0 : op, d_ops, (uint32_t *)d.a.p, (uint32_t *)d.b.p, nullptr, nullptr, d_out, nb, d_tv, tv_per_item);
        else {  // LV0_NOKS: blind rotate + extract, then keep the first n mask entries + body
            if (int r2 = ensure(c, d.lv1, nb * w1 * 4)) return r2;
            r = run_device(c, d, ops ? 0 : op, d_ops, (uint32_t *)d.a.p, (uint32_t *)d.b.p, nullptr, (uint32_t *)d.lv1.p, nullptr, nb, d_tv,
                           tv_per_item);
            if (!r) CU(c, launch_extract2((uint32_t *)d.lv1.p, d_out, (uint32_t)nb, c->prm.n, d.stream, &d.launches));
        }
        if (r) return r;
        if (int r2 = host_copy(c, d, (uint32_t *)out + off * wout, d.out.p, nb * wout * 4, cudaMemcpyDeviceToHost)) return r2;
        CU(c, cudaStreamSynchronize(d.stream));
    }
    return 0;
}

int run_host(tfhe_b200_ctx *c, int op, const int32_t *ops, const uint32_t *a, const uint32_t *b, void *out, Out kind, size_t B,
             const uint32_t *tv, int tv_per_item, int lut_m = 0) {   // lut_m > 0: `tv` holds function tables [B or 1][lut_m]
    if (!c) return TFHE_B200_ERR_INVALID;
    if (!c->has_key) return fail(c, TFHE_B200_ERR_NO_KEY, "no cloud key loaded");
    if (B == 0) return 0;
    if (!a || !out || ((op >= 0 || ops) && !b)) return fail(c, TFHE_B200_ERR_INVALID, "null buffer");
    const int nd = (int)c->devs.size();
    if (nd == 1) return run_host_device(c, c->devs[0], 0, B, op, ops, a, b, out, kind, tv, tv_per_item, lut_m);
    std::vector<int> rc(nd, 0);
    std::vector<std::thread> workers;
    for (int k = 0; k < nd; k++) {
        const size_t lo = B * k / nd, hi = B * (k + 1) / nd;
        if (lo == hi) continue;
        workers.emplace_back([=, &rc] { rc[k] = run_host_device(c, c->devs[k], lo, hi, op, ops, a, b, out, kind, tv, tv_per_item, lut_m); });
    }
    for (auto &w : workers) w.join();
    for (int k = 0; k < nd; k++)
        if (rc[k]) return rc[k];
    return 0;
}

// tensor-core operand image of the packed key-switching key (d.ksk), on the sets the tensor-core kernel covers
int build_tc_key(tfhe_b200_ctx *c, Device &d) {
    const tfhe_b200_params &p = c->prm;
    if (d.ksk_tc) { CU(c, cudaFree(d.ksk_tc)); d.ksk_tc = nullptr; }
    if (!keyswitch_tc_supported(p.basebit, p.iks_t, kN, c->ksk_pitch)) return 0;
    CU(c, cudaMalloc(&d.ksk_tc, keyswitch_tc_key_bytes(c->ksk_pitch, p.basebit, p.iks_t)));
    CU(c, launch_ksk_to_tc(d.ksk, d.ksk_tc, p.basebit, p.iks_t, c->ksk_pitch, d.stream, &c->launches));
    CU(c, cudaStreamSynchronize(d.stream));
    return 0;
}

// exact chunk layout of the key on the device, from its reference layout (d.bsk_ref)
int build_exact_key(tfhe_b200_ctx *c, Device &d) {
    const tfhe_b200_params &p = c->prm;
    if (d.bsk_x) { CU(c, cudaFree(d.bsk_x)); d.bsk_x = nullptr; }
    CU(c, cudaMalloc(&d.bsk_x, (size_t)p.n * 2 * p.L * 2 * kN * 8));
    CU(c, launch_permute_bsk_exact(d.bsk_ref, d.bsk_x, p.n, p.L, d.stream, &c->launches));
    CU(c, cudaStreamSynchronize(d.stream));
    return 0;
}

// device allocation released on scope exit unless handed over (error paths of the key upload)
struct DevTmp {
    void *p = nullptr;
    ~DevTmp() { if (p) cudaFree(p); }
    void *release() { void *q = p; p = nullptr; return q; }
};

int upload_key_device(tfhe_b200_ctx *c, Device &d, const double *src_bsk, bool bsk_on_device, const uint32_t *src_ksk,
                      bool ksk_on_device, size_t ksk_stride_u32) {
    const tfhe_b200_params &p = c->prm;
    CU(c, cudaSetDevice(d.id));
    // the key already on the device stays usable until the new one is completely in place
    const size_t bsk_doubles = (size_t)p.n * 2 * p.L * 2 * kN;
    const double *d_ref = src_bsk;
    DevTmp staging, fast;
    if (!bsk_on_device) {
        CU(c, cudaMalloc(&staging.p, bsk_doubles * 8));
        CU(c, cudaMemcpyAsync(staging.p, src_bsk, bsk_doubles * 8, cudaMemcpyHostToDevice, d.stream));
        d_ref = (const double *)staging.p;
    }
    CU(c, cudaMalloc(&fast.p, bsk_doubles * 8));
    CU(c, launch_permute_bsk(d_ref, (cplx *)fast.p, p.n, p.L, d.stream, &c->launches));
    if (!staging.p) {   // caller's device buffer: keep our own copy of the reference layout for exact mode
        CU(c, cudaMalloc(&staging.p, bsk_doubles * 8));
        CU(c, cudaMemcpyAsync(staging.p, src_bsk, bsk_doubles * 8, cudaMemcpyDeviceToDevice, d.stream));
    }
    CU(c, cudaStreamSynchronize(d.stream));
    if (d.bsk) CU(c, cudaFree(d.bsk));
    d.bsk = (cplx *)fast.release();
    if (d.bsk_ref) CU(c, cudaFree(d.bsk_ref));
    d.bsk_ref = (double *)staging.release();
    if (int r = build_exact_key(c, d)) return r;
    d.has_key = true;

    if (!src_ksk) {
        d.has_ksk = false;
        if (d.ksk) { CU(c, cudaFree(d.ksk)); d.ksk = nullptr; }
        if (d.ksk_tc) { CU(c, cudaFree(d.ksk_tc)); d.ksk_tc = nullptr; }
    } else {
        const int base = 1 << p.basebit;
        const size_t rows = (size_t)kN * p.iks_t * base;
        const uint32_t *d_refk = src_ksk;
        DevTmp stg, packed;
        if (!ksk_on_device) {
            CU(c, cudaMalloc(&stg.p, rows * ksk_stride_u32 * 4));
            CU(c, cudaMemcpyAsync(stg.p, src_ksk, rows * ksk_stride_u32 * 4, cudaMemcpyHostToDevice, d.stream));
            d_refk = (const uint32_t *)stg.p;
        }
        CU(c, cudaMalloc(&packed.p, (size_t)kN * p.iks_t * (base - 1) * c->ksk_pitch * 4));
        CU(c, launch_repack_ksk(d_refk, ksk_stride_u32, (uint32_t *)packed.p, p.n, p.basebit, p.iks_t, c->ksk_pitch, kN, d.stream, &c->launches));
        CU(c, cudaStreamSynchronize(d.stream));
        if (d.ksk) CU(c, cudaFree(d.ksk));
        d.ksk = (uint32_t *)packed.release();
        if (int r = build_tc_key(c, d)) return r;
        d.has_ksk = true;
    }
    return 0;
}


}  // namespace

// ---------------------------------------------------------------------------------------------------
// Gate circuits: a netlist levelised once, every dependency level = one K1 + one K2 launch over
// (gates of the level) x (instances), all wires resident on the device, the level sequence captured in a CUDA graph.
struct tfhe_b200_circuit {
    tfhe_b200_ctx *ctx = nullptr;
    size_t n_inputs = 0, n_gates = 0, n_slots = 0, max_width = 0;
    bool has_consts = false;         // slots n_inputs (true) and n_inputs + 1 (false) hold Gates.constant wires
    struct Level { uint32_t first_slot, G; size_t off; };
    std::vector<Level> levels;
    std::vector<int32_t> ops;        // level-sorted gate tables (slot references, bit 31 = NOT)
    std::vector<uint32_t> wa, wb;
    std::vector<uint32_t> outputs;   // slot references, bit 31 = NOT
    // One entry per (device, lane).  Lanes are independent groups of instances running the same level sequence on their
    // own streams: the tail wave of one lane's level overlaps the head of another's (a level of the 16-bit adder over
    // 1,024 instances is only 1.7 or 3.5 waves of K1 CTAs), and one lane's key switch runs beside another's blind rotation.
    struct PerDev {
        int32_t *d_ops = nullptr;        // gate tables (owned by lane 0 of the device, shared by its other lanes)
        uint32_t *d_a = nullptr, *d_b = nullptr;
        cudaStream_t stream = nullptr;   // lane 0: the device's stream; others: owned
        bool owns_stream = false, owns_tables = false;
        Buf wires, lv1, neg, ksdig;
        cudaGraphExec_t graph = nullptr;
        size_t graph_inst = 0;
        uint64_t graph_epoch = 0;        // ctx->config_epoch at capture time
        uint64_t graph_launches = 0;
        bool graph_failed = false;
    };
    std::vector<PerDev> dev;
    int lanes = 1;
};

namespace {

int circuit_levels(tfhe_b200_ctx *c, tfhe_b200_circuit *q, Device &dev, tfhe_b200_circuit::PerDev &pd, size_t inst) {
    const size_t w0 = (size_t)c->prm.n + 1;
    uint32_t *wires = (uint32_t *)pd.wires.p;
    Device d = dev;              // same keys and tables, this lane's stream
    d.stream = pd.stream;
    int rc = 0;
    for (const auto &lv : q->levels) {
        LevelRef ref{pd.d_ops + lv.off, pd.d_a + lv.off, pd.d_b + lv.off, (uint32_t)inst};
        rc = run_device(c, d, 0, nullptr, wires, nullptr, wires + (size_t)lv.first_slot * inst * w0, (uint32_t *)pd.lv1.p, nullptr,
                        (size_t)lv.G * inst, nullptr, 0, &ref, q->lanes > 1, (uint64_t *)pd.ksdig.p);
        if (rc) break;
    }
    dev.launches = d.launches;   // the copy counted them
    return rc;
}

// enqueue every level of the circuit for `inst` instances whose input wires are already in pd.wires
int circuit_enqueue(tfhe_b200_ctx *c, tfhe_b200_circuit *q, Device &d, tfhe_b200_circuit::PerDev &pd, size_t inst) {
    if (!c->circuit_graph || c->timing || pd.graph_failed) return circuit_levels(c, q, d, pd, inst);
    if (!pd.graph || pd.graph_inst != inst || pd.graph_epoch != c->config_epoch) {
        if (pd.graph) { cudaGraphExecDestroy(pd.graph); pd.graph = nullptr; }
        const uint64_t before = d.launches;
        cudaGraph_t g = nullptr;
        bool ok = cudaStreamBeginCapture(pd.stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
        int r = ok ? circuit_levels(c, q, d, pd, inst) : 0;
        if (ok) ok = cudaStreamEndCapture(pd.stream, &g) == cudaSuccess && r == 0 && g != nullptr;
        if (ok) ok = cudaGraphInstantiate(&pd.graph, g, 0) == cudaSuccess;
        if (g) cudaGraphDestroy(g);
        pd.graph_launches = d.launches - before;
        d.launches = before;
        if (!ok) {               // capture refused: run the levels eagerly from now on
            cudaGetLastError();
            pd.graph = nullptr;
            pd.graph_failed = true;
            return circuit_levels(c, q, d, pd, inst);
        }
        pd.graph_inst = inst;
        pd.graph_epoch = c->config_epoch;
    }
    CU(c, cudaGraphLaunch(pd.graph, pd.stream));
    d.launches += pd.graph_launches;
    return 0;
}

}  // namespace

extern "C" {

const char *tfhe_b200_version(void) { return "tfhe_b200 0.1 (sm_100a)"; }

int tfhe_b200_create(const tfhe_b200_params *params, const int *device_ids, int n_dev, tfhe_b200_ctx **out) {
    if (!params || !out || n_dev < 1) return TFHE_B200_ERR_INVALID;
    *out = nullptr;
    const tfhe_b200_params &p = *params;
    if (!tfhe_b200_keyfile::params_supported(p)) return TFHE_B200_ERR_INVALID;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count < 1) return TFHE_B200_ERR_NO_DEVICE;
    auto *c = new tfhe_b200_ctx();
    c->prm = p;
    c->ksk_pitch = ((p.n + 1) + 3) & ~3;
    cplx tw2[kTw2Len], tw3[kTw3Len];
    make_twiddle_tables(tw2, tw3);
    std::vector<double> xt(6 * 512);
    make_exact_tables(xt.data());
    c->exact_conjugate = exact_tables_conjugate(xt.data());
    std::vector<cplx> xs(kExactSharedTabCplx + 8);
    make_exact_shared_tables(xt.data(), xs.data(), xs.data() + kExactSharedTabCplx);
    for (int k = 0; k < n_dev; k++) {
        Device d;
        d.id = device_ids ? device_ids[k] : k;
        cudaDeviceProp prop{};
        if (d.id < 0 || d.id >= count || cudaGetDeviceProperties(&prop, d.id) != cudaSuccess || prop.major != 10) {
            delete c;
            return TFHE_B200_ERR_NO_DEVICE;   // kernels are built for sm_100a only
        }
        d.sm_count = prop.multiProcessorCount;
        bool ok = cudaSetDevice(d.id) == cudaSuccess && cudaStreamCreateWithFlags(&d.stream, cudaStreamNonBlocking) == cudaSuccess &&
                  cudaMalloc(&d.tw2, sizeof(tw2)) == cudaSuccess && cudaMalloc(&d.tw3, sizeof(tw3)) == cudaSuccess &&
                  cudaMalloc(&d.margin_bits, 8) == cudaSuccess && cudaEventCreate(&d.ev[0]) == cudaSuccess &&
                  cudaEventCreate(&d.ev[1]) == cudaSuccess && cudaEventCreate(&d.ev[2]) == cudaSuccess &&
                  cudaMemcpy(d.tw2, tw2, sizeof(tw2), cudaMemcpyHostToDevice) == cudaSuccess &&
                  cudaMemcpy(d.tw3, tw3, sizeof(tw3), cudaMemcpyHostToDevice) == cudaSuccess &&
                  cudaMalloc(&d.exact_tables, xt.size() * 8) == cudaSuccess &&
                  cudaMemcpy(d.exact_tables, xt.data(), xt.size() * 8, cudaMemcpyHostToDevice) == cudaSuccess &&
                  cudaMalloc(&d.exact_shared, xs.size() * sizeof(cplx)) == cudaSuccess &&
                  cudaMemcpy(d.exact_shared, xs.data(), xs.size() * sizeof(cplx), cudaMemcpyHostToDevice) == cudaSuccess &&
                  cudaMemset(d.margin_bits, 0, 8) == cudaSuccess;
        c->devs.push_back(d);
        if (!ok) {
            tfhe_b200_destroy(c);
            return TFHE_B200_ERR_CUDA;
        }
    }
    *out = c;
    return TFHE_B200_OK;
}

void tfhe_b200_destroy(tfhe_b200_ctx *c) {
    if (!c) return;
    for (Device &d : c->devs) {
        cudaSetDevice(d.id);
        if (d.stream) cudaStreamSynchronize(d.stream);
        for (void *p : {(void *)d.bsk, (void *)d.ksk, (void *)d.ksk_tc, (void *)d.reenc, (void *)d.tw2, (void *)d.tw3, (void *)d.exact_tables, (void *)d.bsk_ref, (void *)d.bsk_x, (void *)d.exact_shared, (void *)d.margin_bits, d.a.p, d.b.p,
                        d.out.p, d.lv1.p, d.ops.p, d.tv.p, d.trlwe.p, d.lut.p, d.ksdig.p})
            if (p) cudaFree(p);
        for (cudaEvent_t e : d.ev) if (e) cudaEventDestroy(e);
        for (int k = 0; k < 2; k++) {
            if (d.stage[k]) cudaFreeHost(d.stage[k]);
            if (d.stage_ev[k]) cudaEventDestroy(d.stage_ev[k]);
            if (d.ev_in[k]) cudaEventDestroy(d.ev_in[k]);
            if (d.ev_done[k]) cudaEventDestroy(d.ev_done[k]);
        }
        for (void *p : {d.a2.p, d.b2.p, d.out2.p, d.ops2.p})
            if (p) cudaFree(p);
        if (d.copy_stream) cudaStreamDestroy(d.copy_stream);
        if (d.stream) cudaStreamDestroy(d.stream);
    }
    delete c;
}

const char *tfhe_b200_last_error(const tfhe_b200_ctx *c) { return c ? c->err.c_str() : "null context"; }
int tfhe_b200_num_devices(const tfhe_b200_ctx *c) { return c ? (int)c->devs.size() : 0; }

int tfhe_b200_load_key(tfhe_b200_ctx *c, const double *bsk, const uint32_t *ksk, size_t ksk_row_stride_bytes, uint32_t offset) {
    if (!c || !bsk) return fail(c, TFHE_B200_ERR_INVALID, "null key");
    size_t stride = (size_t)c->prm.n + 1;
    if (ksk) {
        if (ksk_row_stride_bytes == 0) ksk_row_stride_bytes = stride * 4;
        if (ksk_row_stride_bytes % 4 || ksk_row_stride_bytes < stride * 4) return fail(c, TFHE_B200_ERR_INVALID, "bad KSK row stride");
        stride = ksk_row_stride_bytes / 4;
    }
    for (Device &d : c->devs)
        if (int r = upload_key_device(c, d, bsk, false, ksk, false, stride)) {
            // some devices now hold the new key and some the previous one: a batch sharded over them would mix keys, so the
            // context has no key until a load succeeds on every device
            for (Device &e : c->devs) e.has_key = e.has_ksk = false;
            refresh_key_flags(c);
            return r;
        }
    c->offset = offset;
    refresh_key_flags(c);
    return 0;
}

int tfhe_b200_load_key_file(tfhe_b200_ctx *c, const char *path) {
    if (!c || !path) return fail(c, TFHE_B200_ERR_INVALID, "null argument");
    char msg[512];
    tfhe_b200_keyfile::View v;
    if (int r = tfhe_b200_keyfile::map(path, true, v, msg, sizeof msg)) return fail(c, r, "%s", msg);
    const tfhe_b200_params &f = v.header.params, &p = c->prm;
    int rc;
    if (f.n != p.n || f.N != p.N || f.L != p.L || f.bgbit != p.bgbit || f.basebit != p.basebit || f.iks_t != p.iks_t)
        rc = fail(c, TFHE_B200_ERR_INVALID, "%s: key is for n=%d L=%d bgbit=%d basebit=%d t=%d, the context for n=%d L=%d bgbit=%d basebit=%d t=%d",
                  path, f.n, f.L, f.bgbit, f.basebit, f.iks_t, p.n, p.L, p.bgbit, p.basebit, p.iks_t);
    else
        rc = tfhe_b200_load_key(c, v.bsk, v.ksk, 0, v.header.decomposition_offset);
    tfhe_b200_keyfile::unmap(v);
    return rc;
}

int tfhe_b200_load_key_device(tfhe_b200_ctx *c, int dev, const double *d_bsk, const uint32_t *d_ksk, uint32_t offset) {
    if (!c || !d_bsk || dev < 0 || dev >= (int)c->devs.size()) return fail(c, TFHE_B200_ERR_INVALID, "bad argument");
    const int r = upload_key_device(c, c->devs[dev], d_bsk, true, d_ksk, true, (size_t)c->prm.n + 1);
    if (!r) c->offset = offset;
    refresh_key_flags(c);
    return r;
}


int tfhe_b200_keygen(tfhe_b200_ctx *c, const uint32_t *key_lv0, const uint32_t *key_lv1, uint64_t seed, double ksk_alpha, double bsk_alpha,
                     double *bsk_out, uint32_t *ksk_out) {
    if (!c || !key_lv0 || !key_lv1 || !(ksk_alpha >= 0.0) || !(bsk_alpha >= 0.0)) return fail(c, TFHE_B200_ERR_INVALID, "bad keygen argument");
    const tfhe_b200_params &p = c->prm;
    const int base = 1 << p.basebit;
    const size_t bsk_doubles = (size_t)p.n * 2 * p.L * 2 * kN;
    const size_t ksk_ref_words = (size_t)kN * p.iks_t * base * (p.n + 1);
    for (size_t k = 0; k < c->devs.size(); k++) {
        Device &d = c->devs[k];
        CU(c, cudaSetDevice(d.id));
        uint32_t *d_s0 = nullptr, *d_s1 = nullptr, *d_kref = nullptr;
        CU(c, cudaMalloc(&d_s0, (size_t)p.n * 4));
        CU(c, cudaMalloc(&d_s1, (size_t)kN * 4));
        CU(c, cudaMemcpyAsync(d_s0, key_lv0, (size_t)p.n * 4, cudaMemcpyHostToDevice, d.stream));
        CU(c, cudaMemcpyAsync(d_s1, key_lv1, (size_t)kN * 4, cudaMemcpyHostToDevice, d.stream));
        d.has_key = d.has_ksk = false;
        if (d.bsk) CU(c, cudaFree(d.bsk));
        if (d.bsk_ref) CU(c, cudaFree(d.bsk_ref));
        if (d.ksk) CU(c, cudaFree(d.ksk));
        d.bsk = nullptr; d.bsk_ref = nullptr; d.ksk = nullptr;
        CU(c, cudaMalloc(&d.bsk, bsk_doubles * 8));
        CU(c, cudaMalloc(&d.bsk_ref, bsk_doubles * 8));
        CU(c, cudaMalloc(&d.ksk, (size_t)kN * p.iks_t * (base - 1) * c->ksk_pitch * 4));
        if (k == 0 && ksk_out) {     // reference layout incl. the never-read k = 0 rows (zeroed here; uninitialised upstream)
            CU(c, cudaMalloc(&d_kref, ksk_ref_words * 4));
            CU(c, cudaMemsetAsync(d_kref, 0, ksk_ref_words * 4, d.stream));
        }
        CU(c, launch_keygen_bsk(d_s0, d_s1, seed, bsk_alpha, p.n, p.L, p.bgbit, d.tw2, d.tw3, d.bsk, d.bsk_ref, d.stream, &c->launches));
        CU(c, launch_keygen_ksk(d_s0, d_s1, seed, ksk_alpha, p.n, p.basebit, p.iks_t, c->ksk_pitch, d.ksk, d_kref, d.stream, &c->launches));
        if (int r = build_exact_key(c, d)) return r;
        if (int r = build_tc_key(c, d)) return r;
        if (k == 0 && bsk_out) CU(c, cudaMemcpyAsync(bsk_out, d.bsk_ref, bsk_doubles * 8, cudaMemcpyDeviceToHost, d.stream));
        if (d_kref) CU(c, cudaMemcpyAsync(ksk_out, d_kref, ksk_ref_words * 4, cudaMemcpyDeviceToHost, d.stream));
        CU(c, cudaStreamSynchronize(d.stream));
        CU(c, cudaFree(d_s0));
        CU(c, cudaFree(d_s1));
        if (d_kref) CU(c, cudaFree(d_kref));
    }
    uint32_t offset = 0;             // key.genDecompositionOffset (src/key.zig:121-131)
    for (int i = 0; i < p.L; i++) offset += (1u << (p.bgbit - 1)) << (32 - (i + 1) * p.bgbit);
    c->offset = offset;
    for (Device &d : c->devs) d.has_key = d.has_ksk = true;
    refresh_key_flags(c);
    return 0;
}

uint32_t tfhe_b200_decomposition_offset(const tfhe_b200_ctx *c) { return c ? c->offset : 0u; }

int tfhe_b200_set_mode(tfhe_b200_ctx *c, int mode) {
    if (!c || (mode != TFHE_B200_MODE_FAST && mode != TFHE_B200_MODE_EXACT)) return fail(c, TFHE_B200_ERR_INVALID, "bad mode");
    c->mode = mode;
    c->config_epoch++;
    return 0;
}

int tfhe_b200_gate_batch(tfhe_b200_ctx *c, int op, const uint32_t *a, const uint32_t *b, uint32_t *out, size_t B) {
    if (!c || op < 0 || op > 9) return fail(c, TFHE_B200_ERR_INVALID, "bad gate opcode %d", op);
    return run_host(c, op, nullptr, a, b, out, Out::LV0, B, nullptr, 0);
}

int tfhe_b200_gate_batch_ops(tfhe_b200_ctx *c, const int32_t *ops, const uint32_t *a, const uint32_t *b, uint32_t *out, size_t B) {
    if (!c || !ops) return fail(c, TFHE_B200_ERR_INVALID, "null opcode array");
    for (size_t i = 0; i < B; i++)
        if (ops[i] < 0 || ops[i] > 9) return fail(c, TFHE_B200_ERR_INVALID, "bad gate opcode %d at %zu", ops[i], i);
    return run_host(c, 0, ops, a, b, out, Out::LV0, B, nullptr, 0);
}

int tfhe_b200_bootstrap_batch(tfhe_b200_ctx *c, const uint32_t *in, uint32_t *out, size_t B, const uint32_t *tv, int tv_per_item) {
    return run_host(c, -1, nullptr, in, nullptr, out, Out::LV0, B, tv, tv_per_item);
}

int tfhe_b200_lut_bootstrap_batch(tfhe_b200_ctx *c, const uint32_t *in, uint32_t *out, size_t B, const uint32_t *tables, int message_modulus,
                                  int per_item) {
    if (!c || !tables || message_modulus < 1 || message_modulus > kN) return fail(c, TFHE_B200_ERR_INVALID, "bad lookup table");
    return run_host(c, -1, nullptr, in, nullptr, out, Out::LV0, B, tables, per_item, message_modulus);
}

// One device's share of a many-function bootstrap: K1 with the coarse modulus switch and the interleaved test vector, then
// per function sampleExtractIndex(., f) -> key switch.
static int many_lut_device(tfhe_b200_ctx *c, Device &d, size_t lo, size_t hi, size_t B, const uint32_t *in, uint32_t *out, const uint32_t *tv,
                           int k, int shift) {
    const size_t w0 = (size_t)c->prm.n + 1, w1 = (size_t)kN + 1, wt = (size_t)2 * kN;
    const size_t chunk = std::min<size_t>(c->max_chunk, (size_t)1 << 15);     // 8 KiB of accumulator per ciphertext stays on the device
    CU(c, cudaSetDevice(d.id));
    if (int r = ensure(c, d.tv, wt * 4)) return r;
    CU(c, cudaMemcpyAsync(d.tv.p, tv, wt * 4, cudaMemcpyHostToDevice, d.stream));
    for (size_t off = lo; off < hi; off += chunk) {
        const size_t nb = std::min(chunk, hi - off);
        if (int r = ensure(c, d.a, nb * w0 * 4)) return r;
        if (int r = ensure(c, d.trlwe, nb * wt * 4)) return r;
        if (int r = ensure(c, d.lv1, nb * w1 * 4)) return r;
        if (int r = ensure(c, d.out, nb * w0 * 4)) return r;
        if (int r = host_copy(c, d, d.a.p, in + off * w0, nb * w0 * 4, cudaMemcpyHostToDevice)) return r;
        if (int r = run_device(c, d, -1, nullptr, (uint32_t *)d.a.p, nullptr, nullptr, nullptr, (uint32_t *)d.trlwe.p, nb, (const uint32_t *)d.tv.p, 0,
                               nullptr, false, nullptr, shift))
            return r;
        for (int f = 0; f < k; f++) {
            CU(c, launch_sample_extract((const uint32_t *)d.trlwe.p, (uint32_t *)d.lv1.p, (uint32_t)nb, f, d.stream, &d.launches));
            if (int r = run_keyswitch(c, d, (const uint32_t *)d.lv1.p, (uint32_t *)d.out.p, nb, nullptr, &d.launches)) return r;
            if (int r = host_copy(c, d, out + ((size_t)f * B + off) * w0, d.out.p, nb * w0 * 4, cudaMemcpyDeviceToHost)) return r;
        }
        CU(c, cudaStreamSynchronize(d.stream));
    }
    return 0;
}

int tfhe_b200_lut_bootstrap_many_batch(tfhe_b200_ctx *c, const uint32_t *in, uint32_t *out, size_t B, const uint32_t *tables, int n_functions,
                                       int message_modulus) {
    const int k = n_functions, m = message_modulus;
    if (!c || !in || !out || !tables || m < 1 || k < 1 || (k & (k - 1)) != 0 || (long long)k * 2 * m > kN)
        return fail(c, TFHE_B200_ERR_INVALID, "many-function bootstrap: n_functions must be a power of two with n_functions * 2 * message_modulus <= %d", kN);
    if (!c->has_key || !c->has_ksk) return fail(c, TFHE_B200_ERR_NO_KEY, "no cloud key loaded");
    if (B == 0) return 0;
    int shift = 0;
    while ((1 << shift) < k) shift++;
    // test vector: position k i + f holds function f's lookup table (lut.Generator.generateLookupTableFull, src/lut/generator.zig:150-191)
    // at position k i -- a rotation by a multiple of k then leaves function f at every index congruent to f
    std::vector<uint32_t> tv(2 * kN, 0u), raw(kN), rot(kN);
    const int offset = (kN + m) / (2 * m);                                 // divRound(N, 2m), generator.zig:253-255
    for (int f = 0; f < k; f++) {
        for (int x = 0; x < m; x++) {
            const int start = (int)(((long long)x * kN + m / 2) / m), end = (int)(((long long)(x + 1) * kN + m / 2) / m);
            for (int i = start; i < end && i < kN; i++) raw[i] = tables[(size_t)f * m + x];
        }
        for (int i = 0; i < kN; i++) rot[i] = raw[(i + offset) % kN];
        for (int i = kN - offset; i < kN; i++) rot[i] = 0u - rot[i];
        for (int i = 0; i < kN; i += k) tv[kN + i + f] = rot[i];
    }
    const int nd = (int)c->devs.size();
    if (nd == 1) return many_lut_device(c, c->devs[0], 0, B, B, in, out, tv.data(), k, shift);
    std::vector<int> rc(nd, 0);
    std::vector<std::thread> workers;
    for (int j = 0; j < nd; j++) {
        const size_t lo = B * j / nd, hi = B * (j + 1) / nd;
        if (lo == hi) continue;
        workers.emplace_back([=, &rc, &tv] { rc[j] = many_lut_device(c, c->devs[j], lo, hi, B, in, out, tv.data(), k, shift); });
    }
    for (auto &w : workers) w.join();
    for (int j = 0; j < nd; j++)
        if (rc[j]) return rc[j];
    return 0;
}

int tfhe_b200_lut_generate(tfhe_b200_ctx *c, const uint32_t *table, int message_modulus, uint32_t *testvec_out) {
    if (!c || !table || !testvec_out || message_modulus < 1 || message_modulus > kN) return fail(c, TFHE_B200_ERR_INVALID, "bad lookup table");
    Device &d = c->devs[0];
    CU(c, cudaSetDevice(d.id));
    if (int r = ensure(c, d.tv, (size_t)2 * kN * 4)) return r;
    if (int r = ensure(c, d.lut, (size_t)message_modulus * 4)) return r;
    CU(c, cudaMemcpyAsync(d.lut.p, table, (size_t)message_modulus * 4, cudaMemcpyHostToDevice, d.stream));
    CU(c, launch_build_testvec((const uint32_t *)d.lut.p, message_modulus, (uint32_t *)d.tv.p, 1, d.stream, &c->launches));
    CU(c, cudaMemcpyAsync(testvec_out, d.tv.p, (size_t)2 * kN * 4, cudaMemcpyDeviceToHost, d.stream));
    CU(c, cudaStreamSynchronize(d.stream));
    return 0;
}

int tfhe_b200_bootstrap_no_keyswitch_batch(tfhe_b200_ctx *c, const uint32_t *in, uint32_t *out, size_t B) {
    return run_host(c, -1, nullptr, in, nullptr, out, Out::LV0_NOKS, B, nullptr, 0);
}

int tfhe_b200_blind_rotate_batch(tfhe_b200_ctx *c, const uint32_t *in, uint32_t *trlwe_out, size_t B, const uint32_t *tv, int tv_per_item) {
    return run_host(c, -1, nullptr, in, nullptr, trlwe_out, Out::TRLWE, B, tv, tv_per_item);
}

int tfhe_b200_blind_rotate_extract_batch(tfhe_b200_ctx *c, const uint32_t *in, uint32_t *lv1_out, size_t B) {
    return run_host(c, -1, nullptr, in, nullptr, lv1_out, Out::LV1, B, nullptr, 0);
}

int tfhe_b200_keyswitch_batch(tfhe_b200_ctx *c, const uint32_t *lv1, uint32_t *lv0, size_t B) {
    if (!c || !lv1 || !lv0) return fail(c, TFHE_B200_ERR_INVALID, "null buffer");
    if (!c->has_ksk) return fail(c, TFHE_B200_ERR_NO_KEY, "no key-switching key loaded");
    const size_t w0 = (size_t)c->prm.n + 1, w1 = (size_t)kN + 1;
    const int nd = (int)c->devs.size();
    for (int k = 0; k < nd; k++) {
        Device &d = c->devs[k];
        const size_t lo = B * k / nd, hi = B * (k + 1) / nd;
        for (size_t off = lo; off < hi; off += c->max_chunk) {
            const size_t nb = std::min(c->max_chunk, hi - off);
            CU(c, cudaSetDevice(d.id));
            if (int r = ensure(c, d.lv1, nb * w1 * 4)) return r;
            if (int r = ensure(c, d.out, nb * w0 * 4)) return r;
            CU(c, cudaMemcpyAsync(d.lv1.p, lv1 + off * w1, nb * w1 * 4, cudaMemcpyHostToDevice, d.stream));
            if (int r = run_keyswitch(c, d, (uint32_t *)d.lv1.p, (uint32_t *)d.out.p, nb, nullptr, &c->launches)) return r;
            CU(c, cudaMemcpyAsync(lv0 + off * w0, d.out.p, nb * w0 * 4, cudaMemcpyDeviceToHost, d.stream));
            CU(c, cudaStreamSynchronize(d.stream));
        }
    }
    return 0;
}

int tfhe_b200_load_reencryption_key(tfhe_b200_ctx *c, const uint32_t *key, int basebit, int t) {
    if (!c || !key || basebit < 1 || basebit > 8 || t < 1 || 1 + basebit * t > 32) return fail(c, TFHE_B200_ERR_INVALID, "bad re-encryption key");
    const int n = c->prm.n, base = 1 << basebit;
    const size_t rows = (size_t)n * t * base, w = (size_t)n + 1;
    for (Device &d : c->devs) {
        CU(c, cudaSetDevice(d.id));
        uint32_t *stg = nullptr;
        CU(c, cudaMalloc(&stg, rows * w * 4));
        CU(c, cudaMemcpyAsync(stg, key, rows * w * 4, cudaMemcpyHostToDevice, d.stream));
        if (d.reenc) CU(c, cudaFree(d.reenc));
        CU(c, cudaMalloc(&d.reenc, (size_t)n * t * (base - 1) * c->ksk_pitch * 4));
        CU(c, launch_repack_ksk(stg, w, d.reenc, n, basebit, t, c->ksk_pitch, n, d.stream, &c->launches));
        CU(c, cudaStreamSynchronize(d.stream));
        CU(c, cudaFree(stg));
    }
    c->has_reenc = true;
    c->reenc_basebit = basebit;
    c->reenc_t = t;
    return 0;
}

int tfhe_b200_reencrypt_batch(tfhe_b200_ctx *c, const uint32_t *in, uint32_t *out, size_t B) {
    if (!c || !in || !out) return fail(c, TFHE_B200_ERR_INVALID, "null buffer");
    if (!c->has_reenc) return fail(c, TFHE_B200_ERR_NO_KEY, "no proxy re-encryption key loaded");
    const size_t w = (size_t)c->prm.n + 1;
    const int nd = (int)c->devs.size();
    for (int k = 0; k < nd; k++) {
        Device &d = c->devs[k];
        const size_t lo = B * k / nd, hi = B * (k + 1) / nd;
        for (size_t off = lo; off < hi; off += c->max_chunk) {
            const size_t nb = std::min(c->max_chunk, hi - off);
            CU(c, cudaSetDevice(d.id));
            if (int r = ensure(c, d.a, nb * w * 4)) return r;
            if (int r = ensure(c, d.out, nb * w * 4)) return r;
            CU(c, cudaMemcpyAsync(d.a.p, in + off * w, nb * w * 4, cudaMemcpyHostToDevice, d.stream));
            KsArgs K{(uint32_t *)d.a.p, (uint32_t *)d.out.p, d.reenc, (uint32_t)nb, c->prm.n, c->reenc_basebit, c->reenc_t, c->ksk_pitch, c->prm.n, c->ks_tile, c->ks_vec};
            K.fill = c->ks_fill; K.rot = c->ks_rot;
            CU(c, launch_keyswitch(K, d.sm_count, d.stream, &c->launches));
            CU(c, cudaMemcpyAsync(out + off * w, d.out.p, nb * w * 4, cudaMemcpyDeviceToHost, d.stream));
            CU(c, cudaStreamSynchronize(d.stream));
        }
    }
    return 0;
}

int tfhe_b200_not_batch(tfhe_b200_ctx *c, const uint32_t *a, uint32_t *out, size_t B) {
    if (!c || !a || !out) return fail(c, TFHE_B200_ERR_INVALID, "null buffer");
    if (B == 0) return 0;
    Device &d = c->devs[0];
    const size_t bytes = B * ((size_t)c->prm.n + 1) * 4;
    CU(c, cudaSetDevice(d.id));
    if (int r = ensure(c, d.a, bytes)) return r;
    if (int r = ensure(c, d.out, bytes)) return r;
    CU(c, cudaMemcpyAsync(d.a.p, a, bytes, cudaMemcpyHostToDevice, d.stream));
    CU(c, launch_negate((uint32_t *)d.a.p, (uint32_t *)d.out.p, bytes / 4, d.stream, &c->launches));
    CU(c, cudaMemcpyAsync(out, d.out.p, bytes, cudaMemcpyDeviceToHost, d.stream));
    CU(c, cudaStreamSynchronize(d.stream));
    return 0;
}

int tfhe_b200_gate_batch_device(tfhe_b200_ctx *c, int dev, int op, const int32_t *d_ops, const uint32_t *d_a, const uint32_t *d_b,
                                uint32_t *d_out, size_t B) {
    if (!c || dev < 0 || dev >= (int)c->devs.size() || !d_a || !d_b || !d_out || (!d_ops && (op < 0 || op > 9)))
        return fail(c, TFHE_B200_ERR_INVALID, "bad argument");
    return run_device(c, c->devs[dev], op, d_ops, d_a, d_b, d_out, nullptr, nullptr, B, nullptr, 0);
}

int tfhe_b200_bootstrap_batch_device(tfhe_b200_ctx *c, int dev, const uint32_t *d_in, uint32_t *d_out, size_t B, const uint32_t *d_tv,
                                     int tv_per_item) {
    if (!c || dev < 0 || dev >= (int)c->devs.size() || !d_in || !d_out) return fail(c, TFHE_B200_ERR_INVALID, "bad argument");
    return run_device(c, c->devs[dev], -1, nullptr, d_in, nullptr, d_out, nullptr, nullptr, B, d_tv, tv_per_item);
}

int tfhe_b200_blind_rotate_batch_device(tfhe_b200_ctx *c, int dev, const uint32_t *d_in, uint32_t *d_trlwe, size_t B, const uint32_t *d_tv,
                                        int tv_per_item) {
    if (!c || dev < 0 || dev >= (int)c->devs.size() || !d_in || !d_trlwe) return fail(c, TFHE_B200_ERR_INVALID, "bad argument");
    return run_device(c, c->devs[dev], -1, nullptr, d_in, nullptr, nullptr, nullptr, d_trlwe, B, d_tv, tv_per_item);
}

int tfhe_b200_keyswitch_batch_device(tfhe_b200_ctx *c, int dev, const uint32_t *d_lv1, uint32_t *d_lv0, size_t B) {
    if (!c || dev < 0 || dev >= (int)c->devs.size() || !d_lv1 || !d_lv0) return fail(c, TFHE_B200_ERR_INVALID, "bad argument");
    if (!c->has_ksk) return fail(c, TFHE_B200_ERR_NO_KEY, "no key-switching key loaded");
    Device &d = c->devs[dev];
    CU(c, cudaSetDevice(d.id));
    return run_keyswitch(c, d, d_lv1, d_lv0, B, nullptr, &c->launches);
}


// host-only part of circuit compilation (no device needed): validation + dependency level of every gate
static int circuit_plan(const tfhe_b200_gate_node *gates, size_t n_gates, size_t n_inputs, const uint32_t *outputs, size_t n_outputs,
                        std::vector<uint32_t> &level, uint32_t &depth, std::string &why) {
    char buf[160];
    const size_t n_wires = n_inputs + n_gates;
    if ((!gates && n_gates) || (!outputs && n_outputs)) { why = "null argument"; return TFHE_B200_ERR_INVALID; }
    if (n_wires == 0 || n_wires >= TFHE_B200_WIRE_FALSE) { why = "bad circuit size"; return TFHE_B200_ERR_INVALID; }
    const uint32_t kNot = TFHE_B200_WIRE_NOT;
    level.assign(n_wires, 0);
    depth = 0;
    auto is_const = [](uint32_t w) { return w == TFHE_B200_WIRE_TRUE || w == TFHE_B200_WIRE_FALSE; };
    for (size_t g = 0; g < n_gates; g++) {
        const uint32_t a = gates[g].a & ~kNot, b = gates[g].b & ~kNot;
        if (gates[g].op < 0 || gates[g].op > 9) {
            snprintf(buf, sizeof(buf), "gate %zu: bad opcode %d", g, gates[g].op);
            why = buf;
            return TFHE_B200_ERR_INVALID;
        }
        if ((!is_const(a) && a >= n_inputs + g) || (!is_const(b) && b >= n_inputs + g)) {
            snprintf(buf, sizeof(buf), "gate %zu reads a wire defined later (not topological)", g);
            why = buf;
            return TFHE_B200_ERR_INVALID;
        }
        // constants sit at level 0 like the inputs (Gates.constant needs no bootstrap, src/gates.zig:144-151)
        level[n_inputs + g] = 1 + std::max(is_const(a) ? 0u : level[a], is_const(b) ? 0u : level[b]);
        depth = std::max(depth, level[n_inputs + g]);
    }
    for (size_t o = 0; o < n_outputs; o++)
        if ((outputs[o] & ~kNot) >= n_wires && !is_const(outputs[o] & ~kNot)) {
            snprintf(buf, sizeof(buf), "output %zu: no such wire", o);
            why = buf;
            return TFHE_B200_ERR_INVALID;
        }
    return 0;
}

int tfhe_b200_circuit_plan(const tfhe_b200_gate_node *gates, size_t n_gates, size_t n_inputs, const uint32_t *outputs, size_t n_outputs,
                           size_t *n_levels, size_t *max_level_width, uint32_t *gate_level) {
    std::vector<uint32_t> level;
    uint32_t depth = 0;
    std::string why;
    if (int r = circuit_plan(gates, n_gates, n_inputs, outputs, n_outputs, level, depth, why)) return r;
    std::vector<size_t> width(depth + 1, 0);
    for (size_t g = 0; g < n_gates; g++) {
        width[level[n_inputs + g]]++;
        if (gate_level) gate_level[g] = level[n_inputs + g];
    }
    if (n_levels) *n_levels = depth;
    if (max_level_width) *max_level_width = depth ? *std::max_element(width.begin() + 1, width.end()) : 0;
    return 0;
}

int tfhe_b200_circuit_create(tfhe_b200_ctx *c, const tfhe_b200_gate_node *gates, size_t n_gates, size_t n_inputs,
                             const uint32_t *outputs, size_t n_outputs, tfhe_b200_circuit **out) {
    if (!c || !out) return fail(c, TFHE_B200_ERR_INVALID, "null argument");
    *out = nullptr;
    const size_t n_wires = n_inputs + n_gates;
    const uint32_t kNot = TFHE_B200_WIRE_NOT;
    std::vector<uint32_t> level;
    uint32_t depth = 0;
    {
        std::string why;
        if (int r = circuit_plan(gates, n_gates, n_inputs, outputs, n_outputs, level, depth, why)) return fail(c, r, "%s", why.c_str());
    }
    auto *q = new tfhe_b200_circuit();
    q->ctx = c;
    auto is_const = [](uint32_t w) { return (w & ~TFHE_B200_WIRE_NOT) == TFHE_B200_WIRE_TRUE || (w & ~TFHE_B200_WIRE_NOT) == TFHE_B200_WIRE_FALSE; };
    for (size_t g = 0; g < n_gates; g++) q->has_consts = q->has_consts || is_const(gates[g].a) || is_const(gates[g].b);
    for (size_t o = 0; o < n_outputs; o++) q->has_consts = q->has_consts || is_const(outputs[o]);
    const uint32_t n_const = q->has_consts ? 2u : 0u;
    q->n_inputs = n_inputs; q->n_gates = n_gates; q->n_slots = n_wires + n_const;
    // storage slots: inputs first, then the two Gates.constant wires (if any gate or output uses one), then the gates level
    // by level, so that a level's outputs are one contiguous block
    std::vector<uint32_t> slot(n_wires);
    for (size_t i = 0; i < n_inputs; i++) slot[i] = (uint32_t)i;
    auto slot_of = [&](uint32_t ref) -> uint32_t {
        const uint32_t w = ref & ~kNot;
        const uint32_t sl = w == TFHE_B200_WIRE_TRUE ? (uint32_t)n_inputs : w == TFHE_B200_WIRE_FALSE ? (uint32_t)n_inputs + 1u : slot[w];
        return sl | (ref & kNot);
    };
    std::vector<std::vector<uint32_t>> by_level(depth + 1);
    for (size_t g = 0; g < n_gates; g++) by_level[level[n_inputs + g]].push_back((uint32_t)g);
    uint32_t next = (uint32_t)n_inputs + n_const;
    for (uint32_t l = 1; l <= depth; l++) {
        q->levels.push_back({next, (uint32_t)by_level[l].size(), q->ops.size()});
        q->max_width = std::max(q->max_width, by_level[l].size());
        for (uint32_t g : by_level[l]) {
            slot[n_inputs + g] = next++;
            q->ops.push_back(gates[g].op);
            q->wa.push_back(gates[g].a);     // user wire ids for now
            q->wb.push_back(gates[g].b);
        }
    }
    for (size_t k = 0; k < q->wa.size(); k++) {
        q->wa[k] = slot_of(q->wa[k]);
        q->wb[k] = slot_of(q->wb[k]);
    }
    for (size_t o = 0; o < n_outputs; o++) q->outputs.push_back(slot_of(outputs[o]));
    q->lanes = std::max(1, std::min(c->circuit_lanes, 8));
    q->dev.resize(c->devs.size() * q->lanes);
    for (size_t k = 0; k < c->devs.size(); k++) {
        auto &pd = q->dev[k * q->lanes];
        const size_t nt = std::max<size_t>(q->ops.size(), 1);
        pd.owns_tables = true;
        pd.stream = c->devs[k].stream;
        bool ok = cudaSetDevice(c->devs[k].id) == cudaSuccess && cudaMalloc(&pd.d_ops, nt * 4) == cudaSuccess &&
                  cudaMalloc(&pd.d_a, nt * 4) == cudaSuccess && cudaMalloc(&pd.d_b, nt * 4) == cudaSuccess;
        if (ok && !q->ops.empty())
            ok = cudaMemcpy(pd.d_ops, q->ops.data(), q->ops.size() * 4, cudaMemcpyHostToDevice) == cudaSuccess &&
                 cudaMemcpy(pd.d_a, q->wa.data(), q->wa.size() * 4, cudaMemcpyHostToDevice) == cudaSuccess &&
                 cudaMemcpy(pd.d_b, q->wb.data(), q->wb.size() * 4, cudaMemcpyHostToDevice) == cudaSuccess;
        for (int l = 1; ok && l < q->lanes; l++) {
            auto &pl = q->dev[k * q->lanes + l];
            pl.d_ops = pd.d_ops; pl.d_a = pd.d_a; pl.d_b = pd.d_b;
            ok = cudaStreamCreateWithFlags(&pl.stream, cudaStreamNonBlocking) == cudaSuccess;
            pl.owns_stream = ok;
        }
        if (!ok) {
            tfhe_b200_circuit_destroy(q);
            return fail(c, TFHE_B200_ERR_CUDA, "circuit tables: %s", cudaGetErrorString(cudaGetLastError()));
        }
    }
    *out = q;
    return 0;
}

void tfhe_b200_circuit_destroy(tfhe_b200_circuit *q) {
    if (!q) return;
    for (size_t v = 0; v < q->dev.size(); v++) {
        auto &pd = q->dev[v];
        cudaSetDevice(q->ctx->devs[v / q->lanes].id);
        if (pd.stream) cudaStreamSynchronize(pd.stream);
        if (pd.graph) cudaGraphExecDestroy(pd.graph);
        if (pd.owns_tables)
            for (void *p : {(void *)pd.d_ops, (void *)pd.d_a, (void *)pd.d_b})
                if (p) cudaFree(p);
        for (void *p : {pd.wires.p, pd.lv1.p, pd.neg.p, pd.ksdig.p})
            if (p) cudaFree(p);
        if (pd.owns_stream) cudaStreamDestroy(pd.stream);
    }
    delete q;
}

int tfhe_b200_circuit_info(const tfhe_b200_circuit *q, size_t *n_levels, size_t *max_level_width, size_t *n_gates) {
    if (!q) return TFHE_B200_ERR_INVALID;
    if (n_levels) *n_levels = q->levels.size();
    if (max_level_width) *max_level_width = q->max_width;
    if (n_gates) *n_gates = q->n_gates;
    return 0;
}

// all lanes of ONE device: passes of (inputs + every level in flight on every lane) -> (results back) -> (lanes drained)
static int circuit_run_device(tfhe_b200_ctx *c, tfhe_b200_circuit *q, int k, const uint32_t *inputs, uint32_t *outputs, size_t instances,
                              size_t dlo, size_t dhi, size_t per_pass) {
    const size_t w0 = (size_t)c->prm.n + 1, w1 = (size_t)kN + 1;
    const int nl = q->lanes;
    Device &d = c->devs[k];
    std::vector<size_t> hi(nl), pos(nl), cur(nl);
    for (int l = 0; l < nl; l++) {
        pos[l] = dlo + (dhi - dlo) * l / nl;
        hi[l] = dlo + (dhi - dlo) * (l + 1) / nl;
    }
    CU(c, cudaSetDevice(d.id));
    bool more = true;
    while (more) {
        more = false;
        for (int l = 0; l < nl; l++) {
            auto &pd = q->dev[k * nl + l];
            const size_t inst = std::min(per_pass, hi[l] - pos[l]);
            cur[l] = inst;
            if (inst == 0) continue;
            const void *old_w = pd.wires.p, *old_l = pd.lv1.p, *old_d = pd.ksdig.p;
            if (int r = ensure(c, pd.wires, q->n_slots * inst * w0 * 4)) return r;
            if (int r = ensure(c, pd.lv1, std::max<size_t>(q->max_width, 1) * inst * w1 * 4)) return r;
            if (int r = ensure(c, pd.ksdig, std::max<size_t>(q->max_width, 1) * inst * keyswitch_tc_digit_words(c->prm.basebit, c->prm.iks_t) * 8)) return r;
            if ((old_w != pd.wires.p || old_l != pd.lv1.p || old_d != pd.ksdig.p) && pd.graph) { cudaGraphExecDestroy(pd.graph); pd.graph = nullptr; }
            uint32_t *wires = (uint32_t *)pd.wires.p;
            for (size_t i = 0; i < q->n_inputs; i++)
                CU(c, cudaMemcpyAsync(wires + i * inst * w0, inputs + (i * instances + pos[l]) * w0, inst * w0 * 4, cudaMemcpyHostToDevice, pd.stream));
            if (q->has_consts) {   // Gates.constant(true) = (0, 2^29), Gates.constant(false) = (0, 1 - 2^29) (src/gates.zig:146-147)
                CU(c, launch_fill_constant(wires + q->n_inputs * inst * w0, inst, (int)w0, 0x20000000u, pd.stream, &d.launches));
                CU(c, launch_fill_constant(wires + (q->n_inputs + 1) * inst * w0, inst, (int)w0, 0xE0000001u, pd.stream, &d.launches));
            }
            if (int r = circuit_enqueue(c, q, d, pd, inst)) return r;
        }
        for (int l = 0; l < nl; l++) {          // results back (a pageable D2H blocks this thread until the lane is done)
            auto &pd = q->dev[k * nl + l];
            const size_t inst = cur[l];
            if (inst == 0) continue;
            const uint32_t *wires = (const uint32_t *)pd.wires.p;
            for (size_t o = 0; o < q->outputs.size(); o++) {
                const uint32_t ref = q->outputs[o];
                const uint32_t *src = wires + (size_t)(ref & ~TFHE_B200_WIRE_NOT) * inst * w0;
                if (ref & TFHE_B200_WIRE_NOT) {   // Gates.notGate of the wire (src/gates.zig:131-133)
                    if (int r = ensure(c, pd.neg, inst * w0 * 4)) return r;
                    CU(c, launch_negate(src, (uint32_t *)pd.neg.p, inst * w0, pd.stream, &d.launches));
                    src = (const uint32_t *)pd.neg.p;
                }
                CU(c, cudaMemcpyAsync(outputs + (o * instances + pos[l]) * w0, src, inst * w0 * 4, cudaMemcpyDeviceToHost, pd.stream));
                if (ref & TFHE_B200_WIRE_NOT) CU(c, cudaStreamSynchronize(pd.stream));   // pd.neg is reused by the next output
            }
        }
        for (int l = 0; l < nl; l++) {
            CU(c, cudaStreamSynchronize(q->dev[k * nl + l].stream));
            pos[l] += cur[l];
            if (pos[l] < hi[l]) more = true;
        }
    }
    return 0;
}

int tfhe_b200_circuit_run(tfhe_b200_ctx *c, tfhe_b200_circuit *q, const uint32_t *inputs, uint32_t *outputs, size_t instances) {
    if (!c || !q || q->ctx != c) return fail(c, TFHE_B200_ERR_INVALID, "circuit belongs to another context");
    if (!c->has_key || !c->has_ksk) return fail(c, TFHE_B200_ERR_NO_KEY, "no cloud key loaded");
    if (instances == 0) return 0;
    if ((q->n_inputs && !inputs) || (!q->outputs.empty() && !outputs)) return fail(c, TFHE_B200_ERR_INVALID, "null buffer");
    const size_t w0 = (size_t)c->prm.n + 1;
    const int nd = (int)c->devs.size();
    // instances per pass: bounded by the launch chunk and by ~8 GiB of wire storage per device
    size_t per_pass = std::max<size_t>(1, c->max_chunk / std::max<size_t>(q->max_width, 1));
    per_pass = std::min(per_pass, std::max<size_t>(1, ((size_t)8 << 30) / (q->n_slots * w0 * 4 * q->lanes)));
    // contiguous instance ranges per device (then per lane); one host thread per device, as in run_host
    if (nd == 1) return circuit_run_device(c, q, 0, inputs, outputs, instances, 0, instances, per_pass);
    std::vector<int> rc(nd, 0);
    std::vector<std::thread> workers;
    for (int k = 0; k < nd; k++) {
        const size_t dlo = instances * k / nd, dhi = instances * (k + 1) / nd;
        if (dlo == dhi) continue;
        workers.emplace_back([=, &rc] { rc[k] = circuit_run_device(c, q, k, inputs, outputs, instances, dlo, dhi, per_pass); });
    }
    for (auto &w : workers) w.join();
    for (int k = 0; k < nd; k++)
        if (rc[k]) return rc[k];
    return 0;
}

void *tfhe_b200_stream(tfhe_b200_ctx *c, int dev) {
    if (!c || dev < 0 || dev >= (int)c->devs.size()) return nullptr;
    return (void *)c->devs[dev].stream;
}

int tfhe_b200_sync(tfhe_b200_ctx *c) {
    if (!c) return TFHE_B200_ERR_INVALID;
    for (Device &d : c->devs) {
        CU(c, cudaSetDevice(d.id));
        CU(c, cudaStreamSynchronize(d.stream));
    }
    return 0;
}

int tfhe_b200_track_margin(tfhe_b200_ctx *c, int enable) {
    if (!c) return TFHE_B200_ERR_INVALID;
    c->track_margin = enable != 0;
    c->config_epoch++;
    return 0;
}

double tfhe_b200_max_round_margin(tfhe_b200_ctx *c, int reset) {
    if (!c) return -1.0;
    double m = 0.0;
    for (Device &d : c->devs) {
        unsigned long long bits = 0;
        cudaSetDevice(d.id);
        cudaStreamSynchronize(d.stream);
        if (cudaMemcpy(&bits, d.margin_bits, 8, cudaMemcpyDeviceToHost) != cudaSuccess) return -1.0;
        double v;
        memcpy(&v, &bits, 8);
        m = std::max(m, v);
        if (reset) cudaMemset(d.margin_bits, 0, 8);
    }
    return m;
}

uint64_t tfhe_b200_launch_count(const tfhe_b200_ctx *c) {
    if (!c) return 0;
    uint64_t n = c->launches;
    for (const Device &d : c->devs) n += d.launches;
    return n;
}

int tfhe_b200_set_tuning(tfhe_b200_ctx *c, const char *key, int value) {
    if (!c || !key) return TFHE_B200_ERR_INVALID;
    if (!strcmp(key, "kct")) c->tune.kct = value;
    else if (!strcmp(key, "use_tma")) c->tune.use_tma = value;
    else if (!strcmp(key, "latency_mode")) c->tune.latency_mode = value;
    else if (!strcmp(key, "team")) c->tune.team = value;
    else if (!strcmp(key, "twt")) c->tune.twt = value;
    else if (!strcmp(key, "diag")) c->tune.diag = value;
    else if (!strcmp(key, "timing")) c->timing = value != 0;
    else if (!strcmp(key, "ks_tile")) c->ks_tile = value;
    else if (!strcmp(key, "ks_vec")) c->ks_vec = value;
    else if (!strcmp(key, "ks_fill")) c->ks_fill = value;
    else if (!strcmp(key, "ks_rot")) c->ks_rot = value;
    else if (!strcmp(key, "inject_fault")) c->inject_fault = value;
    else if (!strcmp(key, "host_pipeline")) c->host_pipeline = value;
    else if (!strcmp(key, "host_copy_threads")) c->host_copy_threads = std::max(0, std::min(value, 16));
    else if (!strcmp(key, "ks_tc")) c->ks_tc = value;
    else if (!strcmp(key, "ks_tc_min")) c->ks_tc_min = value;
    else if (!strcmp(key, "exact_legacy")) c->exact_legacy = value != 0;
    else if (!strcmp(key, "exact_kct")) c->exact_kct = value;
    else if (!strcmp(key, "circuit_graph")) c->circuit_graph = value != 0;
    else if (!strcmp(key, "circuit_lanes")) c->circuit_lanes = value;
    else if (!strcmp(key, "max_chunk")) c->max_chunk = value > 0 ? (size_t)value : c->max_chunk;
    else return fail(c, TFHE_B200_ERR_INVALID, "unknown tuning key %s", key);
    c->config_epoch++;
    return 0;
}

double tfhe_b200_last_kernel_ms(tfhe_b200_ctx *c, int dev, int which) {
    if (!c || dev < 0 || dev >= (int)c->devs.size() || which < 0 || which > 1) return -1.0;
    Device &d = c->devs[dev];
    if (!d.ev_valid || cudaSetDevice(d.id) != cudaSuccess || cudaEventSynchronize(d.ev[2]) != cudaSuccess) return -1.0;
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, d.ev[which], d.ev[which + 1]) != cudaSuccess) return -1.0;
    return (double)ms;
}

double tfhe_b200_measure_fp64_tflops(tfhe_b200_ctx *c, int dev) {
    if (!c || dev < 0 || dev >= (int)c->devs.size()) return -1.0;
    Device &d = c->devs[dev];
    if (cudaSetDevice(d.id) != cudaSuccess) return -1.0;
    double tf = 0.0;
    if (run_fp64_peak(d.sm_count, d.stream, &tf, &c->launches) != cudaSuccess) return -1.0;
    return tf;
}

}  // extern "C"

