// capi.cu -- the C ABI declared in include/zng_b200.h: context management, argument validation
// (mirroring the order of checks in deflateInit2 / deflate, deflate.c:292-323,823-863, for the
// frozen parameter set), kernel launches and the pipelined host-buffer entry points.
// No CPU fallback lives here: a missing device or a failed launch is an error.
#include "../../include/zng_b200.h"
#include "common.cuh"
#include "kernels.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>
#include <chrono>

using namespace zb;

namespace {
constexpr int kCounters = 64;
constexpr int kPipeMax = 16;                   // most slabs in flight on the host path (ctx->pipe of them are used)
constexpr uint32_t kSlabChunksDefault = 4096;  // 256 MiB of input per slab at 64 KiB chunks (ctx->slab_chunks, env ZNG_B200_SLAB_CHUNKS)
constexpr uint32_t kAutoCtaChunks = 8192;      // ZNG_B200_K1=auto: launches of at most this many chunks take the CTA-per-chain level-1 parser
constexpr uint32_t kBatchChunks = 16384;       // chunks per K1a/K1b launch pair (token scratch: 4 GiB at 64 KiB chunks)

// K1a -> K1b hand-over: LZ77 token lists (4 B per input byte) + token counts
struct Scratch {
    uint32_t* tokens = nullptr;
    size_t tok_words = 0;
    uint32_t* ntok = nullptr;
    uint32_t ntok_cap = 0;
};

// Streamed level-1 host path: ONE persistent parse kernel per call takes the chunks in order as the copy engine delivers
// them; emit / checksum / gather / D2H run per output slab on high-priority streams as soon as the slab is parsed.
constexpr uint32_t kStreamMinSlab = 256;           // smallest output slab (chunks); the default is 512 = 32 MiB of input (ctx->stream_shift; measured 2^8: 27.8, 2^9: 31.95, 2^10: 31.66, 2^11: 30.96 GB/s)
constexpr uint32_t kStreamPiece = 256;            // chunks per H2D piece (16 MiB)
constexpr uint32_t kStreamMaxChunks = 16384;      // chunks per call of the streamed path (1 GiB); longer inputs run as several
constexpr int kStreamOut = 4;
struct Streamed {
    bool ready = false;
    cudaStream_t copy = nullptr, parse = nullptr, d2h = nullptr, out[kStreamOut] = {};
    cudaEvent_t reset_done = nullptr, slab_done[kStreamMaxChunks / kStreamMinSlab] = {};
    uint8_t* d_in_alloc = nullptr; uint8_t* d_in = nullptr; uint8_t* d_slots = nullptr; uint8_t* d_packed = nullptr;
    uint32_t* tokens = nullptr; uint32_t* ntok = nullptr; uint32_t* d_meta = nullptr; uint64_t* d_offsets = nullptr;
    uint32_t* d_sync = nullptr;                   // [0] chunks delivered, [1] failed, [2] chunk counter, [4..] parsed chunks per slab
    uint32_t* d_res = nullptr;
    uint32_t* h_ready = nullptr; uint64_t* h_meta = nullptr; uint32_t* h_failed = nullptr;
    uint32_t* h_done = nullptr; uint32_t* d_h_done = nullptr;   // mapped: chunks parsed per slab, written by the parse kernel when a slab is complete
    uint32_t cap_chunks = 0;
};

struct Slab {
    cudaStream_t stream = nullptr;
    cudaEvent_t done = nullptr;
    uint8_t* d_in_alloc = nullptr;
    uint8_t* d_in = nullptr;
    uint8_t* d_slots = nullptr;
    uint8_t* d_packed = nullptr;
    uint32_t* d_sizes = nullptr;               // sizes | crcs | adlers, kSlabChunks each
    uint64_t* d_offsets = nullptr;             // kSlabChunks + 1
    uint32_t* d_res = nullptr;                 // [0] crc fold, [1] adler fold
    uint64_t* h_meta = nullptr;                // pinned: [0] total bytes, [1] crc | adler << 32
    Scratch scratch;
    size_t in_bytes = 0;
    bool busy = false;
};
// one in-flight batch of the host-buffer inflate path
struct InfSlab {
    cudaStream_t stream = nullptr;
    cudaEvent_t done = nullptr;
    uint8_t* d_in = nullptr;  size_t in_cap = 0;
    uint8_t* d_out = nullptr; size_t out_cap = 0;
    uint64_t* d_off = nullptr;                 // in offsets (cnt+1) | out offsets (cnt+1), relative to the slab
    uint32_t* d_res = nullptr;                 // sizes | checks | status | in_used | detail, cnt_cap each
    uint64_t* h_off = nullptr;                 // pinned staging of d_off
    uint32_t* h_res = nullptr;                 // pinned staging of d_res
    uint32_t cnt_cap = 0;
    uint32_t first = 0, cnt = 0;
    size_t out_lo = 0, out_bytes = 0;
    bool busy = false;
};
constexpr int kInfPipe = 3;
constexpr uint32_t kInfSlabMembers = 32768;
constexpr size_t kInfSlabIn = (size_t)96 << 20, kInfSlabOut = (size_t)160 << 20;
}  // namespace

struct zng_b200_ctx {
    int device = 0;
    int sms = 0;
    char err[256] = {0};
    uint32_t* counters = nullptr;
    uint8_t* tails = nullptr;                  // kCounters scratch buffers for the padded last chunks
    int next_counter = 0;
    // K1 hash-head slab pool: nsmid x 64 slabs of 128 KiB, handed out per SM by the kernel itself
    uint16_t* heads = nullptr;
    unsigned long long* sm_slots = nullptr;
    uint32_t nsmid = 0;
    uint32_t* heads32 = nullptr;               // primed K1: 256 KiB slabs of absolute positions, same (sm, slot) indexing
    uint8_t* ptails = nullptr;                 // primed K1: kCounters padded copies of the last chunks + their dictionary
    uint16_t* prevs = nullptr;                 // K2: prev[] slab pool and stale-window images, same (sm, slot) indexing
    uint32_t* vtails = nullptr;
    int chains_per_sm = 24;
    int k1_cta = 2;                            // level-1 parser of the device-resident / slab calls (env ZNG_B200_K1): 0 = warp per chain (K1a v3), 1 = CTA per
                                               // chain (K1a v7), 2 = auto: v7 up to kAutoCtaChunks chunks per launch, v3 beyond.  The streamed host path keeps v3.
    int warps_cta = 22;                        // warps per chain of the v5 parser (env ZNG_B200_K1_WARPS: 2, 4, 8)
    unsigned long long* k1_stats = nullptr;    // env ZNG_B200_K1_STATS=1: 16 debug counters of the v6 parser, printed when the context goes
    int chains_cta = 11;                       // chains (CTAs) per SM of the v5 parser (env ZNG_B200_K1_CHAINS)
    int chains_per_sm_l2 = 32;                 // measured on B200: 16 -> 12.4, 24 -> 14.4, 32 -> 15.4 GB/s
    Scratch scratch;                           // for the device-resident entry points; users are ordered by k1_done
    cudaEvent_t k1_done = nullptr;
    bool k1_pending = false;
    uint32_t* ck_scratch = nullptr;            // per-tile crcs | adlers for the flat checksum calls
    size_t ck_tiles = 0;
    uint32_t* d_result = nullptr;              // small result area
    uint32_t* h_result = nullptr;              // pinned
    // host path
    Slab slab[kPipeMax];
    Streamed st;
    int streamed = 1;                          // env ZNG_B200_STREAMED=0: level 1 goes through the slab pipeline as well
    uint32_t stream_shift = 9;                 // log2(chunks per output slab) of the streamed path (env ZNG_B200_STREAM_SHIFT, 8..11)
    int pipe = 4;                              // slabs in flight (env ZNG_B200_PIPE); round 1, level 1: 2048 x 4 -> 23.2 GB/s e2e, 1024 x 6 -> 20.6; round 2, levels 2-6: 4096 x 4
    uint32_t slab_chunks = kSlabChunksDefault;
    bool slabs_ready = false;
    size_t slab_stride = 0;
    InfSlab inf[kInfPipe];
    // parallel stream inflate: whole-stream device buffers (grow only)
    uint8_t* d_sin = nullptr;  size_t sin_cap = 0;
    uint8_t* d_sout = nullptr; size_t sout_cap = 0;
    uint8_t* d_sslots = nullptr; size_t sslots_cap = 0;   // one-pass attempt: a 64 KiB output slot per segment
    uint64_t* d_soff = nullptr; uint32_t* d_sres = nullptr; uint32_t seg_cap = 0;
    unsigned long long* d_marks = nullptr; uint32_t marks_cap = 0;
    cudaStream_t sstream[3] = {nullptr, nullptr, nullptr};   // byte pass + D2H of segment groups
    cudaEvent_t sready = nullptr;
    uint8_t* d_pbuf = nullptr; size_t pbuf_cap = 0;   // zng_b200_deflate_host_primed: [dictionary | input], slots | packed, sizes | crcs | adlers | offsets
    uint8_t* d_pout = nullptr; size_t pout_cap = 0;
    uint32_t* d_pmeta = nullptr; size_t pmeta_cap = 0;
    uint8_t* wins = nullptr;                   // K2w: 64.5 KiB window buffer per (sm, slot)
    uint8_t* blkflags = nullptr; size_t blkflags_cap = 0;   // K2w: 8 flags per chunk
    uint8_t* d_arena = nullptr; size_t arena_cap = 0;   // scratch of the host-callable operator table
    uint8_t* d_hostbuf = nullptr;              // staging for *_host checksums
    size_t hostbuf_cap = 0;
    uint32_t x2n[32];
};

namespace {

int fail(zng_b200_ctx* c, cudaError_t e, const char* what) {
    if (c) snprintf(c->err, sizeof(c->err), "%s: %s", what, cudaGetErrorString(e));
    return ZNG_B200_CUDA_ERROR;
}
int bad(zng_b200_ctx* c, const char* what) {
    if (c) snprintf(c->err, sizeof(c->err), "%s", what);
    return ZNG_B200_STREAM_ERROR;
}

#define CK(call, what) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail(ctx, e_, what); } while (0)

struct DeviceGuard {
    int prev = -1; bool ok = false;
    explicit DeviceGuard(int dev) { if (cudaGetDevice(&prev) == cudaSuccess) { ok = (prev == dev) || cudaSetDevice(dev) == cudaSuccess; } }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

int next_slot(zng_b200_ctx* ctx) {
    int k = ctx->next_counter;
    ctx->next_counter = (ctx->next_counter + 1) % kCounters;
    return k;
}

int ensure_heads(zng_b200_ctx* ctx) {
    if (ctx->heads) return 0;
    CK(query_nsmid(ctx->d_result, &ctx->nsmid), "query nsmid");
    CK(cudaMalloc(&ctx->sm_slots, (size_t)ctx->nsmid * sizeof(unsigned long long)), "cudaMalloc(sm slots)");
    CK(cudaMemset(ctx->sm_slots, 0, (size_t)ctx->nsmid * sizeof(unsigned long long)), "cudaMemset(sm slots)");
    CK(cudaMalloc(&ctx->heads, deflate_quick_head_bytes(ctx->nsmid)), "cudaMalloc(hash-head slab pool)");
    return 0;
}

int ensure_primed(zng_b200_ctx* ctx) {
    if (ctx->heads32) return 0;
    CK(cudaMalloc(&ctx->heads32, deflate_primed_head_bytes(ctx->nsmid)), "cudaMalloc(primed hash-head slab pool)");
    CK(cudaMalloc(&ctx->ptails, (size_t)kCounters * deflate_primed_tail_bytes()), "cudaMalloc(primed tails)");
    return 0;
}

int ensure_prevs(zng_b200_ctx* ctx) {
    if (ctx->prevs) return 0;
    CK(cudaMalloc(&ctx->prevs, deflate_fast_prev_bytes(ctx->nsmid)), "cudaMalloc(prev slab pool)");
    CK(cudaMalloc(&ctx->vtails, deflate_fast_tail_bytes(ctx->nsmid)), "cudaMalloc(window tail images)");
    return 0;
}

int ensure_scratch(zng_b200_ctx* ctx, Scratch& sc, uint32_t batch, uint32_t stride, uint32_t nchunks, bool need_tokens) {
    if (need_tokens && sc.tok_words < (size_t)batch * stride) {
        if (sc.tokens) { cudaDeviceSynchronize(); cudaFree(sc.tokens); sc.tokens = nullptr; sc.tok_words = 0; }
        CK(cudaMalloc(&sc.tokens, (size_t)batch * stride * sizeof(uint32_t)), "cudaMalloc(token scratch)");
        sc.tok_words = (size_t)batch * stride;
    }
    if (sc.ntok_cap < nchunks) {
        if (sc.ntok) { cudaDeviceSynchronize(); cudaFree(sc.ntok); sc.ntok = nullptr; sc.ntok_cap = 0; }
        const uint32_t want = nchunks < 4096u ? 4096u : nchunks;
        CK(cudaMalloc(&sc.ntok, (size_t)want * sizeof(uint32_t)), "cudaMalloc(token counts)");
        sc.ntok_cap = want;
    }
    return 0;
}

// Chunk compression.  Level 1: K1a parse -> token lists in `sc`, K1b static emit (+ the K3 tile kernel when per-chunk
// checksums are wanted).  Inputs larger than kBatchChunks chunks run as several batches that reuse
// the token scratch (4 B per input byte).
// level 2: K2a parse (hash chains) -> token lists, K2b block writer.  have_prev: see kernels.h.
int run_deflate_chunks(zng_b200_ctx* ctx, Scratch& sc, const uint8_t* d_in, size_t n, uint32_t chunk, uint32_t nchunks, int last,
                      uint8_t* d_out, size_t out_stride, uint32_t* d_sizes, uint32_t* d_crcs, uint32_t* d_adlers,
                      cudaStream_t stream, uint32_t* d_tokens, uint32_t tok_stride, int level = 1, int have_prev = 0, int ck_warps = 32) {
    if (nchunks == 0) return 0;
    int r = ensure_heads(ctx);
    if (r) return r;
    if (level >= 2) { r = ensure_prevs(ctx); if (r) return r; }
    const uint32_t own_stride = (chunk + 32u) & ~31u;                 // tokens per chunk incl. end marker, 128-byte rows
    const uint32_t batch = nchunks < kBatchChunks ? nchunks : kBatchChunks;
    r = ensure_scratch(ctx, sc, batch, own_stride, nchunks, d_tokens == nullptr);
    if (r) return r;
    for (uint32_t c0 = 0; c0 < nchunks; c0 += batch) {
        const uint32_t nb = (nchunks - c0) < batch ? (nchunks - c0) : batch;
        const size_t off = (size_t)c0 * chunk;
        const size_t nbytes = (c0 + nb == nchunks) ? n - off : (size_t)nb * chunk;
        uint32_t* toks = d_tokens ? d_tokens + (size_t)c0 * tok_stride : sc.tokens;
        const uint32_t stride = d_tokens ? tok_stride : own_stride;
        const uint32_t grid = deflate_quick_grid(nb, ctx->sms, ctx->chains_per_sm);
        const int slot = next_slot(ctx);
        if (level >= 2) {
            CK(launch_fast_parse(d_in + off, nbytes, chunk, nb, toks, stride, sc.ntok + c0, ctx->counters + slot, ctx->heads, ctx->prevs,
                                 ctx->vtails, ctx->sm_slots, ctx->sms, ctx->chains_per_sm_l2, (c0 > 0 || have_prev) ? 1 : 0, level, stream),
               "fast_parse launch");
            CK(launch_block_emit(d_in + off, toks, stride, sc.ntok + c0, nbytes, chunk, nb, last, d_out + (size_t)c0 * out_stride, out_stride,
                                 d_sizes + c0, ctx->sms, stream, -1, nullptr, ctx->counters + next_slot(ctx)),
               "block_emit launch");
            continue;
        }
        // auto: the CTA-per-chain parser (K1a v7) for batches that leave the warp-per-chain parser's 24 chains per SM under-filled or
        // barely filled -- it is 5-21 % faster up to ~8 Ki chunks and equal beyond (profiles/r2_latency_small_batches.txt)
        if (ctx->k1_cta == 1 || (ctx->k1_cta == 2 && nb <= kAutoCtaChunks))
            CK(launch_quick_parse_cta(d_in + off, nbytes, chunk, nb, toks, stride, sc.ntok + c0, ctx->counters + slot, ctx->heads, ctx->sm_slots,
                                      ctx->sms, ctx->chains_cta, ctx->warps_cta, ctx->tails + (size_t)slot * deflate_quick_tail_bytes(), stream, nullptr, ctx->k1_stats),
               "quick_parse_cta launch");
        else
            CK(launch_quick_parse(d_in + off, nbytes, chunk, nb, toks, stride, sc.ntok + c0, ctx->counters + slot, ctx->heads, ctx->sm_slots,
                                  grid, ctx->tails + (size_t)slot * deflate_quick_tail_bytes(), stream),
               "quick_parse launch");
        CK(launch_static_emit(toks, stride, sc.ntok + c0, nbytes, chunk, nb, last, d_out + (size_t)c0 * out_stride, out_stride,
                              d_sizes + c0, ctx->sms, stream),
           "static_emit launch");
    }
    if (d_crcs || d_adlers)
        CK(launch_checksum_tiles(d_in, n, chunk, nchunks, d_crcs, d_adlers, ctx->sms, stream, ck_warps), "checksum launch");
    return 0;
}

// the ctx-level scratch is shared by all device-resident calls: order them
int run_deflate_shared(zng_b200_ctx* ctx, const uint8_t* d_in, size_t n, uint32_t chunk, uint32_t nchunks, int last,
                             uint8_t* d_out, size_t out_stride, uint32_t* d_sizes, uint32_t* d_crcs, uint32_t* d_adlers,
                             cudaStream_t stream, uint32_t* d_tokens, uint32_t tok_stride, int level) {
    if (nchunks == 0) return 0;
    if (ctx->k1_pending) CK(cudaStreamWaitEvent(stream, ctx->k1_done, 0), "cudaStreamWaitEvent");
    int r = run_deflate_chunks(ctx, ctx->scratch, d_in, n, chunk, nchunks, last, d_out, out_stride, d_sizes, d_crcs, d_adlers, stream,
                              d_tokens, tok_stride, level, 0);
    if (r) return r;
    CK(cudaEventRecord(ctx->k1_done, stream), "cudaEventRecord");
    ctx->k1_pending = true;
    return 0;
}

// pigz's dependent-chunk mode at level 1: as run_deflate_shared, every chunk after the stream's first primed with the 32768
// bytes in front of it (deflate.c:456-512 deflateSetDictionary on a fresh stream).
// first0: index of d_in's first chunk in the stream (> 0: the 32768 bytes in front of d_in are its dictionary)
// levels 2-6: K2w (deflate_window.cu) keeps the reference's own window state per chain; the K2 block writer takes the tokens
int run_deflate_primed_window(zng_b200_ctx* ctx, const uint8_t* d_in, size_t n, uint32_t chunk, uint32_t nchunks, int last, int level,
                              uint8_t* d_out, size_t out_stride, uint32_t* d_sizes, uint32_t* d_crcs, uint32_t* d_adlers, cudaStream_t stream,
                              uint32_t first0) {
    int r = ensure_heads(ctx);
    if (r) return r;
    r = ensure_prevs(ctx);
    if (r) return r;
    if (!ctx->wins) CK(cudaMalloc(&ctx->wins, deflate_window_win_bytes(ctx->nsmid)), "cudaMalloc(window buffer pool)");
    if (ctx->blkflags_cap < (size_t)nchunks * 8u) {
        if (ctx->blkflags) { cudaDeviceSynchronize(); cudaFree(ctx->blkflags); ctx->blkflags = nullptr; ctx->blkflags_cap = 0; }
        const size_t want = ((size_t)nchunks < 4096u ? 4096u : (size_t)nchunks) * 8u;
        CK(cudaMalloc(&ctx->blkflags, want), "cudaMalloc(block flags)");
        ctx->blkflags_cap = want;
    }
    Scratch& sc = ctx->scratch;
    const uint32_t stride = (chunk + 32u) & ~31u;
    const uint32_t batch = nchunks < kBatchChunks ? nchunks : kBatchChunks;
    r = ensure_scratch(ctx, sc, batch, stride, nchunks, true);
    if (r) return r;
    for (uint32_t c0 = 0; c0 < nchunks; c0 += batch) {
        const uint32_t nb = (nchunks - c0) < batch ? (nchunks - c0) : batch;
        const size_t off = (size_t)c0 * chunk;
        const size_t nbytes = (c0 + nb == nchunks) ? n - off : (size_t)nb * chunk;
        const int slot = next_slot(ctx);
        static const int k2w_chains = [] { const char* e = getenv("ZNG_B200_CHAINS_K2W"); int v = e ? atoi(e) : 48; return v >= 1 && v <= 48 ? v : 48; }();
        CK(launch_window_parse(d_in + off, nbytes, chunk, nb, c0 + first0, last, sc.tokens, stride, sc.ntok + c0, ctx->blkflags + (size_t)c0 * 8u,
                               ctx->counters + slot, ctx->heads, ctx->prevs, ctx->wins, ctx->sm_slots, ctx->sms, k2w_chains, level, stream),
           "window_parse launch");
        CK(launch_block_emit(d_in + off, sc.tokens, stride, sc.ntok + c0, nbytes, chunk, nb, last, d_out + (size_t)c0 * out_stride, out_stride,
                             d_sizes + c0, ctx->sms, stream, -1, ctx->blkflags + (size_t)c0 * 8u, ctx->counters + next_slot(ctx)),
           "block_emit launch");
    }
    if (d_crcs || d_adlers)
        CK(launch_checksum_tiles(d_in, n, chunk, nchunks, d_crcs, d_adlers, ctx->sms, stream), "checksum launch");
    CK(cudaEventRecord(ctx->k1_done, stream), "cudaEventRecord");
    ctx->k1_pending = true;
    return 0;
}

int run_deflate_primed(zng_b200_ctx* ctx, const uint8_t* d_in, size_t n, uint32_t chunk, uint32_t nchunks, int last,
                       uint8_t* d_out, size_t out_stride, uint32_t* d_sizes, uint32_t* d_crcs, uint32_t* d_adlers, cudaStream_t stream,
                       uint32_t first0 = 0, int level = 1) {
    if (nchunks == 0) return 0;
    if (ctx->k1_pending) CK(cudaStreamWaitEvent(stream, ctx->k1_done, 0), "cudaStreamWaitEvent");
    if (level >= 2) return run_deflate_primed_window(ctx, d_in, n, chunk, nchunks, last, level, d_out, out_stride, d_sizes, d_crcs, d_adlers, stream, first0);
    int r = ensure_heads(ctx);
    if (r) return r;
    r = ensure_primed(ctx);
    if (r) return r;
    Scratch& sc = ctx->scratch;
    const uint32_t stride = (chunk + 32u) & ~31u;
    const uint32_t batch = nchunks < kBatchChunks ? nchunks : kBatchChunks;
    r = ensure_scratch(ctx, sc, batch, stride, nchunks, true);
    if (r) return r;
    for (uint32_t c0 = 0; c0 < nchunks; c0 += batch) {
        const uint32_t nb = (nchunks - c0) < batch ? (nchunks - c0) : batch;
        const size_t off = (size_t)c0 * chunk;
        const size_t nbytes = (c0 + nb == nchunks) ? n - off : (size_t)nb * chunk;
        const uint32_t grid = deflate_quick_grid(nb, ctx->sms, ctx->chains_per_sm);
        const int slot = next_slot(ctx);
        CK(launch_primed_parse(d_in + off, nbytes, chunk, nb, c0 + first0, sc.tokens, stride, sc.ntok + c0, ctx->counters + slot, ctx->heads32,
                               ctx->sm_slots, grid, ctx->ptails + (size_t)slot * deflate_primed_tail_bytes(), stream),
           "primed_parse launch");
        CK(launch_static_emit(sc.tokens, stride, sc.ntok + c0, nbytes, chunk, nb, last, d_out + (size_t)c0 * out_stride, out_stride,
                              d_sizes + c0, ctx->sms, stream),
           "static_emit launch");
    }
    if (d_crcs || d_adlers)
        CK(launch_checksum_tiles(d_in, n, chunk, nchunks, d_crcs, d_adlers, ctx->sms, stream), "checksum launch");
    CK(cudaEventRecord(ctx->k1_done, stream), "cudaEventRecord");
    ctx->k1_pending = true;
    return 0;
}

int ensure_ck_scratch(zng_b200_ctx* ctx, size_t tiles) {
    if (tiles <= ctx->ck_tiles) return 0;
    if (ctx->ck_scratch) cudaFree(ctx->ck_scratch);
    ctx->ck_scratch = nullptr; ctx->ck_tiles = 0;
    size_t want = tiles < 4096 ? 4096 : tiles;
    CK(cudaMalloc(&ctx->ck_scratch, want * 2 * sizeof(uint32_t)), "cudaMalloc(checksum scratch)");
    ctx->ck_tiles = want;
    return 0;
}

int ensure_slabs(zng_b200_ctx* ctx) {
    if (ctx->slabs_ready) return 0;
    const size_t stride = zng_b200_deflate_bound(ZNG_B200_CHUNK_MAX);
    ctx->slab_stride = stride;
    for (int i = 0; i < ctx->pipe; i++) {
        Slab& s = ctx->slab[i];
        CK(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking), "cudaStreamCreate");
        CK(cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming), "cudaEventCreate");
        // 32 KiB in front of the slab hold the stream bytes that precede it (level 2: stale-window image of a short last chunk)
        CK(cudaMalloc(&s.d_in_alloc, (size_t)ctx->slab_chunks * ZNG_B200_CHUNK_MAX + kWSize), "cudaMalloc(slab in)");
        s.d_in = s.d_in_alloc + kWSize;
        CK(cudaMalloc(&s.d_slots, (size_t)ctx->slab_chunks * stride), "cudaMalloc(slab slots)");
        CK(cudaMalloc(&s.d_packed, (size_t)ctx->slab_chunks * stride), "cudaMalloc(slab packed)");
        CK(cudaMalloc(&s.d_sizes, (size_t)ctx->slab_chunks * 3 * sizeof(uint32_t)), "cudaMalloc(slab sizes)");
        CK(cudaMalloc(&s.d_offsets, ((size_t)ctx->slab_chunks + 1) * sizeof(uint64_t)), "cudaMalloc(slab offsets)");
        CK(cudaMalloc(&s.d_res, 4 * sizeof(uint32_t)), "cudaMalloc(slab res)");
        CK(cudaHostAlloc(&s.h_meta, 4 * sizeof(uint64_t), cudaHostAllocDefault), "cudaHostAlloc(slab meta)");
    }
    ctx->slabs_ready = true;
    return 0;
}

int check_chunk_args(zng_b200_ctx* ctx, const void* d_in, size_t n, uint32_t chunk, int level, int flush,
                     const void* d_out, size_t out_stride, const uint32_t* d_sizes) {
    if (!ctx) return ZNG_B200_STREAM_ERROR;
    if (level < 1 || level > 6) return bad(ctx, "level must be 1 (deflate_quick), 2 (deflate_fast) or 3..6 (deflate_medium)");
    if (flush != ZNG_B200_SYNC_FLUSH && flush != ZNG_B200_FULL_FLUSH && flush != ZNG_B200_FINISH)
        return bad(ctx, "flush must be Z_SYNC_FLUSH, Z_FULL_FLUSH or Z_FINISH");
    if (chunk == 0 || chunk > ZNG_B200_CHUNK_MAX) return bad(ctx, "chunk must be in 1..65536");
    if (n && !d_in) return bad(ctx, "d_in is NULL");
    if (n && (!d_out || !d_sizes)) return bad(ctx, "d_out / d_sizes is NULL");
    if (n && (out_stride < zng_b200_deflate_bound(chunk) || (out_stride & 15u) || (reinterpret_cast<uintptr_t>(d_out) & 15u))) {
        snprintf(ctx->err, sizeof(ctx->err), "out_stride must be >= zng_b200_deflate_bound(chunk), a multiple of 16, d_out 16-byte aligned");
        return ZNG_B200_BUF_ERROR;
    }
    if ((n + chunk - 1) / chunk > 0xffffffffull) return bad(ctx, "too many chunks");
    return 0;
}

}  // namespace

extern "C" {

int zng_b200_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

int zng_b200_ctx_create(zng_b200_ctx** out, int device) {
    if (!out) return ZNG_B200_STREAM_ERROR;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return ZNG_B200_CUDA_ERROR;
    if (device < 0) { if (cudaGetDevice(&device) != cudaSuccess) return ZNG_B200_CUDA_ERROR; }
    if (device >= ndev) return ZNG_B200_STREAM_ERROR;
    zng_b200_ctx* ctx = new (std::nothrow) zng_b200_ctx();
    if (!ctx) return ZNG_B200_MEM_ERROR;
    ctx->device = device;
    DeviceGuard g(device);
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete ctx; return ZNG_B200_CUDA_ERROR; }
    ctx->sms = prop.multiProcessorCount;
    if (prop.major < 10) { delete ctx; return ZNG_B200_STREAM_ERROR; }       // sm_100a kernels only
    if (const char* e = getenv("ZNG_B200_CHAINS")) { int v = atoi(e); if (v >= 1 && v <= 64) ctx->chains_per_sm = v; }
    if (const char* e = getenv("ZNG_B200_K1")) ctx->k1_cta = strcmp(e, "cta") == 0 ? 1 : (strcmp(e, "warp") == 0 ? 0 : 2);
    if (const char* e = getenv("ZNG_B200_K1_STATS")) { if (atoi(e) && cudaMalloc(&ctx->k1_stats, 16 * sizeof(unsigned long long)) == cudaSuccess) cudaMemset(ctx->k1_stats, 0, 16 * sizeof(unsigned long long)); }
    if (const char* e = getenv("ZNG_B200_K1_CHAINS")) { int v = atoi(e); if (v >= 1 && v <= 16) ctx->chains_cta = v; }
    if (const char* e = getenv("ZNG_B200_K1_WARPS")) { int v = atoi(e); if (v >= 1 && v <= 4) ctx->warps_cta = 10 * v + (ctx->warps_cta % 10); }
    if (const char* e = getenv("ZNG_B200_K1_BPW")) { int v = atoi(e); if (v >= 1 && v <= 4) ctx->warps_cta = 10 * (ctx->warps_cta / 10) + v; }
    if (const char* e = getenv("ZNG_B200_SLAB_CHUNKS")) { int v = atoi(e); if (v >= 64 && v <= 16384) ctx->slab_chunks = (uint32_t)v; }
    if (const char* e = getenv("ZNG_B200_STREAMED")) ctx->streamed = atoi(e);
    if (const char* e = getenv("ZNG_B200_STREAM_SHIFT")) { int v = atoi(e); if (v >= 8 && v <= 11) ctx->stream_shift = (uint32_t)v; }
    if (const char* e = getenv("ZNG_B200_PIPE")) { int v = atoi(e); if (v >= 2 && v <= kPipeMax) ctx->pipe = v; }
    if (const char* e = getenv("ZNG_B200_CHAINS_L2")) { int v = atoi(e); if (v >= 1 && v <= 64) ctx->chains_per_sm_l2 = v; }
    if (cudaMalloc(&ctx->counters, kCounters * sizeof(uint32_t)) != cudaSuccess ||
        cudaMalloc(&ctx->tails, (size_t)kCounters * deflate_quick_tail_bytes()) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->k1_done, cudaEventDisableTiming) != cudaSuccess ||
        cudaMalloc(&ctx->d_result, 16 * sizeof(uint32_t)) != cudaSuccess ||
        cudaHostAlloc(&ctx->h_result, 16 * sizeof(uint32_t), cudaHostAllocDefault) != cudaSuccess) {
        zng_b200_ctx_destroy(ctx);
        return ZNG_B200_MEM_ERROR;
    }
    // The hash-head / prev tables are hit with random 2-byte accesses; the default L2 fetch granularity pulls more
    // than one 32-byte sector per miss (measured: ~58 B of DRAM reads per lookup).  Ask for sector-sized fetches.
    {
        size_t gran = 32;
        if (const char* e = getenv("ZNG_B200_L2FETCH")) gran = (size_t)atoi(e);
        if (gran) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran);
    }
    build_x2n(ctx->x2n);
    *out = ctx;
    return ZNG_B200_OK;
}

void zng_b200_ctx_destroy(zng_b200_ctx* ctx) {
    if (!ctx) return;
    DeviceGuard g(ctx->device);
    cudaDeviceSynchronize();
    for (int i = 0; i < kPipeMax; i++) {
        Slab& s = ctx->slab[i];
        if (s.d_in_alloc) cudaFree(s.d_in_alloc);
        if (s.d_slots) cudaFree(s.d_slots);
        if (s.d_packed) cudaFree(s.d_packed);
        if (s.d_sizes) cudaFree(s.d_sizes);
        if (s.d_offsets) cudaFree(s.d_offsets);
        if (s.d_res) cudaFree(s.d_res);
        if (s.h_meta) cudaFreeHost(s.h_meta);
        if (s.scratch.tokens) cudaFree(s.scratch.tokens);
        if (s.scratch.ntok) cudaFree(s.scratch.ntok);
        if (s.done) cudaEventDestroy(s.done);
        if (s.stream) cudaStreamDestroy(s.stream);
    }
    for (int i = 0; i < kInfPipe; i++) {
        InfSlab& s = ctx->inf[i];
        if (s.d_in) cudaFree(s.d_in);
        if (s.d_out) cudaFree(s.d_out);
        if (s.d_off) cudaFree(s.d_off);
        if (s.d_res) cudaFree(s.d_res);
        if (s.h_off) cudaFreeHost(s.h_off);
        if (s.h_res) cudaFreeHost(s.h_res);
        if (s.done) cudaEventDestroy(s.done);
        if (s.stream) cudaStreamDestroy(s.stream);
    }
    if (ctx->d_sin) cudaFree(ctx->d_sin);
    if (ctx->d_sout) cudaFree(ctx->d_sout);
    if (ctx->d_sslots) cudaFree(ctx->d_sslots);
    if (ctx->d_soff) cudaFree(ctx->d_soff);
    if (ctx->d_sres) cudaFree(ctx->d_sres);
    if (ctx->d_marks) cudaFree(ctx->d_marks);
    for (int i = 0; i < 3; i++) if (ctx->sstream[i]) cudaStreamDestroy(ctx->sstream[i]);
    if (ctx->sready) cudaEventDestroy(ctx->sready);
    if (ctx->counters) cudaFree(ctx->counters);
    if (ctx->tails) cudaFree(ctx->tails);
    if (ctx->heads) cudaFree(ctx->heads);
    if (ctx->prevs) cudaFree(ctx->prevs);
    if (ctx->heads32) cudaFree(ctx->heads32);
    if (ctx->ptails) cudaFree(ctx->ptails);
    if (ctx->vtails) cudaFree(ctx->vtails);
    if (ctx->sm_slots) cudaFree(ctx->sm_slots);
    if (ctx->scratch.tokens) cudaFree(ctx->scratch.tokens);
    if (ctx->scratch.ntok) cudaFree(ctx->scratch.ntok);
    if (ctx->k1_done) cudaEventDestroy(ctx->k1_done);
    if (ctx->ck_scratch) cudaFree(ctx->ck_scratch);
    if (ctx->d_result) cudaFree(ctx->d_result);
    if (ctx->h_result) cudaFreeHost(ctx->h_result);
    {
        Streamed& S = ctx->st;
        for (void* p : {(void*)S.d_in_alloc, (void*)S.d_slots, (void*)S.d_packed, (void*)S.tokens, (void*)S.ntok, (void*)S.d_meta, (void*)S.d_offsets,
                        (void*)S.d_sync, (void*)S.d_res}) if (p) cudaFree(p);
        for (void* p : {(void*)S.h_ready, (void*)S.h_meta, (void*)S.h_failed, (void*)S.h_done}) if (p) cudaFreeHost(p);
        if (S.copy) cudaStreamDestroy(S.copy);
        if (S.parse) cudaStreamDestroy(S.parse);
        if (S.d2h) cudaStreamDestroy(S.d2h);
        for (auto st : S.out) if (st) cudaStreamDestroy(st);
        if (S.reset_done) cudaEventDestroy(S.reset_done);
        for (auto e : S.slab_done) if (e) cudaEventDestroy(e);
    }
    if (ctx->d_hostbuf) cudaFree(ctx->d_hostbuf);
    if (ctx->d_arena) cudaFree(ctx->d_arena);
    if (ctx->wins) cudaFree(ctx->wins);
    if (ctx->blkflags) cudaFree(ctx->blkflags);
    if (ctx->k1_stats) {
        unsigned long long h[16] = {0};
        cudaDeviceSynchronize();
        if (cudaMemcpy(h, ctx->k1_stats, sizeof(h), cudaMemcpyDeviceToHost) == cudaSuccess && h[0])
            fprintf(stderr, "[zng_b200 K1 v6] chunks %llu  windows/chunk %.1f  steps/chunk %.1f  walk clk/step %.0f  walker busy %.1f %%  "
                            "producer(warp0) clk/window %.0f  hit steps %.1f %%  fallback steps %.2f %%  cuts %.2f %%  long compares/step %.3f\n",
                    h[0], (double)h[9] / h[0], (double)h[1] / h[0], (double)h[2] / (h[1] ? h[1] : 1), 100.0 * h[2] / (h[3] ? h[3] : 1),
                    (double)h[8] / (h[9] ? h[9] : 1), 100.0 * h[4] / (h[1] ? h[1] : 1), 100.0 * h[5] / (h[1] ? h[1] : 1), 100.0 * h[6] / (h[1] ? h[1] : 1),
                    (double)h[7] / (h[1] ? h[1] : 1));
        cudaFree(ctx->k1_stats);
    }
    if (ctx->d_pbuf) cudaFree(ctx->d_pbuf);
    if (ctx->d_pout) cudaFree(ctx->d_pout);
    if (ctx->d_pmeta) cudaFree(ctx->d_pmeta);
    delete ctx;
}

int zng_b200_ctx_device(const zng_b200_ctx* ctx) { return ctx ? ctx->device : -1; }
int zng_b200_ctx_sm_count(const zng_b200_ctx* ctx) { return ctx ? ctx->sms : 0; }
const char* zng_b200_last_error(const zng_b200_ctx* ctx) { return ctx ? ctx->err : "no context"; }

int zng_b200_sync(zng_b200_ctx* ctx, void* stream) {
    if (!ctx) return ZNG_B200_STREAM_ERROR;
    DeviceGuard g(ctx->device);
    CK(cudaStreamSynchronize((cudaStream_t)stream), "cudaStreamSynchronize");
    return 0;
}

void* zng_b200_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) return nullptr;
    return p;
}
void zng_b200_host_free(void* p) { if (p) cudaFreeHost(p); }

size_t zng_b200_deflate_bound(size_t chunk_len) {
    // 9 bits per literal + 3-bit header + 7-bit EOB + 5-byte flush marker, + 36 bytes of read slack
    // for the gather's word reads; rounded up to 16.
    size_t b = chunk_len + ((chunk_len + 7) >> 3) + 16 + 36;
    return (b + 15) & ~(size_t)15;
}

int zng_b200_deflate_chunks_trace(zng_b200_ctx* ctx, const void* d_in, size_t n, uint32_t chunk, int level, int flush,
                                  void* d_out, size_t out_stride, uint32_t* d_sizes, uint32_t* d_tokens,
                                  uint32_t tok_stride, void* stream) {
    int r = check_chunk_args(ctx, d_in, n, chunk, level, flush, d_out, out_stride, d_sizes);
    if (r) return r;
    if (!d_tokens || tok_stride < chunk + 1u) return bad(ctx, "d_tokens / tok_stride");
    DeviceGuard g(ctx->device);
    const uint32_t nchunks = (uint32_t)((n + chunk - 1) / chunk);
    return run_deflate_shared(ctx, (const uint8_t*)d_in, n, chunk, nchunks, flush == ZNG_B200_FINISH, (uint8_t*)d_out, out_stride,
                                    d_sizes, nullptr, nullptr, (cudaStream_t)stream, d_tokens, tok_stride, level);
}

int zng_b200_deflate_chunks(zng_b200_ctx* ctx, const void* d_in, size_t n, uint32_t chunk, int level, int flush,
                            void* d_out, size_t out_stride, uint32_t* d_sizes, uint32_t* d_crcs,
                            uint32_t* d_adlers, void* stream) {
    int r = check_chunk_args(ctx, d_in, n, chunk, level, flush, d_out, out_stride, d_sizes);
    if (r) return r;
    DeviceGuard g(ctx->device);
    const uint32_t nchunks = (uint32_t)((n + chunk - 1) / chunk);
    return run_deflate_shared(ctx, (const uint8_t*)d_in, n, chunk, nchunks, flush == ZNG_B200_FINISH, (uint8_t*)d_out, out_stride,
                                    d_sizes, d_crcs, d_adlers, (cudaStream_t)stream, nullptr, 0, level);
}

int zng_b200_deflate_chunks_primed(zng_b200_ctx* ctx, const void* d_in, size_t n, uint32_t chunk, int level, int flush,
                                   void* d_out, size_t out_stride, uint32_t* d_sizes, uint32_t* d_crcs,
                                   uint32_t* d_adlers, void* stream) {
    int r = check_chunk_args(ctx, d_in, n, chunk, level, flush, d_out, out_stride, d_sizes);
    if (r) return r;
    if (level < 1 || level > 6 || chunk != 65536u) return bad(ctx, "primed chunks: levels 1..6 and chunk 65536 only");
    DeviceGuard g(ctx->device);
    const uint32_t nchunks = (uint32_t)((n + chunk - 1) / chunk);
    return run_deflate_primed(ctx, (const uint8_t*)d_in, n, chunk, nchunks, flush == ZNG_B200_FINISH, (uint8_t*)d_out, out_stride,
                              d_sizes, d_crcs, d_adlers, (cudaStream_t)stream, 0, level);
}

int zng_b200_deflate_chunks_primed_at(zng_b200_ctx* ctx, const void* d_in, size_t n, uint32_t chunk, int level, int flush, int have_halo,
                                      void* d_out, size_t out_stride, uint32_t* d_sizes, uint32_t* d_crcs,
                                      uint32_t* d_adlers, void* stream) {
    int r = check_chunk_args(ctx, d_in, n, chunk, level, flush, d_out, out_stride, d_sizes);
    if (r) return r;
    if (level < 1 || level > 6 || chunk != 65536u) return bad(ctx, "primed chunks: levels 1..6 and chunk 65536 only");
    DeviceGuard g(ctx->device);
    const uint32_t nchunks = (uint32_t)((n + chunk - 1) / chunk);
    return run_deflate_primed(ctx, (const uint8_t*)d_in, n, chunk, nchunks, flush == ZNG_B200_FINISH, (uint8_t*)d_out, out_stride,
                              d_sizes, d_crcs, d_adlers, (cudaStream_t)stream, have_halo ? 1u : 0u, level);
}

int zng_b200_chunk_offsets(zng_b200_ctx* ctx, const uint32_t* d_sizes, uint32_t nchunks, uint64_t base,
                           uint64_t* d_offsets, void* stream) {
    if (!ctx || !d_offsets || (nchunks && !d_sizes)) return ZNG_B200_STREAM_ERROR;
    DeviceGuard g(ctx->device);
    CK(launch_offsets(d_sizes, nchunks, base, d_offsets, (cudaStream_t)stream), "offsets launch");
    return 0;
}

int zng_b200_gather_chunks(zng_b200_ctx* ctx, const void* d_slots, size_t out_stride, const uint32_t* d_sizes,
                           const uint64_t* d_offsets, uint32_t nchunks, void* d_dst, void* stream) {
    if (!ctx) return ZNG_B200_STREAM_ERROR;
    if (nchunks && (!d_slots || !d_sizes || !d_offsets || !d_dst)) return bad(ctx, "NULL argument");
    if ((out_stride & 15u) || (reinterpret_cast<uintptr_t>(d_slots) & 15u)) return bad(ctx, "slots must be 16-byte aligned");
    DeviceGuard g(ctx->device);
    CK(launch_gather((const uint8_t*)d_slots, out_stride, d_sizes, d_offsets, nchunks, (uint8_t*)d_dst, ctx->sms, (cudaStream_t)stream),
       "gather launch");
    return 0;
}

int zng_b200_checksum_chunks(zng_b200_ctx* ctx, const void* d_in, size_t n, uint32_t tile_bytes,
                             uint32_t* d_crcs, uint32_t* d_adlers, void* stream) {
    if (!ctx) return ZNG_B200_STREAM_ERROR;
    if (tile_bytes == 0 || tile_bytes > ZNG_B200_CHUNK_MAX) return bad(ctx, "tile_bytes must be in 1..65536");
    if (n && !d_in) return bad(ctx, "d_in is NULL");
    const size_t ntiles = (n + tile_bytes - 1) / tile_bytes;
    if (ntiles > 0xffffffffull) return bad(ctx, "too many tiles");
    DeviceGuard g(ctx->device);
    CK(launch_checksum_tiles((const uint8_t*)d_in, n, tile_bytes, (uint32_t)ntiles, d_crcs, d_adlers, ctx->sms, (cudaStream_t)stream),
       "checksum launch");
    return 0;
}

int zng_b200_crc32_fold(zng_b200_ctx* ctx, const uint32_t* d_crcs, uint32_t ntiles, uint32_t tile_bytes, size_t n,
                        uint32_t init, uint32_t* d_result, void* stream) {
    if (!ctx || !d_result || (ntiles && !d_crcs)) return ZNG_B200_STREAM_ERROR;
    if ((size_t)ntiles != (tile_bytes ? (n + tile_bytes - 1) / tile_bytes : 0)) return bad(ctx, "ntiles does not match n / tile_bytes");
    DeviceGuard g(ctx->device);
    CK(launch_crc32_fold(d_crcs, ntiles, tile_bytes, n, init, d_result, (cudaStream_t)stream), "crc32 fold launch");
    return 0;
}

int zng_b200_adler32_fold(zng_b200_ctx* ctx, const uint32_t* d_adlers, uint32_t ntiles, uint32_t tile_bytes, size_t n,
                          uint32_t init, uint32_t* d_result, void* stream) {
    if (!ctx || !d_result || (ntiles && !d_adlers)) return ZNG_B200_STREAM_ERROR;
    if ((size_t)ntiles != (tile_bytes ? (n + tile_bytes - 1) / tile_bytes : 0)) return bad(ctx, "ntiles does not match n / tile_bytes");
    DeviceGuard g(ctx->device);
    CK(launch_adler32_fold(d_adlers, ntiles, tile_bytes, n, init, d_result, (cudaStream_t)stream), "adler32 fold launch");
    return 0;
}

int zng_b200_crc32(zng_b200_ctx* ctx, const void* d_buf, size_t n, uint32_t init, uint32_t* d_result, void* stream) {
    if (!ctx || !d_result) return ZNG_B200_STREAM_ERROR;
    const size_t ntiles = (n + ZNG_B200_CHUNK_MAX - 1) / ZNG_B200_CHUNK_MAX;
    if (ntiles > 0xffffffffull) return bad(ctx, "buffer too large");
    DeviceGuard g(ctx->device);
    int r = ensure_ck_scratch(ctx, ntiles);
    if (r) return r;
    CK(launch_checksum_tiles((const uint8_t*)d_buf, n, ZNG_B200_CHUNK_MAX, (uint32_t)ntiles, ctx->ck_scratch, nullptr, ctx->sms, (cudaStream_t)stream),
       "checksum launch");
    CK(launch_crc32_fold(ctx->ck_scratch, (uint32_t)ntiles, ZNG_B200_CHUNK_MAX, n, init, d_result, (cudaStream_t)stream), "crc32 fold launch");
    return 0;
}

int zng_b200_adler32(zng_b200_ctx* ctx, const void* d_buf, size_t n, uint32_t init, uint32_t* d_result, void* stream) {
    if (!ctx || !d_result) return ZNG_B200_STREAM_ERROR;
    const size_t ntiles = (n + ZNG_B200_CHUNK_MAX - 1) / ZNG_B200_CHUNK_MAX;
    if (ntiles > 0xffffffffull) return bad(ctx, "buffer too large");
    DeviceGuard g(ctx->device);
    int r = ensure_ck_scratch(ctx, ntiles);
    if (r) return r;
    uint32_t* ad = ctx->ck_scratch + ctx->ck_tiles;
    CK(launch_checksum_tiles((const uint8_t*)d_buf, n, ZNG_B200_CHUNK_MAX, (uint32_t)ntiles, nullptr, ad, ctx->sms, (cudaStream_t)stream),
       "checksum launch");
    CK(launch_adler32_fold(ad, (uint32_t)ntiles, ZNG_B200_CHUNK_MAX, n, init, d_result, (cudaStream_t)stream), "adler32 fold launch");
    return 0;
}

// ---------------------------------------------------------------- operator surface
int zng_b200_op_compare256(zng_b200_ctx* ctx, const void* d_a, const void* d_b, size_t stride, uint32_t n_pairs, uint32_t* d_out, void* stream) {
    if (!ctx) return ZNG_B200_STREAM_ERROR;
    if (n_pairs && (!d_a || !d_b || !d_out)) return bad(ctx, "NULL argument");
    DeviceGuard g(ctx->device);
    CK(launch_op_compare256((const uint8_t*)d_a, (const uint8_t*)d_b, stride, n_pairs, d_out, (cudaStream_t)stream), "compare256 launch");
    return 0;
}
int zng_b200_op_longest_match(zng_b200_ctx* ctx, const void* d_window, uint32_t n, const uint16_t* d_prev, const uint32_t* d_pos,
                              const uint32_t* d_cand, uint32_t n_q, uint32_t* d_len, uint32_t* d_start, void* stream) {
    if (!ctx) return ZNG_B200_STREAM_ERROR;
    if (n_q && (!d_window || !d_prev || !d_pos || !d_cand || !d_len || !d_start)) return bad(ctx, "NULL argument");
    if (n > ZNG_B200_CHUNK_MAX) return bad(ctx, "window larger than 65536");
    DeviceGuard g(ctx->device);
    CK(launch_op_longest_match((const uint8_t*)d_window, n, d_prev, d_pos, d_cand, n_q, d_len, d_start, (cudaStream_t)stream), "longest_match launch");
    return 0;
}
int zng_b200_op_longest_match_level(zng_b200_ctx* ctx, const void* d_window, uint32_t n, const uint16_t* d_prev, const uint32_t* d_pos,
                                    const uint32_t* d_cand, uint32_t n_q, int level, uint32_t* d_len, uint32_t* d_start, void* stream) {
    if (!ctx) return ZNG_B200_STREAM_ERROR;
    const bool raw = (level & 0x100) != 0;                     // internal: the host operator table wants the reference's raw return value
    level &= 0xff;
    if (level < 2 || level > 6) return bad(ctx, "longest_match: levels 2..6 (levels 7-9 use longest_match_slow, outside the hot path)");
    if (n_q && (!d_window || !d_prev || !d_pos || !d_cand || !d_len || !d_start)) return bad(ctx, "NULL argument");
    if (n > ZNG_B200_CHUNK_MAX) return bad(ctx, "window larger than 65536");
    DeviceGuard g(ctx->device);
    CK(launch_op_longest_match((const uint8_t*)d_window, n, d_prev, d_pos, d_cand, n_q, d_len, d_start, (cudaStream_t)stream, level, raw), "longest_match launch");
    return 0;
}
int zng_b200_op_quick_insert_string(zng_b200_ctx* ctx, const void* d_window, uint16_t* d_head, uint16_t* d_prev, uint32_t str,
                                    uint32_t* d_old_head, void* stream) {
    if (!ctx) return ZNG_B200_STREAM_ERROR;
    if (!d_window || !d_head || !d_prev || !d_old_head) return bad(ctx, "NULL argument");
    DeviceGuard g(ctx->device);
    CK(launch_op_insert_string((const uint8_t*)d_window, d_head, d_prev, str, 1, (cudaStream_t)stream, d_old_head), "quick_insert_string launch");
    return 0;
}
int zng_b200_op_slide_hash(zng_b200_ctx* ctx, uint16_t* d_head, uint16_t* d_prev, uint32_t wsize, void* stream) {
    if (!ctx) return ZNG_B200_STREAM_ERROR;
    if (!d_head || !d_prev) return bad(ctx, "NULL argument");
    if (wsize == 0 || wsize > 32768u || (wsize & 7u)) return bad(ctx, "slide_hash: wsize must be a multiple of 8 up to 32768");
    DeviceGuard g(ctx->device);
    CK(launch_op_slide_hash(d_head, d_prev, wsize, (cudaStream_t)stream), "slide_hash launch");
    return 0;
}
// device scratch of the host-callable operator table (host/zng_functable.c): grow-only, owned by the context -- no allocation per call
void* zng_b200_ctx_arena(zng_b200_ctx* ctx, size_t bytes) {
    if (!ctx) return nullptr;
    DeviceGuard g(ctx->device);
    if (ctx->arena_cap < bytes) {
        if (ctx->d_arena) { cudaDeviceSynchronize(); cudaFree(ctx->d_arena); ctx->d_arena = nullptr; ctx->arena_cap = 0; }
        size_t want = bytes < ((size_t)1 << 20) ? ((size_t)1 << 20) : bytes + bytes / 4;
        if (cudaMalloc(&ctx->d_arena, want) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        ctx->arena_cap = want;
    }
    return ctx->d_arena;
}
int zng_b200_op_insert_string(zng_b200_ctx* ctx, const void* d_window, uint16_t* d_head, uint16_t* d_prev, uint32_t str, uint32_t count, void* stream) {
    if (!ctx) return ZNG_B200_STREAM_ERROR;
    if (count && (!d_window || !d_head || !d_prev)) return bad(ctx, "NULL argument");
    DeviceGuard g(ctx->device);
    CK(launch_op_insert_string((const uint8_t*)d_window, d_head, d_prev, str, count, (cudaStream_t)stream), "insert_string launch");
    return 0;
}
int zng_b200_op_chunkmemset(zng_b200_ctx* ctx, void* d_out, uint32_t pos, uint32_t dist, uint32_t len, void* stream) {
    if (!ctx) return ZNG_B200_STREAM_ERROR;
    if (len && (!d_out || dist == 0 || dist > pos)) return bad(ctx, "dist must be in 1..pos");
    DeviceGuard g(ctx->device);
    CK(launch_op_chunkmemset((uint8_t*)d_out, pos, dist, len, (cudaStream_t)stream), "chunkmemset launch");
    return 0;
}

// ---------------------------------------------------------------- K4 inflate
int zng_b200_inflate_members(zng_b200_ctx* ctx, const void* d_in, const uint64_t* d_in_off, uint32_t n_members, int window_bits,
                             void* d_out, const uint64_t* d_out_off, uint32_t* d_sizes, uint32_t* d_checks, int32_t* d_status,
                             uint32_t* d_in_used, uint32_t* d_detail, void* stream) {
    if (!ctx) return ZNG_B200_STREAM_ERROR;
    // zng_inflateInit2's own windowBits rules (inflate.c:232-251)
    int wb = window_bits;
    if (wb < 0) { if (wb < -15) return bad(ctx, "windowBits out of range"); wb = -wb; }
    else if (wb < 48) wb &= 15;
    if (wb && (wb < 8 || wb > 15)) return bad(ctx, "windowBits out of range");
    if (n_members && (!d_in || !d_in_off || !d_out || !d_out_off || !d_sizes || !d_status)) return bad(ctx, "NULL argument");
    DeviceGuard g(ctx->device);
    const int slot = next_slot(ctx);
    CK(launch_inflate_members((const uint8_t*)d_in, d_in_off, n_members, window_bits, (uint8_t*)d_out, d_out_off, d_sizes, d_checks,
                              d_status, d_in_used, d_detail, ctx->counters + slot, ctx->sms, (cudaStream_t)stream),
       "inflate launch");
    return 0;
}

const char* zng_b200_inflate_msg(uint32_t detail) { return inflate_msg(detail & 0xffu); }

static int inf_slab_reserve(zng_b200_ctx* ctx, InfSlab& s, uint32_t cnt, size_t in_bytes, size_t out_bytes) {
    if (!s.stream) {
        CK(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking), "cudaStreamCreate");
        CK(cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming), "cudaEventCreate");
    }
    if (s.in_cap < in_bytes + 16) {
        if (s.d_in) cudaFree(s.d_in);
        s.d_in = nullptr; s.in_cap = 0;
        const size_t want = in_bytes + 16 > kInfSlabIn ? in_bytes + 16 : kInfSlabIn;
        CK(cudaMalloc(&s.d_in, want), "cudaMalloc(inflate in)");
        s.in_cap = want;
    }
    if (s.out_cap < out_bytes + 16) {
        if (s.d_out) cudaFree(s.d_out);
        s.d_out = nullptr; s.out_cap = 0;
        const size_t want = out_bytes + 16 > kInfSlabOut ? out_bytes + 16 : kInfSlabOut;
        CK(cudaMalloc(&s.d_out, want), "cudaMalloc(inflate out)");
        s.out_cap = want;
    }
    if (s.cnt_cap < cnt) {
        if (s.d_off) cudaFree(s.d_off);
        if (s.d_res) cudaFree(s.d_res);
        if (s.h_off) cudaFreeHost(s.h_off);
        if (s.h_res) cudaFreeHost(s.h_res);
        s.d_off = nullptr; s.d_res = nullptr; s.h_off = nullptr; s.h_res = nullptr; s.cnt_cap = 0;
        const uint32_t want = cnt > kInfSlabMembers ? cnt : kInfSlabMembers;
        CK(cudaMalloc(&s.d_off, 2 * ((size_t)want + 1) * sizeof(uint64_t)), "cudaMalloc(inflate offsets)");
        CK(cudaMalloc(&s.d_res, 5 * (size_t)want * sizeof(uint32_t)), "cudaMalloc(inflate results)");
        CK(cudaHostAlloc(&s.h_off, 2 * ((size_t)want + 1) * sizeof(uint64_t), cudaHostAllocDefault), "cudaHostAlloc(inflate offsets)");
        CK(cudaHostAlloc(&s.h_res, 5 * (size_t)want * sizeof(uint32_t), cudaHostAllocDefault), "cudaHostAlloc(inflate results)");
        s.cnt_cap = want;
    }
    return 0;
}

static int inf_slab_drain(zng_b200_ctx* ctx, InfSlab& s, uint32_t* h_sizes, uint32_t* h_checks, int32_t* h_status,
                          uint32_t* h_in_used, uint32_t* h_detail) {
    if (!s.busy) return 0;
    s.busy = false;
    CK(cudaEventSynchronize(s.done), "cudaEventSynchronize");
    const size_t c = s.cnt, cap = s.cnt_cap;
    memcpy(h_sizes + s.first, s.h_res, c * 4);
    if (h_checks) memcpy(h_checks + s.first, s.h_res + cap, c * 4);
    memcpy(h_status + s.first, s.h_res + 2 * cap, c * 4);
    if (h_in_used) memcpy(h_in_used + s.first, s.h_res + 3 * cap, c * 4);
    if (h_detail) memcpy(h_detail + s.first, s.h_res + 4 * cap, c * 4);
    return 0;
}

int zng_b200_inflate_members_host(zng_b200_ctx* ctx, const void* h_in, const uint64_t* h_in_off, uint32_t n_members, int window_bits,
                                  void* h_out, const uint64_t* h_out_off, uint32_t* h_sizes, uint32_t* h_checks, int32_t* h_status,
                                  uint32_t* h_in_used, uint32_t* h_detail) {
    if (!ctx) return ZNG_B200_STREAM_ERROR;
    if (n_members && (!h_in || !h_in_off || !h_out || !h_out_off || !h_sizes || !h_status)) return bad(ctx, "NULL argument");
    DeviceGuard g(ctx->device);
    int k = 0, r = 0;
    uint32_t m = 0;
    while (m < n_members) {
        InfSlab& s = ctx->inf[k];
        r = inf_slab_drain(ctx, s, h_sizes, h_checks, h_status, h_in_used, h_detail);
        if (r) break;
        // greedy batch: as many consecutive members as fit the slab limits (at least one)
        uint32_t cnt = 0;
        const uint64_t i0 = h_in_off[m], o0 = h_out_off[m];
        while (m + cnt < n_members && cnt < kInfSlabMembers) {
            const uint64_t i1 = h_in_off[m + cnt + 1], o1 = h_out_off[m + cnt + 1];
            if (i1 < h_in_off[m + cnt] || o1 < h_out_off[m + cnt] || i1 - h_in_off[m + cnt] > 0xffffffffull || o1 - h_out_off[m + cnt] > 0xffffffffull) {
                for (int i = 0; i < kInfPipe; i++) { ctx->inf[i].busy = false; if (ctx->inf[i].stream) cudaStreamSynchronize(ctx->inf[i].stream); }
                return bad(ctx, "offsets must be non-decreasing, members < 4 GiB");
            }
            if (cnt && (i1 - i0 > kInfSlabIn || o1 - o0 > kInfSlabOut)) break;
            cnt++;
        }
        const size_t in_bytes = (size_t)(h_in_off[m + cnt] - i0), out_bytes = (size_t)(h_out_off[m + cnt] - o0);
        r = inf_slab_reserve(ctx, s, cnt, in_bytes, out_bytes);
        if (r) break;
        uint64_t* hi = s.h_off; uint64_t* ho = s.h_off + (cnt + 1);
        for (uint32_t j = 0; j <= cnt; j++) { hi[j] = h_in_off[m + j] - i0; ho[j] = h_out_off[m + j] - o0; }
        if (in_bytes) CK(cudaMemcpyAsync(s.d_in, (const uint8_t*)h_in + i0, in_bytes, cudaMemcpyHostToDevice, s.stream), "H2D members");
        CK(cudaMemcpyAsync(s.d_off, s.h_off, 2 * ((size_t)cnt + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, s.stream), "H2D offsets");
        const size_t cap = s.cnt_cap;
        const int slot = next_slot(ctx);
        // batched members: the whole output range of the slab goes back in one copy; what a member leaves unwritten inside its
        // range (it decoded less than its capacity) must not carry bytes of an earlier call on this context
        if (n_members != 1 && out_bytes) CK(cudaMemsetAsync(s.d_out, 0, out_bytes, s.stream), "cudaMemsetAsync(inflate output)");
        CK(launch_inflate_members(s.d_in, s.d_off, cnt, window_bits, s.d_out, s.d_off + (cnt + 1), s.d_res, s.d_res + cap,
                                  (int32_t*)(s.d_res + 2 * cap), s.d_res + 3 * cap, s.d_res + 4 * cap, ctx->counters + slot, ctx->sms, s.stream),
           "inflate launch");
        if (n_members == 1) {
            // the one-member call of zng_inflate / zng_uncompress: the capacity is the caller's whole avail_out.  Only the bytes
            // the decoder produced are moved (the rest of the caller's buffer stays untouched, no stale device bytes travel).
            CK(cudaMemcpyAsync(s.h_res, s.d_res, 5 * cap * sizeof(uint32_t), cudaMemcpyDeviceToHost, s.stream), "D2H results");
            CK(cudaStreamSynchronize(s.stream), "cudaStreamSynchronize");
            size_t produced = s.h_res[0];
            if (produced > out_bytes) produced = out_bytes;
            if (produced) CK(cudaMemcpyAsync((uint8_t*)h_out + o0, s.d_out, produced, cudaMemcpyDeviceToHost, s.stream), "D2H output");
        } else {
            if (out_bytes) CK(cudaMemcpyAsync((uint8_t*)h_out + o0, s.d_out, out_bytes, cudaMemcpyDeviceToHost, s.stream), "D2H output");
            CK(cudaMemcpyAsync(s.h_res, s.d_res, 5 * cap * sizeof(uint32_t), cudaMemcpyDeviceToHost, s.stream), "D2H results");
        }
        CK(cudaEventRecord(s.done, s.stream), "event record");
        s.first = m; s.cnt = cnt; s.busy = true;
        m += cnt;
        k = (k + 1) % kInfPipe;
    }
    for (int i = 0; i < kInfPipe && !r; i++) r = inf_slab_drain(ctx, ctx->inf[(k + i) % kInfPipe], h_sizes, h_checks, h_status, h_in_used, h_detail);
    for (int i = 0; i < kInfPipe; i++) { ctx->inf[i].busy = false; if (ctx->inf[i].stream) cudaStreamSynchronize(ctx->inf[i].stream); }
    return r;
}

// ---------------------------------------------------------------- parallel inflate of ONE flush-delimited stream
// (the inverse of the chunked compressor.)  A pigz-style stream -- what zng_deflate of this
// library and of the reference produce when every piece ends with Z_FULL_FLUSH -- is a sequence of independent raw
// deflate segments separated by the empty stored block 00 00 FF FF.  The markers are found on the device
// (inflateSync's search, inflate.c:1290-1306), every segment is decoded by its own warp in two passes (sizes, then
// bytes at the scanned offsets), a marker that turns out to lie inside a block is dropped and its segment re-sized,
// and the checksum of the whole output comes from K3.  Anything unusual falls back to the exact one-member path.
namespace {
struct Wrapper { int kind; size_t body; };               // kind 0 raw, 1 zlib, 2 gzip

// 0 = header understood, -1 = leave it to the exact path (errors, FDICT, FHCRC ...)
int parse_wrapper(const uint8_t* p, size_t n, int window_bits, Wrapper& w) {
    w.kind = 0; w.body = 0;
    if (window_bits < 0) return 0;
    const int wrap = (window_bits >> 4) + 5;
    if (n < 2) return -1;
    if ((wrap & 2) && p[0] == 0x1f && p[1] == 0x8b) {
        if (n < 10 || p[2] != 8 || (p[3] & 0xe0) || (p[3] & 0x02)) return -1;
        size_t b = 10;
        if (p[3] & 0x04) { if (n < b + 2) return -1; const size_t xlen = p[b] | (p[b + 1] << 8); b += 2; if (n < b + xlen) return -1; b += xlen; }
        for (int f = 0x08; f <= 0x10; f <<= 1)
            if (p[3] & f) { while (b < n && p[b]) b++; if (b >= n) return -1; b++; }
        w.kind = 2; w.body = b;
        return 0;
    }
    if (!(wrap & 1)) return -1;
    if (((p[0] << 8) + p[1]) % 31 || (p[0] & 0xf) != 8 || (p[0] >> 4) + 8 > 15 || (p[1] & 0x20)) return -1;
    const int wb = window_bits < 48 ? (window_bits & 15) : window_bits;
    if (wb && (p[0] >> 4) + 8 > wb) return -1;
    w.kind = 1; w.body = 2;
    return 0;
}

int grow(zng_b200_ctx* ctx, uint8_t*& p, size_t& cap, size_t want, const char* what) {
    if (cap >= want) return 0;
    if (p) { cudaDeviceSynchronize(); cudaFree(p); p = nullptr; cap = 0; }
    want += want / 8 + 4096;
    CK(cudaMalloc(&p, want), what);
    cap = want;
    return 0;
}
}  // namespace

int zng_b200_inflate_stream_host(zng_b200_ctx* ctx, const void* h_in, size_t n, int window_bits, void* h_out, size_t cap,
                                 size_t* out_len, size_t* in_used, uint32_t* check, int32_t* status, uint32_t* detail) {
    if (!ctx || !out_len || !status || (n && !h_in) || (cap && !h_out)) return ZNG_B200_STREAM_ERROR;
    uint32_t dummy_detail = 0; if (!detail) detail = &dummy_detail;
    size_t dummy_used = 0; if (!in_used) in_used = &dummy_used;
    uint32_t dummy_check = 0; if (!check) check = &dummy_check;
    const uint8_t* src = (const uint8_t*)h_in;
    auto exact = [&]() -> int {                               // one member, one warp: reference-exact codes and messages
        const size_t nn = n > 0xffffffffull ? 0xffffffffull : n, cc = cap > 0xfffffff0ull ? 0xfffffff0ull : cap;
        uint64_t ioff[2] = {0, nn}, ooff[2] = {0, cc};
        uint32_t sz = 0, used = 0; uint8_t d0 = 0, d1 = 0;
        int r = zng_b200_inflate_members_host(ctx, nn ? h_in : &d0, ioff, 1, window_bits, cc ? h_out : &d1, ooff, &sz, check, status, &used, detail);
        *out_len = sz; *in_used = used;
        return r;
    };
    Wrapper wr;
    if (n < ((size_t)1 << 20) || n > 0xffffffffull || parse_wrapper(src, n, window_bits, wr) != 0) return exact();
    DeviceGuard g(ctx->device);
    cudaStream_t st = 0;
    int r = grow(ctx, ctx->d_sin, ctx->sin_cap, n + 64, "cudaMalloc(stream in)");
    if (r) return r;
    CK(cudaMemcpyAsync(ctx->d_sin, src, n, cudaMemcpyHostToDevice, st), "H2D stream");
    CK(cudaMemsetAsync(ctx->d_sin + n, 0, 64, st), "memset");
    // ---- markers
    const uint32_t want_marks = (uint32_t)std::min<size_t>(n / 64 + 1024, 1u << 24);
    if (ctx->marks_cap < want_marks) {
        if (ctx->d_marks) { cudaDeviceSynchronize(); cudaFree(ctx->d_marks); ctx->d_marks = nullptr; ctx->marks_cap = 0; }
        CK(cudaMalloc(&ctx->d_marks, (size_t)want_marks * sizeof(unsigned long long)), "cudaMalloc(markers)");
        ctx->marks_cap = want_marks;
    }
    CK(launch_marker_scan(ctx->d_sin, n, wr.body, ctx->d_marks, ctx->marks_cap, ctx->d_result, ctx->sms, st), "marker scan");
    CK(cudaMemcpyAsync(ctx->h_result, ctx->d_result, sizeof(uint32_t), cudaMemcpyDeviceToHost, st), "D2H count");
    CK(cudaStreamSynchronize(st), "sync");
    const uint32_t nmarks = ctx->h_result[0];
    if (nmarks == 0 || nmarks > ctx->marks_cap) return exact();
    std::vector<unsigned long long> marks(nmarks);
    CK(cudaMemcpy(marks.data(), ctx->d_marks, (size_t)nmarks * sizeof(unsigned long long), cudaMemcpyDeviceToHost), "D2H markers");
    std::sort(marks.begin(), marks.end());
    std::vector<uint64_t> starts;
    starts.push_back(wr.body);
    for (unsigned long long m : marks) if (m > wr.body && m < n) starts.push_back(m);
    const size_t tlen0 = wr.kind == 2 ? 8 : (wr.kind == 1 ? 4 : 0);
    // ---- one pass when every segment fits a 64 KiB slot (the streams this library and pigz -b 64 write): decode straight
    // into slots, scan the sizes, gather.  Any surprise (a segment larger than its slot, a false marker, an error) leaves
    // it to the two-pass path below.
    {
        const uint32_t nseg1 = (uint32_t)starts.size();
        const size_t S = ZNG_B200_CHUNK_MAX;
        bool ok = (size_t)nseg1 * S <= ((size_t)24 << 30);
        std::vector<uint32_t> h1;
        size_t total1 = 0, end1 = 0;
        if (ok) {
            r = grow(ctx, ctx->d_sslots, ctx->sslots_cap, (size_t)nseg1 * S + 64, "cudaMalloc(stream slots)");
            if (r) return r;
            if (ctx->seg_cap < nseg1 + 1u) {
                if (ctx->d_soff) { cudaDeviceSynchronize(); cudaFree(ctx->d_soff); cudaFree(ctx->d_sres); ctx->d_soff = nullptr; ctx->d_sres = nullptr; ctx->seg_cap = 0; }
                const uint32_t want = nseg1 + nseg1 / 4 + 1024;
                CK(cudaMalloc(&ctx->d_soff, 4 * ((size_t)want + 1) * sizeof(uint64_t)), "cudaMalloc(segment offsets)");
                CK(cudaMalloc(&ctx->d_sres, 5 * (size_t)want * sizeof(uint32_t)), "cudaMalloc(segment results)");
                ctx->seg_cap = want;
            }
            std::vector<uint64_t> ho(2 * ((size_t)nseg1 + 1));
            for (uint32_t i = 0; i < nseg1; i++) { ho[i] = starts[i]; ho[nseg1 + 1 + i] = (uint64_t)i * S; }
            ho[nseg1] = n; ho[2 * (size_t)nseg1 + 1] = (uint64_t)nseg1 * S;
            CK(cudaMemcpyAsync(ctx->d_soff, ho.data(), ho.size() * sizeof(uint64_t), cudaMemcpyHostToDevice, st), "H2D segment offsets");
            const size_t capn = ctx->seg_cap;
            const int slot = next_slot(ctx);
            CK(launch_inflate_members(ctx->d_sin, ctx->d_soff, nseg1, -15, ctx->d_sslots, ctx->d_soff + (nseg1 + 1), ctx->d_sres, nullptr,
                                      (int32_t*)(ctx->d_sres + 2 * capn), ctx->d_sres + 3 * capn, ctx->d_sres + 4 * capn, ctx->counters + slot,
                                      ctx->sms, st, 1),
               "inflate slot pass");
            h1.resize(5 * capn);
            CK(cudaMemcpyAsync(h1.data(), ctx->d_sres, 5 * capn * sizeof(uint32_t), cudaMemcpyDeviceToHost, st), "D2H segment results");
            CK(cudaStreamSynchronize(st), "sync");
            uint32_t used_segs = 0;
            for (uint32_t i = 0; i < nseg1 && ok; i++) {
                const int32_t ret = (int32_t)h1[2 * capn + i];
                const uint64_t len_i = ((i + 1 < nseg1) ? starts[i + 1] : n) - starts[i];
                total1 += h1[i]; used_segs = i + 1;
                if (ret == 1) { end1 = (size_t)(starts[i] + h1[3 * capn + i]); break; }
                if (!(ret == 0 && h1[3 * capn + i] == len_i)) ok = false;
            }
            if (ok && end1 == 0) ok = false;                    // no final block
            if (ok && end1 + tlen0 > n) ok = false;
            if (ok && total1 <= cap) {
                r = grow(ctx, ctx->d_sout, ctx->sout_cap, total1 + 64, "cudaMalloc(stream out)");
                if (r) return r;
                uint64_t* d_goff = ctx->d_soff + 2 * ((size_t)nseg1 + 1);             // scanned offsets (used_segs + 1 entries)
                CK(launch_offsets(ctx->d_sres, used_segs, 0, d_goff, st), "offsets launch");
                CK(launch_gather(ctx->d_sslots, S, ctx->d_sres, d_goff, used_segs, ctx->d_sout, ctx->sms, st), "gather launch");
                if (wr.kind == 2) { int rc = zng_b200_crc32(ctx, ctx->d_sout, total1, 0, ctx->d_result + 1, st); if (rc) return rc; }
                if (wr.kind == 1) { int rc = zng_b200_adler32(ctx, ctx->d_sout, total1, 1, ctx->d_result + 1, st); if (rc) return rc; }
                CK(cudaMemcpyAsync(ctx->h_result + 1, ctx->d_result + 1, sizeof(uint32_t), cudaMemcpyDeviceToHost, st), "D2H check");
                if (total1) CK(cudaMemcpyAsync(h_out, ctx->d_sout, total1, cudaMemcpyDeviceToHost, st), "D2H output");
                CK(cudaStreamSynchronize(st), "sync");
                const uint32_t chk1 = wr.kind ? ctx->h_result[1] : 0u;
                *out_len = total1; *check = chk1; *detail = 0; *status = 1; *in_used = end1 + tlen0;
                const uint8_t* t = src + end1;
                if (wr.kind == 2) {
                    const uint32_t c = t[0] | (t[1] << 8) | (t[2] << 16) | ((uint32_t)t[3] << 24), l = t[4] | (t[5] << 8) | (t[6] << 16) | ((uint32_t)t[7] << 24);
                    if (c != chk1) { *status = ZNG_B200_DATA_ERROR; *detail = 17; }
                    else if (l != (uint32_t)total1) { *status = ZNG_B200_DATA_ERROR; *detail = 18; }
                } else if (wr.kind == 1) {
                    const uint32_t c = ((uint32_t)t[0] << 24) | (t[1] << 16) | (t[2] << 8) | t[3];
                    if (c != chk1) { *status = ZNG_B200_DATA_ERROR; *detail = 17; }
                }
                return 0;
            }
            if (ok && total1 > cap) {                            // Z_BUF_ERROR; the size needed is reported
                *status = ZNG_B200_BUF_ERROR; *detail = 0x100u | 0x400u; *out_len = total1; *in_used = 0;
                return 0;
            }
        }
    }
    // ---- size pass (count mode); segments whose end marker proves false are merged with their successor and re-run
    struct Seg { uint32_t size = 0, used = 0, det = 0; int32_t ret = -5; bool dirty = true; };
    std::vector<Seg> segs(starts.size());
    auto ensure_seg = [&](uint32_t cnt) -> int {
        if (ctx->seg_cap >= cnt) return 0;
        if (ctx->d_soff) { cudaDeviceSynchronize(); cudaFree(ctx->d_soff); cudaFree(ctx->d_sres); ctx->d_soff = nullptr; ctx->d_sres = nullptr; ctx->seg_cap = 0; }
        const uint32_t want = cnt + cnt / 4 + 1024;
        CK(cudaMalloc(&ctx->d_soff, 4 * ((size_t)want + 1) * sizeof(uint64_t)), "cudaMalloc(segment offsets)");
        CK(cudaMalloc(&ctx->d_sres, 5 * (size_t)want * sizeof(uint32_t)), "cudaMalloc(segment results)");
        ctx->seg_cap = want;
        return 0;
    };
    std::vector<uint64_t> hoff; std::vector<uint32_t> hres;
    size_t stream_end = 0; bool finished = false;
    for (int round = 0; round < 32 && !finished; round++) {
        std::vector<uint32_t> idx;
        for (uint32_t i = 0; i < segs.size(); i++) if (segs[i].dirty) idx.push_back(i);
        if (!idx.empty()) {
            const uint32_t cnt = (uint32_t)idx.size();
            r = ensure_seg(cnt);
            if (r) return r;
            hoff.assign(4 * (size_t)cnt, 0);                  // (begin, end) pairs: input ranges, then (unused) output ranges
            for (uint32_t k = 0; k < cnt; k++) {
                const uint32_t i = idx[k];
                hoff[2 * k] = starts[i];
                hoff[2 * k + 1] = (i + 1 < starts.size()) ? starts[i + 1] : n;
            }
            CK(cudaMemcpyAsync(ctx->d_soff, hoff.data(), hoff.size() * sizeof(uint64_t), cudaMemcpyHostToDevice, st), "H2D segment offsets");
            const size_t capn = ctx->seg_cap;
            const int slot = next_slot(ctx);
            CK(launch_inflate_members(ctx->d_sin, ctx->d_soff, cnt, -15, ctx->d_sin, ctx->d_soff + 2 * (size_t)cnt, ctx->d_sres, nullptr,
                                      (int32_t*)(ctx->d_sres + 2 * capn), ctx->d_sres + 3 * capn, ctx->d_sres + 4 * capn, ctx->counters + slot,
                                      ctx->sms, st, 1 | 2 | 4),
               "inflate size pass");
            hres.resize(5 * capn);
            CK(cudaMemcpyAsync(hres.data(), ctx->d_sres, 5 * capn * sizeof(uint32_t), cudaMemcpyDeviceToHost, st), "D2H segment results");
            CK(cudaStreamSynchronize(st), "sync");
            for (uint32_t k = 0; k < cnt; k++) {
                Seg& sg = segs[idx[k]];
                sg.size = hres[k]; sg.ret = (int32_t)hres[2 * capn + k]; sg.used = hres[3 * capn + k]; sg.det = hres[4 * capn + k];
                sg.dirty = false;
            }
        }
        // walk in order: clean ends continue, a final block ends the stream, a block cut by its end marker drops that marker
        std::vector<uint64_t> nstarts; std::vector<Seg> nsegs;
        bool merged = false, ended = false;
        for (uint32_t i = 0; i < segs.size() && !ended; i++) {
            Seg sg = segs[i];
            const uint64_t bgn = starts[i], end = (i + 1 < starts.size()) ? starts[i + 1] : n;
            if (sg.ret == 1) { nstarts.push_back(bgn); nsegs.push_back(sg); stream_end = (size_t)(bgn + sg.used); ended = true; }
            else if (sg.ret == 0 && sg.used == end - bgn) { nstarts.push_back(bgn); nsegs.push_back(sg); }
            else if (sg.ret == -5 && (sg.det & 0x200u) && i + 1 < starts.size()) {   // input ran out inside a block: false marker
                sg.dirty = true; merged = true;
                nstarts.push_back(bgn); nsegs.push_back(sg);
                i++;                                          // the next start is dropped
            } else return exact();                            // data error, no final block, ...: the exact path reports it
        }
        if (!ended && !merged) return exact();                // ran off the end of the input without a final block
        starts.swap(nstarts); segs.swap(nsegs);               // (segments behind a final block are cut off)
        finished = ended && !merged;                          // merged segments are re-sized in the next round
    }
    if (!finished) return exact();
    // ---- trailer present?
    const size_t tlen = wr.kind == 2 ? 8 : (wr.kind == 1 ? 4 : 0);
    if (stream_end + tlen > n) return exact();                // truncated trailer: Z_BUF_ERROR from the exact path
    // ---- offsets, capacity
    const uint32_t nseg = (uint32_t)segs.size();
    std::vector<uint64_t> obeg(nseg + 1, 0);
    for (uint32_t i = 0; i < nseg; i++) obeg[i + 1] = obeg[i] + segs[i].size;
    const size_t total = (size_t)obeg[nseg];
    if (total > cap) {                                        // Z_BUF_ERROR; the size needed is reported so the caller can come back
        *status = ZNG_B200_BUF_ERROR; *detail = 0x100u | 0x400u; *out_len = total; *in_used = 0;
        return 0;
    }
    r = grow(ctx, ctx->d_sout, ctx->sout_cap, total + 64, "cudaMalloc(stream out)");
    if (r) return r;
    r = ensure_seg(nseg + 1);
    if (r) return r;
    // ---- byte pass.  Segment i reads [starts[i], starts[i+1]) (the last one up to the end of its final block)
    hoff.assign(2 * ((size_t)nseg + 1), 0);
    for (uint32_t i = 0; i < nseg; i++) hoff[i] = starts[i];
    hoff[nseg] = stream_end;
    for (uint32_t i = 0; i + 1 < nseg; i++) if (starts[i + 1] < starts[i]) return exact();
    for (uint32_t i = 0; i <= nseg; i++) hoff[nseg + 1 + i] = obeg[i];
    CK(cudaMemcpyAsync(ctx->d_soff, hoff.data(), hoff.size() * sizeof(uint64_t), cudaMemcpyHostToDevice, st), "H2D segment offsets");
    {
        // groups of segments (~256 MiB of output each) on three streams: the D2H of one group overlaps the decode of the next
        if (!ctx->sready) {
            for (int i = 0; i < 3; i++) CK(cudaStreamCreateWithFlags(&ctx->sstream[i], cudaStreamNonBlocking), "cudaStreamCreate");
            CK(cudaEventCreateWithFlags(&ctx->sready, cudaEventDisableTiming), "cudaEventCreate");
        }
        CK(cudaEventRecord(ctx->sready, st), "event record");
        const size_t capn = ctx->seg_cap;
        int gi = 0;
        for (uint32_t g0 = 0; g0 < nseg;) {
            uint32_t g1 = g0;
            while (g1 < nseg && (g1 == g0 || obeg[g1 + 1] - obeg[g0] <= ((uint64_t)256 << 20))) g1++;
            cudaStream_t gs = ctx->sstream[gi % 3]; gi++;
            CK(cudaStreamWaitEvent(gs, ctx->sready, 0), "cudaStreamWaitEvent");
            const int slot = next_slot(ctx);
            CK(launch_inflate_members(ctx->d_sin, ctx->d_soff + g0, g1 - g0, -15, ctx->d_sout, ctx->d_soff + (nseg + 1) + g0, ctx->d_sres + g0, nullptr,
                                      (int32_t*)(ctx->d_sres + 2 * capn) + g0, ctx->d_sres + 3 * capn + g0, ctx->d_sres + 4 * capn + g0,
                                      ctx->counters + slot, ctx->sms, gs, 1),
               "inflate byte pass");
            const size_t ob = (size_t)obeg[g0], ol = (size_t)(obeg[g1] - obeg[g0]);
            if (ol) CK(cudaMemcpyAsync((uint8_t*)h_out + ob, ctx->d_sout + ob, ol, cudaMemcpyDeviceToHost, gs), "D2H output");
            g0 = g1;
        }
        for (int i = 0; i < 3; i++) CK(cudaStreamSynchronize(ctx->sstream[i]), "sync");
        hres.resize(5 * capn);
        CK(cudaMemcpyAsync(hres.data(), ctx->d_sres, 5 * capn * sizeof(uint32_t), cudaMemcpyDeviceToHost, st), "D2H segment results");
        // checksum of the whole output (K3)
        if (wr.kind == 2) { int rc = zng_b200_crc32(ctx, ctx->d_sout, total, 0, ctx->d_result + 1, st); if (rc) return rc; }
        if (wr.kind == 1) { int rc = zng_b200_adler32(ctx, ctx->d_sout, total, 1, ctx->d_result + 1, st); if (rc) return rc; }
        CK(cudaMemcpyAsync(ctx->h_result + 1, ctx->d_result + 1, sizeof(uint32_t), cudaMemcpyDeviceToHost, st), "D2H check");
        CK(cudaStreamSynchronize(st), "sync");
        for (uint32_t i = 0; i < nseg; i++) {
            const int32_t ret = (int32_t)hres[2 * capn + i];
            if (hres[i] != segs[i].size || ret != (i + 1 == nseg ? 1 : 0)) return exact();
        }
    }
    const uint32_t chk = wr.kind ? ctx->h_result[1] : 0u;
    *out_len = total; *check = chk; *detail = 0; *status = 1;
    *in_used = stream_end + tlen;
    const uint8_t* t = src + stream_end;
    if (wr.kind == 2) {
        const uint32_t c = t[0] | (t[1] << 8) | (t[2] << 16) | ((uint32_t)t[3] << 24), l = t[4] | (t[5] << 8) | (t[6] << 16) | ((uint32_t)t[7] << 24);
        if (c != chk) { *status = ZNG_B200_DATA_ERROR; *detail = 17; }
        else if (l != (uint32_t)total) { *status = ZNG_B200_DATA_ERROR; *detail = 18; }
    } else if (wr.kind == 1) {
        const uint32_t c = ((uint32_t)t[0] << 24) | (t[1] << 16) | (t[2] << 8) | t[3];
        if (c != chk) { *status = ZNG_B200_DATA_ERROR; *detail = 17; }
    }
    return 0;
}

// ---------------------------------------------------------------- resumable inflate of ONE stream fed piecewise
// What zng_inflate(Z_NO_FLUSH / Z_SYNC_FLUSH) of the host library calls.  The member decoder (K4) runs in RESUME mode: each feed decodes
// from the last block boundary the previous feed reached -- the output so far stays in a device buffer, where back-references
// reach it -- to the last block boundary inside what has arrived; the bytes decoded for good go to the caller at once and the
// input behind them is dropped.  Work is linear in the stream (a feed re-decodes at most the one block it could not finish),
// output arrives piecewise.  The check value (CRC-32 / Adler-32 of all output, K3) and the trailer are verified at the end.
// Streams this layer does not parse itself (FDICT, FHCRC, header errors) are reported as "not resumable": the caller keeps the
// one-shot path for them.
struct zng_b200_inflate_stream {
    zng_b200_ctx* ctx = nullptr;
    int window_bits = 15;
    Wrapper wr; bool header_done = false;
    std::vector<uint8_t> in;                      // input from the last boundary on (after the header)
    uint32_t start_bit = 0;
    uint8_t* d_in = nullptr; size_t d_in_cap = 0;
    uint8_t* d_out = nullptr; size_t d_out_cap = 0;
    uint64_t out_total = 0, delivered = 0;        // decoded for good / handed to the caller: offsets into d_out
    uint64_t base = 0;                            // stream offset of d_out[0]: delivered output older than the 32 KiB of history is dropped
    uint64_t compact_at = (uint64_t)64 << 20;     // ... once that much of it has piled up (env ZNG_B200_INFLATE_COMPACT, bytes)
    uint64_t backlog_max = (uint64_t)256 << 20;   // decoded but not yet taken by the caller: beyond this no more input is accepted (env ZNG_B200_INFLATE_BACKLOG)
    uint32_t run_check = 0;                       // CRC-32 / Adler-32 of the output up to out_total, carried from attempt to attempt
    bool body_done = false, finished = false;
    uint32_t check = 0;
    size_t last_try = 0;                          // retained bytes at the last attempt that found no new boundary
    uint32_t* d_io = nullptr;                     // offsets (u64 x 4) | results (u32 x 8) | resume_io (u32 x 4)
    uint32_t* h_io = nullptr;
};

int zng_b200_inflate_stream_open(zng_b200_ctx* ctx, int window_bits, zng_b200_inflate_stream** out) {
    if (!ctx || !out) return ZNG_B200_STREAM_ERROR;
    int wb = window_bits;
    if (wb < 0) { if (wb < -15) return bad(ctx, "windowBits out of range"); wb = -wb; } else if (wb < 48) wb &= 15;
    if (wb && (wb < 8 || wb > 15)) return bad(ctx, "windowBits out of range");
    DeviceGuard g(ctx->device);
    zng_b200_inflate_stream* st = new (std::nothrow) zng_b200_inflate_stream();
    if (!st) return ZNG_B200_MEM_ERROR;
    st->ctx = ctx; st->window_bits = window_bits;
    if (const char* e = getenv("ZNG_B200_INFLATE_COMPACT")) { const long long v = atoll(e); if (v >= 65536) st->compact_at = (uint64_t)v; }
    if (const char* e = getenv("ZNG_B200_INFLATE_BACKLOG")) { const long long v = atoll(e); if (v >= 65536) st->backlog_max = (uint64_t)v; }
    if (cudaMalloc(&st->d_io, 128) != cudaSuccess || cudaHostAlloc(&st->h_io, 128, cudaHostAllocDefault) != cudaSuccess) {
        cudaGetLastError(); cudaFree(st->d_io); delete st; return ZNG_B200_MEM_ERROR;
    }
    *out = st;
    return 0;
}

void zng_b200_inflate_stream_close(zng_b200_inflate_stream* st) {
    if (!st) return;
    DeviceGuard g(st->ctx->device);
    cudaDeviceSynchronize();
    cudaFree(st->d_in); cudaFree(st->d_out); cudaFree(st->d_io); cudaFreeHost(st->h_io);
    delete st;
}

// status: 0 = fed, more input (or more room for output) needed; 1 = stream end, every byte delivered; 2 = too much output is waiting
// for the caller (more than the backlog limit): some was handed over, NO input was taken, call again; -3 = data error (*detail =
// message id, zng_b200_inflate_msg); -6 (ZNG_B200_NOT_RESUMABLE) = a header this layer leaves to the one-shot path, nothing consumed.
// *in_used: bytes of THIS call's input that belong to the stream (less than n only at the stream's end); *out_len: bytes written to h_out.
int zng_b200_inflate_stream_feed(zng_b200_inflate_stream* st, const void* h_in, size_t n, void* h_out, size_t cap,
                                 size_t* in_used, size_t* out_len, int32_t* status, uint32_t* detail, uint32_t* check) {
    if (!st || !in_used || !out_len || !status || (n && !h_in) || (cap && !h_out)) return ZNG_B200_STREAM_ERROR;
    zng_b200_ctx* ctx = st->ctx;
    DeviceGuard g(ctx->device);
    *in_used = n; *out_len = 0; *status = 0;
    if (detail) *detail = 0;
    auto deliver = [&]() -> int {
        const uint64_t avail = st->out_total - st->delivered;
        const size_t give = (size_t)std::min<uint64_t>(avail, cap - *out_len);
        if (give) {
            CK(cudaMemcpy((uint8_t*)h_out + *out_len, st->d_out + st->delivered, give, cudaMemcpyDeviceToHost), "D2H output");
            st->delivered += give; *out_len += give;
        }
        return 0;
    };
    if (st->finished) { int r = deliver(); if (r) return r; *in_used = 0; *status = st->delivered == st->out_total ? 1 : 0; if (check) *check = st->check; return 0; }
    if (!st->body_done && st->out_total - st->delivered > st->backlog_max) {   // like the reference with avail_out == 0: hand over output, take no input
        int r = deliver(); if (r) return r;
        *in_used = 0; *status = 2;
        return 0;
    }
    if (n) st->in.insert(st->in.end(), (const uint8_t*)h_in, (const uint8_t*)h_in + n);
    if (!st->header_done) {
        const int wrap = st->window_bits < 0 ? 0 : (st->window_bits >> 4) + 5;
        if (wrap == 0) { st->wr.kind = 0; st->wr.body = 0; st->header_done = true; }
        else {
            // Nothing is taken before the header parses: the caller keeps the bytes and passes them again with more.  An incomplete
            // header waits (a gzip header with names can be long); what parse_wrapper refuses must go to the one-shot path.
            if (st->in.size() < 2) { st->in.clear(); *in_used = 0; return 0; }
            const bool gz = (wrap & 2) && st->in[0] == 0x1f && st->in[1] == 0x8b;
            if (parse_wrapper(st->in.data(), st->in.size(), st->window_bits, st->wr) != 0) {
                const bool maybe_short = gz && st->in.size() < 4096 && !(st->in.size() >= 4 && (st->in[2] != 8 || (st->in[3] & 0xe2)));
                st->in.clear(); *in_used = 0;
                if (!maybe_short) *status = ZNG_B200_NOT_RESUMABLE;
                return 0;
            }
            st->in.erase(st->in.begin(), st->in.begin() + (ptrdiff_t)st->wr.body);
            st->header_done = true;
            st->run_check = st->wr.kind == 1 ? 1u : 0u;
        }
    }
    // the check value rides along: bytes [out_total, upto) join it as soon as they are final (a block boundary has been passed)
    auto extend_check = [&](uint64_t upto) -> int {
        if (st->wr.kind && upto > st->out_total) {
            const int rr = st->wr.kind == 2 ? zng_b200_crc32(ctx, st->d_out + st->out_total, (size_t)(upto - st->out_total), st->run_check, ctx->d_result, nullptr)
                                            : zng_b200_adler32(ctx, st->d_out + st->out_total, (size_t)(upto - st->out_total), st->run_check, ctx->d_result, nullptr);
            if (rr) return rr;
            CK(cudaMemcpy(&st->run_check, ctx->d_result, 4, cudaMemcpyDeviceToHost), "D2H check");
        }
        st->out_total = upto;
        return 0;
    };
    // drop delivered output that is no longer history (back-references reach 32 KiB): memory stays bounded, offsets stay 32-bit
    {
        const uint64_t hist = st->out_total > 32768u ? st->out_total - 32768u : 0u;
        const uint64_t from = st->delivered < hist ? st->delivered : hist;
        if (from >= st->compact_at && from >= st->out_total - from) {          // (source and destination do not overlap)
            if (st->out_total > from) CK(cudaMemcpy(st->d_out, st->d_out + from, (size_t)(st->out_total - from), cudaMemcpyDeviceToDevice), "compact history");
            st->base += from; st->delivered -= from; st->out_total -= from;
        }
    }
    // ---- decode from the last boundary, as far as the input goes
    // Every call that brings input is an attempt: the bytes after the stream's end must stay with the caller of THIS call (zlib's
    // next_in / total_in contract), so input is never parked undecoded.  An attempt that crossed no block boundary decodes the
    // open block again next time: the cost is bounded by (block size / piece size) passes over one block, not over the stream.
    while (!st->body_done && !st->in.empty() && (st->last_try == 0 || st->in.size() > st->last_try)) {
        const size_t nin = st->in.size();
        if (nin > 0xfffffff0ull || st->out_total > 0xe0000000ull) return bad(ctx, "resumable inflate: stream too large for one member decoder");
        if (st->d_in_cap < nin + 64) {
            if (st->d_in) cudaFree(st->d_in);
            st->d_in = nullptr; st->d_in_cap = 0;
            size_t want = nin + nin / 2 + (1 << 20);
            CK(cudaMalloc(&st->d_in, want), "cudaMalloc(resumable in)");
            st->d_in_cap = want;
        }
        size_t want_out = (size_t)st->out_total + std::max<size_t>(4 * nin, (size_t)4 << 20) + 65536;
        if (want_out > 0xfffffff0ull) want_out = 0xfffffff0ull;
        if (st->d_out_cap < want_out) {
            uint8_t* nb = nullptr;
            want_out += want_out / 2;
            if (want_out > 0xfffffff0ull) want_out = 0xfffffff0ull;
            CK(cudaMalloc(&nb, want_out), "cudaMalloc(resumable out)");
            if (st->d_out && st->out_total) CK(cudaMemcpy(nb, st->d_out, (size_t)st->out_total, cudaMemcpyDeviceToDevice), "copy history");
            cudaFree(st->d_out);
            st->d_out = nb; st->d_out_cap = want_out;
        }
        CK(cudaMemcpyAsync(st->d_in, st->in.data(), nin, cudaMemcpyHostToDevice, 0), "H2D");
        CK(cudaMemsetAsync(st->d_in + nin, 0, 64, 0), "memset");
        uint64_t* ho = reinterpret_cast<uint64_t*>(st->h_io);                 // [0..1] in offsets, [2..3] out offsets
        uint32_t* hr = st->h_io + 8;                                            // [0..4] results, [8..11] resume_io
        ho[0] = 0; ho[1] = nin; ho[2] = 0; ho[3] = st->d_out_cap;
        hr[8] = st->start_bit; hr[9] = (uint32_t)st->out_total; hr[10] = 0; hr[11] = 0;
        CK(cudaMemcpyAsync(st->d_io, st->h_io, 128, cudaMemcpyHostToDevice, 0), "H2D io");
        uint64_t* d_o = reinterpret_cast<uint64_t*>(st->d_io);
        uint32_t* d_r = st->d_io + 8;
        const int slot = next_slot(ctx);
        CK(launch_inflate_members(st->d_in, d_o, 1, -15, st->d_out, d_o + 2, d_r + 0, d_r + 1, (int32_t*)(d_r + 2), d_r + 3, d_r + 4,
                                  ctx->counters + slot, ctx->sms, 0, 8, d_r + 8), "inflate launch (resume)");
        CK(cudaMemcpyAsync(st->h_io, st->d_io, 128, cudaMemcpyDeviceToHost, 0), "D2H io");
        CK(cudaStreamSynchronize(0), "sync");
        const uint32_t o_len = hr[0]; const int32_t ret = (int32_t)hr[2]; const uint32_t used = hr[3], det = hr[4];
        const uint32_t bb_byte = hr[8], bb_bit = hr[9], bb_out = hr[10];
        if (ret == ZNG_B200_DATA_ERROR) { *status = ZNG_B200_DATA_ERROR; if (detail) *detail = det & 0xffu; int r = extend_check(bb_out); if (r) return r; return deliver(); }
        if (ret == 1) {                                                         // the final block is decoded
            { int r = extend_check(o_len); if (r) return r; }
            st->body_done = true;
            st->in.erase(st->in.begin(), st->in.begin() + (ptrdiff_t)std::min<size_t>(used, st->in.size()));
            st->start_bit = 0; st->last_try = 0;
            break;
        }
        const bool out_full = (det & 0x100u) && !(det & 0x200u);
        if (out_full) {                                                          // the output buffer filled up: double it and go on
            size_t want = st->d_out_cap * 2;
            if (want > 0xfffffff0ull) { if (st->d_out_cap >= 0xfffffff0ull) return bad(ctx, "resumable inflate: output beyond 4 GiB"); want = 0xfffffff0ull; }
            uint8_t* nb = nullptr;
            CK(cudaMalloc(&nb, want), "cudaMalloc(resumable out)");
            CK(cudaMemcpy(nb, st->d_out, (size_t)bb_out, cudaMemcpyDeviceToDevice), "copy history");
            cudaFree(st->d_out); st->d_out = nb; st->d_out_cap = want;
        }
        { int r = extend_check(bb_out); if (r) return r; }
        if (bb_byte) st->in.erase(st->in.begin(), st->in.begin() + (ptrdiff_t)bb_byte);
        st->start_bit = bb_bit;
        if (out_full) { st->last_try = 0; continue; }
        st->last_try = st->in.size();                                            // all of it has been tried: wait for more input
        break;                                                                   // ran out of input: wait for the next feed
    }
    int r = deliver();
    if (r) return r;
    if (st->body_done && !st->finished) {
        const size_t tlen = st->wr.kind == 2 ? 8 : (st->wr.kind == 1 ? 4 : 0);
        if (st->in.size() >= tlen) {
            const uint32_t chk = st->wr.kind ? st->run_check : 0u;
            st->check = chk;
            const uint8_t* t = st->in.data();
            if (st->wr.kind == 2) {
                const uint32_t c = t[0] | (t[1] << 8) | (t[2] << 16) | ((uint32_t)t[3] << 24), l = t[4] | (t[5] << 8) | (t[6] << 16) | ((uint32_t)t[7] << 24);
                if (c != chk) { *status = ZNG_B200_DATA_ERROR; if (detail) *detail = 17; return 0; }
                if (l != (uint32_t)(st->base + st->out_total)) { *status = ZNG_B200_DATA_ERROR; if (detail) *detail = 18; return 0; }
            } else if (st->wr.kind == 1) {
                const uint32_t c = ((uint32_t)t[0] << 24) | (t[1] << 16) | (t[2] << 8) | t[3];
                if (c != chk) { *status = ZNG_B200_DATA_ERROR; if (detail) *detail = 17; return 0; }
            }
            st->in.erase(st->in.begin(), st->in.begin() + (ptrdiff_t)tlen);
            st->finished = true;
            // what is left in `in` follows the stream: it came with this call (earlier calls ended inside the stream)
            const size_t left = st->in.size();
            *in_used = left <= n ? n - left : 0;
            st->in.clear();
        }
    }
    if (st->finished) { if (check) *check = st->check; *status = st->delivered == st->out_total ? 1 : 0; }
    return 0;
}

// ---------------------------------------------------------------- host-buffer entry points
static int host_checksum(zng_b200_ctx* ctx, const void* h_buf, size_t n, uint32_t init, uint32_t* result, bool crc, void* h_copy = nullptr) {
    if (!ctx || !result) return ZNG_B200_STREAM_ERROR;
    if (n && !h_buf) return bad(ctx, "h_buf is NULL");
    DeviceGuard g(ctx->device);
    // slabs of <= 256 MiB through one staging buffer; partial results chain through `init`
    const size_t slab = (size_t)256 << 20;
    if (ctx->hostbuf_cap < (n < slab ? n : slab)) {
        if (ctx->d_hostbuf) cudaFree(ctx->d_hostbuf);
        ctx->d_hostbuf = nullptr; ctx->hostbuf_cap = 0;
        size_t want = n < slab ? n : slab;
        if (want < (1u << 20)) want = 1u << 20;
        CK(cudaMalloc(&ctx->d_hostbuf, want), "cudaMalloc(host staging)");
        ctx->hostbuf_cap = want;
    }
    uint32_t cur = init;
    size_t off = 0;
    do {
        const size_t take = (n - off) < slab ? (n - off) : slab;
        if (take) CK(cudaMemcpyAsync(ctx->d_hostbuf, (const uint8_t*)h_buf + off, take, cudaMemcpyHostToDevice, 0), "H2D");
        int r = crc ? zng_b200_crc32(ctx, ctx->d_hostbuf, take, cur, ctx->d_result, nullptr)
                    : zng_b200_adler32(ctx, ctx->d_hostbuf, take, cur, ctx->d_result, nullptr);
        if (r) return r;
        CK(cudaMemcpyAsync(ctx->h_result, ctx->d_result, sizeof(uint32_t), cudaMemcpyDeviceToHost, 0), "D2H");
        // checksum-while-copy (read_buf, deflate.c:1190-1212): the copy is the bytes' trip back from the staging buffer
        if (h_copy && take) CK(cudaMemcpyAsync((uint8_t*)h_copy + off, ctx->d_hostbuf, take, cudaMemcpyDeviceToHost, 0), "D2H copy");
        CK(cudaStreamSynchronize(0), "sync");
        cur = ctx->h_result[0];
        off += take;
    } while (off < n);
    *result = cur;
    return 0;
}

int zng_b200_crc32_host(zng_b200_ctx* ctx, const void* h_buf, size_t n, uint32_t init, uint32_t* result) {
    return host_checksum(ctx, h_buf, n, init, result, true);
}
int zng_b200_adler32_host(zng_b200_ctx* ctx, const void* h_buf, size_t n, uint32_t init, uint32_t* result) {
    return host_checksum(ctx, h_buf, n, init, result, false);
}
int zng_b200_crc32_copy_host(zng_b200_ctx* ctx, void* h_dst, const void* h_src, size_t n, uint32_t init, uint32_t* result) {
    if (n && !h_dst) return ZNG_B200_STREAM_ERROR;
    return host_checksum(ctx, h_src, n, init, result, true, h_dst);
}
int zng_b200_adler32_copy_host(zng_b200_ctx* ctx, void* h_dst, const void* h_src, size_t n, uint32_t init, uint32_t* result) {
    if (n && !h_dst) return ZNG_B200_STREAM_ERROR;
    return host_checksum(ctx, h_src, n, init, result, false, h_dst);
}

// drain one slab: wait for its kernels, copy the packed bytes out
static int drain_slab(zng_b200_ctx* ctx, Slab& s, uint8_t* h_out, size_t out_cap, size_t& out_pos,
                      uint32_t& crc, uint32_t& adler) {
    if (!s.busy) return 0;
    s.busy = false;
    CK(cudaEventSynchronize(s.done), "cudaEventSynchronize");
    const size_t total = (size_t)s.h_meta[0];
    const uint32_t scrc = (uint32_t)s.h_meta[1], sadler = (uint32_t)(s.h_meta[1] >> 32);
    if (out_pos + total > out_cap) { snprintf(ctx->err, sizeof(ctx->err), "output buffer too small"); return ZNG_B200_BUF_ERROR; }
    CK(cudaMemcpyAsync(h_out + out_pos, s.d_packed, total, cudaMemcpyDeviceToHost, s.stream), "D2H packed");
    out_pos += total;
    crc = crc32_combine_dev(ctx->x2n, crc, scrc, s.in_bytes);
    adler = adler32_combine_dev(adler, sadler, s.in_bytes);
    return 0;
}

static int sync_slabs(zng_b200_ctx* ctx) {
    for (int i = 0; i < ctx->pipe; i++) {
        ctx->slab[i].busy = false;
        if (ctx->slab[i].stream) cudaStreamSynchronize(ctx->slab[i].stream);
    }
    return 0;
}

extern "C++" {
template <typename T>
static int grow(zng_b200_ctx* ctx, T*& p, size_t& cap, size_t want, const char* what) {
    if (cap >= want) return 0;
    if (p) { cudaDeviceSynchronize(); cudaFree(p); p = nullptr; cap = 0; }
    want += want / 4;
    CK(cudaMalloc(&p, want * sizeof(T)), what);
    cap = want;
    return 0;
}
}

// pigz's dependent mode from host buffers (what zng_deflate does after zng_deflateSetDictionary): every 65536-byte piece is
// compressed by a fresh level-1 stream primed with the 32768 bytes in front of it -- h_dict for the first piece (NULL: the
// first piece has no dictionary).  Pieces end with the sync-flush marker; final: Z_FINISH on the last one.
int zng_b200_deflate_host_primed(zng_b200_ctx* ctx, const void* h_dict, const void* h_in, size_t n, int final,
                                 void* h_out, size_t out_cap, size_t* out_len, uint32_t* crc32, uint32_t* adler32) {
    return zng_b200_deflate_host_primed_level(ctx, h_dict, h_in, n, 1, final, h_out, out_cap, out_len, crc32, adler32);
}

int zng_b200_deflate_host_primed_level(zng_b200_ctx* ctx, const void* h_dict, const void* h_in, size_t n, int level, int final,
                                       void* h_out, size_t out_cap, size_t* out_len, uint32_t* crc32, uint32_t* adler32) {
    if (!ctx) return ZNG_B200_STREAM_ERROR;
    if (level < 1 || level > 6) return bad(ctx, "primed chunks: levels 1..6");
    if (!out_len || (n && !h_in) || !h_out) return bad(ctx, "NULL argument");
    if (n > ((size_t)1 << 32)) return bad(ctx, "primed host call: at most 4 GiB per call");
    DeviceGuard g(ctx->device);
    const uint32_t chunk = ZNG_B200_CHUNK_MAX;
    const uint32_t nin = (uint32_t)((n + chunk - 1) / chunk);
    const uint32_t nch = nin ? nin : (final ? 1u : 0u);                  // no input + final: one empty Z_FINISH piece ("03 00")
    if (nch == 0) { *out_len = 0; if (crc32) *crc32 = 0; if (adler32) *adler32 = 1; return 0; }
    const size_t stride = zng_b200_deflate_bound(chunk);
    int r = grow(ctx, ctx->d_pbuf, ctx->pbuf_cap, (size_t)kWSize + n + 4096, "cudaMalloc(primed input)");
    if (r) return r;
    r = grow(ctx, ctx->d_pout, ctx->pout_cap, 2 * (size_t)nch * stride + 4096, "cudaMalloc(primed output)");
    if (r) return r;
    r = grow(ctx, ctx->d_pmeta, ctx->pmeta_cap, 3 * (size_t)nch + 2 * ((size_t)nch + 1) + 16, "cudaMalloc(primed meta)");
    if (r) return r;
    cudaStream_t st = nullptr;
    uint8_t* d_in = ctx->d_pbuf + kWSize;
    uint8_t* d_slots = ctx->d_pout; uint8_t* d_packed = ctx->d_pout + (size_t)nch * stride;
    uint32_t* d_sizes = ctx->d_pmeta; uint32_t* d_crcs = d_sizes + nch; uint32_t* d_adlers = d_crcs + nch;
    uint64_t* d_off = reinterpret_cast<uint64_t*>(ctx->d_pmeta + ((3 * (size_t)nch + 1) & ~(size_t)1));
    if (h_dict) CK(cudaMemcpyAsync(ctx->d_pbuf, h_dict, kWSize, cudaMemcpyHostToDevice, st), "H2D dictionary");
    if (n) CK(cudaMemcpyAsync(d_in, h_in, n, cudaMemcpyHostToDevice, st), "H2D");
    const uint32_t first0 = h_dict ? 1u : 0u;
    const uint32_t body = final ? nch - 1 : nch;
    const size_t body_bytes = final ? (size_t)body * chunk : n;
    if (body) {
        r = run_deflate_primed(ctx, d_in, body_bytes, chunk, body, 0, d_slots, stride, d_sizes, d_crcs, d_adlers, st, first0, level);
        if (r) return r;
    }
    if (final) {                                             // the Z_FINISH piece (possibly empty: "03 00")
        r = run_deflate_primed(ctx, d_in + body_bytes, n - body_bytes, chunk, 1, 1, d_slots + (size_t)body * stride, stride, d_sizes + body,
                               d_crcs + body, d_adlers + body, st, first0 + body, level);
        if (r) return r;
    }
    CK(launch_offsets(d_sizes, nch, 0, d_off, st), "offsets launch");
    CK(launch_gather(d_slots, stride, d_sizes, d_off, nch, d_packed, ctx->sms, st), "gather launch");
    CK(launch_crc32_fold(d_crcs, nin, chunk, n, 0, ctx->d_result, st), "crc fold");
    CK(launch_adler32_fold(d_adlers, nin, chunk, n, 1, ctx->d_result + 1, st), "adler fold");
    uint64_t total = 0; uint32_t res[2] = {0, 1};
    CK(cudaMemcpyAsync(&total, d_off + nch, sizeof(uint64_t), cudaMemcpyDeviceToHost, st), "D2H total");
    CK(cudaMemcpyAsync(res, ctx->d_result, sizeof(res), cudaMemcpyDeviceToHost, st), "D2H checks");
    CK(cudaStreamSynchronize(st), "sync");
    if (total > out_cap) return ZNG_B200_BUF_ERROR;
    CK(cudaMemcpy(h_out, d_packed, (size_t)total, cudaMemcpyDeviceToHost), "D2H packed");
    *out_len = (size_t)total;
    if (crc32) *crc32 = n ? res[0] : 0;
    if (adler32) *adler32 = n ? res[1] : 1;
    return 0;
}

extern "C++" {
static int ensure_streamed(zng_b200_ctx* ctx, uint32_t nch) {
    Streamed& S = ctx->st;
    if (!S.ready) {
        int least = 0, greatest = 0;
        CK(cudaDeviceGetStreamPriorityRange(&least, &greatest), "stream priorities");
        CK(cudaStreamCreateWithPriority(&S.copy, cudaStreamNonBlocking, greatest), "stream");
        CK(cudaStreamCreateWithPriority(&S.parse, cudaStreamNonBlocking, least), "stream");
        CK(cudaStreamCreateWithPriority(&S.d2h, cudaStreamNonBlocking, greatest), "stream");
        for (int i = 0; i < kStreamOut; i++) CK(cudaStreamCreateWithPriority(&S.out[i], cudaStreamNonBlocking, greatest), "stream");
        CK(cudaEventCreateWithFlags(&S.reset_done, cudaEventDisableTiming), "event");
        for (auto& e : S.slab_done) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming), "event");
        CK(cudaMalloc(&S.d_sync, 128 * sizeof(uint32_t)), "cudaMalloc(stream sync)");
        CK(cudaMalloc(&S.d_res, 2 * (kStreamMaxChunks / kStreamMinSlab) * sizeof(uint32_t)), "cudaMalloc(stream res)");
        CK(cudaHostAlloc(&S.h_ready, (kStreamMaxChunks / kStreamPiece + 1) * sizeof(uint32_t), cudaHostAllocDefault), "cudaHostAlloc");
        CK(cudaHostAlloc(&S.h_meta, 2 * (kStreamMaxChunks / kStreamMinSlab) * sizeof(uint64_t), cudaHostAllocDefault), "cudaHostAlloc");
        CK(cudaHostAlloc(&S.h_failed, sizeof(uint32_t), cudaHostAllocDefault), "cudaHostAlloc");
        CK(cudaHostAlloc(&S.h_done, (kStreamMaxChunks / kStreamMinSlab) * sizeof(uint32_t), cudaHostAllocMapped), "cudaHostAlloc(mapped)");
        CK(cudaHostGetDevicePointer(&S.d_h_done, S.h_done, 0), "cudaHostGetDevicePointer");
        S.ready = true;
    }
    if (S.cap_chunks >= nch) return 0;
    cudaDeviceSynchronize();
    for (void* p : {(void*)S.d_in_alloc, (void*)S.d_slots, (void*)S.d_packed, (void*)S.tokens, (void*)S.ntok, (void*)S.d_meta, (void*)S.d_offsets}) if (p) cudaFree(p);
    S.d_in_alloc = S.d_in = S.d_slots = S.d_packed = nullptr; S.tokens = S.ntok = S.d_meta = nullptr; S.d_offsets = nullptr; S.cap_chunks = 0;
    const size_t stride = zng_b200_deflate_bound(ZNG_B200_CHUNK_MAX), tstride = (ZNG_B200_CHUNK_MAX + 32u) & ~31u;
    const size_t nslabs = (nch + kStreamMinSlab - 1) / kStreamMinSlab;
    CK(cudaMalloc(&S.d_in_alloc, (size_t)nch * ZNG_B200_CHUNK_MAX + kWSize + 4096), "cudaMalloc(stream in)");
    S.d_in = S.d_in_alloc + kWSize;                           // 32 KiB in front: the stream bytes that precede this call's first chunk (levels 2+)
    CK(cudaMemset(S.d_in + (size_t)nch * ZNG_B200_CHUNK_MAX, 0, 4096), "cudaMemset(stream pad)");
    CK(cudaMalloc(&S.d_slots, (size_t)nch * stride), "cudaMalloc(stream slots)");
    CK(cudaMalloc(&S.d_packed, (size_t)nch * stride), "cudaMalloc(stream packed)");
    CK(cudaMalloc(&S.tokens, (size_t)nch * tstride * sizeof(uint32_t)), "cudaMalloc(stream tokens)");
    CK(cudaMalloc(&S.ntok, (size_t)nch * sizeof(uint32_t)), "cudaMalloc(stream ntok)");
    CK(cudaMalloc(&S.d_meta, (size_t)nch * 3 * sizeof(uint32_t)), "cudaMalloc(stream meta)");
    CK(cudaMalloc(&S.d_offsets, ((size_t)nch + nslabs + 1) * sizeof(uint64_t)), "cudaMalloc(stream offsets)");
    S.cap_chunks = nch;
    return 0;
}

// one call of the streamed path: n <= kStreamMaxChunks * 65536 bytes, n > 0
static int deflate_host_streamed(zng_b200_ctx* ctx, const uint8_t* h_in, size_t n, int final, uint8_t* h_out, size_t out_cap,
                                 size_t& out_pos, uint32_t& crc, uint32_t& adler, int level, int have_prev) {
    const uint32_t chunk = ZNG_B200_CHUNK_MAX;
    const uint32_t nch = (uint32_t)((n + chunk - 1) / chunk);
    const uint32_t kStreamSlabShift = ctx->stream_shift, kStreamSlab = 1u << kStreamSlabShift;
    const uint32_t nslabs = (nch + kStreamSlab - 1) / kStreamSlab, npieces = (nch + kStreamPiece - 1) / kStreamPiece;
    int r = ensure_heads(ctx);
    if (r) return r;
    if (level >= 2) { r = ensure_prevs(ctx); if (r) return r; }
    r = ensure_streamed(ctx, nch < 4096u ? 4096u : nch);
    if (r) return r;
    Streamed& S = ctx->st;
    static const int k2_dyn = [] { const char* e = getenv("ZNG_B200_STREAM_DYNSMEM_L2"); return e ? atoi(e) : 4096; }();
    static const int k2_carve = [] { const char* e = getenv("ZNG_B200_CO_CARVE_L2"); return e ? atoi(e) : 44; }();
    const size_t stride = ctx->slab_stride, tstride = (chunk + 32u) & ~31u;
    uint32_t* d_sizes = S.d_meta; uint32_t* d_crcs = S.d_meta + nch; uint32_t* d_adlers = S.d_meta + 2 * (size_t)nch;
    const long long patience = 20000000000ll;                 // ~10 s of SM clocks: a wait that long means something is broken
    static const bool trace = getenv("ZNG_B200_TRACE") != nullptr;
    std::vector<cudaEvent_t> tev;
    auto mark = [&](cudaStream_t st) { if (trace) { cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, st); tev.push_back(e); } };
    CK(cudaMemsetAsync(S.d_sync, 0, 128 * sizeof(uint32_t), S.parse), "memset sync");
    mark(S.parse);                                                       // [0] start
    CK(cudaEventRecord(S.reset_done, S.parse), "event");
    CK(cudaStreamWaitEvent(S.copy, S.reset_done, 0), "wait");
    std::vector<cudaEvent_t> d2h_ev;
    for (uint32_t j = 0; j < nslabs; j++) S.h_done[j] = 0;
    StreamSync sy; sy.ready = S.d_sync; sy.failed = S.d_sync + 1; sy.done = S.d_sync + 4; sy.done_shift = kStreamSlabShift; sy.patience = patience;
    sy.host_done = S.d_h_done;
    if (level == 1 && ctx->k1_cta == 1)                                  // only on explicit request: the streamed path is tuned around v3's occupancy
        CK(launch_quick_parse_cta(S.d_in, n, chunk, nch, S.tokens, (uint32_t)tstride, S.ntok, S.d_sync + 2, ctx->heads, ctx->sm_slots,
                                  ctx->sms, ctx->chains_cta, ctx->warps_cta, nullptr, S.parse, &sy, ctx->k1_stats),
           "quick_parse_cta launch");
    else if (level == 1)
        CK(launch_quick_parse(S.d_in, n, chunk, nch, S.tokens, (uint32_t)tstride, S.ntok, S.d_sync + 2, ctx->heads, ctx->sm_slots,
                              deflate_quick_grid(nch, ctx->sms, ctx->chains_per_sm), nullptr, S.parse, &sy),
           "quick_parse launch");
    else                                                                 // 28 chains per SM: 8 CTAs would hold every register of the SM
        CK(launch_fast_parse(S.d_in, n, chunk, nch, S.tokens, (uint32_t)tstride, S.ntok, S.d_sync + 2, ctx->heads, ctx->prevs, ctx->vtails,
                             ctx->sm_slots, ctx->sms, ctx->chains_per_sm_l2 < 28 ? ctx->chains_per_sm_l2 : 28, have_prev, level, S.parse, &sy, k2_dyn),
           "fast_parse launch");
    mark(S.parse);                                                       // [1] parse kernel done
    if (have_prev) CK(cudaMemcpyAsync(S.d_in - kWSize, h_in - kWSize, kWSize, cudaMemcpyHostToDevice, S.copy), "H2D window in front");
    // the copy engine delivers the input piece by piece; chunk ci may start once chunk ci + 1 is there as well (its
    // read-ahead reaches a few hundred bytes into the next chunk, and nothing may be cached before it has arrived)
    for (uint32_t i = 0; i < npieces; i++) {
        const size_t off = (size_t)i * kStreamPiece * chunk;
        const size_t len = (n - off) < (size_t)kStreamPiece * chunk ? (n - off) : (size_t)kStreamPiece * chunk;
        CK(cudaMemcpyAsync(S.d_in + off, h_in + off, len, cudaMemcpyHostToDevice, S.copy), "H2D piece");
        const uint32_t delivered = (i + 1 == npieces) ? nch : (i + 1) * kStreamPiece;
        S.h_ready[i] = (i + 1 == npieces) ? nch : delivered - 1u;
        CK(cudaMemcpyAsync(S.d_sync, &S.h_ready[i], sizeof(uint32_t), cudaMemcpyHostToDevice, S.copy), "H2D ready");
    }
    mark(S.copy);                                                        // [2] last piece delivered
    // The host drives the rest: it launches a slab's emit / checksum / gather work when the parse kernel reports the slab
    // complete, and copies a slab's packed bytes out when that work is done.  Nothing is ever queued behind a wait, so
    // streams that happen to share a hardware queue (CUDA_DEVICE_MAX_CONNECTIONS) cannot hold each other up.
    constexpr int kCoCarve = 44;                                         // the split the streamed parse kernel runs with
    int rc = 0;
    uint32_t next_launch = 0, next_drain = 0;
    auto t_progress = std::chrono::steady_clock::now();                  // reset whenever a slab is launched or drained
    while (next_drain < nslabs) {
        bool progressed = false;
        if (next_launch < nslabs) {
            const uint32_t j = next_launch;
            const uint32_t c0 = j * kStreamSlab, nb = (nch - c0) < kStreamSlab ? (nch - c0) : kStreamSlab;
            if (*(volatile uint32_t*)&S.h_done[j] == nb) {
                cudaStream_t st = S.out[j % kStreamOut];
                const size_t off = (size_t)c0 * chunk, bytes = (c0 + nb == nch) ? n - off : (size_t)nb * chunk;
                mark(st);                                                // [3+2j] slab j parsed
                auto emit = [&](uint32_t first, uint32_t count, size_t nbytes, int last) -> int {     // chunks [c0 + first, c0 + first + count) of the slab
                    const uint32_t* tk = S.tokens + (size_t)(c0 + first) * tstride;
                    if (level == 1)
                        CK(launch_static_emit(tk, (uint32_t)tstride, S.ntok + c0 + first, nbytes, chunk, count, last, S.d_slots + (size_t)(c0 + first) * stride,
                                              stride, d_sizes + c0 + first, ctx->sms, st, kCoCarve), "static_emit launch");
                    else
                        CK(launch_block_emit(S.d_in + off + (size_t)first * chunk, tk, (uint32_t)tstride, S.ntok + c0 + first, nbytes, chunk, count, last,
                                             S.d_slots + (size_t)(c0 + first) * stride, stride, d_sizes + c0 + first, ctx->sms, st, k2_carve, nullptr,
                                             ctx->counters + next_slot(ctx)), "block_emit launch");
                    return 0;
                };
                if (final && c0 + nb == nch) {                           // the stream's last chunk is the Z_FINISH chunk
                    const uint32_t body = nb - 1;
                    if (body) { r = emit(0, body, (size_t)body * chunk, 0); if (r) return r; }
                    r = emit(body, 1, bytes - (size_t)body * chunk, 1);
                    if (r) return r;
                } else {
                    r = emit(0, nb, bytes, 0);
                    if (r) return r;
                }
                CK(launch_checksum_tiles(S.d_in + off, bytes, chunk, nb, d_crcs + c0, d_adlers + c0, ctx->sms, st, 8), "checksum launch");
                uint64_t* offs = S.d_offsets + (size_t)j * (kStreamSlab + 1);
                CK(launch_offsets(d_sizes + c0, nb, 0, offs, st), "offsets launch");
                CK(launch_gather(S.d_slots + (size_t)c0 * stride, stride, d_sizes + c0, offs, nb, S.d_packed + (size_t)c0 * stride, ctx->sms, st), "gather launch");
                CK(launch_crc32_fold(d_crcs + c0, nb, chunk, bytes, 0, S.d_res + 2 * j, st), "crc fold");
                CK(launch_adler32_fold(d_adlers + c0, nb, chunk, bytes, 1, S.d_res + 2 * j + 1, st), "adler fold");
                CK(cudaMemcpyAsync(&S.h_meta[2 * j], offs + nb, sizeof(uint64_t), cudaMemcpyDeviceToHost, st), "D2H total");
                CK(cudaMemcpyAsync(&S.h_meta[2 * j + 1], S.d_res + 2 * j, sizeof(uint64_t), cudaMemcpyDeviceToHost, st), "D2H checks");
                CK(cudaEventRecord(S.slab_done[j], st), "event record");
                mark(st);                                                // [4+2j] slab j packed
                next_launch++; progressed = true;
            }
        }
        if (next_drain < next_launch && cudaEventQuery(S.slab_done[next_drain]) == cudaSuccess) {
            const uint32_t j = next_drain;
            const uint32_t c0 = j * kStreamSlab, nb = (nch - c0) < kStreamSlab ? (nch - c0) : kStreamSlab;
            const size_t bytes = (c0 + nb == nch) ? n - (size_t)c0 * chunk : (size_t)nb * chunk;
            const size_t total = (size_t)S.h_meta[2 * j];
            if (!rc && out_pos + total > out_cap) { snprintf(ctx->err, sizeof(ctx->err), "output buffer too small"); rc = ZNG_B200_BUF_ERROR; }
            if (!rc) {
                CK(cudaMemcpyAsync(h_out + out_pos, S.d_packed + (size_t)c0 * stride, total, cudaMemcpyDeviceToHost, S.d2h), "D2H packed");
                if (trace) { cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, S.d2h); d2h_ev.push_back(e); }
                out_pos += total;
                crc = crc32_combine_dev(ctx->x2n, crc, (uint32_t)S.h_meta[2 * j + 1], bytes);
                adler = adler32_combine_dev(adler, (uint32_t)(S.h_meta[2 * j + 1] >> 32), bytes);
            }
            next_drain++; progressed = true;
        }
        if (progressed) t_progress = std::chrono::steady_clock::now();
        if (!progressed) {
            if (std::chrono::steady_clock::now() - t_progress > std::chrono::seconds(30)) {
                snprintf(ctx->err, sizeof(ctx->err), "streamed pipeline: no progress for 30 s");
                cudaDeviceSynchronize();
                return ZNG_B200_CUDA_ERROR;
            }
            if (cudaStreamQuery(S.parse) == cudaSuccess && next_launch < nslabs &&
                *(volatile uint32_t*)&S.h_done[next_launch] == 0u && cudaStreamQuery(S.parse) == cudaSuccess) {
                // the parse kernel is gone and never reported this slab: a device-side wait gave up
                cudaDeviceSynchronize();
                if (*(volatile uint32_t*)&S.h_done[next_launch] == 0u) {
                    snprintf(ctx->err, sizeof(ctx->err), "streamed pipeline: the parse kernel ended early");
                    return ZNG_B200_CUDA_ERROR;
                }
            }
        }
    }
    CK(cudaMemcpyAsync(S.h_failed, S.d_sync + 1, sizeof(uint32_t), cudaMemcpyDeviceToHost, S.d2h), "D2H failed flag");
    CK(cudaStreamSynchronize(S.parse), "sync");
    CK(cudaStreamSynchronize(S.d2h), "sync");
    CK(cudaStreamSynchronize(S.copy), "sync");
    for (int i = 0; i < kStreamOut; i++) CK(cudaStreamSynchronize(S.out[i]), "sync");
    if (trace && !tev.empty()) {
        cudaEvent_t end; cudaEventCreate(&end); cudaEventRecord(end, S.d2h); cudaEventSynchronize(end);
        auto ms = [&](cudaEvent_t e) { float t = 0; cudaEventElapsedTime(&t, tev[0], e); return t; };
        fprintf(stderr, "streamed: parse kernel done %.2f, input delivered %.2f, all drained %.2f ms\n", ms(tev[1]), ms(tev[2]), ms(end));
        for (size_t j = 0; 4 + 2 * j < tev.size(); j++) fprintf(stderr, "  slab %2zu: parsed %7.2f  packed %7.2f\n", j, ms(tev[3 + 2 * j]), ms(tev[4 + 2 * j]));
        fprintf(stderr, "  packed bytes on the host at:");
        for (cudaEvent_t e : d2h_ev) { fprintf(stderr, " %.2f", ms(e)); cudaEventDestroy(e); }
        fprintf(stderr, " ms\n");
        for (cudaEvent_t e : tev) cudaEventDestroy(e);
        cudaEventDestroy(end);
    }
    if (rc) return rc;
    if (*S.h_failed) { snprintf(ctx->err, sizeof(ctx->err), "streamed pipeline: a device-side wait ran out of patience"); return ZNG_B200_CUDA_ERROR; }
    return 0;
}
}

int zng_b200_deflate_host(zng_b200_ctx* ctx, const void* h_in, size_t n, uint32_t chunk, int level, int final,
                          void* h_out, size_t out_cap, size_t* out_len, uint32_t* crc32, uint32_t* adler32) {
    if (!ctx) return ZNG_B200_STREAM_ERROR;
    if (level < 1 || level > 6) return bad(ctx, "level must be 1..6");
    if (chunk == 0 || chunk > ZNG_B200_CHUNK_MAX) return bad(ctx, "chunk must be in 1..65536");
    if (!out_len || (n && !h_in) || !h_out) return bad(ctx, "NULL argument");
    DeviceGuard g(ctx->device);
    int r = ensure_slabs(ctx);
    if (r) return r;
    if ((level == 1 || ctx->streamed >= 2) && chunk == ZNG_B200_CHUNK_MAX && ctx->streamed && n >= ((size_t)32 << 20)) {
        // large inputs: one persistent parse kernel fed by the copy engine (see Streamed)
        size_t pos = 0, done = 0; uint32_t c = 0, a = 1;
        const size_t super = (size_t)kStreamMaxChunks * chunk;
        while (done < n) {
            const size_t take = (n - done) < super ? (n - done) : super;
            r = deflate_host_streamed(ctx, (const uint8_t*)h_in + done, take, (final && done + take == n) ? 1 : 0, (uint8_t*)h_out, out_cap, pos, c, a,
                                      level, (level >= 2 && done > 0) ? 1 : 0);
            if (r) return r;
            done += take;
        }
        *out_len = pos;
        if (crc32) *crc32 = c;
        if (adler32) *adler32 = a;
        return 0;
    }
    const size_t stride = ctx->slab_stride;
    const uint32_t kSlabChunks = ctx->slab_chunks;
    const size_t slab_in = (size_t)kSlabChunks * chunk;
    uint8_t* out = (uint8_t*)h_out;
    size_t out_pos = 0, off = 0;
    uint32_t crc = 0, adler = 1;
    int k = 0;
    // n == 0 with final: one empty Z_FINISH chunk ("03 00"); n == 0 without final: nothing to emit.
    // Slabs run on their own streams: H2D, K1a/K1b, K3, gather and D2H of different slabs overlap.
    bool emitted_final = false;
    // ZNG_B200_TRACE=1: per-slab stage times (ms since the first H2D was issued) on stderr -- a debugging aid for the pipeline
    static const bool trace = getenv("ZNG_B200_TRACE") != nullptr;
    std::vector<cudaEvent_t> tev;
    auto mark = [&](cudaStream_t st) { if (trace) { cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, st); tev.push_back(e); } };
    while (off < n || (final && !emitted_final)) {
        Slab& s = ctx->slab[k];
        r = drain_slab(ctx, s, out, out_cap, out_pos, crc, adler);     // the previous occupant's bytes leave first
        if (r) { sync_slabs(ctx); return r; }
        mark(s.stream);                                                 // [4i+0] slab stream free again (behind the old occupant's D2H)
        // Full slabs only: at the slow levels a slab must hold most of a wave of chains (4 736 on B200) or the GPU idles behind each
        // slab's stragglers -- 4096 chunks x 4 slabs is the measured optimum at levels 2, 4 and 6; short first / last slabs (to start
        // and finish early) were tried and lose (profiles/r2_sweep_slab_levels.txt).
        const size_t take = (n - off) < slab_in ? (n - off) : slab_in;
        const bool is_last = (off + take == n);
        const size_t pre = (level >= 2) ? (off < kWSize ? off : (size_t)kWSize) : 0;
        const int have_prev = off > 0 ? 1 : 0;
        if (take) CK(cudaMemcpyAsync(s.d_in - pre, (const uint8_t*)h_in + off - pre, take + pre, cudaMemcpyHostToDevice, s.stream), "H2D");
        mark(s.stream);                                                 // [4i+1] H2D done
        uint32_t nch = (uint32_t)((take + chunk - 1) / chunk);
        uint32_t* d_sizes = s.d_sizes; uint32_t* d_crcs = s.d_sizes + kSlabChunks; uint32_t* d_adlers = s.d_sizes + 2 * kSlabChunks;
        if (final && is_last) {
            // all but the last chunk end with the full-flush marker; the last one is the Z_FINISH chunk
            const uint32_t body = nch ? nch - 1 : 0;
            const size_t body_bytes = (size_t)body * chunk;
            if (body) {
                r = run_deflate_chunks(ctx, s.scratch, s.d_in, body_bytes, chunk, body, 0, s.d_slots, stride, d_sizes, d_crcs, d_adlers, s.stream, nullptr, 0,
                                      level, have_prev, 8);
                if (r) { sync_slabs(ctx); return r; }
            }
            const size_t tail = take - body_bytes;
            if (tail) {
                r = run_deflate_chunks(ctx, s.scratch, s.d_in + body_bytes, tail, chunk, 1, 1, s.d_slots + (size_t)body * stride, stride,
                                      d_sizes + body, d_crcs + body, d_adlers + body, s.stream, nullptr, 0, level, (body || have_prev) ? 1 : 0, 8);
                if (r) { sync_slabs(ctx); return r; }
            } else {
                // zng_deflate(Z_FINISH) with no input: "03 00" (empty static block, BFINAL) -- deflate_quick.c:53-58
                static const uint8_t fin[2] = {0x03, 0x00};
                static const uint32_t meta[3] = {2u, 0u, 1u};
                CK(cudaMemcpyAsync(s.d_slots + (size_t)body * stride, fin, 2, cudaMemcpyHostToDevice, s.stream), "H2D fin");
                CK(cudaMemcpyAsync(d_sizes + body, &meta[0], 4, cudaMemcpyHostToDevice, s.stream), "H2D fin size");
                CK(cudaMemcpyAsync(d_crcs + body, &meta[1], 4, cudaMemcpyHostToDevice, s.stream), "H2D fin crc");
                CK(cudaMemcpyAsync(d_adlers + body, &meta[2], 4, cudaMemcpyHostToDevice, s.stream), "H2D fin adler");
                nch = body + 1;
            }
            emitted_final = true;
        } else {
            r = run_deflate_chunks(ctx, s.scratch, s.d_in, take, chunk, nch, 0, s.d_slots, stride, d_sizes, d_crcs, d_adlers, s.stream, nullptr, 0,
                                  level, have_prev, 8);
            if (r) { sync_slabs(ctx); return r; }
        }
        mark(s.stream);                                                 // [4i+2] parse + emit + checksums done
        CK(launch_offsets(d_sizes, nch, 0, s.d_offsets, s.stream), "offsets launch");
        CK(launch_gather(s.d_slots, stride, d_sizes, s.d_offsets, nch, s.d_packed, ctx->sms, s.stream), "gather launch");
        const uint32_t ntiles = (uint32_t)((take + chunk - 1) / chunk);
        CK(launch_crc32_fold(d_crcs, ntiles, chunk, take, 0, s.d_res, s.stream), "crc fold");
        CK(launch_adler32_fold(d_adlers, ntiles, chunk, take, 1, s.d_res + 1, s.stream), "adler fold");
        CK(cudaMemcpyAsync(&s.h_meta[0], s.d_offsets + nch, sizeof(uint64_t), cudaMemcpyDeviceToHost, s.stream), "D2H total");
        CK(cudaMemcpyAsync(&s.h_meta[1], s.d_res, sizeof(uint64_t), cudaMemcpyDeviceToHost, s.stream), "D2H checks");
        CK(cudaEventRecord(s.done, s.stream), "event record");
        mark(s.stream);                                                 // [4i+3] gather + folds done (the slab is drainable)
        s.in_bytes = take;
        s.busy = true;
        off += take;
        k = (k + 1) % ctx->pipe;
        // eager drain, oldest first (output order is slab order): start the D2H of finished slabs now
        for (int i = 0; i < ctx->pipe; i++) {
            Slab& o = ctx->slab[(k + i) % ctx->pipe];
            if (!o.busy) continue;
            if (cudaEventQuery(o.done) != cudaSuccess) break;
            r = drain_slab(ctx, o, out, out_cap, out_pos, crc, adler);
            if (r) { sync_slabs(ctx); return r; }
        }
    }
    // drain in issue order
    for (int i = 0; i < ctx->pipe; i++) {
        r = drain_slab(ctx, ctx->slab[(k + i) % ctx->pipe], out, out_cap, out_pos, crc, adler);
        if (r) { sync_slabs(ctx); return r; }
    }
    for (int i = 0; i < ctx->pipe; i++) CK(cudaStreamSynchronize(ctx->slab[i].stream), "final sync");
    if (trace && !tev.empty()) {
        cudaEvent_t end; cudaEventCreate(&end); cudaEventRecord(end, ctx->slab[0].stream); cudaEventSynchronize(end);
        for (size_t i = 0; i + 3 < tev.size(); i += 4) {
            float a = 0, b = 0, c = 0, d = 0;
            cudaEventElapsedTime(&a, tev[1], tev[i]); cudaEventElapsedTime(&b, tev[1], tev[i + 1]);
            cudaEventElapsedTime(&c, tev[1], tev[i + 2]); cudaEventElapsedTime(&d, tev[1], tev[i + 3]);
            fprintf(stderr, "slab %2zu: stream free %7.2f  h2d done %7.2f  kernels done %7.2f  packed %7.2f ms\n", i / 4, a, b, c, d);
        }
        float e = 0; cudaEventElapsedTime(&e, tev[1], end);
        fprintf(stderr, "all drained %7.2f ms (relative to the end of the first H2D)\n", e);
        for (cudaEvent_t ev : tev) cudaEventDestroy(ev);
        cudaEventDestroy(end);
    }
    *out_len = out_pos;
    if (crc32) *crc32 = crc;
    if (adler32) *adler32 = adler;
    return 0;
}

}  // extern "C"
