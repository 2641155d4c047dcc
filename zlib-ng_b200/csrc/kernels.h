// kernels.h -- host-callable launchers of the sm_100a kernels (internal to the shared library;
// the public C ABI is include/zng_b200.h).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace zb {

// Streamed input (host pipeline): `ready` = chunks delivered so far (written by the copy stream), `done[ci >> done_shift]`
// counts parsed chunks per output slab; `patience` clocks bound every wait, `failed` is raised when one runs out.
struct StreamSync {
    const uint32_t* ready = nullptr;
    uint32_t* done = nullptr;
    uint32_t done_shift = 0;
    uint32_t* host_done = nullptr;      // mapped pinned memory: host_done[slab] = chunks of the slab once all are parsed
    long long patience = 0;
    uint32_t* failed = nullptr;
};

// K1: level-1 chunk deflate (deflate_quick.cu): K1a parse (one warp per chain) -> token lists, K1b
// static-Huffman emit.  `heads` = 128 KiB hash-head slab per warp of the parse grid, `tail` =
// deflate_quick_tail_bytes() of scratch for the padded copy of the last chunks.
size_t deflate_quick_head_bytes(uint32_t nsmid);       // pool of nsmid x 64 slabs, handed out per SM via sm_slots[nsmid]
cudaError_t query_nsmid(uint32_t* d_scratch, uint32_t* nsmid);
size_t deflate_quick_tail_bytes();
uint32_t deflate_quick_grid(uint32_t nchunks, int num_sms, int chains_per_sm);
cudaError_t launch_quick_parse(const uint8_t* in, size_t n, uint32_t chunk, uint32_t nchunks,
                               uint32_t* tokens, uint32_t tok_stride, uint32_t* ntok, uint32_t* counter,
                               uint16_t* heads, unsigned long long* sm_slots, uint32_t grid, uint8_t* tail, cudaStream_t stream,
                               const StreamSync* sync = nullptr);
// K1a v6 (deflate_quick_cta.cu): one CTA per chain -- producer warps one window ahead of a walker warp, few chains per SM so that
// the head tables stay in L2.  Same inputs / outputs / slab pool / StreamSync protocol as launch_quick_parse.
cudaError_t launch_quick_parse_cta(const uint8_t* in, size_t n, uint32_t chunk, uint32_t nchunks,
                                   uint32_t* tokens, uint32_t tok_stride, uint32_t* ntok, uint32_t* counter,
                                   uint16_t* heads, unsigned long long* sm_slots, int num_sms, int chains_per_sm, int warps, uint8_t* tail,
                                   cudaStream_t stream, const StreamSync* sync = nullptr,
                                   unsigned long long* stats = nullptr);   // warps = producer warps per chain: 4, 6, 8 or 12; stats: 16 debug counters
// K1 primed (pigz's dependent-chunk mode): every chunk but the stream's first has the 32768 bytes in front of it as its
// preset dictionary.  heads = deflate_primed_head_bytes() pool of 256 KiB slabs (32-bit absolute positions).
size_t deflate_primed_head_bytes(uint32_t nsmid);
size_t deflate_primed_tail_bytes();
cudaError_t launch_primed_parse(const uint8_t* in, size_t n, uint32_t chunk, uint32_t nchunks, uint32_t first,
                                uint32_t* tokens, uint32_t tok_stride, uint32_t* ntok, uint32_t* counter,
                                uint32_t* heads, unsigned long long* sm_slots, uint32_t grid, uint8_t* tail, cudaStream_t stream);
cudaError_t launch_static_emit(const uint32_t* tokens, uint32_t tok_stride, const uint32_t* ntok, size_t n, uint32_t chunk,
                               uint32_t nchunks, int last, uint8_t* out, size_t out_stride, uint32_t* sizes,
                               int num_sms, cudaStream_t stream, int co_carve = -1);

// K2: level-2 chunk deflate (deflate_fast.cu): K2a parse with hash chains -> token lists, K2b block writer
// (dynamic / static / stored blocks).  Shares the head slab pool and sm_slots with K1; prevs / tails are the
// per-slab prev[] tables and stale-window images.  have_prev: chunk 0 of this launch is preceded in memory by
// the 32 KiB of input that came before it in the stream (what a short last chunk's over-reads see).
size_t deflate_fast_prev_bytes(uint32_t nsmid);
size_t deflate_fast_tail_bytes(uint32_t nsmid);
cudaError_t launch_fast_parse(const uint8_t* in, size_t n, uint32_t chunk, uint32_t nchunks, uint32_t* tokens, uint32_t tok_stride,
                              uint32_t* ntok, uint32_t* counter, uint16_t* heads, uint16_t* prevs, uint32_t* tails,
                              unsigned long long* sm_slots, int num_sms, int chains_per_sm, int have_prev, int level, cudaStream_t stream,
                              const StreamSync* sync = nullptr, int dyn_smem = 0);
cudaError_t launch_block_emit(const uint8_t* in, const uint32_t* tokens, uint32_t tok_stride, const uint32_t* ntok, size_t n,
                              uint32_t chunk, uint32_t nchunks, int last, uint8_t* out, size_t out_stride, uint32_t* sizes,
                              int num_sms, cudaStream_t stream, int co_carve = -1,
                              const uint8_t* blkflags = nullptr,    // blkflags: 8 bytes per chunk, != 0: block k may not be stored (K2w)
                              uint32_t* counter = nullptr);         // counter: a device word the kernel claims chunks from (zeroed here)
// K2w (deflate_window.cu): dictionary-primed chunks at levels 2-6 on the reference's own window / head / prev state (real slides).
// Shares the K2 slab pools; wins = deflate_window_win_bytes() of window buffers, blkflags as above (written by the parser).
size_t deflate_window_win_bytes(uint32_t nsmid);
cudaError_t launch_window_parse(const uint8_t* in, size_t n, uint32_t chunk, uint32_t nchunks, uint32_t first, int last, uint32_t* tokens,
                                uint32_t tok_stride, uint32_t* ntok, uint8_t* blkflags, uint32_t* counter, uint16_t* heads, uint16_t* prevs,
                                uint8_t* wins, unsigned long long* sm_slots, int num_sms, int chains_per_sm, int level, cudaStream_t stream);

// K3: checksums (checksum.cu)
cudaError_t launch_checksum_tiles(const uint8_t* in, size_t n, uint32_t tile_bytes, uint32_t ntiles,
                                  uint32_t* crcs, uint32_t* adlers, int num_sms, cudaStream_t stream, int cta_warps = 32);
cudaError_t launch_crc32_fold(const uint32_t* crcs, uint32_t ntiles, uint32_t tile_bytes, size_t n, uint32_t init,
                              uint32_t* result, cudaStream_t stream);
cudaError_t launch_adler32_fold(const uint32_t* adlers, uint32_t ntiles, uint32_t tile_bytes, size_t n, uint32_t init,
                                uint32_t* result, cudaStream_t stream);

// K4: batched inflate (inflate.cu), one warp per member
cudaError_t launch_inflate_members(const uint8_t* in, const uint64_t* in_off, uint32_t n_members, int window_bits,
                                   uint8_t* out, const uint64_t* out_off, uint32_t* sizes, uint32_t* checks, int32_t* status,
                                   uint32_t* in_used, uint32_t* detail, uint32_t* counter, int num_sms, cudaStream_t stream,
                                   int mode = 0,    // bit 0: flush-delimited segment of a raw stream; bit 1: sizes only; bit 2: (begin, end) offset pairs;
                                   uint32_t* resume_io = nullptr);   // bit 3: resume at a block boundary: 4 words per member, in {start_bit, out_start}, out {bb_byte, bb_bit, bb_out}
cudaError_t launch_marker_scan(const uint8_t* in, size_t n, size_t from, unsigned long long* pos, uint32_t cap, uint32_t* count,
                               int num_sms, cudaStream_t stream);
const char* inflate_msg(uint32_t id);

// operator surface (ops.cu): the device operators of K1/K2/K4 one at a time
cudaError_t launch_op_compare256(const uint8_t* a, const uint8_t* b, size_t stride, uint32_t n_pairs, uint32_t* out, cudaStream_t s);
cudaError_t launch_op_longest_match(const uint8_t* window, uint32_t n, const uint16_t* prev, const uint32_t* pos, const uint32_t* cand,
                                    uint32_t n_q, uint32_t* len, uint32_t* start, cudaStream_t s, int level = 2,
                                    bool raw = false);   // raw: lengths below 4 are reported as the reference returns them (2, 3)
cudaError_t launch_op_insert_string(const uint8_t* window, uint16_t* head, uint16_t* prev, uint32_t str, uint32_t count, cudaStream_t s,
                                    uint32_t* old_head = nullptr);
cudaError_t launch_op_slide_hash(uint16_t* head, uint16_t* prev, uint32_t wsize, cudaStream_t s);
cudaError_t launch_op_chunkmemset(uint8_t* out, uint32_t pos, uint32_t dist, uint32_t len, cudaStream_t s);

// stream assembly (assemble.cu)
cudaError_t launch_offsets(const uint32_t* sizes, uint32_t n, uint64_t base, uint64_t* offsets, cudaStream_t stream);
cudaError_t launch_gather(const uint8_t* slots, size_t stride, const uint32_t* sizes, const uint64_t* offsets,
                          uint32_t nchunks, uint8_t* dst, int num_sms, cudaStream_t stream);

}  // namespace zb
