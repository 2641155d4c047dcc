// multi.cu -- the N-GPU step of the chunked gzip writer behind the C ABI.
//
// Rank r of G owns a contiguous chunk range of one stream and has compressed it on its own GPU.  The ONLY exchange is an
// NCCL allgather of the per-chunk (compressed size, crc32) pairs -- 8 bytes per chunk -- after which every rank runs, on its
// device, the exclusive scan of all sizes (the byte offset of each of its chunks in the final stream; chunk outputs are byte
// aligned because each ends with the empty stored block of Z_FULL_FLUSH, deflate.c:1064-1065) and the crc32_combine fold
// (crc32_braid_comb.c:16-24) that the gzip trailer needs (deflate.c:1091-1096: CRC-32 LE, ISIZE = total_in mod 2^32).
// Payload bytes never cross NVLink: each rank packs its own chunks and writes them at its own offset.
//
// NCCL is bound at run time (dlopen of libnccl.so.2: inside a PyTorch process that is the copy torch already loaded, in a C
// program the system's), so the library itself does not link against it and loads on boxes without NCCL; the
// communicator entry points then fail with ZNG_B200_STREAM_ERROR.  Only ncclAllGather, ncclGetUniqueId, ncclCommInitRank,
// ncclCommDestroy, ncclCommCount, ncclCommUserRank and ncclGetErrorString are used.
#include "common.cuh"
#include "kernels.h"
#include "zng_b200.h"
#include <dlfcn.h>
#include <mutex>
#include <stdio.h>
#include <string.h>

namespace {

// the part of nccl.h this file needs (NCCL 2.x ABI: ncclUniqueId is 128 bytes, ncclUint8 = 1)
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
constexpr int kNcclUint8 = 1;

struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*CommCount)(const ncclComm_t, int*) = nullptr;
    ncclResult_t (*CommUserRank)(const ncclComm_t, int*) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    bool ok = false;
};

NcclApi& nccl() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* nm : names) { api.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL); if (api.lib) break; }
        if (!api.lib) return;
        auto sym = [&](const char* s) { return dlsym(api.lib, s); };
        api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
        api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
        api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
        api.CommCount = (decltype(api.CommCount))sym("ncclCommCount");
        api.CommUserRank = (decltype(api.CommUserRank))sym("ncclCommUserRank");
        api.AllGather = (decltype(api.AllGather))sym("ncclAllGather");
        api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
        api.Send = (decltype(api.Send))sym("ncclSend");
        api.Recv = (decltype(api.Recv))sym("ncclRecv");
        api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
        api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
        api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.CommCount && api.CommUserRank && api.AllGather;
    });
    return api;
}

// gathered[r][0][i] = size, gathered[r][1][i] = crc of rank r's chunk i (row stride `width`) -> all_sizes / all_crcs in global
// chunk order; first[r] = global index of rank r's first chunk
__global__ void compact_pairs_kernel(const uint32_t* __restrict__ gathered, uint32_t width, const uint32_t* __restrict__ first,
                                     uint32_t nranks, uint32_t total, uint32_t* __restrict__ all_sizes, uint32_t* __restrict__ all_crcs) {
    for (uint32_t g = blockIdx.x * blockDim.x + threadIdx.x; g < total; g += gridDim.x * blockDim.x) {
        uint32_t r = 0;
        while (r + 1u < nranks && first[r + 1u] <= g) r++;
        const uint32_t i = g - first[r];
        all_sizes[g] = gathered[((size_t)r * 2u + 0u) * width + i];
        all_crcs[g] = gathered[((size_t)r * 2u + 1u) * width + i];
    }
}

}  // namespace

struct zng_b200_comm {
    zng_b200_ctx* ctx = nullptr;
    ncclComm_t nc = nullptr;
    int nranks = 1, rank = 0;
    bool own = false;
    char err[200] = {0};
    // scratch (grow only)
    unsigned long long* d_hdr = nullptr;   // nranks x {nchunks_local, n_local}
    unsigned long long* h_hdr = nullptr;   // pinned copy + results
    uint32_t* d_first = nullptr;
    uint32_t* d_mine = nullptr;  size_t mine_cap = 0;       // 2 x width
    uint32_t* d_gath = nullptr;  size_t gath_cap = 0;       // nranks x 2 x width
    uint32_t* d_all = nullptr;   size_t all_cap = 0;        // sizes | crcs, total each
    uint64_t* d_off = nullptr;   size_t off_cap = 0;        // total + 1
    uint32_t* d_res = nullptr;
};

namespace {
int cfail(zng_b200_comm* c, const char* what, const char* detail) {
    snprintf(c->err, sizeof(c->err), "%s: %s", what, detail ? detail : "");
    return ZNG_B200_CUDA_ERROR;
}
#define CCK(call, what) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return cfail(c, what, cudaGetErrorString(e_)); } while (0)
#define NCK(call, what) do { ncclResult_t r_ = (call); if (r_ != 0) return cfail(c, what, nccl().GetErrorString ? nccl().GetErrorString(r_) : "nccl error"); } while (0)

template <typename T>
int grow(zng_b200_comm* c, T*& p, size_t& cap, size_t want, const char* what) {
    if (cap >= want) return 0;
    if (p) { cudaDeviceSynchronize(); cudaFree(p); p = nullptr; cap = 0; }
    want += want / 4 + 1024;
    CCK(cudaMalloc(&p, want * sizeof(T)), what);
    cap = want;
    return 0;
}

int comm_finish_init(zng_b200_comm* c) {
    CCK(cudaSetDevice(zng_b200_ctx_device(c->ctx)), "cudaSetDevice");
    CCK(cudaMalloc(&c->d_hdr, (size_t)c->nranks * 2 * sizeof(unsigned long long)), "cudaMalloc(comm header)");
    CCK(cudaHostAlloc(&c->h_hdr, ((size_t)c->nranks * 2 + 8) * sizeof(unsigned long long), cudaHostAllocDefault), "cudaHostAlloc(comm header)");
    CCK(cudaMalloc(&c->d_first, ((size_t)c->nranks + 1) * sizeof(uint32_t)), "cudaMalloc(comm first)");
    CCK(cudaMalloc(&c->d_res, 4 * sizeof(uint32_t)), "cudaMalloc(comm result)");
    return 0;
}
}  // namespace

extern "C" {

int zng_b200_comm_unique_id(void* id, size_t cap) {
    if (!id || cap < ZNG_B200_COMM_ID_BYTES || !nccl().ok) return ZNG_B200_STREAM_ERROR;
    ncclUniqueId u;
    if (nccl().GetUniqueId(&u) != 0) return ZNG_B200_CUDA_ERROR;
    memcpy(id, &u, sizeof(u));
    return ZNG_B200_OK;
}

int zng_b200_comm_create(zng_b200_ctx* ctx, int nranks, int rank, const void* id, zng_b200_comm** out) {
    if (!ctx || !out || nranks < 1 || rank < 0 || rank >= nranks || (nranks > 1 && !id)) return ZNG_B200_STREAM_ERROR;
    zng_b200_comm* c = new (std::nothrow) zng_b200_comm();
    if (!c) return ZNG_B200_MEM_ERROR;
    c->ctx = ctx; c->nranks = nranks; c->rank = rank;
    int r = 0;
    if (nranks > 1) {
        if (!nccl().ok) { delete c; return ZNG_B200_STREAM_ERROR; }
        if (cudaSetDevice(zng_b200_ctx_device(ctx)) != cudaSuccess) { delete c; return ZNG_B200_CUDA_ERROR; }
        ncclUniqueId u; memcpy(&u, id, sizeof(u));
        if (nccl().CommInitRank(&c->nc, nranks, u, rank) != 0) { delete c; return ZNG_B200_CUDA_ERROR; }
        c->own = true;
    }
    r = comm_finish_init(c);
    if (r) { zng_b200_comm_destroy(c); return r; }
    *out = c;
    return ZNG_B200_OK;
}

int zng_b200_comm_adopt(zng_b200_ctx* ctx, void* nccl_comm, zng_b200_comm** out) {
    if (!ctx || !out || !nccl_comm || !nccl().ok) return ZNG_B200_STREAM_ERROR;
    zng_b200_comm* c = new (std::nothrow) zng_b200_comm();
    if (!c) return ZNG_B200_MEM_ERROR;
    c->ctx = ctx; c->nc = (ncclComm_t)nccl_comm; c->own = false;
    if (nccl().CommCount(c->nc, &c->nranks) != 0 || nccl().CommUserRank(c->nc, &c->rank) != 0) { delete c; return ZNG_B200_CUDA_ERROR; }
    int r = comm_finish_init(c);
    if (r) { zng_b200_comm_destroy(c); return r; }
    *out = c;
    return ZNG_B200_OK;
}

void zng_b200_comm_destroy(zng_b200_comm* c) {
    if (!c) return;
    cudaSetDevice(zng_b200_ctx_device(c->ctx));
    cudaDeviceSynchronize();
    if (c->own && c->nc) nccl().CommDestroy(c->nc);
    cudaFree(c->d_hdr); cudaFreeHost(c->h_hdr); cudaFree(c->d_first); cudaFree(c->d_mine); cudaFree(c->d_gath);
    cudaFree(c->d_all); cudaFree(c->d_off); cudaFree(c->d_res);
    delete c;
}

int zng_b200_comm_size(const zng_b200_comm* c) { return c ? c->nranks : 0; }
int zng_b200_comm_rank(const zng_b200_comm* c) { return c ? c->rank : -1; }
const char* zng_b200_comm_error(const zng_b200_comm* c) { return c ? c->err : "no communicator"; }

int zng_b200_stream_index_multi(zng_b200_comm* c, const uint32_t* d_sizes, const uint32_t* d_crcs, uint32_t nchunks_local,
                                uint32_t chunk, size_t n_local, uint64_t base, uint64_t* d_offsets_local,
                                uint64_t* h_stream_end, uint32_t* h_crc32, uint64_t* h_total_in, void* stream) {
    if (!c || (nchunks_local && (!d_sizes || !d_crcs)) || !d_offsets_local || chunk == 0 || chunk > ZNG_B200_CHUNK_MAX) return ZNG_B200_STREAM_ERROR;
    if (nchunks_local != (uint32_t)((n_local + chunk - 1) / chunk)) return ZNG_B200_STREAM_ERROR;
    cudaStream_t st = (cudaStream_t)stream;
    zng_b200_ctx* ctx = c->ctx;
    CCK(cudaSetDevice(zng_b200_ctx_device(ctx)), "cudaSetDevice");
    const int G = c->nranks;
    // ---- who owns how much (16 bytes per rank)
    unsigned long long* hh = c->h_hdr;
    hh[2 * G + 0] = nchunks_local; hh[2 * G + 1] = n_local;
    if (G > 1) {
        CCK(cudaMemcpyAsync(c->d_hdr + 2 * c->rank, hh + 2 * G, 16, cudaMemcpyHostToDevice, st), "H2D header");
        NCK(nccl().AllGather(c->d_hdr + 2 * c->rank, c->d_hdr, 16, kNcclUint8, c->nc, st), "ncclAllGather(header)");
        CCK(cudaMemcpyAsync(hh, c->d_hdr, (size_t)G * 16, cudaMemcpyDeviceToHost, st), "D2H header");
        CCK(cudaStreamSynchronize(st), "cudaStreamSynchronize");
    } else {
        hh[0] = nchunks_local; hh[1] = n_local;
    }
    uint64_t total = 0, n_total = 0, width = 0;
    uint32_t first[1025];
    if (G > 1024) return ZNG_B200_STREAM_ERROR;
    for (int r = 0; r < G; r++) {
        first[r] = (uint32_t)total;
        // every chunk of the stream but its very last is `chunk` bytes long: the fold and the scan rely on it
        if (r + 1 < G && hh[2 * r + 1] != hh[2 * r] * (unsigned long long)chunk) {
            bool later = false;
            for (int q = r + 1; q < G; q++) later |= hh[2 * q] != 0;
            if (later) { snprintf(c->err, sizeof(c->err), "rank %d ends with a short chunk but is not the last rank with data", r); return ZNG_B200_STREAM_ERROR; }
        }
        total += hh[2 * r]; n_total += hh[2 * r + 1];
        if (hh[2 * r] > width) width = hh[2 * r];
    }
    first[G] = (uint32_t)total;
    if (total > 0xfffffff0ull) return ZNG_B200_STREAM_ERROR;
    int r = 0;
    if ((r = grow(c, c->d_all, c->all_cap, 2 * (size_t)total + 2, "cudaMalloc(all pairs)"))) return r;
    if ((r = grow(c, c->d_off, c->off_cap, (size_t)total + 1, "cudaMalloc(all offsets)"))) return r;
    uint32_t* all_sizes = c->d_all; uint32_t* all_crcs = c->d_all + total;
    if (G > 1) {
        // ---- THE collective: (size, crc32) of every chunk, 8 bytes per chunk (rows padded to the widest rank)
        if ((r = grow(c, c->d_mine, c->mine_cap, 2 * (size_t)width + 2, "cudaMalloc(my pairs)"))) return r;
        if ((r = grow(c, c->d_gath, c->gath_cap, 2 * (size_t)width * G + 2, "cudaMalloc(gathered pairs)"))) return r;
        if (nchunks_local) {
            CCK(cudaMemcpyAsync(c->d_mine, d_sizes, (size_t)nchunks_local * 4, cudaMemcpyDeviceToDevice, st), "copy sizes");
            CCK(cudaMemcpyAsync(c->d_mine + width, d_crcs, (size_t)nchunks_local * 4, cudaMemcpyDeviceToDevice, st), "copy crcs");
        }
        if (width) NCK(nccl().AllGather(c->d_mine, c->d_gath, 2 * (size_t)width * 4, kNcclUint8, c->nc, st), "ncclAllGather(size, crc32)");
        CCK(cudaMemcpyAsync(c->d_first, first, ((size_t)G + 1) * 4, cudaMemcpyHostToDevice, st), "H2D first");
        if (total) {
            compact_pairs_kernel<<<(unsigned)((total + 255) / 256 > 1184 ? 1184 : (total + 255) / 256), 256, 0, st>>>(
                c->d_gath, (uint32_t)width, c->d_first, (uint32_t)G, (uint32_t)total, all_sizes, all_crcs);
            CCK(cudaGetLastError(), "compact_pairs_kernel");
        }
    } else if (total) {
        CCK(cudaMemcpyAsync(all_sizes, d_sizes, (size_t)total * 4, cudaMemcpyDeviceToDevice, st), "copy sizes");
        CCK(cudaMemcpyAsync(all_crcs, d_crcs, (size_t)total * 4, cudaMemcpyDeviceToDevice, st), "copy crcs");
    }
    // ---- every rank: scan of all sizes, fold of all CRCs (both on the device), then its own slice of the offsets
    if ((r = zng_b200_chunk_offsets(ctx, all_sizes, (uint32_t)total, base, c->d_off, st))) return r;
    if ((r = zng_b200_crc32_fold(ctx, all_crcs, (uint32_t)total, chunk, (size_t)n_total, 0, c->d_res, st))) return r;
    CCK(cudaMemcpyAsync(d_offsets_local, c->d_off + first[c->rank], ((size_t)nchunks_local + 1) * 8, cudaMemcpyDeviceToDevice, st), "copy offsets");
    CCK(cudaMemcpyAsync(hh + 2 * G + 2, c->d_off + total, 8, cudaMemcpyDeviceToHost, st), "D2H stream end");
    CCK(cudaMemcpyAsync(hh + 2 * G + 3, c->d_res, 4, cudaMemcpyDeviceToHost, st), "D2H crc");
    CCK(cudaStreamSynchronize(st), "cudaStreamSynchronize");
    if (h_stream_end) *h_stream_end = hh[2 * G + 2];
    if (h_crc32) *h_crc32 = (uint32_t)hh[2 * G + 3];
    if (h_total_in) *h_total_in = n_total;
    return ZNG_B200_OK;
}

// Dependent (primed) chunks across ranks: rank r's first chunk is primed with the last 32768 bytes of rank r-1's shard.  Collective:
// every rank sends its last 32 KiB to its successor and receives its predecessor's into d_halo (which the caller places directly in
// front of its own input, so that [d_halo, d_halo + 32768 + n_local) is contiguous).  Rank 0 receives nothing.  One ncclSend /
// ncclRecv pair per neighbour over NVLink; 32 KiB per rank is all the payload that ever crosses it.
int zng_b200_halo_exchange(zng_b200_comm* c, const void* d_in, size_t n_local, void* d_halo, void* stream) {
    if (!c || !d_halo || (n_local && !d_in)) return ZNG_B200_STREAM_ERROR;
    if (c->nranks == 1) return ZNG_B200_OK;
    if (n_local < 32768u) { snprintf(c->err, sizeof(c->err), "halo exchange: every rank needs at least 32768 bytes of input"); return ZNG_B200_STREAM_ERROR; }
    if (!nccl().Send || !nccl().Recv || !nccl().GroupStart || !nccl().GroupEnd) return ZNG_B200_STREAM_ERROR;
    cudaStream_t st = (cudaStream_t)stream;
    CCK(cudaSetDevice(zng_b200_ctx_device(c->ctx)), "cudaSetDevice");
    NCK(nccl().GroupStart(), "ncclGroupStart");
    if (c->rank + 1 < c->nranks) NCK(nccl().Send((const uint8_t*)d_in + n_local - 32768u, 32768u, kNcclUint8, c->rank + 1, c->nc, st), "ncclSend(halo)");
    if (c->rank > 0) NCK(nccl().Recv(d_halo, 32768u, kNcclUint8, c->rank - 1, c->nc, st), "ncclRecv(halo)");
    NCK(nccl().GroupEnd(), "ncclGroupEnd");
    return ZNG_B200_OK;
}

int zng_b200_gzip_multi(zng_b200_comm* c, const void* d_in, size_t n_local, int level,
                        void* d_slots, size_t out_stride, uint32_t* d_sizes, uint32_t* d_crcs, uint64_t* d_offsets_local,
                        void* d_packed, size_t packed_cap, uint64_t* h_my_offset, uint64_t* h_my_bytes,
                        uint64_t* h_file_bytes, uint8_t* h_header10, uint8_t* h_trailer10, void* stream) {
    if (!c || !d_slots || !d_sizes || !d_crcs || !d_offsets_local || !d_packed || (n_local && !d_in)) return ZNG_B200_STREAM_ERROR;
    if (level < 1 || level > 6) return ZNG_B200_STREAM_ERROR;
    zng_b200_ctx* ctx = c->ctx;
    cudaStream_t st = (cudaStream_t)stream;
    const uint32_t chunk = ZNG_B200_CHUNK_MAX;
    const uint32_t nch = (uint32_t)((n_local + chunk - 1) / chunk);
    int r = zng_b200_deflate_chunks(ctx, d_in, n_local, chunk, level, ZNG_B200_FULL_FLUSH, d_slots, out_stride, d_sizes, d_crcs, nullptr, st);
    if (r) { snprintf(c->err, sizeof(c->err), "deflate_chunks: %s", zng_b200_last_error(ctx)); return r; }
    uint64_t stream_end = 0, total_in = 0; uint32_t crc = 0;
    r = zng_b200_stream_index_multi(c, d_sizes, d_crcs, nch, chunk, n_local, 10u, d_offsets_local, &stream_end, &crc, &total_in, st);
    if (r) return r;
    // my bytes: [offsets_local[0], offsets_local[nch]) of the file; packed into d_packed from 0
    uint64_t* hh = (uint64_t*)c->h_hdr + 2 * c->nranks + 4;
    CCK(cudaMemcpyAsync(hh, d_offsets_local, 8, cudaMemcpyDeviceToHost, st), "D2H my offset");
    CCK(cudaMemcpyAsync(hh + 1, d_offsets_local + nch, 8, cudaMemcpyDeviceToHost, st), "D2H my end");
    CCK(cudaStreamSynchronize(st), "cudaStreamSynchronize");
    const uint64_t my_off = hh[0], my_end = hh[1];
    if (my_end - my_off > packed_cap) return ZNG_B200_BUF_ERROR;
    if (nch) {
        r = zng_b200_gather_chunks(ctx, d_slots, out_stride, d_sizes, d_offsets_local, nch, (uint8_t*)d_packed - my_off, st);
        if (r) { snprintf(c->err, sizeof(c->err), "gather_chunks: %s", zng_b200_last_error(ctx)); return r; }
    }
    if (h_my_offset) *h_my_offset = my_off;
    if (h_my_bytes) *h_my_bytes = my_end - my_off;
    if (h_file_bytes) *h_file_bytes = stream_end + 10u;
    if (h_header10) {           // deflate.c:902-921 with no gz_header: 1f 8b 08 00 <mtime 0> XFL OS; XFL 4 for level 1, OS_CODE 3
        const uint8_t hdr[10] = {0x1f, 0x8b, 8, 0, 0, 0, 0, 0, (uint8_t)(level < 2 ? 4 : 0), 3};
        memcpy(h_header10, hdr, 10);
    }
    if (h_trailer10) {          // zng_deflate(Z_FINISH) with no input left: "03 00", then CRC-32 and ISIZE little endian
        h_trailer10[0] = 3; h_trailer10[1] = 0;
        for (int k = 0; k < 4; k++) { h_trailer10[2 + k] = (uint8_t)(crc >> (8 * k)); h_trailer10[6 + k] = (uint8_t)((uint32_t)total_in >> (8 * k)); }
    }
    return ZNG_B200_OK;
}

}  // extern "C"
