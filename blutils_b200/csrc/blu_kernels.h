// blu_kernels.h -- launch interface between the host runtime (blu_api.cpp) and the sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "blu_core.cuh"

namespace blu {

// Device-side counters of one run (zeroed by the host before the first chunk).
struct Counters {
    unsigned long long rec_slots;  // packed reservation counter: records << 32 | top-row slots
    unsigned int n_defer;      // deferred runs of the current chunk
    unsigned int err_code;     // first DevErr
    unsigned long long err_off;    // byte offset (in the current device buffer) of the first error
    unsigned long long tail_start; // !final chunks: start of the last (unfinished) run
    unsigned long long pool_used;  // bytes of the string pool in use
    unsigned int dup_found;    // a query id occurs in two separate runs
    unsigned int cap_overflow; // some output capacity was exceeded (host grows and retries)
    unsigned int work_ticket;  // dynamic work distribution of the long-run kernel
    unsigned int pad;
    unsigned long long n_rows; // hit rows of the finished queries (summed by the gather kernel)
};

struct RunParams {
    const uint8_t* text;   // device buffer (16-byte aligned, padded to a multiple of 128 bytes)
    uint64_t begin, end;   // valid text is [begin, end)
    int final_chunk;       // 1: `end` is the end of the input; 0: the run touching `end` is carried over
    int strategy;
    LinTables T;
    blu_record* records;
    uint32_t rec_cap;
    blu_bean* beans;
    blu_acc* accs;
    TopRowRaw* toprows;    // top bit-score rows of every query (fields 1..4 folded to integers, or unparsed references;
                           // lineage not joined yet), slot-indexed
    uint32_t slot_cap;
    uint64_t* defer;       // (offset << 1) | check_prev
    uint32_t defer_cap;
    Counters* ctr;
};

struct ConsParams {
    blu_record* records;
    uint32_t rec_begin, rec_end;
    const TopRowRaw* toprows;
    uint64_t text_end;     // valid text is [0, text_end) of `text` (references beyond it are an internal error)
    blu_bean* beans;
    blu_acc* accs;
    const uint8_t* text;
    LinTables T;
    int strategy;
    Counters* ctr;
};

struct GatherParams {
    const uint8_t* text;
    blu_record* records;
    blu_acc* accs;
    uint32_t rec_begin, rec_end;  // records produced by the current chunk
    uint8_t* pool;
    uint64_t pool_cap;
    Counters* ctr;
};

struct DupParams {
    const blu_record* records;
    uint32_t n_rec;
    const uint8_t* pool;
    uint64_t pool_cap;          // bytes of `pool` (ids outside it: the gather pass overflowed, the host reruns with a larger pool)
    unsigned long long* table;  // open addressing, 0 = empty
    uint32_t mask;
    Counters* ctr;
};

// Geometry of the streaming tile kernel: every CTA walks one contiguous segment of the text in
// kTile-byte steps (double-buffered TMA), carrying the unfinished query from one step to the next.
#ifndef BLU_TILE_BYTES
#define BLU_TILE_BYTES 32768
#endif
#ifndef BLU_TILE_CTAS
#define BLU_TILE_CTAS 2
#endif
#ifndef BLU_TILE_THREADS
#define BLU_TILE_THREADS 512
#endif
constexpr int kTile = BLU_TILE_BYTES;   // bytes staged per step
constexpr int kBack = 1024;             // look-behind of a segment's first window (predecessor of its first row)
constexpr int kTileThreads = BLU_TILE_THREADS;
constexpr int kTileCtasPerSm = BLU_TILE_CTAS;
constexpr int kWin = 60416;             // window of the block path (long-run kernel)

int tile_kernel_grid(int device);
cudaError_t launch_tile_kernel(const RunParams& p, int grid, cudaStream_t s);
cudaError_t launch_longrun_kernel(const RunParams& p, int grid, cudaStream_t s);
cudaError_t launch_consensus_kernel(const ConsParams& p, cudaStream_t s);
cudaError_t launch_gather_kernel(const GatherParams& p, cudaStream_t s);
cudaError_t launch_dup_kernel(const DupParams& p, cudaStream_t s);
cudaError_t kernels_set_attributes();

}  // namespace blu
