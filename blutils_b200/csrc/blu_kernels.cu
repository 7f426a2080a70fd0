// blu_kernels.cu -- hand-written sm_100a kernels of the consensus-identity path.
//
//   tile_kernel    : the dominant kernel (streaming).  Every CTA walks one contiguous segment of the outfmt-6 text in
//                    32 KB windows staged by TMA bulk copies (cp.async.bulk + mbarrier, double-buffered), classifies
//                    the bytes with SIMD-in-register tests (newline / tab / digit bitmasks, dp4a packing), indexes
//                    the rows, validates / parses every row, finds query runs and their top bit-score group --
//                    carrying the unfinished query from window to window -- and emits one record header plus the
//                    parsed top rows per query.  Text is read from HBM once; nothing per-row is written back.
//   consensus_kernel: one warp per query: taxid join, multi-taxa consensus, cutoffs (reads the tile kernel's output).
//   longrun_kernel : block-per-query path for the few queries the tile kernel hands over (top group of more than 32
//                    rows, bit score beyond int32, first row without a predecessor in its window).
//   gather_kernel  : copies query ids and accessions of finished queries into the result's string pool.
//   dup_kernel     : detects a query id that occurs in two separate runs (non-contiguous input).
//
// Reference semantics: see blu_core.cuh.  Geometry and roofline accounting: DESIGN.md.
#include <cuda_runtime.h>

#include <climits>
#include <cstdio>

#include "blu_kernels.h"

namespace blu {

namespace {

constexpr int kRowCap = kWin / 26 + 16;   // a valid row is >= 26 bytes (13 one-byte fields, 12 tabs, '\n')
constexpr int kLongTopCap = 1024;         // largest top bit-score group the block path sorts
constexpr int kLongThreads = 512;
constexpr int kLongWarps = kLongThreads / 32;

// ---------------------------------------------------------------------------------------------------------------
// small PTX wrappers (TMA 1-D bulk copy + mbarrier)
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// asks the TMA unit to pull [src, src+bytes) into L2 (no destination): the tile after the current one
__device__ __forceinline__ void tma_prefetch_l2(const void* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}

__device__ __forceinline__ void report(Counters* ctr, uint32_t err, unsigned long long off) {
    if (atomicCAS(&ctr->err_code, 0u, err) == 0u) ctr->err_off = off;
}

// ---------------------------------------------------------------------------------------------------------------
// shared-memory window + row index (used by both kernels)
// ---------------------------------------------------------------------------------------------------------------
constexpr int kChunks = (kWin + 128) / 16;  // 16-byte chunks of the window (incl. the spare tail)

struct WindowIndex {
    alignas(128) uint8_t win[kWin + 128];
    // byte-class bitmasks, one bit per window byte (u16 per 16-byte chunk; read back as 64-bit words)
    alignas(8) uint16_t tabm[kChunks + 8];
    alignas(8) uint16_t digm[kChunks + 8];
    // row starts / row ends per chunk (pass 1 of the row scan -> pass 2).  Dead once the row table exists: the tile
    // kernel re-uses the space for the per-row bit scores.
    alignas(8) uint16_t startm[kChunks + 8];
    uint16_t endm[kChunks + 8];
    uint16_t row_s[kRowCap];
    uint16_t row_e[kRowCap + 1];
    alignas(8) unsigned long long mbar;
    int n_starts, n_ends;
    int bad_byte;  // window offset of the first '"' / '\r' byte inside [qlo, qhi), or INT_MAX
    int warp_cnt[32];
};

struct WinGeom {
    unsigned long long lo;  // absolute offset of win[0]
    int vb;                 // first valid byte of the window (everything before it is not text / not loaded)
    int rb;                 // == vb when a row is known to start there (virtual newline in front), else -1
    int re;                 // end - lo   (may exceed the window)
    int loaded;             // bytes staged
    int L;                  // bytes scanned (text end + optional virtual newline)
    bool covers_eof;        // the window contains the end of the text
    int qlo, qhi;           // window range in which a '"' / '\r' byte is this CTA's to report
};

// Stage [lo, lo+bytes) of the text into the window with TMA bulk copies.  All threads call; returns when the
// bytes are visible.  `phase` is the barrier parity, flipped by the caller after every use.
__device__ __forceinline__ void load_window(WindowIndex& W, const uint8_t* text, unsigned long long lo, int bytes, uint32_t& phase,
                                            bool synced = false) {
    if (!synced) __syncthreads();  // everyone is done with the previous contents
    if (threadIdx.x == 0) {
        fence_proxy_async();
        mbar_expect_tx(&W.mbar, (uint32_t)bytes);
        for (int o = 0; o < bytes; o += 16384) {
            int n = bytes - o < 16384 ? bytes - o : 16384;
            tma_bulk_g2s(W.win + o, text + lo + o, (uint32_t)n, &W.mbar);
        }
    }
    while (!mbar_try_wait(&W.mbar, phase)) {
    }
    phase ^= 1;
}

// Window geometry for text [begin, end) seen through a window starting at `lo` of at most `max_bytes`.
__device__ __forceinline__ WinGeom make_geom(unsigned long long lo, int max_bytes, unsigned long long begin, unsigned long long end) {
    WinGeom g;
    g.lo = lo;
    unsigned long long up = (end + 15ull) & ~15ull;
    unsigned long long hi = lo + (unsigned long long)max_bytes;
    if (hi > up) hi = up;
    g.loaded = hi > lo ? (int)(hi - lo) : 0;
    g.rb = begin >= lo ? (int)((begin - lo) > 0x7fffffffull ? 0x7fffffff : (begin - lo)) : -1;
    g.vb = g.rb < 0 ? 0 : g.rb;
    long long re = (long long)end - (long long)lo;
    g.re = re > (long long)kWin + 64 ? kWin + 64 : (int)re;
    g.covers_eof = end <= lo + (unsigned long long)g.loaded;
    g.L = g.re < g.loaded ? g.re : g.loaded;
    g.qlo = g.vb;
    g.qhi = g.L;
    return g;
}

// SIMD-in-register byte classification: 0x80 in every byte of the result where the predicate holds
__device__ __forceinline__ uint32_t bytes_eq(uint32_t x, uint32_t c4) {
    uint32_t t = x ^ c4;  // zero byte <=> equal
    return ~(((t & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | t) & 0x80808080u;
}
__device__ __forceinline__ uint32_t bytes_digit(uint32_t x) {
    uint32_t t = x ^ 0x30303030u;  // '0'..'9' -> 0..9
    return ~(((t & 0x7F7F7F7Fu) + 0x76767676u) | t) & 0x80808080u;
}
// bytes below 0x23 ('#'): control characters, space, '!' and '"'.  In clean BLAST text only tab and newline are.
__device__ __forceinline__ uint32_t bytes_below_23(uint32_t x) {
    uint32_t t = (x | 0x80808080u) - 0x23232323u;  // no borrow between bytes; bit 7 survives iff (x & 0x7f) >= 0x23
    return ~t & ~x & 0x80808080u;
}
// Gathers the sixteen 0x80 flags of four words into a 16-bit mask with four dot products:
// a flag byte is 0x80 = 128, so dp4a(flags, weights) = 128 * (sum of the weights of the set bytes).
__device__ __forceinline__ uint32_t pack16(uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    const uint32_t lo = __dp4a(a, 0x08040201u, __dp4a(b, 0x80402010u, 0u));
    const uint32_t hi = __dp4a(c, 0x08040201u, __dp4a(d, 0x80402010u, 0u));
    return (lo >> 7) | (hi << 1);
}

__device__ __forceinline__ uint32_t range_mask16(int pos0, int lo, int hi) {  // bits k with lo <= pos0+k < hi
    int a = lo - pos0, b = hi - pos0;
    a = a < 0 ? 0 : (a > 16 ? 16 : a);
    b = b < 0 ? 0 : (b > 16 ? 16 : b);
    if (b <= a) return 0;
    return ((1u << b) - 1u) & ~((1u << a) - 1u);
}

// Rare path of the row scan: some byte below 0x23 is neither tab nor newline -- look for '"' / '\r' exactly.
__device__ __noinline__ void quote_check(WindowIndex& W, const WinGeom& g, const uint4 v, int pos0) {
    const uint32_t qm = pack16(bytes_eq(v.x, 0x22222222u) | bytes_eq(v.x, 0x0D0D0D0Du), bytes_eq(v.y, 0x22222222u) | bytes_eq(v.y, 0x0D0D0D0Du),
                               bytes_eq(v.z, 0x22222222u) | bytes_eq(v.z, 0x0D0D0D0Du), bytes_eq(v.w, 0x22222222u) | bytes_eq(v.w, 0x0D0D0D0Du)) &
                        range_mask16(pos0, g.qlo, g.qhi);
    if (qm) atomicMin(&W.bad_byte, pos0 + __ffs(qm) - 1);
}

// Row scan, pass 1: classifies one 16-byte chunk (newline / tab / digit / quote-or-CR) with SIMD-in-register byte
// tests, publishes the tab and digit masks, and derives the row-start / row-end masks of the chunk
// (DESIGN.md "row index").  `carry` = the byte in front of the chunk is a newline.
__device__ __forceinline__ void classify_chunk(WindowIndex& W, const WinGeom& g, int c, uint32_t carry, uint32_t& nl_out, uint32_t& start,
                                               uint32_t& end) {
    const int pos0 = c << 4;
    const uint4 v = *reinterpret_cast<const uint4*>(W.win + pos0);
    const int rb = g.vb;
    const int tend = g.re < g.loaded ? g.re : g.loaded;  // end of the text inside the window
    const uint32_t nx = bytes_eq(v.x, 0x0A0A0A0Au), ny = bytes_eq(v.y, 0x0A0A0A0Au), nz = bytes_eq(v.z, 0x0A0A0A0Au), nw = bytes_eq(v.w, 0x0A0A0A0Au);
    const uint32_t tx = bytes_eq(v.x, 0x09090909u), ty = bytes_eq(v.y, 0x09090909u), tz = bytes_eq(v.z, 0x09090909u), tw = bytes_eq(v.w, 0x09090909u);
    uint32_t nl = pack16(nx, ny, nz, nw);
    W.tabm[c] = (uint16_t)pack16(tx, ty, tz, tw);
    W.digm[c] = (uint16_t)pack16(bytes_digit(v.x), bytes_digit(v.y), bytes_digit(v.z), bytes_digit(v.w));
    // any byte below 0x23 that is neither tab nor newline is suspicious; only then look for '"' / '\r' exactly
    const uint32_t sus = (bytes_below_23(v.x) & ~(nx | tx)) | (bytes_below_23(v.y) & ~(ny | ty)) | (bytes_below_23(v.z) & ~(nz | tz)) |
                         (bytes_below_23(v.w) & ~(nw | tw));
    if (__builtin_expect(sus != 0, 0)) quote_check(W, g, v, pos0);
    if (__builtin_expect(pos0 <= rb || pos0 + 16 > tend, 0)) {
        // chunk on an edge of the text (the first / last chunk of the whole text only)
        nl &= range_mask16(pos0, rb, g.L);
        const uint32_t tmask = range_mask16(pos0, rb, tend);
        if (g.rb >= 0 && pos0 == g.rb) carry = 1;  // virtual newline in front of the text
        if (pos0 < rb) carry = 0;
        uint32_t prev = ((nl << 1) | carry) & 0xFFFFu;
        if (g.rb > pos0 && g.rb < pos0 + 16) prev |= 1u << (g.rb - pos0);
        start = prev & ~nl & tmask;
        end = nl & ~prev;
    } else {
        const uint32_t prev = ((nl << 1) | carry) & 0xFFFFu;
        start = prev & ~nl;
        end = nl & ~prev;
    }
    nl_out = nl;
}

// Builds row_s / row_e for the staged window.  Returns false (uniformly) when the row table would overflow,
// which can only happen when some row is shorter than 26 bytes, i.e. malformed.
template <int NWARPS>
__device__ bool scan_rows(WindowIndex& W, const WinGeom& g) {
    const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
    const int cbeg = g.vb >> 4;  // chunks in front of the first valid byte are not scanned
    const int nchunks = (g.L + 15) >> 4;
    const int cpw = ((((nchunks - cbeg) + NWARPS - 1) / NWARPS) + 31) & ~31;  // chunks per warp, whole 32-lane rounds
    const int c0 = cbeg + w * cpw;
    const int c1 = c0 + cpw < nchunks ? c0 + cpw : nchunks;
    const int rb = g.vb;
    // ---- pass 1: classify; every warp compacts the row starts / ends of its chunk range ------------------------------
    // A valid row is >= 26 bytes, so a 16-byte chunk holds at most one row start and one row end; a chunk with more
    // proves a malformed (too short) row and is reported as such.  So a warp's k-th row start can be parked at index
    // c0 + k of the (chunk-indexed) scratch array: it never overtakes the chunk it came from.
    const uint32_t lt = (1u << lane) - 1u;
    int ns = 0, ne = 0;
    bool crowded = false;
    uint32_t carry_in = 0;
    if (c0 < c1 && (c0 << 4) > rb) carry_in = W.win[(c0 << 4) - 1] == '\n';
    for (int base = c0; base < c1; base += 32) {
        const int c = base + lane;
        uint32_t nl = 0, s = 0, e = 0;
        const bool live = c < c1;
        if (live) classify_chunk(W, g, c, 0u, nl, s, e);  // provisional carry 0, bit 0 is patched below
        const uint32_t up = __shfl_up_sync(0xffffffffu, nl, 1);
        const uint32_t carry = lane == 0 ? carry_in : (up >> 15) & 1u;
        carry_in = (__shfl_sync(0xffffffffu, nl, 31) >> 15) & 1u;
        const int pos0 = c << 4;
        if (live && carry && pos0 >= rb && !(g.rb >= 0 && pos0 == g.rb)) {
            // byte 0 of the chunk follows a newline: it starts a row unless it is a newline itself, and a newline
            // there does not end a row (empty line)
            const int tend = g.re < g.loaded ? g.re : g.loaded;
            if (!(nl & 1u) && pos0 < tend) s |= 1u;
            e &= ~1u;
        }
        crowded |= (s & (s - 1)) != 0 || (e & (e - 1)) != 0;
        const uint32_t bs = __ballot_sync(0xffffffffu, s != 0), be = __ballot_sync(0xffffffffu, e != 0);
        // (all chunks of this round are classified before anything is parked: the parking slots c0+ns.. lie at or
        // below the chunks of this round, whose own masks are not needed any more)
        if (s) W.startm[c0 + ns + __popc(bs & lt)] = (uint16_t)(pos0 + __ffs(s) - 1);
        if (e) W.endm[c0 + ne + __popc(be & lt)] = (uint16_t)(pos0 + __ffs(e) - 1);
        ns += __popc(bs);
        ne += __popc(be);
    }
    crowded = __any_sync(0xffffffffu, crowded);
    if (lane == 0) W.warp_cnt[w] = ns | (ne << 16) | (crowded ? 0x80000000 : 0);
    __syncthreads();
    int so = 0, eo = 0, ts = 0, te = 0;
    bool any_crowded = false;
#pragma unroll
    for (int i = 0; i < NWARPS; i++) {
        const int v = W.warp_cnt[i];
        any_crowded |= v < 0;
        if (i < w) {
            so += v & 0xFFFF;
            eo += (v >> 16) & 0x7FFF;
        }
        ts += v & 0xFFFF;
        te += (v >> 16) & 0x7FFF;
    }
    if (tid == 0) {
        W.n_starts = ts;
        W.n_ends = te;
    }
    if (any_crowded || ts > kRowCap || te > kRowCap) {
        __syncthreads();
        return false;
    }
    // ---- pass 2: move every warp's compacted positions to their place in the row table -------------------------------
    for (int i = lane; i < ns; i += 32) W.row_s[so + i] = W.startm[c0 + i];
    for (int i = lane; i < ne; i += 32) W.row_e[eo + i] = W.endm[c0 + i];
    __syncthreads();
    return true;
}

// Writes the virtual newline that terminates an unterminated last row (final chunk only) and fixes g.L.
__device__ __forceinline__ void finish_geom(WindowIndex& W, WinGeom& g, bool final_chunk) {
    if (final_chunk && g.covers_eof && g.re > 0 && g.re <= g.loaded) {
        bool needs = (g.re - 1 >= 0) && W.win[g.re - 1] != '\n' && g.re - 1 >= g.vb;
        if (needs) {
            __syncthreads();
            if (threadIdx.x == 0) W.win[g.re] = '\n';  // win has 128 spare bytes
            g.L = g.re + 1;
            __syncthreads();
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// tile kernel (streaming)
//
// The text [begin, end) is cut into gridDim.x contiguous segments; CTA c owns every query run whose first row starts
// inside segment c.  It walks its segment in windows of kTile bytes, staged into shared memory by TMA bulk copies
// into two alternating buffers (the copy of window k+1 runs under the row / run phases of window k).  A window
// starts 16-byte aligned just in front of the last complete row of the previous window (that row is the
// "look-behind" row: it is only there so that the first new row can be compared with its predecessor), so no row
// and no query ever has to fit a window: the unfinished query is carried in shared memory (row count, best bit
// score, its top rows so far) from window to window, and a CTA keeps walking past the end of its segment until the
// query it has open ends.  Phases of one window (three CTA-wide barriers):
//   B  classify   every lane owns 32 bytes: SIMD-in-register byte tests (ASCII fast path) -> newline / tab / digit
//                 bitmasks (dp4a packing); row starts compacted per 1 KB round with two ballots; the last warp out
//                 computes the prefix over the rounds' row counts
//   D  rows       one thread per row (global row index -> round by a shuffle search): validation + truncated bit score
//                 (parse_row_lean, falling back to the full grammar), packed field positions, head flag (query id
//                 differs from the previous row's); meanwhile the last warp of the CTA finds the last newline and
//                 requests the next window's bytes
//   E  runs       one warp per query run (static assignment): extent, best bit score, top rows (ballot compaction);
//                 the warp decides the run's fate (finished / still open / block path), merges it with the carried
//                 query, takes top-row slots from the CTA's slab and a place in the record buffer (shared-memory
//                 atomics), folds its top rows' digits (top_row_from_info) and writes them to HBM / the carry
//      then       thread 0: a new slot slab when the current one runs low, record flush (one global atomic per >= 128
//                 headers), where the CTA goes next
// ---------------------------------------------------------------------------------------------------------------
constexpr int kSWarps = kTileThreads / 32;
constexpr int kUnits = kTile / 32 + 1;       // 32-byte units of a window (+1: the unit that can hold a virtual final newline)
constexpr int kRounds = kTile / 1024;        // phase B works in rounds of 32 units = 1 KB (the extra unit of a virtual final
                                             // newline goes to the last round)
constexpr int kSegCap = 32 * 2 + 4;           // row starts one round can find (<= 1 per 16 bytes, else malformed)
constexpr int kSRowCap = kTile / 26 + 8;     // a valid row is >= 26 bytes
constexpr int kRecFlush = 128;                // buffered record headers that trigger a flush (one global atomic)
constexpr int kRecBuf = 256;                  // capacity of the record buffer (beyond: the run reserves its record itself)
constexpr uint32_t kSlotSlab = 1024;          // top-row slots a CTA reserves at a time (one global atomic)
constexpr uint32_t kSlotLow = 64;             // a new slab is fetched at the end of a window that leaves fewer free slots
constexpr int kCarryTop = 32;                 // top rows of the open query kept in shared memory (beyond: block path)

static_assert(kSRowCap < 0x8000, "row indices are 15-bit");
static_assert(kTile % 32 == 0 && kTile + 128 < 65536, "window offsets are 16-bit");
static_assert(kTile % 1024 == 0 && kRounds <= 32, "per-round tables are read by one warp");

enum : uint32_t { RK_EMIT = 1, RK_DEFER = 2, RK_PSEUDO = 4, RK_OPEN = 8, RK_NEWCARRY = 16 };

struct CarryRun {
    unsigned long long head_abs;  // offset of the open query's first row
    uint32_t qlen;
    uint32_t nrows;
    int32_t mx;        // best truncated bit score so far
    int32_t g;         // top rows so far (kept in tops[])
    int32_t open;
    int32_t deferred;  // the query goes to the block path when it ends (top group too large / bit score beyond int32)
    TopRowRaw tops[kCarryTop];  // its top rows so far
};

struct StagedRec {
    unsigned long long abs;
    uint32_t qlen, nrows;
    int32_t mx;
    uint32_t slot;  // first top-row slot
    uint32_t gtot;
    uint32_t pad;
};

struct WinDesc {  // what a window covers; written one window ahead, together with the request for its bytes
    unsigned long long lo;  // text offset of win[0] (16-byte aligned)
    int loaded, tend, vb;   // bytes staged; end / begin of the text inside the window
    int flags;              // 1: the window reaches the end of the text, 2: it contains the first byte of the text
};

struct WinGeo {  // row geometry of a window; written by the last warp that leaves phase B
    unsigned long long next_lo;
    int n_starts, n_complete, last_nl, may_continue, crowded;
    int inc[32], nfirst[32], plast[32];  // per warp share: rows up to and including it; first row start behind / last before it
};

struct StreamSmem {
    alignas(128) uint8_t win[2][kTile + 128];
    // byte-class bitmasks, bit i of word u = byte 32u+i of the window (read by the row parsers as 32/64-bit words)
    alignas(8) uint32_t tabm[kUnits + 7];
    alignas(8) uint32_t digm[kUnits + 7];
    alignas(8) uint32_t nlm[kUnits + 7];
    uint16_t seg[kRounds][kSegCap];  // row starts found in round i, in order
    uint16_t row_s[kSRowCap + 2];    // all row starts of the window (written by phase D)
    int32_t bits[kSRowCap];
    uint8_t flags[kSRowCap];         // bit1: bit score does not fit int32
    uint32_t rowinfo[kSRowCap];      // packed tab positions of the row (parse_row_lean), 0: unknown
    uint32_t headw[(kSRowCap + kTileThreads) / 32 + 2];  // head flags, one bit per row (the row loop writes whole rounds)
    uint32_t runs[kSRowCap + 1];     // head rows in arrival order: row | id length << 16 (bit 15: continuation of the carried query)
    StagedRec rec_buf[kRecBuf];  // record headers of finished queries, waiting for the next flush
    uint16_t stage[kSWarps][32];
    CarryRun carry[2];
    alignas(8) unsigned long long mbar[2];
    int warp_cnt[32], warp_first[32], warp_last[32];  // per round of phase B: row starts, the first and the last one
    WinDesc wd[2];
    WinGeo geo;
    int b_done;  // warps that have finished phase B of the current window
    int bad_byte, has_blank, crowded;
    int n_runs, next_run, n_skip, term, new_open;
    uint32_t rec_base;
    uint32_t slot_cur, slot_end;  // the CTA's current slab of top-row slots: [slot_cur, slot_end) is free
    int rec_cnt;                  // record headers in rec_buf
    uint32_t slots_taken;         // slots this CTA has reserved in slabs so far (sizes the next slab)
    int out_ok;                   // 0: an output capacity was exceeded (the host grows the arrays and reruns)
};

static_assert((sizeof(StreamSmem) + 1024) * kTileCtasPerSm <= 227 * 1024, "tile CTAs must fit one SM");

__device__ __forceinline__ void push_defer(const RunParams& p, unsigned long long off, unsigned check_prev) {
    unsigned i = atomicAdd(&p.ctr->n_defer, 1u);
    if (i < p.defer_cap)
        p.defer[i] = (off << 1) | check_prev;
    else
        p.ctr->cap_overflow = 1;
}

__device__ __forceinline__ void stream_issue_load(StreamSmem& S, int buf, const uint8_t* text, unsigned long long lo, int bytes) {
    fence_proxy_async();
    mbar_expect_tx(&S.mbar[buf], (uint32_t)bytes);
    for (int o = 0; o < bytes; o += 16384) {
        const int n = bytes - o < 16384 ? bytes - o : 16384;
        tma_bulk_g2s(S.win[buf] + o, text + lo + o, (uint32_t)n, &S.mbar[buf]);
    }
}

// dp4a gathers the 0x80 flags of two words into 128 * (8-bit mask)
__device__ __forceinline__ uint32_t pack8(uint32_t a, uint32_t b) { return __dp4a(a, 0x08040201u, __dp4a(b, 0x80402010u, 0u)); }
__device__ __forceinline__ uint32_t pack32(const uint32_t (&f)[8]) {
    const uint32_t b0 = pack8(f[0], f[1]) + (pack8(f[2], f[3]) << 8);  // 128 * 16-bit mask
    const uint32_t b1 = pack8(f[4], f[5]) + (pack8(f[6], f[7]) << 8);
    return __funnelshift_r(b0 + (b1 << 16), b1 >> 16, 7);
}

// Exact byte-wise classification of one 32-byte unit: text edges (bytes outside [vb, tend) are not text), units with
// non-ASCII bytes, and the exact '"' / '\r' search behind the cheap "suspicious byte" test.  Publishes the unit's tab /
// digit / newline masks when `publish` is set and returns the newline mask.
__device__ __noinline__ uint32_t classify_unit_slow(StreamSmem& S, const uint8_t* w, int pos0, int vb, int tend, int virt_nl_at, bool publish) {
    uint32_t nl = 0, tab = 0, dig = 0;
    int bad = INT_MAX;
    for (int k = 0; k < 32; k++) {
        const int pos = pos0 + k;
        if (pos < vb || pos >= tend) continue;
        const uint32_t c = w[pos];
        if (c == '\n') nl |= 1u << k;
        if (c == '\t') tab |= 1u << k;
        if (c - '0' <= 9u) dig |= 1u << k;
        if ((c == '"' || c == '\r') && pos < bad) bad = pos;
    }
    if (bad != INT_MAX) atomicMin(&S.bad_byte, bad);
    if (virt_nl_at >= pos0 && virt_nl_at < pos0 + 32) nl |= 1u << (virt_nl_at - pos0);
    if (publish) {
        const int u = pos0 >> 5;
        S.tabm[u] = tab;
        S.digm[u] = dig;
        S.nlm[u] = nl;
    }
    return nl;
}

// first head row at index >= from, or -1 (whole warp)
__device__ __forceinline__ int next_head(const StreamSmem& S, int from, int n_rows, int lane) {
    const int nwords = (n_rows + 31) >> 5;
    for (int k0 = from >> 5; k0 < nwords; k0 += 32) {
        const int k = k0 + lane;
        uint32_t w = k < nwords ? S.headw[k] : 0u;
        if (k == (from >> 5)) w &= ~0u << (from & 31);
        const unsigned bal = __ballot_sync(0xffffffffu, w != 0);
        if (bal) {
            const int src = __ffs(bal) - 1;
            const uint32_t wv = __shfl_sync(0xffffffffu, w, src);
            return ((k0 + src) << 5) + __ffs(wv) - 1;
        }
    }
    return -1;
}

// end (position of the newline) of the row that starts at window offset s, through the newline mask
__device__ __noinline__ int row_end_search(const StreamSmem& S, int s, int limit) {
    for (int u = s >> 5; (u << 5) <= limit; u++) {
        uint32_t w = S.nlm[u];
        if (u == (s >> 5)) w &= ~0u << (s & 31);
        if (w) return (u << 5) + __ffs(w) - 1;
    }
    return limit;
}

// Phase B for one round: units [u0, u1) of the window (32 of them; the last round may have one more).  kInterior: every
// unit is whole text (no edge of the input in the window), so the per-unit edge tests are compiled out.
template <bool kInterior>
__device__ __forceinline__ void classify_round(StreamSmem& S, const uint8_t* win, int round, int u0, int u1, int vb, int tend, bool has_begin,
                                               bool virt_nl, int lane) {
    const unsigned FULL = 0xffffffffu;
    const uint32_t lt = (1u << lane) - 1u;
    int cnt = 0;
    uint32_t crowd = 0, blank = 0;
    uint32_t carry_in = 0;
    if (u0 < u1 && u0 > 0 && (u0 << 5) - 1 >= vb) carry_in = win[(u0 << 5) - 1] == '\n';
    uint16_t* const seg = S.seg[round];
    for (int base = u0; base < u1; base += 32) {
        const int u = base + lane;
        const int pos0 = u << 5;
        uint32_t nl = 0;
        const bool live = kInterior || u < u1;
        const bool edge = !kInterior && ((has_begin && pos0 <= vb) || pos0 + 32 > tend);  // first / last bytes of the text
        if (live) {
            const uint4 v0 = *reinterpret_cast<const uint4*>(win + pos0);
            const uint4 v1 = *reinterpret_cast<const uint4*>(win + pos0 + 16);
            const uint32_t x[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
            const uint32_t hi = (x[0] | x[1] | x[2] | x[3] | x[4] | x[5] | x[6] | x[7]) & 0x80808080u;
            if (__builtin_expect(edge || hi != 0, 0)) {
                nl = classify_unit_slow(S, win, pos0, vb, tend, virt_nl ? tend : -1, true);
            } else {
                // ASCII bytes: per-byte sums stay below 0x100, so plain 32-bit adds classify four bytes at once
                uint32_t fn[8], ft[8], fd[8], sus = 0;
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    uint32_t tu = (x[k] ^ 0x09090909u) + 0x7F7F7F7Fu;  // bit 7 clear <=> tab
                    uint32_t nu = (x[k] ^ 0x0A0A0A0Au) + 0x7F7F7F7Fu;  // bit 7 clear <=> newline
                    // (opaque to the optimiser: it would otherwise re-derive ~tu / ~nu with a second add per word)
                    asm("" : "+r"(tu));
                    asm("" : "+r"(nu));
                    const uint32_t ge30 = x[k] + 0x50505050u, ge3a = x[k] + 0x46464646u, ge23 = x[k] + 0x5D5D5D5Du;
                    ft[k] = ~tu & 0x80808080u;
                    fn[k] = ~nu & 0x80808080u;
                    fd[k] = ge30 & ~ge3a & 0x80808080u;
                    sus |= ~ge23 & tu & nu;  // below '#' and neither tab nor newline: look closer
                }
                nl = pack32(fn);
                S.tabm[u] = pack32(ft);
                S.digm[u] = pack32(fd);
                S.nlm[u] = nl;
                if (__builtin_expect((sus & 0x80808080u) != 0, 0)) classify_unit_slow(S, win, pos0, vb, tend, -1, false);  // exact '"' / '\r' search
            }
        }
        // row starts: a non-newline byte behind a newline (or behind the virtual newline in front of the text)
        const uint32_t upv = __shfl_up_sync(FULL, nl, 1);
        const uint32_t carry = lane == 0 ? carry_in : (upv >> 31);
        carry_in = __shfl_sync(FULL, nl, 31) >> 31;
        uint32_t prev = (nl << 1) | carry;
        uint32_t st;
        if (__builtin_expect(edge, 0)) {
            const uint32_t valid_lo = pos0 >= vb ? ~0u : (vb - pos0 >= 32 ? 0u : ~0u << (vb - pos0));
            const uint32_t valid_hi = pos0 + 32 <= tend ? ~0u : (tend <= pos0 ? 0u : ~0u >> (32 - (tend - pos0)));
            if (pos0 <= vb) prev &= valid_lo;  // the byte in front of vb is not text
            if (has_begin && vb >= pos0 && vb < pos0 + 32) prev |= 1u << (vb - pos0);
            st = prev & ~nl & valid_lo & valid_hi;
        } else
            st = prev & ~nl;
        blank |= nl & prev;
        // a valid row is >= 26 bytes: at most one start per 16 bytes, compacted with one ballot per half
        const uint32_t sl = st & 0xFFFFu, sh = st >> 16;
        crowd |= (sl & (sl - 1u)) | (sh & (sh - 1u));
        const unsigned bl = __ballot_sync(FULL, sl != 0), bh = __ballot_sync(FULL, sh != 0);
        int idx = cnt + __popc(bl & lt) + __popc(bh & lt);
        if (sl) seg[idx++] = (uint16_t)(pos0 + __ffs(sl) - 1);
        if (sh) seg[idx] = (uint16_t)(pos0 + 16 + __ffs(sh) - 1);
        cnt += __popc(bl) + __popc(bh);
    }
    if (__any_sync(FULL, blank != 0) && lane == 0) S.has_blank = 1;
    if (__any_sync(FULL, crowd != 0) && lane == 0) S.crowded = 1;
    __syncwarp();
    if (lane == 0) {
        S.warp_cnt[round] = cnt;
        S.warp_first[round] = cnt ? (int)seg[0] : -1;
        S.warp_last[round] = cnt ? (int)seg[cnt - 1] : -1;
    }
}

__device__ __forceinline__ void write_record(blu_record* dst, const StagedRec& sr) {
    blu_record rec;
    rec.query_off = sr.abs;
    rec.query_len = sr.qlen;
    rec.n_rows = sr.nrows;
    rec.keep_mask = 0;
    rec.perc_identity = 0.0;
    rec.bit_score = (int64_t)sr.mx;
    rec.ref_lineage = 0;
    rec.slot_base = sr.slot;
    rec.n_beans = 0;
    rec.n_accessions = sr.gtot;  // size of the top group until the consensus kernel overwrites it
    rec.status = 2;              // waiting for the consensus kernel
    rec.single_match = 0;
    rec.mutated = 0;
    rec.reached_pos = 0;
    rec.allowed_pos = -1;
    rec.bean_level = 0;
    rec.pad[0] = rec.pad[1] = 0;
    *dst = rec;
}

// Flushes the buffered record headers: ONE global atomic reserves their (dense) place in the record array, then every
// thread writes headers.  All threads of the CTA call; two barriers.
__device__ __forceinline__ void flush_records(const RunParams& p, StreamSmem& S, int tid) {
    const int n = S.rec_cnt < kRecBuf ? S.rec_cnt : kRecBuf;
    if (n == 0) return;
    if (tid == 0) {
        const unsigned long long rs = atomicAdd(&p.ctr->rec_slots, (unsigned long long)n << 32);
        const uint32_t rec_base = (uint32_t)(rs >> 32);
        S.rec_base = rec_base;
        if ((unsigned long long)rec_base + (unsigned long long)n > p.rec_cap) {
            S.out_ok = 0;
            p.ctr->cap_overflow = 1;
        }
    }
    __syncthreads();
    if (S.out_ok) {
        const uint32_t rec_base = S.rec_base;
        for (int i = tid; i < n; i += kTileThreads) write_record(p.records + rec_base + i, S.rec_buf[i]);
    }
    __syncthreads();
    if (tid == 0) S.rec_cnt = 0;
}

// One thread: describes the window that starts at `lo` in S.wd[buf] and asks the TMA unit for its bytes.
__device__ __forceinline__ void describe_and_load(StreamSmem& S, int buf, const RunParams& p, unsigned long long lo, unsigned long long up) {
    const int loaded = (int)(up - lo < (unsigned long long)kTile ? up - lo : (unsigned long long)kTile);
    WinDesc d;
    d.lo = lo;
    d.loaded = loaded;
    d.tend = (int)(p.end - lo < (unsigned long long)loaded ? p.end - lo : (unsigned long long)loaded);
    const bool has_begin = lo <= p.begin;
    d.vb = has_begin ? (int)(p.begin - lo) : 0;
    d.flags = (p.end <= lo + (unsigned long long)loaded ? 1 : 0) | (has_begin ? 2 : 0);
    S.wd[buf] = d;
    stream_issue_load(S, buf, p.text, lo, loaded);
}

#ifdef BLU_PHASE_CLOCKS
#define PCLK(i)                                  \
    {                                            \
        long long _t;                            \
        (void)*(volatile int*)&S.n_runs; /* a shared-memory access: the warp really is past the barrier */ \
        asm volatile("mov.u64 %0, %%clock64;" : "=l"(_t)::"memory"); \
        pc[i] += _t - pt;                        \
        pt = _t;                                 \
    }
#define RCLK(i)                                  \
    {                                            \
        long long _t;                            \
        asm volatile("mov.u64 %0, %%clock64;" : "=l"(_t)::"memory"); \
        rc[i] += _t - rt;                        \
        rt = _t;                                 \
    }
#else
#define PCLK(i)
#define RCLK(i)
#endif

__global__ void __launch_bounds__(kTileThreads, kTileCtasPerSm) tile_kernel(const __grid_constant__ RunParams p) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    StreamSmem& S = *reinterpret_cast<StreamSmem*>(smem_raw);
#ifdef BLU_PHASE_CLOCKS
    long long pc[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    long long pt = clock64();
    int n_win = 0;
    long long rc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long rt = 0;
#endif
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const unsigned FULL = 0xffffffffu;
    if (p.end <= p.begin) return;
    // ---- this CTA's segment --------------------------------------------------------------------------------------
    const unsigned long long total = p.end - p.begin;
    unsigned long long seg = (total + gridDim.x - 1) / gridDim.x;
    if (seg < 4ull * kTile) seg = 4ull * kTile;
    const unsigned long long seg_lo = p.begin + (unsigned long long)blockIdx.x * seg;
    if (seg_lo >= p.end) return;
    const unsigned long long seg_hi = (seg_lo + seg < p.end) ? seg_lo + seg : p.end;
    const unsigned long long b16 = p.begin & ~15ull;
    const unsigned long long up = (p.end + 15ull) & ~15ull;

    if (tid == 0) {
        mbar_init(&S.mbar[0], 1);
        mbar_init(&S.mbar[1], 1);
        S.carry[0].open = S.carry[1].open = 0;
        S.bad_byte = INT_MAX;
        S.has_blank = S.crowded = 0;
        S.n_runs = S.next_run = S.n_skip = S.term = S.new_open = 0;
        S.b_done = 0;
        {
            // the CTA's first slab of top-row slots (about 40 per window of the segment, at most kSlotSlab)
            const unsigned long long est = ((seg_hi - seg_lo) / (unsigned long long)kTile + 1ull) * 40ull + (unsigned long long)kSlotLow;
            const uint32_t slab = est < (unsigned long long)kSlotSlab ? (uint32_t)est : kSlotSlab;
            const unsigned long long rs = atomicAdd(&p.ctr->rec_slots, (unsigned long long)slab);
            S.slot_cur = (uint32_t)rs;
            S.slot_end = (uint32_t)rs + slab;
            S.slots_taken = slab;
        }
        S.rec_cnt = 0;
        S.out_ok = 1;
        if ((unsigned long long)S.slot_end > (unsigned long long)p.slot_cap) {
            S.out_ok = 0;
            p.ctr->cap_overflow = 1;
        }
    }
    for (int i = tid; i < 7; i += kTileThreads) S.tabm[kUnits + i] = S.digm[kUnits + i] = S.nlm[kUnits + i] = 0u;
    __syncthreads();
    uint32_t ph0 = 0, ph1 = 0;
    int cur = 0;  // which carry buffer holds the open query
    int win_idx = 1;  // number of the window (epoch of S.new_open)
    unsigned long long own_from = seg_lo;  // rows that start at or after this offset have not been processed yet
    int buf = 0;
    if (tid == 0) {
        unsigned long long lo0 = seg_lo > p.begin + kBack ? (seg_lo - kBack) & ~15ull : b16;
        if (lo0 < b16) lo0 = b16;
        describe_and_load(S, 0, p, lo0, up);
    }
    __syncthreads();

    while (true) {
        const uint8_t* const win = S.win[buf];
        const unsigned long long lo = S.wd[buf].lo;
        const int tend = S.wd[buf].tend, vb = S.wd[buf].vb;
        const bool covers_eof = (S.wd[buf].flags & 1) != 0, has_begin = (S.wd[buf].flags & 2) != 0;
        if (tid == 32) {
            // the window after the next one: start moving it from HBM to L2
            const unsigned long long nb = lo + 2ull * kTile - 512ull;
            if (nb < up && (nb < seg_hi + kTile)) {
                const unsigned long long n = up - nb < (unsigned long long)kTile ? up - nb : (unsigned long long)kTile;
                tma_prefetch_l2(p.text + nb, (uint32_t)n);
            }
        }
        // wait for this window's bytes
        {
            uint32_t& ph = buf ? ph1 : ph0;
            while (!mbar_try_wait(&S.mbar[buf], ph)) {
            }
            ph ^= 1;
        }
        PCLK(0)
        // an unterminated last row of the input is closed by a virtual newline at `tend`
        const bool virt_nl = p.final_chunk && covers_eof && tend > vb && win[tend - 1] != '\n';
        const int scan_len = tend + (virt_nl ? 1 : 0);
        const int n_units = (scan_len + 31) >> 5;

        // ---- phase B: classify + row starts, 1 KB rounds ------------------------------------------------------------------
        {
            const bool interior = !has_begin && tend == kTile && !virt_nl;
            for (int rd = warp; rd < kRounds; rd += kSWarps) {
                const int u0 = rd << 5;
                if (interior)
                    classify_round<true>(S, win, rd, u0, u0 + 32, 0, tend, false, false, lane);
                else {
                    const int u1 = (rd == kRounds - 1 || u0 + 32 > n_units) ? n_units : u0 + 32;
                    classify_round<false>(S, win, rd, u0, u1 > u0 ? u1 : u0, vb, tend, has_begin, virt_nl, lane);
                }
            }
        }
        PCLK(1)
        // ---- row tables of the window: the prefix over the rounds is computed ONCE, by the last warp that leaves
        //      phase B; everything else (last newline, next window) is worked out behind the barrier by the last warp of the
        //      CTA, which normally has no rows in phase D, while the others already parse rows
        {
            int ticket = 0;
            __syncwarp();
            if (lane == 0) {
                __threadfence_block();
                ticket = atomicAdd(&S.b_done, 1);
            }
            ticket = __shfl_sync(FULL, ticket, 0);
            if (ticket == kSWarps - 1) {
                __threadfence_block();
                // lane i < kRounds holds the tables of round i
                const int c_i = lane < kRounds ? S.warp_cnt[lane] : 0;
                int inc_i = c_i;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const int t = __shfl_up_sync(FULL, inc_i, d);
                    if (lane >= d) inc_i += t;
                }
                const int n_starts = __shfl_sync(FULL, inc_i, 31);
                const unsigned ne = __ballot_sync(FULL, c_i > 0);
                const unsigned below = ne & ((1u << lane) - 1u), above = lane >= 31 ? 0u : (ne & (~0u << (lane + 1)));
                const int wl = lane < kRounds ? S.warp_last[lane] : -1, wf = lane < kRounds ? S.warp_first[lane] : -1;
                int plast_i = __shfl_sync(FULL, wl, below ? 31 - __clz(below) : 0);
                if (!below) plast_i = -1;
                int nfirst_i = __shfl_sync(FULL, wf, above ? __ffs(above) - 1 : 0);
                if (!above) nfirst_i = -1;
                if (lane < kRounds) {
                    S.geo.inc[lane] = inc_i;
                    S.geo.nfirst[lane] = nfirst_i;
                    S.geo.plast[lane] = plast_i;
                }
                if (lane < 4) S.tabm[n_units + lane] = S.digm[n_units + lane] = S.nlm[n_units + lane] = 0u;
                if (lane == 0) {
                    S.geo.n_starts = n_starts;
                    S.geo.crowded = (S.crowded != 0 || n_starts > kSRowCap) ? 1 : 0;
                    S.b_done = 0;
                }
            }
        }
        __syncthreads();
        PCLK(2)
        const int n_starts = S.geo.n_starts;
        const bool blank = S.has_blank != 0;
        if (S.geo.crowded) {
            // a row shorter than 16 bytes / more rows than 26-byte rows fit: malformed input
            if (tid == 0) report(p.ctr, DE_BAD_FIELD_COUNT, lo);
            break;  // (no copy in flight: the next window has not been requested)
        }
        if (tid == 0 && S.bad_byte != INT_MAX) report(p.ctr, DE_QUOTE_OR_CR, lo + (unsigned)S.bad_byte);
        const int c_i = lane < kRounds ? S.warp_cnt[lane] : 0;
        const int inc_i = lane < kRounds ? S.geo.inc[lane] : n_starts;
        const int nfirst_i = lane < kRounds ? S.geo.nfirst[lane] : -1, plast_i = lane < kRounds ? S.geo.plast[lane] : -1;
        if (warp == kSWarps - 1) {
            // ---- last newline, last complete row, the next window (its bytes are requested now) ---------------------------
            const unsigned ne = __ballot_sync(FULL, c_i > 0);
            int last_start = __shfl_sync(FULL, lane < kRounds ? S.warp_last[lane] : -1, ne ? 31 - __clz(ne) : 0);
            if (!ne) last_start = -1;
            int last_nl = -1;
            for (int ub = n_units - 1; ub >= 0 && last_nl < 0; ub -= 32) {
                const int u = ub - lane;
                const uint32_t w = u >= 0 ? S.nlm[u] : 0u;
                const unsigned bal = __ballot_sync(FULL, w != 0);
                if (bal) {
                    const int src = __ffs(bal) - 1;
                    const uint32_t wv = __shfl_sync(FULL, w, src);
                    last_nl = ((ub - src) << 5) + 31 - __clz(wv);
                }
            }
            const int n_complete = (n_starts > 0 && last_start > last_nl) ? n_starts - 1 : n_starts;
            // the last complete row: the look-behind row of the next window
            int lc = last_start;
            if (n_complete != n_starts) {
                const int wlast = 31 - __clz(ne);  // (ne != 0: n_starts > 0)
                const int cl = __shfl_sync(FULL, c_i, wlast);
                const unsigned rest = ne & ~(1u << wlast);
                const int w2 = rest ? 31 - __clz(rest) : 0;
                const int c2 = __shfl_sync(FULL, c_i, w2);
                lc = cl >= 2 ? (int)S.seg[wlast][cl - 2] : (rest ? (int)S.seg[w2][c2 - 1] : -1);
            }
            if (lane == 0) {
                const bool progress = lc >= 0 && lo + (unsigned long long)lc >= own_from;  // a complete row this CTA had not seen yet
                unsigned long long next_lo = lo;
                if (progress) {
                    const unsigned long long la = lo + (unsigned long long)lc;
                    next_lo = la > p.begin ? (la - 1) & ~15ull : b16;
                    if (next_lo < b16) next_lo = b16;
                }
                const bool may_continue = progress && !covers_eof && next_lo > lo;
                S.geo.n_complete = n_complete;
                S.geo.last_nl = last_nl;
                S.geo.may_continue = may_continue ? 1 : 0;
                if (n_complete == n_starts) S.row_s[n_starts] = (uint16_t)(last_nl + 1);  // sentinel: end of the last row
                if (may_continue) describe_and_load(S, buf ^ 1, p, next_lo, up);
            }
        }
        const uint64_t* tabw = reinterpret_cast<const uint64_t*>(S.tabm);
        const uint64_t* digw = reinterpret_cast<const uint64_t*>(S.digm);
        PCLK(3)
        // ---- phase D: one thread per row -------------------------------------------------------------------------------
        for (int rb = 0; rb < n_starts; rb += kTileThreads) {
            const int r = rb + tid;
            const bool live = r < n_starts;
            const int rr = live ? r : n_starts - 1;
            // which round found row rr: number of rounds that end at or before it
            int w = 0;
#pragma unroll
            for (int step = 16; step; step >>= 1) {
                const int cand = w + step;
                const int v = __shfl_sync(FULL, inc_i, (cand - 1) & 31);
                if (cand <= kRounds && v <= rr) w = cand;
            }
            const int cw = __shfl_sync(FULL, c_i, w);
            const int k = rr - (__shfl_sync(FULL, inc_i, w) - cw);
            const int nf = __shfl_sync(FULL, nfirst_i, w), pl = __shfl_sync(FULL, plast_i, w);
            bool head = false, skip = false;
            // part 1: validate + parse the row (the lanes of a warp part ways here: evalue shapes, the rare full-grammar path)
            bool parsed = false;
            int s = 0, ql = 0, prv = -1;
            unsigned long long abs = 0;
            if (live) {
                const uint16_t* const seg_w = S.seg[w];
                s = seg_w[k];
                S.row_s[r] = (uint16_t)s;
                abs = lo + (unsigned long long)s;
                if (abs < own_from)
                    skip = true;  // look-behind rows: seen by the previous window / owned by the previous segment
                else {
                    // the row's newline: in front of the next row start; the window's last start (and any row of a window
                    // with blank lines) looks it up in the newline mask -- none: an unterminated row, left to the next window
                    const int nxt = k + 1 < cw ? (int)seg_w[k + 1] : nf;
                    const int e = (nxt < 0 || blank) ? row_end_search(S, s, scan_len) : nxt - 1;
                    if (e < scan_len) {
                        int64_t bits;
                        uint32_t info;
                        if (!parse_row_lean(win, S.tabm, S.digm, s, e, bits, ql, info)) {
                            const LightRow lr = parse_row_masked(win, tabw, digw, s, e);
                            if (lr.err) report(p.ctr, lr.err, abs);
                            bits = lr.bits;
                            ql = lr.q_len;
                            info = 0;
                        }
                        S.rowinfo[r] = info;
                        const int32_t b32 = (int32_t)bits;
                        S.flags[r] = ((int64_t)b32 != bits) ? 2 : 0;
                        S.bits[r] = b32;
                        prv = k > 0 ? (int)seg_w[k - 1] : pl;
                        parsed = true;
                    }
                }
            }
            // part 2, with the warp back together (without this the two halves of a warp that took different branches above run
            // the compare one after the other): head flag = the query id differs from the previous row's
            __syncwarp();
            if (parsed) {
                if (prv < 0) {
                    head = has_begin;  // first row of the text; else: predecessor not in the window
                    if (!head) push_defer(p, abs, 1);  // the block path decides whether it starts a query
                } else
                    head = !same_qid_lean(win, S.tabm, s, ql, prv);
                if (head) {
                    if (abs < seg_hi) {
                        const int i = atomicAdd(&S.n_runs, 1);
                        S.runs[i] = (uint32_t)r | ((uint32_t)(ql > 0xFFFF ? 0xFFFF : ql) << 16);
                    } else
                        S.term = 1;  // a query of the next segment starts here: this CTA ends with this window
                }
            }
            const unsigned hb = __ballot_sync(FULL, head);
            if (lane == 0) S.headw[r >> 5] = hb;
            const unsigned sb = __ballot_sync(FULL, skip);
            if (sb && lane == 0) atomicAdd(&S.n_skip, __popc(sb));
        }
        PCLK(4)
        __syncthreads();
        PCLK(5)
        // ---- phase E: one warp per query run (static assignment) -----------------------------------------------------------
        //   extent, best bit score, top rows (ballot compaction); the warp decides the run's fate (finished / still open /
        //   block path), merges it with the carried query, reserves its output and splits + parses its own top rows
        const int n_complete = S.geo.n_complete, last_nl = S.geo.last_nl;
        const bool may_continue = S.geo.may_continue != 0;
        const int n_runs = S.n_runs;
        const int r0 = S.n_skip;  // first row of this window the CTA had not seen
        const bool term = S.term != 0;
        const bool closes = covers_eof && p.final_chunk;  // the end of this window ends the open query
        PCLK(6)
#ifdef BLU_PHASE_CLOCKS
        rt = clock64();
#endif
        for (int j = warp; j < n_runs; j += kSWarps) {
            RCLK(0)
            const uint32_t entry = S.runs[j];
            const bool pseudo = (entry & 0x8000u) != 0;
            const int h = pseudo ? r0 : (int)(entry & 0x7FFFu);
            int e = next_head(S, pseudo ? h : h + 1, n_complete, lane);
            const bool open = e < 0;
            if (open) e = n_complete;
            if (e < h) e = h;
            RCLK(1)
            int mx = INT32_MIN;
            bool ovf = false;
            int g = 0;
            if (e - h <= 64) {
                // the usual case: the run's bit scores in two registers per lane, one pass over shared memory
                const int ra = h + lane, rb2 = ra + 32;
                int ba = INT32_MIN, bb = INT32_MIN;
                if (ra < e) {
                    ba = S.bits[ra];
                    ovf = (S.flags[ra] & 2) != 0;
                }
                if (rb2 < e) {
                    bb = S.bits[rb2];
                    ovf |= (S.flags[rb2] & 2) != 0;
                }
                mx = __reduce_max_sync(FULL, ba > bb ? ba : bb);
                ovf = __any_sync(FULL, ovf);
                const bool ta = ra < e && ba == mx, tb = rb2 < e && bb == mx;
                const unsigned bala = __ballot_sync(FULL, ta), balb = __ballot_sync(FULL, tb);
                const unsigned lt = (1u << lane) - 1u;
                const int na = __popc(bala);
                const int pa = __popc(bala & lt), pb = na + __popc(balb & lt);
                if (ta && pa < 32) S.stage[warp][pa] = (uint16_t)ra;
                if (tb && pb < 32) S.stage[warp][pb] = (uint16_t)rb2;
                g = na + __popc(balb);
            } else {
                for (int r = h + lane; r < e; r += 32) {
                    const int b = S.bits[r];
                    mx = b > mx ? b : mx;
                    ovf |= (S.flags[r] & 2) != 0;
                }
                mx = __reduce_max_sync(FULL, mx);
                ovf = __any_sync(FULL, ovf);
                for (int b = h; b < e; b += 32) {
                    const int r = b + lane;
                    const bool top = r < e && S.bits[r] == mx;
                    const unsigned bal = __ballot_sync(FULL, top);
                    const int pos = g + __popc(bal & ((1u << lane) - 1u));
                    if (top && pos < 32) S.stage[warp][pos] = (uint16_t)r;
                    g += __popc(bal);
                }
            }
            __syncwarp();
            if (g > 32) ovf = true;
            RCLK(2)
            // ---- lane 0: what happens to the run ------------------------------------------------------------------------
            uint32_t kind = 0, slot = 0;
            int g_part = 0, dst = 0, n_old = 0;
            if (lane == 0) {
                const int rows_t = e - h;
                g_part = ovf ? 0 : g;
                int g_tot = 0;
                kind = pseudo ? RK_PSEUDO : 0;
                unsigned long long abs;
                uint32_t qlen, nrows;
                if (pseudo) {
                    CarryRun& C = S.carry[cur];
                    if (ovf) C.deferred = 1;
                    if (rows_t > 0 && !C.deferred) {
                        if (mx > C.mx) {
                            C.mx = mx;
                            C.g = 0;
                        } else if (mx < C.mx)
                            g_part = 0;
                        if (C.g + g_part > kCarryTop) C.deferred = 1;
                    }
                    if (C.deferred || rows_t == 0) g_part = 0;
                    abs = C.head_abs, qlen = C.qlen, nrows = C.nrows + (uint32_t)rows_t;
                    mx = C.mx;
                    if (!open || closes) {
                        kind |= C.deferred ? RK_DEFER : RK_EMIT;
                        g_tot = C.g + g_part;
                        dst = C.g;
                        n_old = C.deferred ? 0 : C.g;
                        C.open = 0;
                    } else if (covers_eof) {
                        // the input continues in the next chunk: the whole query is carried over by the host
                        atomicMin(&p.ctr->tail_start, C.head_abs);
                        g_part = 0;
                        C.open = 0;
                    } else {
                        kind |= RK_OPEN;
                        C.nrows = nrows;
                        dst = -1 - C.g;
                        C.g += g_part;
                    }
                } else {
                    const int s = S.row_s[h];
                    qlen = entry >> 16;
                    if (qlen == 0xFFFFu) {
                        const int re = blank ? row_end_search(S, s, last_nl) : (int)S.row_s[h + 1] - 1;
                        qlen = (uint32_t)(next_tab(tabw, s, re) - s);
                    }
                    abs = lo + (unsigned long long)s;
                    nrows = (uint32_t)rows_t;
                    if (!open || closes) {
                        kind |= ovf ? RK_DEFER : RK_EMIT;
                        g_tot = g_part;
                    } else if (covers_eof) {
                        atomicMin(&p.ctr->tail_start, abs);
                        g_part = 0;
                    } else {
                        // a query that starts in this window and does not end in it: it becomes the carried query
                        CarryRun& C = S.carry[cur ^ 1];
                        kind |= RK_OPEN | RK_NEWCARRY;
                        C.open = 1;
                        C.deferred = ovf ? 1 : 0;
                        C.head_abs = abs;
                        C.qlen = qlen;
                        C.nrows = nrows;
                        C.mx = mx;
                        C.g = g_part;
                        dst = -1;
                        S.new_open = win_idx;
                    }
                }
                if (kind & RK_EMIT) {
                    // top-row slots from the CTA's slab (or, when it is exhausted, straight from the global counter)
                    if (g_tot > 0) {
                        slot = atomicAdd(&S.slot_cur, (uint32_t)g_tot);
                        if (slot + (uint32_t)g_tot > S.slot_end || slot + (uint32_t)g_tot < slot) {
                            const unsigned long long rs = atomicAdd(&p.ctr->rec_slots, (unsigned long long)g_tot);
                            slot = (uint32_t)rs;
                            if ((rs & 0xFFFFFFFFull) + (unsigned long long)g_tot > (unsigned long long)p.slot_cap) {
                                S.out_ok = 0;
                                p.ctr->cap_overflow = 1;
                            }
                        }
                    }
                    StagedRec sr;
                    sr.abs = abs, sr.qlen = qlen, sr.nrows = nrows, sr.mx = mx, sr.gtot = (uint32_t)g_tot, sr.slot = slot, sr.pad = 0;
                    const int ri = atomicAdd(&S.rec_cnt, 1);
                    if (ri < kRecBuf)
                        S.rec_buf[ri] = sr;
                    else {
                        // record buffer full (a window of very short queries): this record is reserved on its own
                        const unsigned long long rs = atomicAdd(&p.ctr->rec_slots, 1ull << 32);
                        if ((rs >> 32) < (unsigned long long)p.rec_cap)
                            write_record(p.records + (rs >> 32), sr);
                        else
                            p.ctr->cap_overflow = 1;
                    }
                }
                if (kind & RK_DEFER) push_defer(p, abs, 0);
            }
            RCLK(3)
            kind = __shfl_sync(FULL, kind, 0);
            slot = __shfl_sync(FULL, slot, 0);
            g_part = __shfl_sync(FULL, g_part, 0);
            dst = __shfl_sync(FULL, dst, 0);
            n_old = __shfl_sync(FULL, n_old, 0);
            const bool ok = S.out_ok != 0;
            // ---- the run's top rows of this window: field split + number parse, one lane each ----------------------------------
            if (lane < g_part && (kind & (RK_EMIT | RK_OPEN))) {
                const int r = S.stage[warp][lane];
                const int s = S.row_s[r];
                const int re = blank ? row_end_search(S, s, last_nl) : (int)S.row_s[r + 1] - 1;
                // a top row leaves the tile kernel as a reference into the text: the consensus kernel splits and parses it
                // fields 1..4 through the tab positions the row phase found: digit folds only, no second tab search; a row of any
                // other shape leaves unparsed (offset + length) and is split by the consensus kernel's full parser
                TopRowRaw ref;
                if (!top_row_from_info(win, s, S.rowinfo[r], lo, ref)) {
                    ref.acc_off = lo + (unsigned long long)s;
                    ref.acc_len = (uint32_t)(re - s);
                    ref.taxid = 0, ref.alnlen = 0, ref.pident = 0.0;
                    ref.dec_frac = kTopRowUnparsed;
                }
                if (kind & RK_EMIT) {
                    if (ok) p.toprows[slot + (uint32_t)(dst + lane)] = ref;
                } else
                    S.carry[(kind & RK_NEWCARRY) ? (cur ^ 1) : cur].tops[-1 - dst + lane] = ref;
            }
            RCLK(4)
            // a carried query that ends here: its earlier top rows go in front of this window's
            if (lane < n_old && ok) p.toprows[slot + lane] = S.carry[cur].tops[lane];
            __syncwarp();
            RCLK(5)
        }
        PCLK(7)
        __syncthreads();
        PCLK(8)
        const bool new_open = S.new_open == win_idx;  // (an epoch, not a flag: nothing to reset)
        const bool need_flush = S.rec_cnt >= kRecFlush;
        if (tid == 0) {
            // a new slab of slots when this one runs low (one global atomic every few dozen windows)
            if (S.slot_end - S.slot_cur < kSlotLow || S.slot_cur > S.slot_end) {
                // sized by what the rest of the segment is likely to need (slots used per window so far x windows left), so that
                // the unused tail of a CTA's last slab -- a hole in slot space that is downloaded with the results -- stays small
                const unsigned long long left = seg_hi > lo ? (seg_hi - lo) / (unsigned long long)kTile + 1ull : 1ull;
                const unsigned long long est = (unsigned long long)(S.slots_taken / (uint32_t)win_idx + 8u) * left + (unsigned long long)kSlotLow;
                const uint32_t slab = est < (unsigned long long)kSlotSlab ? (uint32_t)est : kSlotSlab;
                S.slots_taken += slab;
                const unsigned long long rs = atomicAdd(&p.ctr->rec_slots, (unsigned long long)slab);
                S.slot_cur = (uint32_t)rs;
                S.slot_end = (uint32_t)rs + slab;
                if ((rs & 0xFFFFFFFFull) + (unsigned long long)slab > (unsigned long long)p.slot_cap) {
                    S.out_ok = 0;
                    p.ctr->cap_overflow = 1;
                }
            }
        }
        if (need_flush) flush_records(p, S, tid);
        PCLK(9)
        // ---- where next? -----------------------------------------------------------------------------------------------
        // The CTA is done when the window reached the end of the text, when a query that belongs to the next segment
        // started in it, or when nothing is open and the segment is exhausted.
        if (new_open) cur ^= 1;
        const bool open_now = S.carry[cur].open != 0;
        bool done = covers_eof || term;
        if (!done && !open_now && lo + (unsigned long long)(last_nl + 1) >= seg_hi) done = true;
        if (!done && !may_continue) {
            // no complete new row in a whole window: a row longer than this kernel stages
            if (tid == 0) report(p.ctr, DE_CARRY_TOO_BIG, own_from);
            done = true;
        }
        if (done) {
            if (may_continue) {
                // a copy into the other buffer is in flight: it must land before the CTA exits
                uint32_t& ph = (buf ^ 1) ? ph1 : ph0;
                while (!mbar_try_wait(&S.mbar[buf ^ 1], ph)) {
                }
            }
            break;
        }
#ifdef BLU_PHASE_CLOCKS
        n_win++;
#endif
        own_from = lo + (unsigned long long)(last_nl + 1);
        buf ^= 1;
        win_idx++;
        // reset the per-window state (published by the barrier behind phase B of the next window)
        if (tid == 0) {
            S.bad_byte = INT_MAX;
            S.has_blank = 0;
            S.n_skip = 0;
            S.term = 0;
            S.n_runs = open_now ? 1 : 0;
            S.runs[0] = 0x8000;  // the carried query continues (or ends) at the first new row
        }
    }
#ifdef BLU_PHASE_CLOCKS
    if ((blockIdx.x == 7 || blockIdx.x == 200) && (lane == 0))
        printf("cta %d warp %d windows %d | tma %lld B %lld bar1 %lld geom %lld D %lld bar2 %lld flush %lld E %lld bar3 %lld F %lld bar4 %lld\n", blockIdx.x, warp, n_win, pc[0] / (n_win + 1),
               pc[1] / (n_win + 1), pc[2] / (n_win + 1), pc[3] / (n_win + 1), pc[4] / (n_win + 1), pc[5] / (n_win + 1), pc[6] / (n_win + 1), pc[7] / (n_win + 1), pc[8] / (n_win + 1),
               pc[9] / (n_win + 1), pc[10] / (n_win + 1));
    if ((blockIdx.x == 7) && (lane == 0))
        printf("rcl %d warp %d windows %d | queue %lld head %lld stats %lld decide %lld parse %lld old %lld\n", blockIdx.x, warp, n_win, rc[0] / (n_win + 1), rc[1] / (n_win + 1),
               rc[2] / (n_win + 1), rc[3] / (n_win + 1), rc[4] / (n_win + 1), rc[5] / (n_win + 1));
#endif
    // ---- the record headers still in the buffer -------------------------------------------------------------------------
    __syncthreads();
    flush_records(p, S, tid);
}

// ---------------------------------------------------------------------------------------------------------------
// long-run kernel: one CTA per deferred run
// ---------------------------------------------------------------------------------------------------------------
constexpr int kQidCache = 256;  // bytes of the run's query id kept in shared memory

struct LongSmem {
    WindowIndex W;
    long long bits[kRowCap];
    uint8_t same[kRowCap];
    unsigned long long cand_off[kLongTopCap];  // rows whose bit score equals the running maximum, in file order
    uint32_t cand_len[kLongTopCap];
    TopRow top[kLongTopCap];
    uint16_t tmp[5 * kLongTopCap];
    uint8_t qid[kQidCache];  // first field (+ its tab) of the run's first row
    long long red_max[kLongWarps];
    int red_cnt[kLongWarps];
    int red_first[kLongWarps];
    int bcast[8];
    unsigned long long bcast64[4];
};

static_assert(sizeof(LongSmem) <= 227 * 1024, "long-run CTA exceeds shared memory");

// first field of window row `a` == the run's query id (`qid` holds the id and its terminating tab, qn bytes)
__device__ __forceinline__ bool same_query_cached(const uint8_t* a, int alen, const uint8_t* qid, int qn) {
    if (alen < qn) return false;
    for (int i = 0; i < qn; i++)
        if (a[i] != qid[i]) return false;
    return true;
}

__device__ __forceinline__ bool same_query(const uint8_t* a, int alen, const uint8_t* text, unsigned long long s, unsigned long long end) {
    // compares the first field of row `a` (window) with the first field of the row at absolute offset s (global)
    for (int i = 0; i < alen; i++) {
        if (s + i >= end) return false;
        uint8_t b = text[s + i];
        if (a[i] != b) return false;
        if (b == '\t') return true;
    }
    return false;
}

struct LongScan {
    int ncomplete, eskip, d;     // d = first row that does not belong to the run (== ncomplete when none)
    unsigned long long next;     // where the next window starts
    bool at_end;                 // no more text after this window
    bool giant;                  // a row does not fit the window
};

// Stages the window that starts at `cur`, indexes and parses its rows, and marks which rows belong to the run
// whose first row starts at absolute offset `s`.
__device__ LongScan long_scan(LongSmem& S, const RunParams& p, unsigned long long s, unsigned long long cur, uint32_t& phase, int qn) {
    WindowIndex& W = S.W;
    LongScan r;
    const unsigned long long lo = cur & ~15ull;
    WinGeom g = make_geom(lo, kWin, cur, p.end);
    __syncthreads();
    if (threadIdx.x == 0) W.bad_byte = INT_MAX;
    load_window(W, p.text, lo, g.loaded, phase, true);
    if (threadIdx.x == 32) {
        // a long run continues right behind this window: pull the next one into L2 meanwhile
        const unsigned long long nb = lo + (unsigned long long)g.loaded;
        const unsigned long long up = (p.end + 15ull) & ~15ull;
        if (nb < up) tma_prefetch_l2(p.text + nb, (uint32_t)(up - nb < (unsigned long long)kWin ? up - nb : (unsigned long long)kWin));
    }
    finish_geom(W, g, p.final_chunk != 0);
    r.giant = false;
    if (!scan_rows<kLongWarps>(W, g)) {
        if (threadIdx.x == 0) report(p.ctr, DE_BAD_FIELD_COUNT, lo);
        r.ncomplete = 0, r.eskip = 0, r.d = 0, r.next = p.end, r.at_end = true;
        return r;
    }
    const int n_starts = W.n_starts, n_ends = W.n_ends;
    r.eskip = (n_ends > 0 && (n_starts == 0 || W.row_e[0] < W.row_s[0])) ? 1 : 0;
    r.ncomplete = n_starts < n_ends - r.eskip ? n_starts : n_ends - r.eskip;
    int first_other = r.ncomplete;
    for (int i = threadIdx.x; i < r.ncomplete; i += kLongThreads) {
        const int st = W.row_s[i];
        const int len = (int)W.row_e[i + r.eskip] - st;
        LightRow lr;
        {
            int64_t fb;
            int fq;
            if (parse_row_fast(W.win, reinterpret_cast<const uint32_t*>(W.tabm), reinterpret_cast<const uint32_t*>(W.digm), st, st + len, fb, fq)) {
                lr.bits = fb, lr.q_len = (uint16_t)fq, lr.err = DE_NONE;
            } else
                lr = parse_row_masked(W.win, reinterpret_cast<const uint64_t*>(W.tabm), reinterpret_cast<const uint64_t*>(W.digm), st, st + len);
        }
        const bool sm = qn > 0 ? same_query_cached(W.win + st, len, S.qid, qn) : same_query(W.win + st, len, p.text, s, p.end);
        if (lr.err && sm) report(p.ctr, lr.err, lo + st);
        S.bits[i] = lr.bits;
        S.same[i] = sm;
        if (!sm && i < first_other) first_other = i;
    }
#pragma unroll
    for (int dd = 16; dd > 0; dd >>= 1) {
        int o = __shfl_xor_sync(0xffffffffu, first_other, dd);
        first_other = o < first_other ? o : first_other;
    }
    if ((threadIdx.x & 31) == 0) S.red_first[threadIdx.x >> 5] = first_other;
    __syncthreads();
    int d = r.ncomplete;
    for (int i = 0; i < kLongWarps; i++) d = S.red_first[i] < d ? S.red_first[i] : d;
    r.d = d;
    if (W.bad_byte != INT_MAX && threadIdx.x == 0) {
        // only bytes of this run's rows are ours to report
        const int run_end = d < r.ncomplete ? (int)W.row_s[d] : (n_starts > r.ncomplete ? (int)W.row_s[r.ncomplete] : g.L);
        if (W.bad_byte < run_end) report(p.ctr, DE_QUOTE_OR_CR, lo + (unsigned)W.bad_byte);
    }
    if (n_starts > r.ncomplete) {
        r.next = lo + W.row_s[r.ncomplete];
        r.at_end = false;
        if (r.ncomplete == 0) r.giant = !g.covers_eof || !p.final_chunk;  // partial row only: too long, or chunk tail
        if (g.covers_eof) r.at_end = true;
    } else {
        r.next = r.ncomplete > 0 ? lo + W.row_e[r.ncomplete - 1 + r.eskip] + 1 : p.end;
        r.at_end = g.covers_eof || r.next >= p.end;
    }
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(kLongThreads, 1) longrun_kernel(const __grid_constant__ RunParams p) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    LongSmem& S = *reinterpret_cast<LongSmem*>(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) mbar_init(&S.W.mbar, 1);
    __syncthreads();
    uint32_t phase = 0;
    const unsigned n_defer = p.ctr->n_defer < p.defer_cap ? p.ctr->n_defer : p.defer_cap;

    while (true) {
        __syncthreads();
        if (tid == 0) S.bcast[0] = (int)atomicAdd(&p.ctr->work_ticket, 1u);
        __syncthreads();
        const unsigned wi = (unsigned)S.bcast[0];
        if (wi >= n_defer) break;
        const unsigned long long entry = p.defer[wi];
        const unsigned long long s = entry >> 1;
        // ---- is it really the first row of a run? ------------------------------------------------------------
        if (entry & 1) {
            if (tid == 0) {
                int head = 1;
                if (s > p.begin) {
                    long long q = (long long)s - 1;
                    while (q >= (long long)p.begin && p.text[q] == '\n') q--;  // skip the newline(s) before s
                    if (q >= (long long)p.begin) {
                        long long st = q;
                        while (st > (long long)p.begin && p.text[st - 1] != '\n') st--;
                        // compare first fields of rows at st and s
                        head = 0;
                        for (unsigned long long i = 0;; i++) {
                            if (s + i >= p.end || (unsigned long long)st + i > (unsigned long long)q) {
                                head = 1;
                                break;
                            }
                            uint8_t a = p.text[st + i], b = p.text[s + i];
                            if (a != b) {
                                head = 1;
                                break;
                            }
                            if (a == '\t') break;
                        }
                    }
                }
                S.bcast[1] = head;
            }
            __syncthreads();
            if (!S.bcast[1]) continue;
        }
        // ---- the run's query id (with its tab) is compared against every row: keep it in shared memory ---------------
        if (tid == 0) {
            int qn = 0;
            while (qn < kQidCache && s + qn < p.end) {
                const uint8_t b = p.text[s + qn];
                S.qid[qn++] = b;
                if (b == '\t') break;
            }
            S.bcast[4] = (qn > 0 && S.qid[qn - 1] == '\t') ? qn : 0;  // 0: id longer than the cache -> compare in global memory
        }
        __syncthreads();
        const int qn = S.bcast[4];
        // ---- single pass: extent, max bit score, and the rows that carry it (candidate list reset on a new maximum) --
        long long mx = LLONG_MIN;
        long long cnt = 0, nrows = 0;
        unsigned long long cur = s;
        bool ended = false, hit_end = false, fail = false;
        while (true) {
            LongScan sc = long_scan(S, p, s, cur, phase, qn);
            if (sc.giant && sc.d == 0 && sc.ncomplete == 0) {
                // an unterminated row: either the chunk tail (carried over) or a row longer than the window
                if (!sc.at_end || p.final_chunk) {
                    if (tid == 0) report(p.ctr, DE_CARRY_TOO_BIG, cur);
                    fail = true;
                }
                hit_end = true;
                break;
            }
            const unsigned long long lo = cur & ~15ull;
            // maximum of this window's rows of the run
            long long lm = LLONG_MIN;
            for (int i = tid; i < sc.d; i += kLongThreads) {
                const long long b = S.bits[i];
                lm = b > lm ? b : lm;
            }
#pragma unroll
            for (int dd = 16; dd > 0; dd >>= 1) {
                const long long om = __shfl_xor_sync(0xffffffffu, lm, dd);
                lm = om > lm ? om : lm;
            }
            if (lane == 0) S.red_max[warp] = lm;
            __syncthreads();
            long long wm = LLONG_MIN;
            for (int i = 0; i < kLongWarps; i++) wm = S.red_max[i] > wm ? S.red_max[i] : wm;
            if (sc.d > 0 && wm > mx) {
                mx = wm;
                cnt = 0;  // earlier candidates are beaten
            }
            // append this window's rows with bits == mx, in file order
            for (int b0 = 0; b0 < sc.d; b0 += kLongThreads) {
                const int i = b0 + tid;
                const bool top = i < sc.d && S.bits[i] == mx;
                const unsigned bal = __ballot_sync(0xffffffffu, top);
                if (lane == 0) S.red_cnt[warp] = __popc(bal);
                __syncthreads();
                long long before = cnt;
                int total = 0;
                for (int k = 0; k < kLongWarps; k++) {
                    if (k < warp) before += S.red_cnt[k];
                    total += S.red_cnt[k];
                }
                if (top) {
                    const long long pos = before + __popc(bal & ((1u << lane) - 1u));
                    if (pos < kLongTopCap) {
                        S.cand_off[pos] = lo + S.W.row_s[i];
                        S.cand_len[pos] = (uint32_t)((int)S.W.row_e[i + sc.eskip] - (int)S.W.row_s[i]);
                    }
                }
                cnt += total;
                __syncthreads();
            }
            nrows += sc.d;
            if (sc.d < sc.ncomplete) {
                ended = true;
                break;
            }
            if (sc.at_end) {
                hit_end = true;
                break;
            }
            cur = sc.next;
        }
        if (fail) continue;
        if (!ended && hit_end && !p.final_chunk) {
            if (tid == 0) atomicMin(&p.ctr->tail_start, s);  // carried into the next chunk
            continue;
        }
        if (nrows == 0) continue;
        if (cnt > kLongTopCap) {
            if (tid == 0) report(p.ctr, DE_TOPGROUP_TOO_BIG, s);
            continue;
        }
        const int gcount = (int)cnt;
        if (tid == 0) {
            unsigned long long rs = atomicAdd(&p.ctr->rec_slots, (1ull << 32) | (unsigned long long)gcount);
            S.bcast[2] = (int)(unsigned)(rs >> 32);
            S.bcast[3] = (int)(unsigned)rs;
        }
        // ---- join the top rows (read back from the text; they are few) -----------------------------------------------
        int any_err = 0;
        for (int i = tid; i < gcount; i += kLongThreads) {
            const unsigned long long off = S.cand_off[i];
            uint32_t err = heavy_parse_row(p.text + off, (int)S.cand_len[i], off, p.T, S.top[i]);
            if (err) {
                report(p.ctr, err, off);
                any_err = 1;
            }
        }
        any_err = __syncthreads_or(any_err);
        const unsigned rec_i = (unsigned)S.bcast[2], slot = (unsigned)S.bcast[3];
        if (rec_i >= p.rec_cap || slot + (unsigned)gcount > p.slot_cap) {
            if (tid == 0) p.ctr->cap_overflow = 1;
            continue;
        }
        if (any_err) {
            if (tid == 0) p.records[rec_i].status = 0, p.records[rec_i].n_rows = 0, p.records[rec_i].query_len = 0, p.records[rec_i].n_accessions = 0;
            continue;
        }
        if (tid == 0) {
            blu_record* rec = p.records + rec_i;
            QueryOut out{rec, p.beans + slot, p.accs + slot};
            uint32_t ql = 0;
            while (s + ql < p.end && p.text[s + ql] != '\t') ql++;
            rec->query_off = s;
            rec->query_len = ql;
            rec->n_rows = (uint32_t)nrows;
            rec->bit_score = mx;
            rec->slot_base = slot;
            rec->pad[0] = rec->pad[1] = 0;
            uint32_t ce = gcount == 1 ? consensus_single(S.top[0], p.T, out)
                                      : consensus_multi(S.top, gcount, p.text, p.T, p.strategy, S.tmp, out);
            if (ce) report(p.ctr, ce, s);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// consensus kernel: one warp per query, lanes = rows of the top bit-score group (<= 32)
//   find_multi_taxa_consensus.rs:39-214 + build_blast_consensus_identity.rs:9-105 + consensus_result.rs:65-88
//   (same semantics as the serial consensus_multi() in blu_core.cuh, which the block path and the host tests use)
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int cmp_u64(unsigned long long a, unsigned long long b) { return a < b ? -1 : (a > b ? 1 : 0); }

// bytewise comparison of two accessions given their first 16 bytes as big-endian keys (zero padded)
__device__ __forceinline__ int cmp_acc(unsigned long long a0, unsigned long long a1, unsigned alen, unsigned long long aoff, unsigned long long b0,
                                       unsigned long long b1, unsigned blen, unsigned long long boff, const uint8_t* text) {
    int c = cmp_u64(a0, b0);
    if (c) return c;
    c = cmp_u64(a1, b1);
    if (c) return c;
    if (alen > 16 && blen > 16) {
        const unsigned n = alen < blen ? alen : blen;
        for (unsigned i = 16; i < n; i++) {
            int d = (int)text[aoff + i] - (int)text[boff + i];
            if (d) return d;
        }
    }
    return (int)alen - (int)blen;
}

// rank selection for the reference lineage, lanes = lineage positions (two rounds cover the 64-position limit)
__device__ __forceinline__ void warp_apply_cutoffs(const LinTables& T, uint32_t ref_lin, uint32_t o, int k, double identity, bool whole, int idx,
                                                   blu_record* rec, int lane) {
    unsigned long long ge = 0, notgt = 0;
#pragma unroll
    for (int h = 0; h < 2; h++) {
        const int j = lane + 32 * h;
        double c = 0.0;
        const bool in = j < k;
        if (in) c = T.cut[o + j];
        const unsigned b_ge = __ballot_sync(0xffffffffu, in && identity >= c);     // linnaean_ranks.rs:208
        const unsigned b_ng = __ballot_sync(0xffffffffu, in && !(identity > c));   // linnaean_ranks.rs:188
        ge |= (unsigned long long)b_ge << (32 * h);
        notgt |= (unsigned long long)b_ng << (32 * h);
    }
    if (lane == 0) {
        const int allowed = notgt ? (__ffsll((long long)notgt) - 1) : -1;
        unsigned long long mask = ge;
        if (!whole) {
            // keep the first idx+1 survivors: enumerate AFTER the filter, take_while(index <= bean_index)
            unsigned long long m = 0, rest = ge;
            for (int n = 0; n <= idx && rest; n++) {
                unsigned long long low = rest & (~rest + 1);
                m |= low;
                rest ^= low;
            }
            mask = m;
        }
        const int last = mask ? 63 - __clzll((long long)mask) : -1;
        rec->keep_mask = mask;
        rec->allowed_pos = (int8_t)allowed;
        rec->reached_pos = (int8_t)(last >= 0 ? last : idx);
        rec->mutated = (allowed >= 0 && T.rank_cls[o + idx] != T.allowed_cls[o + allowed]) ? 1 : 0;
        rec->ref_lineage = ref_lin;
    }
}

__global__ void __launch_bounds__(256) consensus_kernel(const ConsParams p) {
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const unsigned qi = p.rec_begin + (blockIdx.x * blockDim.x + threadIdx.x) / 32;
    if (qi >= p.rec_end) return;
    blu_record* rec = p.records + qi;
    if (rec->status != 2) return;
    const LinTables& T = p.T;
    const int g = (int)rec->n_accessions;
    const unsigned slot = rec->slot_base;
    const bool on = lane < g;
    const unsigned act = g >= 32 ? FULL : ((1u << g) - 1u);
    TopRow r;
    r.pident = 0.0, r.alnlen = 0, r.acc_off = 0, r.lin = 0, r.acc_len = 0, r.lin_len = 0;
    uint32_t pos0 = 0;
    uint32_t jerr = 0;
    if (on) {
        // the row as the tile kernel left it -- a reference into the text: fields 1..4 are split and parsed here (the row was
        // validated by the tile kernel), then the join (left_join on subject_taxid == taxid, mod.rs:72-76) probes the taxid table
        const TopRowRaw raw = p.toprows[slot + lane];
        unsigned long long off = raw.acc_off;
        if (raw.dec_frac == kTopRowUnparsed) {
            // a row of unusual shape: split and parsed here by the full parser (acc_off / acc_len are the row's)
            jerr = off + (unsigned long long)raw.acc_len <= p.text_end ? heavy_parse_row(p.text + off, (int)raw.acc_len, off, T, r) : (uint32_t)DE_INTERNAL;
        } else
            jerr = join_top_row(raw, T, r);
        if (jerr)
            report(p.ctr, jerr, off);
        else
            pos0 = T.lin_off[r.lin];
    }
    if (__any_sync(FULL, jerr != 0)) {
        // (the run is over: the host reports the error; the record is emptied so that the gather / duplicate kernels that
        // are already queued behind this one do not follow accession slots that were never filled)
        if (lane == 0) rec->status = 0, rec->n_accessions = 0;
        return;
    }
    if (g == 1) {
        // single match (find_single_query_consensus.rs:74-150)
        const uint32_t o = __shfl_sync(FULL, pos0, 0);
        const int k = __shfl_sync(FULL, (int)r.lin_len, 0);
        const double pid = __shfl_sync(FULL, r.pident, 0);
        unsigned long long mask = 0;
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int j = lane + 32 * h;
            const bool in = j < k;
            double c = 0.0;
            if (in) c = T.cut[o + j];
            mask |= (unsigned long long)__ballot_sync(FULL, in && pid >= c) << (32 * h);
        }
        if (lane == 0) {
            if (!mask) {
                report(p.ctr, DE_EMPTY_ADJUSTED, rec->query_off);
                rec->status = 0;
                return;
            }
            const int last = 63 - __clzll((long long)mask);
            rec->keep_mask = mask;
            rec->perc_identity = r.pident;
            rec->ref_lineage = r.lin;
            rec->n_beans = 1;
            rec->n_accessions = 1;
            rec->single_match = 1;
            rec->mutated = 0;
            rec->reached_pos = (int8_t)last;
            rec->allowed_pos = -1;
            rec->bean_level = (int8_t)last;
            blu_bean b;
            b.first_lineage = r.lin, b.occurrences = 1, b.acc_begin = 0, b.n_acc = 1;
            p.beans[slot] = b;
            blu_acc a;
            a.off = r.acc_off, a.len = r.acc_len, a.pad = 0;
            p.accs[slot] = a;
            rec->status = 1;
        }
        return;
    }
    // ---- first 16 accession bytes as big-endian keys ---------------------------------------------------------------
    unsigned long long k0 = 0, k1 = 0;
    if (on) {
        // five aligned words cover the 16 bytes at any alignment (the row goes on for >= 20 bytes behind saccver, so the
        // reads stay inside the text); realigned by funnel shifts, byte-swapped to big-endian, bytes beyond the accession zeroed
        const unsigned n = r.acc_len;
        const uint32_t* wp = reinterpret_cast<const uint32_t*>(p.text + (r.acc_off & ~3ull));
        const uint32_t sh8 = (uint32_t)(r.acc_off & 3ull) * 8u;
        const uint32_t w0 = wp[0], w1 = wp[1], w2 = wp[2], w3 = wp[3], w4 = wp[4];
        const uint32_t q0 = __byte_perm(__funnelshift_r(w0, w1, sh8), 0u, 0x0123), q1 = __byte_perm(__funnelshift_r(w1, w2, sh8), 0u, 0x0123),
                       q2 = __byte_perm(__funnelshift_r(w2, w3, sh8), 0u, 0x0123), q3 = __byte_perm(__funnelshift_r(w3, w4, sh8), 0u, 0x0123);
        k0 = ((unsigned long long)q0 << 32) | (unsigned long long)q1;
        k1 = ((unsigned long long)q2 << 32) | (unsigned long long)q3;
        if (n < 16) k1 = n > 8 ? (k1 & (~0ull << (8u * (16u - n)))) : 0ull;
        if (n < 8) k0 = n > 0 ? (k0 & (~0ull << (8u * (8u - n)))) : 0ull;
    }
    // ---- S = stable sort by (lineage length, pident, align length, accession)   fmtc.rs:39-54 ----------------------
    int rank = 0;
    for (int j = 0; j < g; j++) {
        const int len_j = __shfl_sync(FULL, (int)r.lin_len, j);
        const double pid_j = __shfl_sync(FULL, r.pident, j);
        const long long aln_j = __shfl_sync(FULL, (long long)r.alnlen, j);
        const unsigned long long k0_j = __shfl_sync(FULL, k0, j), k1_j = __shfl_sync(FULL, k1, j);
        const unsigned alen_j = __shfl_sync(FULL, (unsigned)r.acc_len, j);
        const unsigned long long off_j = __shfl_sync(FULL, (unsigned long long)r.acc_off, j);
        bool lt;
        if (len_j != (int)r.lin_len)
            lt = len_j < (int)r.lin_len;
        else if (pid_j < r.pident)
            lt = true;
        else if (pid_j > r.pident)
            lt = false;
        else if (aln_j != (long long)r.alnlen)
            lt = aln_j < (long long)r.alnlen;
        else {
            const int c = on ? cmp_acc(k0_j, k1_j, alen_j, off_j, k0, k1, r.acc_len, r.acc_off, p.text) : 0;
            lt = c < 0 || (c == 0 && j < lane);
        }
        if (on && j != lane && lt) rank++;
    }
    // move row of rank s into lane s
    int src = lane;
    for (int j = 0; j < g; j++) {
        const int rj = __shfl_sync(FULL, rank, j);
        if (rj == lane) src = j;
    }
    r.pident = __shfl_sync(FULL, r.pident, src);
    r.alnlen = __shfl_sync(FULL, (long long)r.alnlen, src);
    r.acc_off = __shfl_sync(FULL, (unsigned long long)r.acc_off, src);
    r.lin = __shfl_sync(FULL, r.lin, src);
    r.acc_len = (uint16_t)__shfl_sync(FULL, (unsigned)r.acc_len, src);
    r.lin_len = (uint16_t)__shfl_sync(FULL, (unsigned)r.lin_len, src);
    pos0 = __shfl_sync(FULL, pos0, src);
    k0 = __shfl_sync(FULL, k0, src);
    k1 = __shfl_sync(FULL, k1, src);
    // ---- reference row + level walk (fmtc.rs:60-63,137-214) ----------------------------------------------------------
    const int ref_lane = p.strategy == BLU_STRATEGY_CAUTIOUS ? 0 : g - 1;
    const int m = __shfl_sync(FULL, (int)r.lin_len, 0);  // shortest lineage
    const uint32_t lin0 = __shfl_sync(FULL, r.lin, 0);
    int diverge = -1;
    if (__ballot_sync(FULL, on && r.lin != lin0)) {
        for (int i = 0; i < m; i++) {
            uint32_t key = 0;
            if (on) key = T.lvl_key[pos0 + i];
            const uint32_t key0 = __shfl_sync(FULL, key, 0);
            if (__ballot_sync(FULL, on && key != key0)) {
                diverge = i;
                break;
            }
        }
    }
    if (diverge == 0) {
        if (lane == 0) {
            report(p.ctr, DE_ROOT_DISAGREE, rec->query_off);
            rec->status = 0;
        }
        return;
    }
    double identity;
    int idx, level;
    bool single;
    const double ref_pid = __shfl_sync(FULL, r.pident, ref_lane);
    if (diverge > 0) {
        double mx = on ? r.pident : 0.0;  // fold from 0.0 (fmtc.rs:182-185)
        if (!(mx > 0.0)) mx = 0.0;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            const double o = __shfl_xor_sync(FULL, mx, d);
            mx = o > mx ? o : mx;
        }
        identity = mx, idx = diverge - 1, level = diverge, single = false;
    } else {
        identity = ref_pid, idx = m - 1, level = m - 1, single = true;
    }
    // ---- fold beans at `level` in S order (consensus_result.rs:65-88) -----------------------------------------------
    uint32_t bkey = 0xFFFFFF00u + (uint32_t)lane, irank = 0;
    if (on) {
        bkey = T.bean_key[pos0 + level];
        irank = T.ident_rank[pos0 + level];
    }
    const unsigned grp = __match_any_sync(FULL, bkey) & act;
    const unsigned below = grp & ((1u << lane) - 1u);
    const bool leader = on && below == 0;
    const int occ = __popc(grp);
    // Vec::dedup: an accession equal to its predecessor in the bean's S-ordered list is dropped
    const int pl = below ? 31 - __clz(below) : lane;
    const unsigned long long pk0 = __shfl_sync(FULL, k0, pl), pk1 = __shfl_sync(FULL, k1, pl);
    const unsigned plen = __shfl_sync(FULL, (unsigned)r.acc_len, pl);
    const unsigned long long poff = __shfl_sync(FULL, (unsigned long long)r.acc_off, pl);
    bool dropped = false;
    if (on && below) dropped = plen == r.acc_len && cmp_acc(pk0, pk1, plen, poff, k0, k1, r.acc_len, r.acc_off, p.text) == 0;
    const unsigned kept = __ballot_sync(FULL, on && !dropped);
    const int nacc_bean = __popc(grp & kept);
    const int my_idx = __popc(grp & kept & ((1u << lane) - 1u));
    // sort beans: occurrences desc, identifier asc (bbci.rs:50-60); deterministic tie-break on the key id
    const unsigned lm = __ballot_sync(FULL, leader);
    int brank = 0, acc_begin = 0;
    for (unsigned rest = lm; rest;) {
        const int j = __ffs(rest) - 1;
        rest &= rest - 1;
        const int occ_j = __shfl_sync(FULL, occ, j);
        const uint32_t ir_j = __shfl_sync(FULL, irank, j), bk_j = __shfl_sync(FULL, bkey, j);
        const int na_j = __shfl_sync(FULL, nacc_bean, j);
        if (leader && j != lane) {
            const bool better = occ_j != occ ? occ_j > occ : (ir_j != irank ? ir_j < irank : bk_j < bkey);
            if (better) {
                brank++;
                acc_begin += na_j;
            }
        }
    }
    if (leader) {
        blu_bean b;
        b.first_lineage = r.lin, b.occurrences = (uint32_t)occ, b.acc_begin = (uint32_t)acc_begin, b.n_acc = (uint32_t)nacc_bean;
        p.beans[slot + brank] = b;
    }
    const int my_begin = __shfl_sync(FULL, acc_begin, grp ? __ffs(grp) - 1 : 0);
    if (on && !dropped) {
        blu_acc a;
        a.off = r.acc_off, a.len = r.acc_len, a.pad = 0;
        p.accs[slot + my_begin + my_idx] = a;
    }
    const int nb = __popc(lm);
    // ---- rank selection on the reference lineage (bbci.rs:22-37,66-95) -----------------------------------------------
    const uint32_t ref_lin = __shfl_sync(FULL, r.lin, ref_lane);
    const uint32_t ref_o = __shfl_sync(FULL, pos0, ref_lane);
    const int ref_k = __shfl_sync(FULL, (int)r.lin_len, ref_lane);
    warp_apply_cutoffs(T, ref_lin, ref_o, ref_k, identity, single && nb == 1, idx, rec, lane);
    if (lane == 0) {
        rec->perc_identity = ref_pid;
        rec->n_beans = (uint32_t)nb;
        rec->n_accessions = (uint32_t)__popc(kept);
        rec->single_match = 0;
        rec->bean_level = (int8_t)level;
        rec->status = 1;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// gather: query ids + accessions -> string pool
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gather_kernel(const GatherParams p) {
    __shared__ unsigned long long warp_tot[8];
    __shared__ unsigned long long base_sh;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned i = p.rec_begin + blockIdx.x * blockDim.x + tid;
    const bool live = i < p.rec_end;
    unsigned long long bytes = 0;
    blu_record* rec = nullptr;
    unsigned rows = 0;
    if (live) {
        rec = p.records + i;
        rows = rec->n_rows;
        bytes = rec->query_len;
        for (unsigned a = 0; a < rec->n_accessions; a++) bytes += p.accs[rec->slot_base + a].len;
    }
    unsigned long long inc = bytes;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        unsigned long long t = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += t;
    }
    if (lane == 31) warp_tot[warp] = inc;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) rows += __shfl_xor_sync(0xffffffffu, rows, d);
    if (lane == 0 && rows) atomicAdd(&p.ctr->n_rows, (unsigned long long)rows);
    __syncthreads();
    unsigned long long before = 0, total = 0;
    for (int k = 0; k < 8; k++) {
        if (k < warp) before += warp_tot[k];
        total += warp_tot[k];
    }
    if (tid == 0) base_sh = total ? atomicAdd(&p.ctr->pool_used, total) : 0ull;
    __syncthreads();
    if (!live) return;
    unsigned long long o = base_sh + before + inc - bytes;
    if (base_sh + total > p.pool_cap) {
        p.ctr->cap_overflow = 1;
        return;
    }
    const uint8_t* src = p.text + rec->query_off;
    for (unsigned k = 0; k < rec->query_len; k++) p.pool[o + k] = src[k];
    rec->query_off = o;
    o += rec->query_len;
    for (unsigned a = 0; a < rec->n_accessions; a++) {
        blu_acc& ac = p.accs[rec->slot_base + a];
        const uint8_t* sa = p.text + ac.off;
        for (unsigned k = 0; k < ac.len; k++) p.pool[o + k] = sa[k];
        ac.off = o;
        o += ac.len;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// duplicate query detection (reference groups by a HashMap<String,_>, mod.rs:145,192: rows of one query need
// not be contiguous).  A repeated id means the fast contiguous path is not applicable.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) dup_kernel(const DupParams p) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.n_rec) return;
    const blu_record& r = p.records[i];
    // (this kernel is queued behind the gather pass before the host has looked at its overflow flag: when the string pool
    // was too small the ids were not moved and query_off still is a text offset)
    if (r.query_off + (unsigned long long)r.query_len > p.pool_cap) return;
    unsigned long long h = 0xcbf29ce484222325ull;
    const uint8_t* q = p.pool + r.query_off;
    for (unsigned k = 0; k < r.query_len; k++) {
        h ^= q[k];
        h *= 0x100000001b3ull;
    }
    h = mix64(h ^ r.query_len);
    if (h == 0) h = 1;
    unsigned slot = (unsigned)h & p.mask;
    for (unsigned n = 0; n <= p.mask; n++) {
        unsigned long long old = atomicCAS(p.table + slot, 0ull, h);
        if (old == 0ull) return;
        if (old == h) {
            p.ctr->dup_found = 1;
            return;
        }
        slot = (slot + 1) & p.mask;
    }
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------------------------
cudaError_t kernels_set_attributes() {
    cudaError_t e = cudaFuncSetAttribute(tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(StreamSmem));
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(longrun_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(LongSmem));
}

int tile_kernel_grid(int device) {
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    return kTileCtasPerSm * sms;  // one streaming CTA per resident slot
}

cudaError_t launch_tile_kernel(const RunParams& p, int grid, cudaStream_t s) {
    tile_kernel<<<grid, kTileThreads, sizeof(StreamSmem), s>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_longrun_kernel(const RunParams& p, int grid, cudaStream_t s) {
    longrun_kernel<<<grid, kLongThreads, sizeof(LongSmem), s>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_consensus_kernel(const ConsParams& p, cudaStream_t s) {
    if (p.rec_end <= p.rec_begin) return cudaSuccess;
    const unsigned n = p.rec_end - p.rec_begin;  // one warp per record; two per CTA: the warps of a CTA finish at very different
                                                 // times (top groups of 1..32 rows), and a slot is only refilled when its whole CTA is gone
    consensus_kernel<<<(n + 1) / 2, 64, 0, s>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_gather_kernel(const GatherParams& p, cudaStream_t s) {
    if (p.rec_end <= p.rec_begin) return cudaSuccess;
    unsigned n = p.rec_end - p.rec_begin;
    gather_kernel<<<(n + 255) / 256, 256, 0, s>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_dup_kernel(const DupParams& p, cudaStream_t s) {
    if (p.n_rec == 0) return cudaSuccess;
    dup_kernel<<<(p.n_rec + 255) / 256, 256, 0, s>>>(p);
    return cudaGetLastError();
}

}  // namespace blu
