// blu_api.cpp -- C ABI (include/blu_consensus.h): context, taxonomy upload, chunked streaming of the outfmt-6
// text through the CUDA kernels, result download/decoding and the reference-compatible writer.
//
// There is deliberately NO CPU implementation of the consensus here: everything that computes goes through
// blu_kernels.cu; without a CUDA device every compute entry point fails with BLU_ERR_CUDA.
#include <cuda_runtime.h>

#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cerrno>
#include <cstdio>
#include <cstring>
#include <filesystem>
#include <fstream>
#include <memory>
#include <mutex>
#include <random>
#include <string>
#include <string_view>
#include <thread>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "../../include/blu_consensus.h"
#include "blu_decode.h"
#include "blu_json.h"
#include "blu_kernels.h"
#include "blu_tabular.h"
#include "blu_taxonomy.h"

using namespace blu;

namespace {

thread_local std::string g_create_error;

struct CudaErr : std::runtime_error {
    using std::runtime_error::runtime_error;
};
struct UnsupportedErr : std::runtime_error {
    using std::runtime_error::runtime_error;
};

#define CK(expr)                                                                                        \
    do {                                                                                                \
        cudaError_t _e = (expr);                                                                        \
        if (_e != cudaSuccess) throw CudaErr(std::string(#expr) + ": " + cudaGetErrorString(_e));       \
    } while (0)

template <class T>
struct DevBuf {
    T* p = nullptr;
    size_t cap = 0;  // elements
    void ensure(size_t n) {
        if (n <= cap) return;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        CK(cudaMalloc((void**)&p, n * sizeof(T)));
        cap = n;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

struct PinnedBuf {
    void* p = nullptr;
    size_t cap = 0;
};

// Pinned result buffers are recycled between runs.  The pool is shared by the context and every result it
// produced, so a result may be freed after its context.
struct PinnedPool {
    std::mutex mu;
    std::vector<PinnedBuf> free_list;
    bool closed = false;
    PinnedBuf acquire(size_t bytes) {
        {
            std::lock_guard<std::mutex> g(mu);
            size_t best = SIZE_MAX;
            for (size_t i = 0; i < free_list.size(); i++)
                if (free_list[i].cap >= bytes && (best == SIZE_MAX || free_list[i].cap < free_list[best].cap)) best = i;
            if (best != SIZE_MAX) {
                PinnedBuf b = free_list[best];
                free_list.erase(free_list.begin() + best);
                return b;
            }
        }
        PinnedBuf b;
        size_t cap = std::max<size_t>(bytes + bytes / 8, 4096);
        CK(cudaHostAlloc(&b.p, cap, cudaHostAllocDefault));
        b.cap = cap;
        return b;
    }
    void release(PinnedBuf b) {
        if (!b.p) return;
        {
            std::lock_guard<std::mutex> g(mu);
            if (!closed && free_list.size() < 16) {
                free_list.push_back(b);
                return;
            }
        }
        cudaFreeHost(b.p);
    }
    void close() {
        std::vector<PinnedBuf> v;
        {
            std::lock_guard<std::mutex> g(mu);
            closed = true;
            v.swap(free_list);
        }
        for (auto& b : v) cudaFreeHost(b.p);
    }
};

}  // namespace

struct blu_ctx {
    blu_opts opts{};
    Cutoffs cut;
    int device = 0;
    int sms = 148;
    cudaStream_t stream = nullptr, copy_stream = nullptr, d2h_stream = nullptr;
    cudaEvent_t ev[6]{};
    cudaEvent_t ev_h2d[2]{}, ev_free[2]{};
    std::shared_ptr<HostTaxonomy> tax;
    // device taxonomy
    DevBuf<uint32_t> d_lin_off, d_lvl, d_bean, d_irank;
    DevBuf<double> d_cut;
    DevBuf<uint16_t> d_rcls, d_acls;
    DevBuf<uint8_t> d_linok;
    DevBuf<HashSlot> d_slots;
    LinTables dT{};
    // run buffers
    DevBuf<uint8_t> d_text[2];
    DevBuf<blu_record> d_rec;
    DevBuf<blu_bean> d_beans;
    DevBuf<blu_acc> d_accs;
    DevBuf<TopRowRaw> d_top;
    DevBuf<uint64_t> d_defer;
    DevBuf<uint8_t> d_pool;
    DevBuf<unsigned long long> d_dup;
    Counters* d_ctr = nullptr;
    Counters* h_ctr = nullptr;  // pinned, two entries
    std::shared_ptr<PinnedPool> pool = std::make_shared<PinnedPool>();
    std::string err;
    blu_timings tm{};
    uint64_t carry_bytes = 64ull << 20;

    PinnedBuf acquire(size_t bytes) { return pool->acquire(bytes); }
};

struct blu_result {
    std::shared_ptr<PinnedPool> pinned;
    std::shared_ptr<HostTaxonomy> tax;
    Cutoffs cut;
    PinnedBuf b_rec, b_beans, b_accs, b_pool;
    uint64_t n_rec = 0, n_slots = 0, pool_len = 0, n_rows = 0;
    std::vector<std::string> hitless;  // NoConsensusFound (mod.rs:84-102)
    const blu_record* rec() const { return (const blu_record*)b_rec.p; }
    const blu_bean* beans() const { return (const blu_bean*)b_beans.p; }
    const blu_acc* accs() const { return (const blu_acc*)b_accs.p; }
    const char* pool() const { return (const char*)b_pool.p; }
};

namespace {

ResultView make_view(const blu_result* r) {
    ResultView v;
    v.tax = r->tax.get();
    v.cut = r->cut;
    v.rec_ = r->rec();
    v.beans_ = r->beans();
    v.accs_ = r->accs();
    v.pool_ = r->pool();
    v.n_rec = r->n_rec;
    v.hitless_ = &r->hitless;
    return v;
}

int fail(blu_ctx* ctx, int code, const std::string& msg) {
    if (ctx)
        ctx->err = msg;
    else
        g_create_error = msg;
    return code;
}

template <class F>
int guarded(blu_ctx* ctx, F&& f) {
    try {
        f();
        if (ctx) ctx->err.clear();
        return BLU_OK;
    } catch (const IoErr& e) {
        return fail(ctx, BLU_ERR_IO, e.what());
    } catch (const DataErr& e) {
        return fail(ctx, BLU_ERR_DATA, e.what());
    } catch (const CudaErr& e) {
        return fail(ctx, BLU_ERR_CUDA, e.what());
    } catch (const UnsupportedErr& e) {
        return fail(ctx, BLU_ERR_UNSUPPORTED, e.what());
    } catch (const std::invalid_argument& e) {
        return fail(ctx, BLU_ERR_UNSUPPORTED, e.what());
    } catch (const std::bad_alloc&) {
        return fail(ctx, BLU_ERR_INTERNAL, "out of host memory");
    } catch (const std::exception& e) {
        return fail(ctx, BLU_ERR_INTERNAL, e.what());
    }
}

const char* dev_err_text(uint32_t e) {
    switch (e) {
        case DE_BAD_FIELD_COUNT: return "row does not have 13 tab-separated fields";
        case DE_BAD_NUMBER: return "malformed numeric field";
        case DE_EMPTY_STRING: return "empty qseqid / saccver";
        case DE_QUOTE_OR_CR: return "'\"' or '\\r' byte in the blast output (not supported)";
        case DE_UNMAPPED_TAXID: return "subject taxid of a top bit-score hit has no lineage in the taxonomy file (reference: panic on `null` lineage)";
        case DE_BAD_LINEAGE: return "Unexpected error on parse taxonomy (lineage of a top bit-score hit)";
        case DE_EMPTY_ADJUSTED: return "No taxonomy found for result (single match below every identity cutoff)";
        case DE_ROOT_DISAGREE: return "top hits disagree at the first lineage level (reference: index underflow panic)";
        case DE_BITS_RANGE: return "bit score outside the i64 range";
        case DE_NUM_UNSUPPORTED: return "number outside the exactly-parsed range (more than 19 significant digits or |exponent| > 22)";
        case DE_TOPGROUP_TOO_BIG: return "top bit-score group larger than 1024 rows";
        case DE_CARRY_TOO_BIG: return "a single row does not fit the 60 KB window";
        default: return "internal device error";
    }
}

void upload_taxonomy(blu_ctx* c) {
    HostTaxonomy& T = *c->tax;
    auto up = [&](auto& dbuf, const auto& vec) {
        dbuf.ensure(std::max<size_t>(vec.size(), 1));
        if (!vec.empty()) CK(cudaMemcpy(dbuf.p, vec.data(), vec.size() * sizeof(vec[0]), cudaMemcpyHostToDevice));
    };
    up(c->d_lin_off, T.lin_off);
    up(c->d_lvl, T.lvl_key);
    up(c->d_bean, T.bean_key);
    up(c->d_irank, T.ident_rank);
    up(c->d_cut, T.cut);
    up(c->d_rcls, T.rank_cls);
    up(c->d_acls, T.allowed_cls);
    up(c->d_linok, T.lin_ok);
    up(c->d_slots, T.slots);
    c->dT.lin_off = c->d_lin_off.p;
    c->dT.lvl_key = c->d_lvl.p;
    c->dT.bean_key = c->d_bean.p;
    c->dT.ident_rank = c->d_irank.p;
    c->dT.cut = c->d_cut.p;
    c->dT.rank_cls = c->d_rcls.p;
    c->dT.allowed_cls = c->d_acls.p;
    c->dT.lin_ok = c->d_linok.p;
    c->dT.slots = c->d_slots.p;
    c->dT.hash_mask = T.hash_mask;
    c->dT.n_lin = (uint32_t)T.n_lin();
}

struct Caps {
    size_t rec, slots, defer, pool;
};

inline uint64_t n_rec_of(const Counters& h) { return h.rec_slots >> 32; }
inline uint64_t n_slots_of(const Counters& h) { return h.rec_slots & 0xFFFFFFFFull; }

Caps initial_caps(uint64_t n_bytes) {
    Caps c;
    c.rec = n_bytes / 160 + 4096;
    c.slots = n_bytes / 96 + 8192;
    c.defer = c.rec;
    c.pool = n_bytes / 24 + 65536;
    return c;
}

void ensure_out(blu_ctx* c, const Caps& k) {
    c->d_rec.ensure(k.rec);
    c->d_beans.ensure(k.slots);
    c->d_accs.ensure(k.slots);
    c->d_top.ensure(k.slots);
    c->d_defer.ensure(k.defer);
    c->d_pool.ensure(k.pool);
}

void check_device_error(blu_ctx*, const Counters& h, uint64_t stream_base) {
    if (!h.err_code) return;
    std::string m = std::string(dev_err_text(h.err_code)) + " (near byte " + std::to_string(stream_base + h.err_off) + " of the blast output)";
    if (h.err_code >= DE_INTERNAL) throw std::runtime_error(m);
    if (h.err_code >= DE_NUM_UNSUPPORTED) throw UnsupportedErr(m);
    throw DataErr(m);
}

// BLU_SYNC_DEBUG=1: synchronise behind every kernel so that a device fault is attributed to the kernel that caused it
inline void dbg_sync(cudaStream_t s, const char* what) {
    static const bool on = getenv("BLU_SYNC_DEBUG") != nullptr;
    if (!on) return;
    cudaError_t e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) throw std::runtime_error(std::string(what) + ": " + cudaGetErrorString(e));
}

// Runs the kernels on one resident chunk.  rec_begin = number of records before this chunk.
void launch_chunk(blu_ctx* c, const uint8_t* dtext, uint64_t begin, uint64_t end, bool final_chunk, const Caps& k, cudaStream_t s,
                  uint32_t rec_begin_hint, bool time_it) {
    RunParams p{};
    p.text = dtext;
    p.begin = begin;
    p.end = end;
    p.final_chunk = final_chunk ? 1 : 0;
    p.strategy = c->opts.strategy;
    p.T = c->dT;
    p.records = c->d_rec.p;
    p.rec_cap = (uint32_t)std::min<size_t>(k.rec, 0xFFFFFFFFu);
    p.beans = c->d_beans.p;
    p.accs = c->d_accs.p;
    p.toprows = c->d_top.p;
    p.slot_cap = (uint32_t)std::min<size_t>(k.slots, 0xFFFFFFFFu);
    p.defer = c->d_defer.p;
    p.defer_cap = (uint32_t)std::min<size_t>(k.defer, 0xFFFFFFFFu);
    p.ctr = c->d_ctr;
    (void)rec_begin_hint;
    if (time_it) CK(cudaEventRecord(c->ev[0], s));
    CK(launch_tile_kernel(p, tile_kernel_grid(c->device), s));
    dbg_sync(s, "tile_kernel");
    if (time_it) CK(cudaEventRecord(c->ev[1], s));
    CK(launch_longrun_kernel(p, c->sms, s));
    dbg_sync(s, "longrun_kernel");
    if (time_it) CK(cudaEventRecord(c->ev[2], s));
    c->tm.n_kernel_launches += 2;
}

void reset_counters_async(blu_ctx* c, cudaStream_t s, bool whole) {
    if (whole) {
        CK(cudaMemsetAsync(c->d_ctr, 0, sizeof(Counters), s));
    } else {
        // per chunk: n_defer, work_ticket back to zero; keep the output counters
        CK(cudaMemsetAsync(&c->d_ctr->n_defer, 0, sizeof(unsigned), s));
        CK(cudaMemsetAsync(&c->d_ctr->work_ticket, 0, sizeof(unsigned), s));
    }
    CK(cudaMemsetAsync(&c->d_ctr->tail_start, 0xFF, sizeof(unsigned long long), s));
}

// Incremental download of the finished part of the result arrays on the context's download stream, so that the
// device->host copies of one range / chunk run under the kernels (and host->device copies) of the next one.
struct Downloader {
    blu_ctx* c;
    blu_result* r;
    uint64_t rec_done = 0, slot_done = 0, pool_done = 0;
    uint64_t bytes = 0;

    Downloader(blu_ctx* ctx, blu_result* res) : c(ctx), r(res) {}

    static void grow(blu_ctx* c, PinnedBuf& b, size_t need, size_t keep, cudaStream_t ds) {
        if (b.cap >= need && b.p) return;
        PinnedBuf nb = c->acquire(std::max<size_t>(need, 64));
        if (keep) {
            CK(cudaStreamSynchronize(ds));  // the prefix may still be in flight
            memcpy(nb.p, b.p, keep);
        }
        if (b.p) c->pool->release(b);
        b = nb;
    }
    // capacity for at least these totals (called with estimates first, exact numbers at the end)
    void reserve(uint64_t rec, uint64_t slots, uint64_t pool) {
        cudaStream_t ds = c->d2h_stream;
        grow(c, r->b_rec, rec * sizeof(blu_record), rec_done * sizeof(blu_record), ds);
        grow(c, r->b_beans, slots * sizeof(blu_bean), slot_done * sizeof(blu_bean), ds);
        grow(c, r->b_accs, slots * sizeof(blu_acc), slot_done * sizeof(blu_acc), ds);
        grow(c, r->b_pool, pool, pool_done, ds);
    }
    // everything up to the counters in `h` is final on the device (the compute stream has been synchronised)
    void push(const Counters& h) {
        cudaStream_t ds = c->d2h_stream;
        const uint64_t rec = n_rec_of(h), slots = n_slots_of(h), pool = h.pool_used;
        reserve(rec, slots, pool);
        if (rec > rec_done)
            CK(cudaMemcpyAsync((blu_record*)r->b_rec.p + rec_done, c->d_rec.p + rec_done, (rec - rec_done) * sizeof(blu_record),
                               cudaMemcpyDeviceToHost, ds));
        if (slots > slot_done) {
            CK(cudaMemcpyAsync((blu_bean*)r->b_beans.p + slot_done, c->d_beans.p + slot_done, (slots - slot_done) * sizeof(blu_bean),
                               cudaMemcpyDeviceToHost, ds));
            CK(cudaMemcpyAsync((blu_acc*)r->b_accs.p + slot_done, c->d_accs.p + slot_done, (slots - slot_done) * sizeof(blu_acc),
                               cudaMemcpyDeviceToHost, ds));
        }
        if (pool > pool_done)
            CK(cudaMemcpyAsync((char*)r->b_pool.p + pool_done, c->d_pool.p + pool_done, pool - pool_done, cudaMemcpyDeviceToHost, ds));
        bytes += (rec - rec_done) * sizeof(blu_record) + (slots - slot_done) * (sizeof(blu_bean) + sizeof(blu_acc)) + (pool - pool_done);
        rec_done = rec, slot_done = slots, pool_done = pool;
    }
    void finish(const Counters& h) {
        push(h);
        CK(cudaStreamSynchronize(c->d2h_stream));
        r->n_rec = rec_done;
        r->n_slots = slot_done;
        r->pool_len = pool_done;
        r->n_rows = h.n_rows;
        c->tm.d2h_bytes += bytes + sizeof(Counters);
        c->tm.result_bytes = r->n_rec * sizeof(blu_record) + r->n_slots * (sizeof(blu_bean) + sizeof(blu_acc));
        c->tm.n_queries = r->n_rec;
        c->tm.n_rows = r->n_rows;
    }
    void abandon() {  // retry with larger device capacities: drop what was downloaded
        cudaStreamSynchronize(c->d2h_stream);
        rec_done = slot_done = pool_done = 0;
        bytes = 0;
    }
};

float ev_ms(cudaEvent_t a, cudaEvent_t b) {
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    return ms;
}

// Post-pass on all records of the run: string gather for [rec_begin, rec_end) + (at the end) duplicate check.
void launch_gather(blu_ctx* c, const uint8_t* dtext, uint64_t text_end, uint32_t rec_begin, uint32_t rec_end, const Caps& k, cudaStream_t s) {
    {
        // warp-per-query consensus over the top-row table the tile kernel produced for these records
        ConsParams q{};
        q.records = c->d_rec.p;
        q.rec_begin = rec_begin;
        q.rec_end = rec_end;
        q.toprows = c->d_top.p;
        q.beans = c->d_beans.p;
        q.accs = c->d_accs.p;
        q.text = dtext;
        q.text_end = text_end;
        q.T = c->dT;
        q.strategy = c->opts.strategy;
        q.ctr = c->d_ctr;
        CK(launch_consensus_kernel(q, s));
        dbg_sync(s, "consensus_kernel");
        if (rec_end > rec_begin) c->tm.n_kernel_launches += 1;
    }
    GatherParams g{};
    g.text = dtext;
    g.records = c->d_rec.p;
    g.accs = c->d_accs.p;
    g.rec_begin = rec_begin;
    g.rec_end = rec_end;
    g.pool = c->d_pool.p;
    g.pool_cap = k.pool;
    g.ctr = c->d_ctr;
    CK(launch_gather_kernel(g, s));
    dbg_sync(s, "gather_kernel");
    if (rec_end > rec_begin) c->tm.n_kernel_launches += 1;
}

void launch_dup(blu_ctx* c, uint32_t n_rec, cudaStream_t s) {
    if (!n_rec) return;
    size_t cap = 1024;
    while (cap < 2ull * n_rec) cap <<= 1;
    c->d_dup.ensure(cap);
    CK(cudaMemsetAsync(c->d_dup.p, 0, cap * sizeof(unsigned long long), s));
    DupParams d{};
    d.records = c->d_rec.p;
    d.n_rec = n_rec;
    d.pool = c->d_pool.p;
    d.pool_cap = c->d_pool.cap;
    d.table = c->d_dup.p;
    d.mask = (uint32_t)(cap - 1);
    d.ctr = c->d_ctr;
    CK(launch_dup_kernel(d, s));
    dbg_sync(s, "dup_kernel");
    c->tm.n_kernel_launches += 1;
}

// Tried and dropped (profiles/README.md): a one-thread kernel writing the counters to mapped host memory with the host
// spinning on a sequence word, to avoid the copy engine and the stream-synchronise wake-up: no measurable difference.
void read_counters(blu_ctx* c, cudaStream_t s) {
    CK(cudaMemcpyAsync(c->h_ctr, c->d_ctr, sizeof(Counters), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
}

// `scale` = whole input / part of it the counters cover (>= 1): records, slots and pool bytes are cumulative over the
// chunks / ranges processed so far, so the new capacity is extrapolated to the whole input -- otherwise a table of tiny
// rows in many chunks needs one retry per chunk.  `n_bytes` bounds the extrapolation (a row has >= 26 bytes).
bool grow_caps(const Counters& h, Caps& k, uint64_t defer_seen, double scale, uint64_t n_bytes) {
    bool grew = false;
    scale = std::max(1.0, scale);
    auto want = [&](uint64_t seen, uint64_t slack, uint64_t bound) {
        const double w = (double)seen * scale * 1.125 + (double)slack;
        return (size_t)std::max<double>((double)seen + (double)slack, std::min<double>(w, (double)bound + (double)slack));
    };
    if (n_rec_of(h) > k.rec) k.rec = want(n_rec_of(h), 1024, n_bytes / 26), grew = true;
    if (n_slots_of(h) > k.slots) k.slots = want(n_slots_of(h), 1024, n_bytes / 26), grew = true;
    if (defer_seen > k.defer) k.defer = (size_t)defer_seen + defer_seen / 8 + 1024, grew = true;
    if (h.pool_used > k.pool) k.pool = want(h.pool_used, 4096, n_bytes), grew = true;
    if (!grew) {  // overflow flagged but counts look fine (reservation raced past the cap): grow everything
        k.rec *= 2, k.slots *= 2, k.defer *= 2, k.pool *= 2;
    }
    return true;
}

void require_ready(blu_ctx* c) {
    if (!c->tax) throw std::invalid_argument("no taxonomy loaded (call blu_taxonomy_load_json first)");
}

// Non-contiguous hit tables (the reference groups rows through a HashMap<String, Vec<_>>, mod.rs:145,192, so the
// rows of one query need not be adjacent).  Rare (BLAST emits queries contiguously, blutils appends whole chunks),
// so it is handled by data movement only: rows are regrouped by query id -- first-appearance order of the queries,
// file order inside a query, which is all the consensus depends on -- and the GPU pipeline runs on the regrouped
// text.  No consensus arithmetic happens on the host.
std::string regroup_by_query(const char* text, uint64_t n) {
    struct Row {
        uint64_t off;
        uint32_t len;
    };
    std::unordered_map<std::string_view, uint32_t> gid;
    std::vector<std::vector<Row>> groups;
    for (uint64_t p = 0; p < n;) {
        const char* nl = (const char*)memchr(text + p, '\n', n - p);
        const uint64_t e = nl ? (uint64_t)(nl - text) : n;
        if (e > p) {
            const char* tab = (const char*)memchr(text + p, '\t', e - p);
            std::string_view q(text + p, tab ? (size_t)(tab - (text + p)) : (size_t)(e - p));
            auto it = gid.find(q);
            uint32_t g;
            if (it == gid.end()) {
                g = (uint32_t)groups.size();
                gid.emplace(q, g);
                groups.emplace_back();
            } else
                g = it->second;
            groups[g].push_back({p, (uint32_t)(e - p)});
        }
        p = e + 1;
    }
    std::string out;
    out.reserve(n + 1);
    for (auto& g : groups)
        for (auto& r : g) {
            out.append(text + r.off, r.len);
            out.push_back('\n');
        }
    return out;
}

struct NonContiguous : std::runtime_error {
    using std::runtime_error::runtime_error;
};

// BLU_TIMELINE=1: device-side event timeline of one resident run (ms from its start), printed to stderr.  A debugging
// aid for the overlap of result downloads with the next range's kernels; events are only created when it is on.
struct Timeline {
    struct Mark {
        const char* what;
        int range;
        cudaEvent_t ev;
        double host_ms;
    };
    bool on = getenv("BLU_TIMELINE") != nullptr;
    std::vector<Mark> marks;
    cudaEvent_t t0 = nullptr;
    std::chrono::steady_clock::time_point h0;
    void start(cudaStream_t s) {
        if (!on) return;
        cudaEventCreate(&t0);
        cudaEventRecord(t0, s);
        h0 = std::chrono::steady_clock::now();
    }
    void mark(const char* what, int range, cudaStream_t s) {
        if (!on) return;
        cudaEvent_t e;
        cudaEventCreate(&e);
        cudaEventRecord(e, s);
        marks.push_back({what, range, e, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - h0).count()});
    }
    void host(const char* what, int range) {
        if (!on) return;
        marks.push_back({what, range, nullptr, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - h0).count()});
    }
    void dump() {
        if (!on) return;
        cudaDeviceSynchronize();
        for (auto& m : marks) {
            float ms = -1;
            if (m.ev) cudaEventElapsedTime(&ms, t0, m.ev);
            fprintf(stderr, "[timeline] range %d %-22s device %8.3f ms   host(issue) %8.3f ms\n", m.range, m.what, ms, m.host_ms);
            if (m.ev) cudaEventDestroy(m.ev);
        }
        if (t0) cudaEventDestroy(t0);
        marks.clear();
    }
};

// --- text resident on the device ------------------------------------------------------------------------------
void run_device_serial(blu_ctx* c, const uint8_t* dtext, uint64_t n, cudaStream_t s, blu_result* r) {
    require_ready(c);
    if (n == 0) throw DataErr("empty blast output (the reference's CsvReader fails on an empty file)");
    if (((uintptr_t)dtext & 15) != 0) throw std::invalid_argument("device text must be 16-byte aligned");
    c->tm = blu_timings{};
    Caps k = initial_caps(n);
    // Large resident tables are processed in a few query-aligned ranges so that the download of one range's records
    // overlaps the kernels of the next (the unfinished last query of a range simply starts the next range: the
    // buffer is contiguous, nothing is copied).
    // Equal ranges measured best (tools/range_split.py, profiles/README.md): the step is K_1 + ... + K_n + D_n + ~0.1 ms of
    // host round trips per range; a smaller last range makes an earlier download spill past its kernels instead.
    // BLU_RANGE_FRACS="0.3,0.3,0.25,0.15" overrides the split (measurement knob).
    std::vector<double> fracs(n >= (512ull << 20) ? 4 : (n >= (128ull << 20) ? 2 : 1), 1.0);
    if (const char* ev = getenv("BLU_RANGE_FRACS")) {
        std::vector<double> f;
        for (const char* q = ev; *q;) {
            char* e2 = nullptr;
            double v = strtod(q, &e2);
            if (e2 == q || !(v > 0)) break;
            f.push_back(v);
            q = *e2 == ',' ? e2 + 1 : e2;
        }
        if (!f.empty() && f.size() <= 64) fracs = f;
    }
    const uint64_t n_ranges = fracs.size();
    std::vector<uint64_t> range_end(n_ranges);
    {
        double tot = 0, acc = 0;
        for (double f : fracs) tot += f;
        for (uint64_t i = 0; i < n_ranges; i++) {
            acc += fracs[i];
            const uint64_t e = (uint64_t)((double)n * (acc / tot));
            range_end[i] = std::min<uint64_t>(n, (e + (uint64_t)kTile - 1) / (uint64_t)kTile * (uint64_t)kTile);
        }
        range_end[n_ranges - 1] = n;
    }
    Downloader dl(c, r);
    for (int attempt = 0; attempt < 8; attempt++) {
        ensure_out(c, k);
        reset_counters_async(c, s, true);
        bool retry = false;
        uint64_t begin = 0;
        uint32_t rec_done = 0;
        uint64_t defer_total = 0;
        double ms_tile = 0, ms_long = 0, ms_post = 0;
        uint64_t launches = 0;
        Counters h{};
        Timeline tl;
        tl.start(s);
        for (uint64_t ri = 0; ri < n_ranges && !retry; ri++) {
            const bool final_range = ri + 1 == n_ranges;
            const uint64_t end = range_end[ri];
            if (end <= begin && !final_range) continue;
            if (ri) reset_counters_async(c, s, false);
            tl.mark("tile+longrun begin", (int)ri, s);
            launch_chunk(c, dtext, begin, end, final_range, k, s, rec_done, true);
            tl.mark("tile+longrun end", (int)ri, s);
            launches++;
            read_counters(c, s);
            tl.host("counters #1 on host", (int)ri);
            h = *c->h_ctr;
            ms_tile += ev_ms(c->ev[0], c->ev[1]);
            ms_long += ev_ms(c->ev[1], c->ev[2]);
            defer_total = std::max<uint64_t>(defer_total, h.n_defer);
            c->tm.n_deferred_runs += h.n_defer;
            if (h.cap_overflow || n_rec_of(h) > k.rec || n_slots_of(h) > k.slots || h.n_defer > k.defer) {
                grow_caps(h, k, h.n_defer, (double)n / (double)std::max<uint64_t>(end, 1), n);
                retry = true;
                break;
            }
            check_device_error(c, h, 0);
            CK(cudaEventRecord(c->ev[3], s));
            launch_gather(c, dtext, end, rec_done, (uint32_t)n_rec_of(h), k, s);
            if (final_range) launch_dup(c, (uint32_t)n_rec_of(h), s);
            CK(cudaEventRecord(c->ev[4], s));
            tl.mark("consensus+gather end", (int)ri, s);
            rec_done = (uint32_t)n_rec_of(h);
            read_counters(c, s);
            tl.host("counters #2 on host", (int)ri);
            h = *c->h_ctr;
            ms_post += ev_ms(c->ev[3], c->ev[4]);
            if (h.cap_overflow || h.pool_used > k.pool) {
                grow_caps(h, k, defer_total, (double)n / (double)std::max<uint64_t>(end, 1), n);
                retry = true;
                break;
            }
            check_device_error(c, h, 0);
            if (!final_range) {
                if (ri == 0 && end > 0) {
                    // size the pinned result buffers from the density of the first range
                    const double f = 1.15 * (double)n / (double)end;
                    dl.reserve((uint64_t)(n_rec_of(h) * f) + 4096, (uint64_t)(n_slots_of(h) * f) + 8192, (uint64_t)(h.pool_used * f) + 65536);
                }
                tl.mark("download begin", (int)ri, c->d2h_stream);
                dl.push(h);  // runs on the download stream under the next range's kernels
                tl.mark("download end", (int)ri, c->d2h_stream);
                begin = h.tail_start != ~0ull ? h.tail_start : end;
            }
        }
        if (retry) {
            dl.abandon();
            c->tm = blu_timings{};
            continue;
        }
        if (h.dup_found) {
            dl.abandon();
            throw NonContiguous("a query id occurs in two non-adjacent groups of rows");
        }
        if (n_rec_of(h) == 0) throw DataErr("the blast output holds no rows");
        c->tm.ms_tile_kernel = ms_tile;
        c->tm.ms_longrun_kernel = ms_long;
        c->tm.ms_gather_kernel = ms_post;
        c->tm.ms_total_device = ms_tile + ms_long + ms_post;
        c->tm.text_bytes = n;
        c->tm.taxonomy_bytes = c->tax->device_bytes();
        c->tm.n_tile_launches = launches;
        tl.mark("last download begin", (int)n_ranges - 1, c->d2h_stream);
        dl.finish(h);
        tl.mark("last download end", (int)n_ranges - 1, c->d2h_stream);
        tl.host("run complete", (int)n_ranges - 1);
        tl.dump();
        return;
    }
    throw std::runtime_error("output capacity did not converge");
}

void run_device_pipelined(blu_ctx* c, const uint8_t* dtext, uint64_t n, cudaStream_t s, blu_result* r) {
    require_ready(c);
    if (n == 0) throw DataErr("empty blast output (the reference's CsvReader fails on an empty file)");
    if (((uintptr_t)dtext & 15) != 0) throw std::invalid_argument("device text must be 16-byte aligned");
    c->tm = blu_timings{};
    Caps k = initial_caps(n);
    // Large resident tables are processed in a few query-aligned ranges so that the download of one range's records
    // overlaps the kernels of the next (the unfinished last query of a range simply starts the next range: the
    // buffer is contiguous, nothing is copied).
    // Equal ranges measured best (tools/range_split.py, profiles/README.md): the step is K_1 + ... + K_n + D_n + ~0.1 ms of
    // host round trips per range; a smaller last range makes an earlier download spill past its kernels instead.
    // BLU_RANGE_FRACS="0.3,0.3,0.25,0.15" overrides the split (measurement knob).
    std::vector<double> fracs(n >= (512ull << 20) ? 4 : (n >= (128ull << 20) ? 2 : 1), 1.0);
    if (const char* ev = getenv("BLU_RANGE_FRACS")) {
        std::vector<double> f;
        for (const char* q = ev; *q;) {
            char* e2 = nullptr;
            double v = strtod(q, &e2);
            if (e2 == q || !(v > 0)) break;
            f.push_back(v);
            q = *e2 == ',' ? e2 + 1 : e2;
        }
        if (!f.empty() && f.size() <= 64) fracs = f;
    }
    const uint64_t n_ranges = fracs.size();
    std::vector<uint64_t> range_end(n_ranges);
    {
        double tot = 0, acc = 0;
        for (double f : fracs) tot += f;
        for (uint64_t i = 0; i < n_ranges; i++) {
            acc += fracs[i];
            const uint64_t e = (uint64_t)((double)n * (acc / tot));
            range_end[i] = std::min<uint64_t>(n, (e + (uint64_t)kTile - 1) / (uint64_t)kTile * (uint64_t)kTile);
        }
        range_end[n_ranges - 1] = n;
    }
    Downloader dl(c, r);
    for (int attempt = 0; attempt < 8; attempt++) {
        ensure_out(c, k);
        reset_counters_async(c, s, true);
        bool retry = false;
        uint32_t rec_done = 0;
        uint64_t defer_total = 0;
        double ms_tile = 0, ms_long = 0, ms_post = 0;
        uint64_t launches = 0;
        Counters h{}, h2{};
        Timeline tl;
        tl.start(s);
        auto launch_range = [&](uint64_t ri, uint64_t begin) {
            if (ri) reset_counters_async(c, s, false);
            tl.mark("tile+longrun begin", (int)ri, s);
            launch_chunk(c, dtext, begin, range_end[ri], ri + 1 == n_ranges, k, s, rec_done, true);
            tl.mark("tile+longrun end", (int)ri, s);
            launches++;
        };
        // The kernels of range i+1 are queued BEFORE the host looks at the post-pass of range i: the GPU goes from the
        // gather of one range straight into the tile kernel of the next, and the download of range i still starts the
        // moment its post-pass is done (ev[5] + a second pinned copy of the counters, taken between the two).
        uint64_t ri = 0;
        launch_range(0, 0);
        for (;;) {
            const bool final_range = ri + 1 == n_ranges;
            const uint64_t end = range_end[ri];
            read_counters(c, s);
            tl.host("counters #1 on host", (int)ri);
            h = c->h_ctr[0];
            ms_tile += ev_ms(c->ev[0], c->ev[1]);
            ms_long += ev_ms(c->ev[1], c->ev[2]);
            defer_total = std::max<uint64_t>(defer_total, h.n_defer);
            c->tm.n_deferred_runs += h.n_defer;
            if (h.cap_overflow || n_rec_of(h) > k.rec || n_slots_of(h) > k.slots || h.n_defer > k.defer || h.pool_used > k.pool) {
                grow_caps(h, k, h.n_defer, (double)n / (double)std::max<uint64_t>(end, 1), n);
                retry = true;
                break;
            }
            check_device_error(c, h, 0);
            CK(cudaEventRecord(c->ev[3], s));
            launch_gather(c, dtext, end, rec_done, (uint32_t)n_rec_of(h), k, s);
            if (final_range) launch_dup(c, (uint32_t)n_rec_of(h), s);
            CK(cudaEventRecord(c->ev[4], s));
            tl.mark("consensus+gather end", (int)ri, s);
            CK(cudaMemcpyAsync(&c->h_ctr[1], c->d_ctr, sizeof(Counters), cudaMemcpyDeviceToHost, s));
            CK(cudaEventRecord(c->ev[5], s));
            rec_done = (uint32_t)n_rec_of(h);
            uint64_t next = ri;
            if (!final_range) {
                const uint64_t begin = h.tail_start != ~0ull ? h.tail_start : end;  // final after tile + long-run
                next = ri + 1;
                while (next + 1 < n_ranges && range_end[next] <= begin) next++;  // a query longer than a whole range
                launch_range(next, begin);
            }
            CK(cudaEventSynchronize(c->ev[5]));
            tl.host("counters #2 on host", (int)ri);
            h2 = c->h_ctr[1];
            ms_post += ev_ms(c->ev[3], c->ev[4]);
            if (h2.cap_overflow || h2.pool_used > k.pool) {
                grow_caps(h2, k, defer_total, (double)n / (double)std::max<uint64_t>(end, 1), n);
                retry = true;
                break;
            }
            check_device_error(c, h2, 0);
            if (final_range) break;
            if (ri == 0 && end > 0) {
                // size the pinned result buffers from the density of the first range
                const double f = 1.15 * (double)n / (double)end;
                dl.reserve((uint64_t)(n_rec_of(h2) * f) + 4096, (uint64_t)(n_slots_of(h2) * f) + 8192, (uint64_t)(h2.pool_used * f) + 65536);
            }
            tl.mark("download begin", (int)ri, c->d2h_stream);
            dl.push(h2);  // runs on the download stream under the next range's kernels
            tl.mark("download end", (int)ri, c->d2h_stream);
            ri = next;
        }
        if (retry) {
            CK(cudaStreamSynchronize(s));  // the next range may already be running
            dl.abandon();
            c->tm = blu_timings{};
            continue;
        }
        if (h2.dup_found) {
            dl.abandon();
            throw NonContiguous("a query id occurs in two non-adjacent groups of rows");
        }
        if (n_rec_of(h2) == 0) throw DataErr("the blast output holds no rows");
        c->tm.ms_tile_kernel = ms_tile;
        c->tm.ms_longrun_kernel = ms_long;
        c->tm.ms_gather_kernel = ms_post;
        c->tm.ms_total_device = ms_tile + ms_long + ms_post;
        c->tm.text_bytes = n;
        c->tm.taxonomy_bytes = c->tax->device_bytes();
        c->tm.n_tile_launches = launches;
        tl.mark("last download begin", (int)n_ranges - 1, c->d2h_stream);
        dl.finish(h2);
        tl.mark("last download end", (int)n_ranges - 1, c->d2h_stream);
        tl.host("run complete", (int)n_ranges - 1);
        tl.dump();
        return;
    }
    throw std::runtime_error("output capacity did not converge");
}

// BLU_RESIDENT_SERIAL=1 selects the loop with two host round trips per range (A/B measurement).  On any failure nothing
// may still be reading the caller's device text when the call returns.
void run_device(blu_ctx* c, const uint8_t* dtext, uint64_t n, cudaStream_t s, blu_result* r) {
    static const bool serial = getenv("BLU_RESIDENT_SERIAL") != nullptr;
    try {
        if (serial)
            run_device_serial(c, dtext, n, s, r);
        else
            run_device_pipelined(c, dtext, n, s, r);
    } catch (...) {
        cudaStreamSynchronize(s);
        cudaStreamSynchronize(c->d2h_stream);
        throw;
    }
}

// --- where the streamed path takes the text of chunk `ci` from -------------------------------------------------
struct ChunkSource {
    virtual ~ChunkSource() = default;
    virtual const char* acquire(uint64_t ci) = 0;  // host pointer to chunk ci (may block until it is there)
    virtual void release(uint64_t) {}              // the host->device copy of chunk ci has completed
    virtual void restart() {}                      // the run starts over from chunk 0 (no copy is in flight)
};

struct MemorySource : ChunkSource {
    const char* text;
    uint64_t chunk;
    MemorySource(const char* t, uint64_t ch) : text(t), chunk(ch) {}
    const char* acquire(uint64_t ci) override { return text + ci * chunk; }
};

// A file, read by parallel pread()s into a ring of three pinned staging buffers while earlier chunks are copied to
// the device and processed: the file never has to fit in (pinned) host memory, and reading overlaps everything else.
class FileSource : public ChunkSource {
    static constexpr uint64_t kRing = 3;
    blu_ctx* c_;
    int fd_;
    uint64_t n_, chunk_, n_chunks_;
    int n_readers_;
    PinnedBuf buf_[kRing];
    std::mutex mu_;
    std::condition_variable cv_;
    uint64_t staged_ = 0;    // chunks [0, staged_) of this epoch are in their buffers
    uint64_t released_ = 0;  // chunks [0, released_) may be overwritten
    uint64_t epoch_ = 0;
    bool stop_ = false;
    std::string error_;
    std::thread coordinator_;

    bool read_span(char* dst, uint64_t off, uint64_t len, std::string& err) const {
        while (len) {
            ssize_t got = pread(fd_, dst, (size_t)std::min<uint64_t>(len, 1ull << 30), (off_t)off);
            if (got < 0 && errno == EINTR) continue;
            if (got <= 0) {
                err = got == 0 ? "the blast output shrank while it was read" : std::string("read error on the blast output: ") + strerror(errno);
                return false;
            }
            dst += got, off += (uint64_t)got, len -= (uint64_t)got;
        }
        return true;
    }
    bool read_chunk(uint64_t ci, std::string& err) const {
        char* dst = (char*)buf_[ci % kRing].p;
        const uint64_t off = ci * chunk_, len = std::min(chunk_, n_ - off);
        const uint64_t slice = std::max<uint64_t>(((len + n_readers_ - 1) / n_readers_ + 4095) & ~4095ull, 1ull << 20);
        std::vector<std::thread> th;
        std::vector<std::string> errs((size_t)n_readers_);
        std::atomic<bool> ok{true};
        int t = 0;
        try {
            for (uint64_t o = slice; o < len; o += slice, t++)
                th.emplace_back([&, o, t] {
                    if (!read_span(dst + o, off + o, std::min(slice, len - o), errs[(size_t)t])) ok = false;
                });
        } catch (...) {  // could not start a thread: the ones already running must be joined before unwinding
            for (auto& x : th) x.join();
            throw;
        }
        std::string e0;
        if (!read_span(dst, off, std::min(slice, len), e0)) ok = false;
        for (auto& x : th) x.join();
        if (!ok) {
            err = e0;
            for (auto& e : errs)
                if (err.empty()) err = e;
        }
        return ok;
    }
    void run() {
        uint64_t my_epoch = 0, next = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return stop_ || epoch_ != my_epoch || (next < n_chunks_ && next < released_ + kRing); });
                if (stop_) return;
                if (epoch_ != my_epoch) {
                    my_epoch = epoch_, next = 0;
                    continue;
                }
            }
            std::string err;
            bool ok = false;
            try {
                ok = read_chunk(next, err);
            } catch (const std::exception& e) {  // e.g. no more threads: reported through acquire(), not std::terminate
                err = std::string("reading the blast output failed: ") + e.what();
            }
            {
                std::lock_guard<std::mutex> g(mu_);
                if (epoch_ == my_epoch) {
                    if (ok)
                        staged_ = next + 1;
                    else
                        error_ = err;
                }
            }
            cv_.notify_all();
            next = ok ? next + 1 : n_chunks_;  // after an error: idle until restart() or the destructor
        }
    }

   public:
    FileSource(blu_ctx* c, int fd, uint64_t n, uint64_t chunk) : c_(c), fd_(fd), n_(n), chunk_(chunk), n_chunks_((n + chunk - 1) / chunk) {
        const unsigned hc = std::thread::hardware_concurrency();
        n_readers_ = (int)std::min<unsigned>(8, std::max<unsigned>(1, hc ? hc : 4));
        if (const char* ev = getenv("BLU_READ_THREADS")) n_readers_ = std::max(1, std::min(64, atoi(ev)));
        const uint64_t used = std::min<uint64_t>(kRing, n_chunks_);
        try {
            for (uint64_t i = 0; i < used; i++) buf_[i] = c->acquire(std::min(chunk, n));
            coordinator_ = std::thread([this] { run(); });
        } catch (...) {  // a constructor that throws gets no destructor call
            for (auto& b : buf_) c_->pool->release(b);
            throw;
        }
    }
    ~FileSource() override {
        {
            std::lock_guard<std::mutex> g(mu_);
            stop_ = true;
        }
        cv_.notify_all();
        coordinator_.join();
        for (auto& b : buf_) c_->pool->release(b);
    }
    const char* acquire(uint64_t ci) override {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return staged_ > ci || !error_.empty(); });
        if (staged_ <= ci) throw IoErr(error_);
        return (const char*)buf_[ci % kRing].p;
    }
    void release(uint64_t ci) override {
        {
            std::lock_guard<std::mutex> g(mu_);
            released_ = std::max(released_, ci + 1);
        }
        cv_.notify_all();
    }
    void restart() override {
        {
            std::lock_guard<std::mutex> g(mu_);
            epoch_++;
            staged_ = released_ = 0;
            error_.clear();
        }
        cv_.notify_all();
    }
};

// --- text on the host: chunked, double-buffered H2D overlapped with the kernels --------------------------------
inline uint64_t chunk_bytes_of(const blu_ctx* c, uint64_t dflt) { return c->opts.chunk_bytes ? ((c->opts.chunk_bytes + 127) & ~127ull) : dflt; }

void run_host_chunks(blu_ctx* c, ChunkSource& src, uint64_t n, const uint64_t chunk, blu_result* r) {
    require_ready(c);
    if (n == 0) throw DataErr("empty blast output (the reference's CsvReader fails on an empty file)");
    c->tm = blu_timings{};
    const uint64_t carry = c->carry_bytes;
    const uint64_t n_chunks = (n + chunk - 1) / chunk;
    const bool single = n_chunks == 1;
    const uint64_t buf_bytes = (single ? 0 : carry) + std::min(chunk, n) + 256;
    const uint64_t text_off = single ? 0 : carry;  // where H2D data lands inside a buffer
    c->d_text[0].ensure(buf_bytes);
    if (!single) c->d_text[1].ensure(buf_bytes);
    Caps k = initial_caps(n);
    cudaStream_t s = c->stream, cs = c->copy_stream;
    Downloader dl(c, r);
    auto t0 = std::chrono::steady_clock::now();
    for (int attempt = 0; attempt < 8; attempt++) {
        ensure_out(c, k);
        reset_counters_async(c, s, true);
        bool retry = false;
        uint64_t tail_len = 0;           // bytes carried from the previous chunk
        uint32_t rec_done = 0;           // records already gathered
        uint64_t defer_total = 0;
        double ms_tile = 0, ms_long = 0, ms_gather = 0;
        uint64_t launches = 0;
        auto issue_h2d = [&](uint64_t ci) {
            const uint64_t off = ci * chunk, len = std::min(chunk, n - off);
            CK(cudaMemcpyAsync(c->d_text[ci & 1].p + text_off, src.acquire(ci), len, cudaMemcpyHostToDevice, cs));
            CK(cudaEventRecord(c->ev_h2d[ci & 1], cs));
            c->tm.h2d_bytes += len;
        };
        issue_h2d(0);
        for (uint64_t ci = 0; ci < n_chunks && !retry; ci++) {
            const uint64_t off = ci * chunk, len = std::min(chunk, n - off);
            const bool final_chunk = ci + 1 == n_chunks;
            uint8_t* buf = c->d_text[ci & 1].p;
            CK(cudaStreamWaitEvent(s, c->ev_h2d[ci & 1], 0));
            if (tail_len) {
                // the unfinished last query of the previous chunk goes in front of this chunk's text
                const uint8_t* prev = c->d_text[(ci - 1) & 1].p;
                const uint64_t prev_end = text_off + std::min(chunk, n - (ci - 1) * chunk);
                CK(cudaMemcpyAsync(buf + text_off - tail_len, prev + prev_end - tail_len, tail_len, cudaMemcpyDeviceToDevice, s));
            }
            CK(cudaEventRecord(c->ev_free[(ci + 1) & 1], s));  // previous buffer no longer read after this point
            if (ci) reset_counters_async(c, s, false);
            const uint64_t begin = text_off - tail_len, end = text_off + len;
            launch_chunk(c, buf, begin, end, final_chunk, k, s, rec_done, true);
            if (!final_chunk) {  // behind the launches: a file source may block here until the next chunk has been read
                CK(cudaStreamWaitEvent(cs, c->ev_free[(ci + 1) & 1], 0));
                issue_h2d(ci + 1);
            }
            read_counters(c, s);
            src.release(ci);  // `s` waited for this chunk's copy, so it has completed
            Counters h = *c->h_ctr;
            ms_tile += ev_ms(c->ev[0], c->ev[1]);
            ms_long += ev_ms(c->ev[1], c->ev[2]);
            defer_total = std::max<uint64_t>(defer_total, h.n_defer);
            if (h.cap_overflow || n_rec_of(h) > k.rec || n_slots_of(h) > k.slots || h.n_defer > k.defer) {
                grow_caps(h, k, h.n_defer, (double)n / (double)(off + len), n);
                retry = true;
                break;
            }
            check_device_error(c, h, off - text_off);  // err_off is in buffer coordinates
            CK(cudaEventRecord(c->ev[3], s));
            launch_gather(c, buf, end, rec_done, n_rec_of(h), k, s);
            CK(cudaEventRecord(c->ev[4], s));
            rec_done = n_rec_of(h);
            if (!final_chunk) {
                if (h.tail_start == ~0ull)
                    tail_len = 0;
                else {
                    tail_len = end - h.tail_start;
                    if (tail_len > carry) throw UnsupportedErr("a single query spans more than the 64 MiB carry buffer between streamed chunks");
                }
            }
            read_counters(c, s);  // (also: gather must finish before this buffer is overwritten two chunks later)
            ms_gather += ev_ms(c->ev[3], c->ev[4]);
            c->tm.n_deferred_runs += h.n_defer;
            launches++;
            if (!final_chunk) {
                const Counters hg = *c->h_ctr;
                if (hg.cap_overflow || hg.pool_used > k.pool) {
                    grow_caps(hg, k, defer_total, (double)n / (double)(off + len), n);
                    retry = true;
                    break;
                }
                if (ci == 0) {
                    const double f = 1.15 * (double)n / (double)len;
                    dl.reserve((uint64_t)(n_rec_of(hg) * f) + 4096, (uint64_t)(n_slots_of(hg) * f) + 8192, (uint64_t)(hg.pool_used * f) + 65536);
                }
                dl.push(hg);  // result download of this chunk runs beside the next chunks' host->device copies
            }
        }
        if (retry) {
            CK(cudaStreamSynchronize(cs));
            CK(cudaStreamSynchronize(s));
            dl.abandon();
            src.restart();
            c->tm = blu_timings{};
            continue;
        }
        launch_dup(c, rec_done, s);
        read_counters(c, s);
        Counters h = *c->h_ctr;
        if (h.cap_overflow || h.pool_used > k.pool) {
            grow_caps(h, k, defer_total, 1.0, n);
            dl.abandon();
            src.restart();
            c->tm = blu_timings{};
            continue;
        }
        check_device_error(c, h, 0);
        if (h.dup_found) {
            dl.abandon();
            throw NonContiguous("a query id occurs in two non-adjacent groups of rows");
        }
        if (n_rec_of(h) == 0) throw DataErr("the blast output holds no rows");
        c->tm.ms_tile_kernel = ms_tile;
        c->tm.ms_longrun_kernel = ms_long;
        c->tm.ms_gather_kernel = ms_gather;
        c->tm.text_bytes = n;
        c->tm.taxonomy_bytes = c->tax->device_bytes();
        c->tm.n_tile_launches = launches;
        dl.finish(h);
        c->tm.ms_total_device = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        return;
    }
    throw std::runtime_error("output capacity did not converge");
}

// On any failure nothing may still be reading the caller's text (or a staging buffer) when the call returns.
void run_host_source(blu_ctx* c, ChunkSource& src, uint64_t n, uint64_t chunk, blu_result* r) {
    try {
        run_host_chunks(c, src, n, chunk, r);
    } catch (...) {
        cudaStreamSynchronize(c->copy_stream);
        cudaStreamSynchronize(c->stream);
        cudaStreamSynchronize(c->d2h_stream);
        throw;
    }
}

void run_host(blu_ctx* c, const char* text, uint64_t n, blu_result* r) {
    const uint64_t chunk = chunk_bytes_of(c, 256ull << 20);
    MemorySource src(text, chunk);
    run_host_source(c, src, n, chunk, r);
}

}  // namespace

// =================================================================================================================
// C ABI
// =================================================================================================================
extern "C" {

int blu_abi_version(void) { return BLU_ABI_VERSION; }

const char* blu_last_error(const blu_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int blu_ctx_create(const blu_opts* opts, blu_ctx** out) {
    if (!opts || !out) return fail(nullptr, BLU_ERR_ARG, "null argument");
    *out = nullptr;
    if (opts->taxon < 0 || opts->taxon > 3) return fail(nullptr, BLU_ERR_ARG, "bad taxon");
    if (opts->strategy < 0 || opts->strategy > 1) return fail(nullptr, BLU_ERR_ARG, "bad strategy");
    if (opts->taxon == BLU_TAXON_CUSTOM && !opts->has_custom)
        return fail(nullptr, BLU_ERR_DATA, "Custom taxon values are required when the custom taxon option is selected.");
    auto c = std::make_unique<blu_ctx>();
    c->opts = *opts;
    c->cut.taxon = opts->taxon;
    c->cut.has_custom = opts->has_custom != 0;
    for (int i = 0; i < 8; i++) c->cut.custom[i] = opts->custom[i];
    c->device = opts->device;
    int rc = guarded(nullptr, [&] {
        make_backbone(c->cut);  // validates the custom values
        int n = 0;
        cudaError_t e = cudaGetDeviceCount(&n);
        if (e != cudaSuccess || n == 0) throw CudaErr(std::string("no CUDA device available: ") + cudaGetErrorString(e));
        if (c->device < 0 || c->device >= n) throw CudaErr("device ordinal out of range");
        CK(cudaSetDevice(c->device));
        cudaDeviceProp prop;
        CK(cudaGetDeviceProperties(&prop, c->device));
        if (prop.major < 10) throw CudaErr(std::string("device '") + prop.name + "' is not sm_100-class; this library only carries sm_100a code");
        c->sms = prop.multiProcessorCount;
        CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&c->d2h_stream, cudaStreamNonBlocking));
        for (auto& e2 : c->ev) CK(cudaEventCreate(&e2));
        for (auto& e2 : c->ev_h2d) CK(cudaEventCreateWithFlags(&e2, cudaEventDisableTiming));
        for (auto& e2 : c->ev_free) CK(cudaEventCreateWithFlags(&e2, cudaEventDisableTiming));
        CK(cudaMalloc((void**)&c->d_ctr, sizeof(Counters)));
        CK(cudaHostAlloc((void**)&c->h_ctr, 2 * sizeof(Counters), cudaHostAllocDefault));  // [0]: read_counters, [1]: post-pass snapshot
        CK(kernels_set_attributes());
    });
    if (rc != BLU_OK) {
        std::string msg = g_create_error;
        blu_ctx_destroy(c.release());
        g_create_error = msg;
        return rc;
    }
    *out = c.release();
    return BLU_OK;
}

void blu_ctx_destroy(blu_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->copy_stream) cudaStreamSynchronize(c->copy_stream);
    if (c->d2h_stream) cudaStreamSynchronize(c->d2h_stream);
    c->d_lin_off.release(), c->d_lvl.release(), c->d_bean.release(), c->d_irank.release(), c->d_cut.release();
    c->d_rcls.release(), c->d_acls.release(), c->d_linok.release(), c->d_slots.release();
    c->d_text[0].release(), c->d_text[1].release(), c->d_rec.release(), c->d_beans.release(), c->d_accs.release();
    c->d_defer.release(), c->d_pool.release(), c->d_dup.release(), c->d_top.release();
    if (c->d_ctr) cudaFree(c->d_ctr);
    if (c->h_ctr) cudaFreeHost(c->h_ctr);
    c->pool->close();
    for (auto& e : c->ev)
        if (e) cudaEventDestroy(e);
    for (auto& e : c->ev_h2d)
        if (e) cudaEventDestroy(e);
    for (auto& e : c->ev_free)
        if (e) cudaEventDestroy(e);
    if (c->stream) cudaStreamDestroy(c->stream);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    if (c->d2h_stream) cudaStreamDestroy(c->d2h_stream);
    delete c;
}

int blu_custom_cutoffs_from_file(const char* path, blu_opts* opts, char* err, size_t errlen) {
    if (!path || !opts) return BLU_ERR_ARG;
    try {
        Cutoffs c;
        read_custom_cutoffs(path, c);
        opts->has_custom = 1;
        for (int i = 0; i < 8; i++) opts->custom[i] = c.custom[i];
        return BLU_OK;
    } catch (const std::exception& e) {
        if (err && errlen) snprintf(err, errlen, "%s", e.what());
        return BLU_ERR_DATA;  // the reference panics on every failure of CustomTaxon::from_file
    }
}

int blu_taxonomy_load_arrays(blu_ctx* c, const int64_t* taxids, const uint64_t* off, const char* blob, uint64_t n) {
    if (!c || (n && (!taxids || !off || !blob))) return fail(c, BLU_ERR_ARG, "null argument");
    return guarded(c, [&] {
        CK(cudaSetDevice(c->device));
        auto T = std::make_shared<HostTaxonomy>();
        static const uint64_t zero = 0;
        build_taxonomy(taxids, n ? off : &zero, blob, n, c->cut, *T);
        c->tax = T;
        upload_taxonomy(c);
    });
}

int blu_taxonomy_load_json(blu_ctx* c, const char* path) {
    if (!c || !path) return fail(c, BLU_ERR_ARG, "null argument");
    std::vector<int64_t> ids;
    std::vector<uint64_t> off;
    std::string blob;
    int rc = guarded(c, [&] { read_taxonomy_json(path, c->opts.use_taxid != 0, ids, off, blob); });
    if (rc != BLU_OK) return rc;
    return blu_taxonomy_load_arrays(c, ids.data(), off.data(), blob.data(), ids.size());
}

int blu_taxonomy_load_json_cached(blu_ctx* c, const char* path, const char* cache_path, int* cache_state) {
    if (!c || !path) return fail(c, BLU_ERR_ARG, "null argument");
    if (cache_state) *cache_state = 0;
    const std::string cpath = cache_path ? std::string(cache_path) : std::string(path) + ".blucache";
    return guarded(c, [&] {
        CK(cudaSetDevice(c->device));
        const TaxCacheKey key = make_cache_key(path, c->opts.use_taxid != 0, c->cut);  // IoErr if the JSON is unreadable
        auto T = std::make_shared<HostTaxonomy>();
        int state = 1;
        if (!load_taxonomy_cache(cpath.c_str(), key, *T)) {
            std::vector<int64_t> ids;
            std::vector<uint64_t> off;
            std::string blob;
            read_taxonomy_json(path, c->opts.use_taxid != 0, ids, off, blob);
            static const uint64_t zero = 0;
            build_taxonomy(ids.data(), ids.empty() ? &zero : off.data(), blob.data(), ids.size(), c->cut, *T);
            state = 0;
            try {
                save_taxonomy_cache(cpath.c_str(), key, *T);
            } catch (const IoErr&) {
                state = -1;
            }
        }
        c->tax = T;
        upload_taxonomy(c);
        if (cache_state) *cache_state = state;
    });
}

int blu_consensus_run_device(blu_ctx* c, const void* dtext, uint64_t n, void* stream, blu_result** out) {
    if (!c || !out || (!dtext && n)) return fail(c, BLU_ERR_ARG, "null argument");
    *out = nullptr;
    auto r = std::make_unique<blu_result>();
    r->pinned = c->pool;
    int rc = guarded(c, [&] {
        CK(cudaSetDevice(c->device));
        r->tax = c->tax;
        r->cut = c->cut;
        try {
            run_device(c, (const uint8_t*)dtext, n, stream ? (cudaStream_t)stream : c->stream, r.get());
        } catch (const NonContiguous&) {
            // regroup on the host (data movement only), then the normal streamed path
            std::string host(n, '\0');
            CK(cudaMemcpy(host.data(), dtext, n, cudaMemcpyDeviceToHost));
            std::string re = regroup_by_query(host.data(), n);
            host.clear();
            host.shrink_to_fit();
            run_host(c, re.data(), re.size(), r.get());
            c->tm.n_regrouped = 1;
        }
    });
    if (rc != BLU_OK) {
        blu_result_free(r.release());
        return rc;
    }
    *out = r.release();
    return BLU_OK;
}

int blu_consensus_run_host(blu_ctx* c, const char* text, uint64_t n, blu_result** out) {
    if (!c || !out || (!text && n)) return fail(c, BLU_ERR_ARG, "null argument");
    *out = nullptr;
    auto r = std::make_unique<blu_result>();
    r->pinned = c->pool;
    int rc = guarded(c, [&] {
        CK(cudaSetDevice(c->device));
        r->tax = c->tax;
        r->cut = c->cut;
        try {
            run_host(c, text, n, r.get());
        } catch (const NonContiguous&) {
            std::string re = regroup_by_query(text, n);
            run_host(c, re.data(), re.size(), r.get());
            c->tm.n_regrouped = 1;
        }
    });
    if (rc != BLU_OK) {
        blu_result_free(r.release());
        return rc;
    }
    *out = r.release();
    return BLU_OK;
}

// Streams the file (FileSource); only a non-contiguous table (regrouping needs all rows at once) is read whole.
int blu_consensus_run_file(blu_ctx* c, const char* path, blu_result** out) {
    if (!c || !path || !out) return fail(c, BLU_ERR_ARG, "null argument");
    *out = nullptr;
    struct Fd {
        int fd;
        ~Fd() {
            if (fd >= 0) close(fd);
        }
    } f{open(path, O_RDONLY | O_CLOEXEC)};
    struct stat st;
    if (f.fd < 0 || fstat(f.fd, &st) != 0 || !S_ISREG(st.st_mode))
        return fail(c, BLU_ERR_IO, "Unexpected error occurred on load table.");  // mod.rs:357-364
    const uint64_t n = (uint64_t)st.st_size;
    auto r = std::make_unique<blu_result>();
    r->pinned = c->pool;
    bool regroup = false;
    int rc = guarded(c, [&] {
        CK(cudaSetDevice(c->device));
        r->tax = c->tax;
        r->cut = c->cut;
        require_ready(c);
        if (n == 0) throw DataErr("empty blast output (the reference's CsvReader fails on an empty file)");
        try {
            const uint64_t chunk = chunk_bytes_of(c, 64ull << 20);
            FileSource src(c, f.fd, n, chunk);
            run_host_source(c, src, n, chunk, r.get());
        } catch (const NonContiguous&) {
            regroup = true;
        }
    });
    if (rc == BLU_OK && regroup) {
        rc = guarded(c, [&] {
            std::string all((size_t)n, '\0');
            for (uint64_t off = 0; off < n;) {
                ssize_t got = pread(f.fd, all.data() + off, (size_t)std::min<uint64_t>(n - off, 1ull << 30), (off_t)off);
                if (got < 0 && errno == EINTR) continue;
                if (got <= 0) throw IoErr("Unexpected error occurred on load table.");
                off += (uint64_t)got;
            }
            std::string re = regroup_by_query(all.data(), n);
            std::string().swap(all);
            blu_result_free(r.release());  // whatever the streamed attempt had downloaded
            r = std::make_unique<blu_result>();
            r->pinned = c->pool;
            r->tax = c->tax;
            r->cut = c->cut;
            run_host(c, re.data(), re.size(), r.get());
            c->tm.n_regrouped = 1;
        });
    }
    if (rc != BLU_OK) {
        blu_result_free(r.release());
        return rc;
    }
    *out = r.release();
    return BLU_OK;
}

int blu_result_add_headers(blu_result* r, const char* headers_nl, uint64_t len) {
    if (!r) return BLU_ERR_ARG;
    std::unordered_set<std::string_view> have;
    have.reserve(r->n_rec * 2);
    for (uint64_t i = 0; i < r->n_rec; i++) have.insert(std::string_view(r->pool() + r->rec()[i].query_off, r->rec()[i].query_len));
    const char* p = headers_nl;
    const char* e = headers_nl + len;
    while (p < e) {
        const char* nl = (const char*)memchr(p, '\n', e - p);
        const char* le = nl ? nl : e;
        std::string_view h(p, le - p);
        if (!have.count(h)) r->hitless.emplace_back(h);  // duplicates in `headers` stay duplicated, as in the reference
        p = le + 1;
    }
    return BLU_OK;
}

uint64_t blu_result_num_queries(const blu_result* r) { return r ? r->n_rec + r->hitless.size() : 0; }
uint64_t blu_result_num_rows(const blu_result* r) { return r ? r->n_rows : 0; }
const blu_record* blu_result_records(const blu_result* r) { return r ? r->rec() : nullptr; }
const blu_bean* blu_result_beans(const blu_result* r) { return r ? r->beans() : nullptr; }
const blu_acc* blu_result_accessions(const blu_result* r) { return r ? r->accs() : nullptr; }
const char* blu_result_pool(const blu_result* r, uint64_t* len) {
    if (len) *len = r ? r->pool_len : 0;
    return r ? r->pool() : nullptr;
}

uint64_t blu_result_checksum(const blu_result* r) {
    if (!r) return 0;
    ResultView v = make_view(r);
    return view_checksum(&v);
}

int blu_result_to_jsonl(const blu_result* r, char** out, uint64_t* len) {
    if (!r || !out || !len) return BLU_ERR_ARG;
    try {
        ResultView v = make_view(r);
        std::string s = view_to_jsonl(&v);
        char* buf = (char*)malloc(s.size() + 1);
        if (!buf) return BLU_ERR_INTERNAL;
        memcpy(buf, s.data(), s.size());
        buf[s.size()] = 0;
        *out = buf;
        *len = s.size();
        return BLU_OK;
    } catch (const std::exception&) {
        return BLU_ERR_INTERNAL;
    }
}

int blu_result_write(const blu_result* r, const char* path, int format, const char* run_id_in) {
    if (!r || format < 0 || format > 2) return BLU_ERR_ARG;
    try {
        ResultView v = make_view(r);
        return view_write(&v, path, format, run_id_in);
    } catch (const std::exception&) {
        return BLU_ERR_INTERNAL;
    }
}

int blu_result_write_tabular(const blu_result* r, const char* path, const char* run_id) {
    if (!r) return BLU_ERR_ARG;
    try {
        ResultView v = make_view(r);
        return view_write_tabular(&v, path, run_id);
    } catch (const std::exception&) {
        return BLU_ERR_INTERNAL;
    }
}

int blu_result_file_to_tabular(const char* in_path, const char* out_path, int input_format, const char* run_id, char* err, size_t errlen) {
    std::string msg;
    int rc;
    try {
        rc = result_file_to_tabular(in_path, out_path, input_format, run_id, msg);
    } catch (const std::exception& e) {
        msg = e.what();
        rc = BLU_ERR_INTERNAL;
    }
    if (rc != BLU_OK && err && errlen) snprintf(err, errlen, "%s", msg.c_str());
    return rc;
}

void blu_result_free(blu_result* r) {
    if (!r) return;
    if (r->pinned) {
        r->pinned->release(r->b_rec);
        r->pinned->release(r->b_beans);
        r->pinned->release(r->b_accs);
        r->pinned->release(r->b_pool);
    }
    delete r;
}

void blu_free(void* p) { free(p); }

int blu_ctx_last_timings(const blu_ctx* c, blu_timings* out) {
    if (!c || !out) return BLU_ERR_ARG;
    *out = c->tm;
    return BLU_OK;
}

int blu_ctx_measure_h2d(blu_ctx* c, uint64_t bytes, double* gbps) {
    if (!c || !gbps || !bytes) return BLU_ERR_ARG;
    return guarded(c, [&] {
        CK(cudaSetDevice(c->device));
        void* h = nullptr;
        void* d = nullptr;
        CK(cudaHostAlloc(&h, bytes, cudaHostAllocDefault));
        if (cudaMalloc(&d, bytes) != cudaSuccess) {
            cudaFreeHost(h);
            throw CudaErr("cudaMalloc failed");
        }
        memset(h, 1, bytes);
        double best = 0;
        for (int i = 0; i < 5; i++) {
            cudaEventRecord(c->ev[0], c->stream);
            cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, c->stream);
            cudaEventRecord(c->ev[1], c->stream);
            cudaStreamSynchronize(c->stream);
            double ms = ev_ms(c->ev[0], c->ev[1]);
            if (ms > 0) best = std::max(best, bytes / ms / 1e6);
        }
        cudaFree(d);
        cudaFreeHost(h);
        *gbps = best;
    });
}

int blu_shard_cuts(const char* text, uint64_t n, int n_shards, uint64_t* cuts) {
    if (!cuts || n_shards < 1 || (!text && n)) return BLU_ERR_ARG;
    auto first_field = [&](uint64_t p) {
        const char* e = (const char*)memchr(text + p, '\n', n - p);
        const uint64_t re = e ? (uint64_t)(e - text) : n;
        const char* t = (const char*)memchr(text + p, '\t', re - p);
        return std::string_view(text + p, t ? (size_t)(t - (text + p)) : (size_t)(re - p));
    };
    auto next_row = [&](uint64_t p) {  // start of the next non-empty row after the row containing p
        const char* e = (const char*)memchr(text + p, '\n', n - p);
        uint64_t q = e ? (uint64_t)(e - text) + 1 : n;
        while (q < n && text[q] == '\n') q++;
        return q;
    };
    cuts[0] = 0;
    cuts[n_shards] = n;
    for (int k = 1; k < n_shards; k++) {
        uint64_t p = std::max<uint64_t>(cuts[k - 1], n / (uint64_t)n_shards * (uint64_t)k);
        if (p >= n) {
            cuts[k] = n;
            continue;
        }
        // first row start at/after p
        if (p > 0 && text[p - 1] != '\n') p = next_row(p);
        while (p < n && text[p] == '\n') p++;
        if (p >= n) {
            cuts[k] = n;
            continue;
        }
        // previous non-empty row
        if (p > 0) {
            uint64_t q = p - 1;
            while (q > 0 && text[q] == '\n') q--;
            if (text[q] != '\n') {
                uint64_t st = q;
                while (st > 0 && text[st - 1] != '\n') st--;
                const std::string_view run = first_field(st);
                while (p < n && first_field(p) == run) p = next_row(p);  // never split a query
            }
        }
        cuts[k] = p;
    }
    return BLU_OK;
}

void* blu_host_alloc(uint64_t bytes) {
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) return nullptr;
    return p;
}
void blu_host_free(void* p) {
    if (p) cudaFreeHost(p);
}

}  // extern "C"
