// blu_synth.cpp -- seeded synthetic workloads for bench.py and the size-scaled parity tests:
// a lineage map (`.blutils.json` schema, reference core/src/domain/dtos/taxonomies_map.rs:6-32) and BLASTN
// outfmt-6 hit tables (column order: reference core/src/domain/dtos/blast_builder.rs:87).
// Counter-based: every query is a pure function of (seed, query index), so any range can be generated
// independently on any thread / rank.  Not part of the consensus path; separate libblu_synth.so.
//
// By construction the generated inputs avoid the reference's abort cases: every staxid exists in the map, all
// lineages share the root, and the domain cutoff (60 / 50) is below every generated pident (>= 80).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <unordered_set>
#include <vector>

namespace {

inline uint64_t splitmix(uint64_t& x) {
    uint64_t z = (x += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
inline uint64_t hash2(uint64_t a, uint64_t b) {
    uint64_t x = a * 0x9E3779B97F4A7C15ull ^ (b + 0x7F4A7C15ull);
    return splitmix(x);
}

struct Rng {
    uint64_t s;
    explicit Rng(uint64_t seed) : s(seed) {}
    uint64_t next() { return splitmix(s); }
    uint32_t below(uint32_t n) { return (uint32_t)((next() >> 32) * (uint64_t)n >> 32); }
    double unit() { return (next() >> 11) * (1.0 / 9007199254740992.0); }
};

// leaf index <-> digits (p,c,o,f,g,s); leaves are numbered depth-first so close indices are close relatives
struct Tree {
    uint32_t b[6];       // branching at p,c,o,f,g,s
    uint64_t n_leaves;   // prod(b) (>= requested taxa; the first n_taxa leaves are used)
    uint64_t below[7];   // leaves below one node of level l: below[6] = 1, below[5] = b_s, ...
};

Tree make_tree(uint64_t n_taxa) {
    Tree t;
    // roughly geometric growth towards the leaves
    const double w[6] = {0.10, 0.12, 0.16, 0.18, 0.20, 0.24};
    double ln = std::log((double)std::max<uint64_t>(n_taxa, 2));
    uint64_t prod = 1;
    for (int i = 0; i < 6; i++) {
        t.b[i] = (uint32_t)std::max(1.0, std::floor(std::exp(ln * w[i])));
        prod *= t.b[i];
    }
    int i = 5;
    while (prod < n_taxa) {  // top up from the species level upwards
        prod = prod / t.b[i] * (t.b[i] + 1);
        t.b[i]++;
        i = i == 0 ? 5 : i - 1;
    }
    t.n_leaves = prod;
    t.below[6] = 1;
    for (int l = 5; l >= 0; l--) t.below[l] = t.below[l + 1] * t.b[l];
    return t;
}

struct Synth {
    uint64_t n_taxa, seed;
    Tree tree;
    std::vector<int64_t> taxid;          // per leaf
    std::vector<std::string> text, num;  // per leaf lineage strings
};

const char* kRank[6] = {"p", "c", "o", "f", "g", "s"};

void build_lineages(Synth& S) {
    const Tree& t = S.tree;
    S.taxid.resize(S.n_taxa);
    S.text.resize(S.n_taxa);
    S.num.resize(S.n_taxa);
    // distinct pseudo-random taxids < 2^31
    std::unordered_set<uint32_t> used;
    used.reserve(S.n_taxa * 2);
    uint64_t st = S.seed ^ 0xABCDEF1234ull;
    for (uint64_t i = 0; i < S.n_taxa; i++) {
        uint32_t v;
        do v = (uint32_t)(splitmix(st) % 2147483000ull) + 2;
        while (!used.insert(v).second);
        S.taxid[i] = v;
    }
    char buf[64];
    for (uint64_t leaf = 0; leaf < S.n_taxa; leaf++) {
        uint64_t node[6];  // node id at each level = leaf / below[l+1]
        for (int l = 0; l < 6; l++) node[l] = leaf / t.below[l + 1];
        std::string& tx = S.text[leaf];
        std::string& nm = S.num[leaf];
        auto add = [&](const char* rank, const char* name, uint64_t id, uint64_t numeric) {
            if (!tx.empty()) {
                tx.push_back(';');
                nm.push_back(';');
            }
            snprintf(buf, sizeof buf, "%s__%s%llu", rank, name, (unsigned long long)id);
            tx += buf;
            snprintf(buf, sizeof buf, "%s__%llu", rank, (unsigned long long)numeric);
            nm += buf;
        };
        add("d", "bacteria", 0, 2);
        // 20 % of phyla sit under a clade
        if (hash2(S.seed ^ 1, node[0]) % 100 < 20) add("clade", "clade", node[0] / 2, 3000000 + node[0] / 2);
        // 2 % of lineages are truncated at family or genus
        const uint64_t hl = hash2(S.seed ^ 2, leaf);
        int depth = 6;
        if (hl % 100 < 2) depth = (hl >> 8) & 1 ? 5 : 4;
        for (int l = 0; l < depth; l++) {
            if (l == 5) {
                // 15 % of species sit in a species-group (half of those also in a species-subgroup)
                const uint64_t hg = hash2(S.seed ^ 3, node[5] / 3);
                if (hg % 100 < 15) {
                    add("species group", "sg", node[5] / 3, 4000000 + node[5] / 3);
                    if ((hg >> 10) & 1) add("species subgroup", "ssg", node[5] / 3, 5000000 + node[5] / 3);
                }
            }
            add(kRank[l], kRank[l], node[l], 10000000ull * (l + 1) + node[l]);
        }
        // 10 % carry a trailing strain
        if (depth == 6 && (hl >> 16) % 100 < 10) add("strain", "strain", leaf, 90000000ull + leaf);
    }
}

inline char* put_uint(char* p, uint64_t v) {
    char tmp[24];
    int n = 0;
    do tmp[n++] = (char)('0' + v % 10);
    while (v /= 10);
    while (n) *p++ = tmp[--n];
    return p;
}
inline char* put_fixed3(char* p, uint32_t milli) {  // ddd.ddd
    p = put_uint(p, milli / 1000);
    *p++ = '.';
    uint32_t f = milli % 1000;
    *p++ = (char)('0' + f / 100);
    *p++ = (char)('0' + f / 10 % 10);
    *p++ = (char)('0' + f % 10);
    return p;
}

// hits per query: mode 0 = fixed `hits`, mode 1 = Zipf(s=1.1) truncated to [1, hits]
uint32_t hits_for(uint64_t seed, uint64_t q, int mode, uint32_t hits) {
    if (mode == 0) return hits;
    Rng r(hash2(seed ^ 0x51F, q));
    // inverse CDF of the continuous power law x^-1.1 on [1, hits+1)
    const double a = 0.1;  // s - 1
    double u = r.unit();
    double hi = std::pow((double)hits + 1.0, -a);
    double x = std::pow(1.0 - u * (1.0 - hi), -1.0 / a);
    uint32_t k = (uint32_t)x;
    return std::min(std::max(k, 1u), hits);
}

void gen_query(const Synth& S, uint64_t q, int mode, uint32_t hits, std::string& out, uint64_t& n_rows) {
    const Tree& t = S.tree;
    Rng r(hash2(S.seed, q));
    const uint32_t H = hits_for(S.seed, q, mode, hits);
    const uint64_t leaf0 = r.next() % S.n_taxa;
    // class of the top group: 0 single, 1 all rows same taxon, 2..7 siblings under the same g,f,o,c,p / anywhere
    const uint32_t cls = r.below(8);
    uint32_t G = 1;
    if (cls != 0) {
        G = 2;
        while (G < 8 && (r.next() & 1)) G++;
    }
    G = std::min(G, H);
    const uint32_t len = 200 + r.below(1301);
    const uint32_t pid0 = 80000 + r.below(20001);  // milli-percent
    const bool low = r.below(100) < 5;             // low-score query with fractional bit scores
    const uint32_t top_bits = low ? 60 + r.below(39) : (uint32_t)std::llround(1.85 * len * (pid0 / 100000.0));
    char line[256];
    char qid[16];
    {
        char* p = qid;
        *p++ = 'q';
        char d[9];
        uint64_t v = q;
        for (int i = 8; i >= 0; i--) {
            d[i] = (char)('0' + v % 10);
            v /= 10;
        }
        memcpy(p, d, 9);
        p += 9;
        *p = 0;
    }
    const size_t qlen = q > 999999999ull ? (size_t)snprintf(qid, sizeof qid, "q%llu", (unsigned long long)q) : 10;
    for (uint32_t h = 0; h < H; h++) {
        uint64_t leaf;
        uint32_t pid, bits10;  // bits in tenths
        if (h < G) {
            if (h == 0 || cls == 1)
                leaf = leaf0;
            else {
                // share the ancestor at level (7 - cls): cls 2 -> same genus (level 4 node), ... cls 7 -> any leaf
                int lvl = 6 - (int)(cls - 1);  // 5..0 ; leaves sharing node at level lvl-1
                uint64_t span = lvl >= 0 ? t.below[lvl] : t.n_leaves;
                uint64_t base = leaf0 / span * span;
                leaf = base + r.next() % span;
                if (leaf >= S.n_taxa) leaf = leaf0;
            }
            pid = pid0 - std::min(pid0 - 80000u, r.below(3) * 37u);
            bits10 = top_bits * 10 + (low ? r.below(10) : 0);
        } else {
            // lower-scoring hits drift away in the tree
            uint64_t span = t.below[std::max(0, 5 - (int)(h * 6 / std::max(H, 1u)))];
            uint64_t base = leaf0 / span * span;
            leaf = base + r.next() % span;
            if (leaf >= S.n_taxa) leaf = leaf0;
            uint32_t drop = 1 + r.below(std::max(1u, top_bits / 3));
            uint32_t b = top_bits > drop ? top_bits - drop : 1;
            if (b >= top_bits) b = top_bits - 1;
            if (b == 0) b = 1;
            bits10 = b * 10 + ((b < 100 && r.below(10) == 0) ? r.below(10) : 0);
            if (bits10 / 10 >= top_bits) bits10 = (top_bits - 1) * 10;
            pid = 80000 + r.below(std::max(1u, pid0 - 80000u + 1));
        }
        char* p = line;
        memcpy(p, qid, qlen);
        p += qlen;
        *p++ = '\t';
        memcpy(p, "NR_", 3);
        p += 3;
        {
            uint64_t an = (leaf * 4 + r.below(4)) % 1000000ull;
            char d[6];
            for (int i = 5; i >= 0; i--) {
                d[i] = (char)('0' + an % 10);
                an /= 10;
            }
            memcpy(p, d, 6);
            p += 6;
        }
        *p++ = '.';
        *p++ = '1';
        *p++ = '\t';
        p = put_uint(p, (uint64_t)S.taxid[leaf]);
        *p++ = '\t';
        p = put_fixed3(p, pid);
        *p++ = '\t';
        const uint32_t alen = h < G ? len : 200 + r.below(1301);
        p = put_uint(p, alen);
        *p++ = '\t';
        p = put_uint(p, (uint64_t)((100000 - pid) / 1000.0 * alen / 100.0));  // mismatches
        *p++ = '\t';
        p = put_uint(p, r.below(4));  // gap openings
        *p++ = '\t';
        *p++ = '1';
        *p++ = '\t';
        p = put_uint(p, alen);
        *p++ = '\t';
        const uint32_t ss = 1 + r.below(300);
        p = put_uint(p, ss);
        *p++ = '\t';
        p = put_uint(p, ss + alen - 1);
        *p++ = '\t';
        {
            static const char* ev[6] = {"0.0", "1e-180", "2.51e-117", "4e-50", "3.4e-08", "0.001"};
            const char* e = ev[r.below(6)];
            size_t n = strlen(e);
            memcpy(p, e, n);
            p += n;
        }
        *p++ = '\t';
        p = put_uint(p, bits10 / 10);
        if (bits10 % 10) {
            *p++ = '.';
            *p++ = (char)('0' + bits10 % 10);
        }
        *p++ = '\n';
        out.append(line, p - line);
    }
    n_rows += H;
}

}  // namespace

extern "C" {

void* blu_synth_create(uint64_t n_taxa, uint64_t seed) {
    if (n_taxa == 0) return nullptr;
    auto* S = new Synth();
    S->n_taxa = n_taxa;
    S->seed = seed;
    S->tree = make_tree(n_taxa);
    build_lineages(*S);
    return S;
}

void blu_synth_destroy(void* h) { delete (Synth*)h; }

uint64_t blu_synth_num_taxa(void* h) { return ((Synth*)h)->n_taxa; }

// Copies the lineage table out: taxids[n]; concatenated strings with offsets off[n+1].  Pass blob == NULL to query sizes.
uint64_t blu_synth_lineages(void* h, int numeric, int64_t* taxids, uint64_t* off, char* blob) {
    Synth& S = *(Synth*)h;
    const auto& v = numeric ? S.num : S.text;
    uint64_t tot = 0;
    for (uint64_t i = 0; i < S.n_taxa; i++) {
        if (off) off[i] = tot;
        if (blob) memcpy(blob + tot, v[i].data(), v[i].size());
        if (taxids) taxids[i] = S.taxid[i];
        tot += v[i].size();
    }
    if (off) off[S.n_taxa] = tot;
    return tot;
}

// Writes the map in the `.blutils.json` schema.
int blu_synth_write_json(void* h, const char* path) {
    Synth& S = *(Synth*)h;
    FILE* f = fopen(path, "wb");
    if (!f) return 1;
    fprintf(f, "{\"blutilsVersion\":\"8.3.1\",\"ignoreTaxids\":null,\"replaceRank\":null,\"dropNonLinnaeanTaxonomies\":false,"
               "\"sourceDatabase\":\"/synthetic/blast_db/synthetic_16S\",\"taxonomies\":[");
    for (uint64_t i = 0; i < S.n_taxa; i++) {
        const std::string& tx = S.text[i];
        size_t k = tx.rfind(';');
        std::string last = tx.substr(k == std::string::npos ? 0 : k + 1);
        std::string rank = last.substr(0, last.find("__"));
        fprintf(f, "%s{\"taxid\":%lld,\"rank\":\"%s\",\"numericLineage\":\"%s\",\"textLineage\":\"%s\",\"accessions\":[{\"accession\":\"NR_%06llu.1\","
                   "\"oid\":\"%llu\"}]}",
                i ? "," : "", (long long)S.taxid[i], rank.c_str(), S.num[i].c_str(), tx.c_str(), (unsigned long long)((i * 4) % 1000000ull),
                (unsigned long long)i);
    }
    fprintf(f, "]}");
    return fclose(f) ? 1 : 0;
}

// Generates queries [q_begin, q_begin + n_queries) into dst (capacity cap).  mode 0: `hits` rows per query;
// mode 1: Zipf(1.1) on [1, hits].  Returns 0 ok, 2 if cap is too small (len then holds the needed size).
int blu_synth_hits(void* h, uint64_t q_begin, uint64_t n_queries, int mode, uint32_t hits, int threads, char* dst, uint64_t cap, uint64_t* len,
                   uint64_t* n_rows) {
    Synth& S = *(Synth*)h;
    threads = std::max(1, threads);
    if ((uint64_t)threads > n_queries) threads = (int)std::max<uint64_t>(1, n_queries);
    std::vector<std::string> parts(threads);
    std::vector<uint64_t> rows(threads, 0);
    std::vector<std::thread> th;
    auto work = [&](int t) {
        uint64_t a = n_queries * t / threads, b = n_queries * (t + 1) / threads;
        std::string& o = parts[t];
        o.reserve((size_t)((b - a) * (mode == 0 ? hits : 64) * 72 + 1024));
        for (uint64_t q = a; q < b; q++) gen_query(S, q_begin + q, mode, hits, o, rows[t]);
    };
    for (int t = 1; t < threads; t++) th.emplace_back(work, t);
    work(0);
    for (auto& x : th) x.join();
    uint64_t tot = 0, nr = 0;
    std::vector<uint64_t> offs(threads);
    for (int t = 0; t < threads; t++) {
        offs[t] = tot;
        tot += parts[t].size();
        nr += rows[t];
    }
    *len = tot;
    if (n_rows) *n_rows = nr;
    if (tot > cap || !dst) return 2;
    th.clear();
    auto cp = [&](int t) { memcpy(dst + offs[t], parts[t].data(), parts[t].size()); };
    for (int t = 1; t < threads; t++) th.emplace_back(cp, t);
    cp(0);
    for (auto& x : th) x.join();
    return 0;
}
}
