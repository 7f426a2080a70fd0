// TEST INFRASTRUCTURE ONLY.
// CPU restatement ("port") of blutils 8.3.1's consensus-identity hot path, used as (1) the parity
// oracle for the CUDA path and (2) the `cpu_baseline` / `--impl reference` arm of bench.py.
// Nothing under blutils_b200/ may include, link or call this file.
//
// PARITY STATUS: parity unpinned for parse/join/group (the reference has no tests and cannot be
// built here -- no Rust toolchain; polars 0.37, serde_json 1.0, slugify 0.1 are not vendored).
// The cutoff interpolation / rank selection / bean folding arithmetic is pinned by
// tests/test_golden_derived.py against the reference's own golden output, and this file is
// cross-checked against the independent Python mirror oracle/pyoracle.py.
//
// All `file:line` citations are relative to /root/reference.
//
// Accepted input grammar (anything else -> data error == "reference aborts or is unpinned"):
//   int   := -?[0-9]{1,18}
//   float := -?([0-9]+(\.[0-9]*)?|\.[0-9]+)([eE][+-]?[0-9]+)?   (value: strtod, correctly rounded)
//   rows '\n'-terminated (last newline optional), 13 '\t'-separated fields, empty lines skipped,
//   no '"' and no '\r' bytes, non-empty qseqid/saccver.
//
// Differences from the real reference that make this port FASTER than it (so GPU/CPU ratios are
// conservative): lineages are parsed and interpolated once per taxon instead of once per row per
// query (blast_result.rs:38 / linnaean_ranks.rs:154), no 100-byte lineage string is copied per
// row (mod.rs:72-76), the row walk (mod.rs:147-209) is parallel.

#include <algorithm>
#include <atomic>
#include <charconv>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <string_view>
#include <thread>
#include <unordered_map>
#include <vector>

namespace {

struct DataError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

// ---------------------------------------------------------------------------------------------
// LinnaeanRank (core/src/domain/dtos/linnaean_ranks.rs:14-107)
// ---------------------------------------------------------------------------------------------
static const char* kFull[9] = {"undefined", "domain", "kingdom", "phylum", "class", "order", "family", "genus", "species"};

struct Rank {
    int def = -1;      // 0..8 index into kFull, or -1 for Other
    std::string slug;  // Other(slug)
    bool operator==(const Rank& o) const { return def == o.def && slug == o.slug; }
    bool operator!=(const Rank& o) const { return !(*this == o); }
    std::string display() const { return def >= 0 ? std::string(1, kFull[def][0]) : slug; }  // :74-89
    std::string full() const { return def >= 0 ? std::string(kFull[def]) : slug; }           // :92-106 / serde
};

static std::string slugify(const std::string& in) {  // slugify 0.1.0, sep '-', ASCII only
    for (unsigned char c : in)
        if (c > 127) throw DataError("non-ASCII rank name (unidecode not restated)");
    std::string s;
    for (char c : in) s.push_back((char)tolower((unsigned char)c));
    auto ws = [](char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == '\v' || c == '\f'; };
    size_t a = 0, b = s.size();
    while (a < b && ws(s[a])) a++;
    while (b > a && ws(s[b - 1])) b--;
    while (a < b && s[a] == '-') a++;
    while (b > a && s[b - 1] == '-') b--;
    std::string out;
    bool is_sep = true;
    for (size_t i = a; i < b; i++) {
        char c = s[i];
        if ((c >= 'a' && c <= 'z') || (c >= '0' && c <= '9')) {
            is_sep = false;
            out.push_back(c);
        } else if (!is_sep) {
            is_sep = true;
            out.push_back('-');
        }
    }
    if (out.empty()) throw DataError("empty rank slug");
    if (out.back() == '-') out.pop_back();
    return out;
}

static Rank rank_from_str(const std::string& in) {  // :55-71
    std::string t;
    for (char c : in) t.push_back((char)tolower((unsigned char)c));
    auto ws = [](char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == '\v' || c == '\f'; };
    size_t a = 0, b = t.size();
    while (a < b && ws(t[a])) a++;
    while (b > a && ws(t[b - 1])) b--;
    t = t.substr(a, b - a);
    Rank r;
    for (int i = 0; i < 9; i++)
        if (t == kFull[i] || (t.size() == 1 && t[0] == kFull[i][0])) {
            r.def = i;
            return r;
        }
    r.slug = slugify(t);
    return r;
}

// ---------------------------------------------------------------------------------------------
// Cutoffs (core/src/domain/dtos/taxon.rs:105-185) + interpolation (linnaean_ranks.rs:220-383)
// ---------------------------------------------------------------------------------------------
struct BB {
    int def;
    double cut;
};

static std::vector<BB> backbone_for(int taxon /*0 fungi 1 bacteria 2 eukaryotes 3 custom*/, const int* custom8 /* -1 = absent */) {
    std::vector<BB> v;
    if (taxon == 3) {
        if (!custom8) throw DataError("Custom taxon values are required");
        if (custom8[0] == -1 || custom8[7] == -1) throw DataError("custom cutoffs: domain and species are mandatory");
        for (int i = 0; i < 8; i++) v.push_back({i + 1, custom8[i] == -1 ? 0.0 : (double)custom8[i]});  // domain-first :123-139
        return v;
    }
    const double b[7] = {99, 97, 92, 85, 80, 75, 60}, f[7] = {97, 95, 90, 85, 80, 75, 60};
    const int order[7] = {8, 7, 6, 5, 4, 3, 1};  // species..phylum, domain (no kingdom)
    for (int i = 0; i < 7; i++) v.push_back({order[i], taxon == 1 ? b[i] : f[i]});
    return v;
}

static double round3(double v) {  // domain/utils/mod.rs:1-4
    double y = 1000.0;
    return std::round(v * y) / y;
}

struct RLI {  // RankedLinnaeanIdentity
    bool is_default;
    Rank rank;        // DefaultRank(rank, _)
    std::string name; // NonDefaultRank(name, _)
    double cut;
    bool operator==(const RLI& o) const {
        if (is_default != o.is_default) return false;
        return is_default ? (rank == o.rank && cut == o.cut) : (name == o.name && cut == o.cut);
    }
};

static std::vector<RLI> interpolate(const std::vector<Rank>& ranks, const std::vector<BB>& bb) {
    std::vector<RLI> m;
    bool all_def = true;
    for (auto& r : ranks) {
        const BB* hit = nullptr;
        if (r.def >= 0)
            for (auto& b : bb)
                if (b.def == r.def) {
                    hit = &b;
                    break;
                }
        if (hit)
            m.push_back({true, r, "", hit->cut});
        else {
            m.push_back({false, Rank(), r.display(), 0.0});
            all_def = false;
        }
    }
    if (all_def) return m;
    std::vector<RLI> out = m;
    const size_t n_ = m.size();
    for (size_t n = 0; n < n_; n++) {
        if (m[n].is_default) continue;
        const RLI* prev = &m[0];
        for (size_t i = n; i-- > 0;)
            if (m[i].is_default) {
                prev = &m[i];
                break;
            }
        size_t p = 0;
        for (size_t i = 0; i < n_; i++)
            if (m[i] == *prev) {
                p = i;
                break;
            }
        const RLI* next = &m[n_ - 1];
        for (size_t i = n; i < n_; i++)
            if (m[i].is_default) {
                next = &m[i];
                break;
            }
        size_t q = n_ - 1;
        for (size_t i = 0; i < n_; i++)
            if (m[i] == *next) {
                q = i;
                break;
            }
        size_t wlen = std::min(q + 1, n_ - p);  // skip_while(!=prev).take(q+1)
        const RLI& w0 = m[p];
        const RLI& wl = m[p + wlen - 1];
        double first = w0.is_default ? w0.cut : bb[0].cut;
        double last = wl.is_default ? wl.cut : 100.0;
        double weight = last - first;
        double size = (double)(wlen - 1);
        double t = (double)(n - p);
        out[n].cut = round3(first + (t * (weight / size)));
    }
    return out;
}

// ---------------------------------------------------------------------------------------------
// Lineage (blast_result.rs:38-120), parsed once per taxon
// ---------------------------------------------------------------------------------------------
struct Bean {
    Rank rank;
    std::string ident;
    std::string level_key;  // rank.to_string() + identifier (find_multi_taxa_consensus.rs:153-157)
    std::string bean_key;   // "{rank}__{identifier}" (consensus_result.rs:70-73)
    uint32_t key_ord = 0;   // order of first appearance of bean_key in the taxonomy map (the tie-break of fold_beans)
};

struct Lineage {
    std::vector<Bean> beans;
    std::vector<RLI> interp;
    std::string str;  // Taxonomy::taxonomy_beans_to_string
    bool ok = false;
    std::string err;
};

static void split(const std::string& s, const std::string& sep, std::vector<std::string>& out) {
    out.clear();
    size_t pos = 0;
    while (true) {
        size_t k = s.find(sep, pos);
        if (k == std::string::npos) {
            out.push_back(s.substr(pos));
            return;
        }
        out.push_back(s.substr(pos, k - pos));
        pos = k + sep.size();
    }
}

static void parse_lineage(const std::string& s, const std::vector<BB>& bb, Lineage& L) {
    try {
        std::vector<std::string> parts, pieces;
        split(s, ";", parts);
        std::vector<Rank> ranks;
        for (auto& p : parts) {
            split(p, "__", pieces);
            if (pieces.size() != 2) throw DataError("Unexpected error on parse taxonomy");
            Bean b;
            b.rank = rank_from_str(pieces[0]);
            b.ident = pieces[1];
            b.level_key = b.rank.display() + b.ident;
            b.bean_key = b.rank.display() + "__" + b.ident;
            ranks.push_back(b.rank);
            L.beans.push_back(std::move(b));
        }
        L.interp = interpolate(ranks, bb);
        for (size_t i = 0; i < L.beans.size(); i++) {
            if (i) L.str += ";";
            L.str += L.beans[i].bean_key;
        }
        L.ok = true;
    } catch (const DataError& e) {
        L.ok = false;
        L.err = e.what();
    }
}

// ---------------------------------------------------------------------------------------------
// Number parsing
// ---------------------------------------------------------------------------------------------
static bool parse_int(const char* p, const char* e, int64_t& v) {
    bool neg = false;
    if (p < e && *p == '-') {
        neg = true;
        p++;
    }
    if (p == e || e - p > 18) return false;
    int64_t x = 0;
    for (; p < e; p++) {
        if (*p < '0' || *p > '9') return false;
        x = x * 10 + (*p - '0');
    }
    v = neg ? -x : x;
    return true;
}

// returns 0 ok, 1 bad grammar, 2 valid grammar but outside the supported exact path
static int parse_float(const char* p, const char* e, bool need_value, double& v) {
    const char* s = p;
    if (p < e && *p == '-') p++;
    int nint = 0, nfrac = 0;
    int sig = 0;
    bool lead = true;
    auto feed = [&](char c) {
        if (lead && c == '0') return;
        lead = false;
        sig++;
    };
    while (p < e && *p >= '0' && *p <= '9') {
        feed(*p);
        nint++;
        p++;
    }
    if (p < e && *p == '.') {
        p++;
        while (p < e && *p >= '0' && *p <= '9') {
            feed(*p);
            nfrac++;
            p++;
        }
        if (nint == 0 && nfrac == 0) return 1;
    } else if (nint == 0)
        return 1;
    long ex = 0;
    if (p < e && (*p == 'e' || *p == 'E')) {
        p++;
        bool eneg = false;
        if (p < e && (*p == '+' || *p == '-')) {
            eneg = *p == '-';
            p++;
        }
        int nd = 0;
        while (p < e && *p >= '0' && *p <= '9') {
            ex = ex * 10 + (*p - '0');
            nd++;
            p++;
        }
        if (nd == 0) return 1;
        if (eneg) ex = -ex;
    }
    if (p != e) return 1;
    if (!need_value) return 0;
    (void)sig;  // any number of digits: strtod is correctly rounded (the CUDA path reports BLU_ERR_UNSUPPORTED beyond its exact range)
    std::string tmp(s, e);
    v = strtod(tmp.c_str(), nullptr);  // correctly rounded; independent of the GPU's Clinger arithmetic
    return 0;
}

// ---------------------------------------------------------------------------------------------
// Rows
// ---------------------------------------------------------------------------------------------
struct Row {
    std::string_view query, acc;
    int64_t taxid, alnlen, bits;
    double pident;
};

struct Handle {
    std::vector<BB> bb;
    int strategy;  // 0 cautious, 1 relaxed
    std::unordered_map<int64_t, uint32_t> by_taxid;
    std::vector<Lineage> lineages;
};

static void parse_row(const char* p, const char* e, Row& r) {
    const char* f[14];
    int nf = 0;
    f[nf++] = p;
    for (const char* c = p; c < e; c++)
        if (*c == '\t') {
            if (nf >= 13) throw DataError("row has more than 13 fields");
            f[nf++] = c + 1;
        }
    if (nf != 13) throw DataError("row does not have 13 fields");
    f[13] = e + 1;
    auto fb = [&](int i) { return f[i]; };
    auto fe = [&](int i) { return f[i + 1] - 1; };
    if (fe(0) == fb(0) || fe(1) == fb(1)) throw DataError("empty query/accession field");
    r.query = std::string_view(fb(0), fe(0) - fb(0));
    r.acc = std::string_view(fb(1), fe(1) - fb(1));
    if (!parse_int(fb(2), fe(2), r.taxid)) throw DataError("bad staxid");
    int rc = parse_float(fb(3), fe(3), true, r.pident);
    if (rc) throw DataError(rc == 1 ? "bad pident" : "unsupported pident number");
    if (!parse_int(fb(4), fe(4), r.alnlen)) throw DataError("bad length");
    int64_t dummy;
    for (int k = 5; k < 11; k++)
        if (!parse_int(fb(k), fe(k), dummy)) throw DataError("bad integer column");
    double d;
    if (parse_float(fb(11), fe(11), false, d)) throw DataError("bad evalue");
    rc = parse_float(fb(12), fe(12), true, d);
    if (rc) throw DataError(rc == 1 ? "bad bitscore" : "unsupported bitscore number");
    if (!(std::fabs(d) < 9223372036854775808.0)) throw DataError("bit score out of i64 range");
    r.bits = (int64_t)d;  // truncation toward zero (mod.rs:162,184)
}

// ---------------------------------------------------------------------------------------------
// JSON output (serde_json compact; write_blutils_output.rs:166-215)
// ---------------------------------------------------------------------------------------------
static void jstr(std::string& o, std::string_view s) {
    o.push_back('"');
    for (unsigned char c : s) {
        switch (c) {
            case '"': o += "\\\""; break;
            case '\\': o += "\\\\"; break;
            case '\b': o += "\\b"; break;
            case '\f': o += "\\f"; break;
            case '\n': o += "\\n"; break;
            case '\r': o += "\\r"; break;
            case '\t': o += "\\t"; break;
            default:
                if (c < 0x20) {
                    char b[8];
                    snprintf(b, sizeof b, "\\u%04x", c);
                    o += b;
                } else
                    o.push_back((char)c);
        }
    }
    o.push_back('"');
}

static void jf64(std::string& o, double v) {  // ryu pretty (d2s) layout over shortest digits
    if (!std::isfinite(v)) {
        o += "null";
        return;
    }
    if (v == 0) {
        o += std::signbit(v) ? "-0.0" : "0.0";
        return;
    }
    char buf[64];
    auto res = std::to_chars(buf, buf + sizeof buf, std::fabs(v), std::chars_format::scientific);
    std::string s(buf, res.ptr);  // d[.ddd]e[+-]XX
    size_t epos = s.find('e');
    int ex = atoi(s.c_str() + epos + 1);
    std::string digits;
    for (size_t i = 0; i < epos; i++)
        if (s[i] != '.') digits.push_back(s[i]);
    while (digits.size() > 1 && digits.back() == '0') digits.pop_back();
    int length = (int)digits.size();
    int kk = ex + 1;       // position of the decimal point
    int k = kk - length;   // exponent of the last digit
    if (std::signbit(v)) o.push_back('-');
    if (0 <= k && kk <= 16) {
        o += digits;
        o.append(k, '0');
        o += ".0";
    } else if (0 < kk && kk <= 16) {
        o.append(digits, 0, kk);
        o.push_back('.');
        o.append(digits, kk, std::string::npos);
    } else if (-5 < kk && kk <= 0) {
        o += "0.";
        o.append(-kk, '0');
        o += digits;
    } else if (length == 1) {
        o += digits;
        o.push_back('e');
        o += std::to_string(kk - 1);
    } else {
        o.push_back(digits[0]);
        o.push_back('.');
        o.append(digits, 1, std::string::npos);
        o.push_back('e');
        o += std::to_string(kk - 1);
    }
}

// ---------------------------------------------------------------------------------------------
// Consensus
// ---------------------------------------------------------------------------------------------
struct FoldBean {  // ConsensusBean (consensus_result.rs:37-45)
    const Bean* bean;
    int occurrences;
    const std::string* taxonomy;
    std::vector<std::string_view> accessions;
};

static void fold_beans(const std::vector<std::pair<const Lineage*, const Row*>>& S, size_t level,
                       std::vector<FoldBean>& out) {  // consensus_result.rs:65-88 + bbci.rs:50-60
    out.clear();
    std::unordered_map<std::string_view, size_t> idx;
    for (auto& [L, r] : S) {
        if (level >= L->beans.size()) break;  // take_while
        const Bean& b = L->beans[level];
        auto it = idx.find(b.bean_key);
        size_t k;
        if (it == idx.end()) {
            k = out.size();
            idx.emplace(b.bean_key, k);
            out.push_back({&b, 0, &L->str, {}});
        } else
            k = it->second;
        auto& fb = out[k];
        if (fb.accessions.empty() || fb.accessions.back() != r->acc) fb.accessions.push_back(r->acc);  // extend + dedup
        fb.occurrences++;
    }
    // Full ties (same occurrences and identifier, different rank) leave the reference in HashMap order: non-deterministic there.
    // Every implementation of this repository breaks them by the order in which the keys first appear in the taxonomy map.
    std::stable_sort(out.begin(), out.end(), [](const FoldBean& a, const FoldBean& b) {
        if (a.occurrences != b.occurrences) return a.occurrences > b.occurrences;
        if (a.bean->ident != b.bean->ident) return a.bean->ident < b.bean->ident;
        return a.bean->key_ord < b.bean->key_ord;
    });
}

static void emit_taxon(std::string& o, const Rank& reached, const Rank* max_allowed, const std::string& ident, double pident,
                       int64_t bits, const std::string& taxonomy, bool mutated, bool single, const std::vector<FoldBean>& beans) {
    o += "{\"reachedRank\":";
    jstr(o, reached.full());
    o += ",\"maxAllowedRank\":";
    if (max_allowed)
        jstr(o, max_allowed->full());
    else
        o += "null";
    o += ",\"identifier\":";
    jstr(o, ident);
    o += ",\"percIdentity\":";
    jf64(o, pident);
    o += ",\"bitScore\":";
    jf64(o, (double)bits);
    o += ",\"taxonomy\":";
    jstr(o, taxonomy);
    o += ",\"mutated\":";
    o += mutated ? "true" : "false";
    o += ",\"singleMatch\":";
    o += single ? "true" : "false";
    o += ",\"consensusBeans\":[";
    for (size_t i = 0; i < beans.size(); i++) {
        if (i) o.push_back(',');
        o += "{\"rank\":";
        jstr(o, beans[i].bean->rank.full());
        o += ",\"identifier\":";
        jstr(o, beans[i].bean->ident);
        o += ",\"occurrences\":" + std::to_string(beans[i].occurrences);
        o += ",\"taxonomy\":";
        jstr(o, *beans[i].taxonomy);
        o += ",\"accessions\":[";
        for (size_t a = 0; a < beans[i].accessions.size(); a++) {
            if (a) o.push_back(',');
            jstr(o, beans[i].accessions[a]);
        }
        o += "]}";
    }
    o += "]}";
}

static void consensus_for_query(const Handle& H, const std::vector<const Row*>& rows, std::string& o) {
    // find_single_query_consensus.rs:28-64: only the max bit-score group is ever evaluated
    int64_t top = rows[0]->bits;
    for (auto r : rows) top = std::max(top, r->bits);
    std::vector<std::pair<const Lineage*, const Row*>> S;
    for (auto r : rows) {
        if (r->bits != top) continue;
        auto it = H.by_taxid.find(r->taxid);
        if (it == H.by_taxid.end()) throw DataError("unmapped taxid in top bit-score group");
        const Lineage& L = H.lineages[it->second];
        if (!L.ok) throw DataError(L.err);
        S.push_back({&L, r});
    }
    std::vector<FoldBean> beans;
    if (S.size() == 1) {  // :74-150
        const Lineage& L = *S[0].first;
        const Row& r = *S[0].second;
        std::string tax;
        int last = -1;
        for (size_t k = 0; k < L.beans.size(); k++)
            if (r.pident >= L.interp[k].cut) {
                if (last >= 0) tax += ";";
                tax += L.beans[k].bean_key;
                last = (int)k;
            }
        if (last < 0) throw DataError("No taxonomy found for result");
        beans.push_back({&L.beans[last], 1, &L.str, {r.acc}});
        emit_taxon(o, L.beans[last].rank, nullptr, L.beans[last].ident, r.pident, r.bits, tax, false, true, beans);
        return;
    }
    // find_multi_taxa_consensus.rs:39-54
    std::stable_sort(S.begin(), S.end(), [](const auto& a, const auto& b) {
        size_t la = a.first->beans.size(), lb = b.first->beans.size();
        if (la != lb) return la < lb;
        if (a.second->pident < b.second->pident) return true;
        if (a.second->pident > b.second->pident) return false;
        if (a.second->alnlen != b.second->alnlen) return a.second->alnlen < b.second->alnlen;
        return a.second->acc < b.second->acc;
    });
    const auto& ref = H.strategy == 0 ? S.front() : S.back();  // :60-63
    const Lineage& R = *ref.first;
    const Row& rr = *ref.second;
    const size_t shortest = S.front().first->beans.size();
    // level walk :137-214
    bool have = false, single = true;
    size_t idx = 0, bean_level = 0;
    double identity = 0;
    for (size_t i = 0; i < R.beans.size(); i++) {
        if (i >= shortest) continue;  // take_while over a length-ascending list is all-or-nothing
        bool agree = true;
        const std::string& k0 = S[0].first->beans[i].level_key;
        for (auto& s : S)
            if (s.first->beans[i].level_key != k0) {
                agree = false;
                break;
            }
        if (!agree) {
            if (i == 0) throw DataError("root-level disagreement (index - 1 underflow)");
            double mx = 0.0;
            for (auto& s : S)
                if (s.second->pident > mx) mx = s.second->pident;
            have = true, single = false, idx = i - 1, bean_level = i, identity = mx;
            break;
        }
        have = true, single = true, idx = i, bean_level = i, identity = rr.pident;
    }
    if (!have) throw DataError("internal: no level evaluated");
    // build_blast_consensus_identity.rs:9-105
    int allowed = -1;
    for (size_t j = 0; j < R.interp.size(); j++)
        if (!(identity > R.interp[j].cut)) {
            allowed = (int)j;
            break;
        }
    Rank allowed_rank;
    bool mutated = false;
    if (allowed >= 0) {
        if (R.interp[allowed].is_default)
            allowed_rank = R.interp[allowed].rank;
        else
            allowed_rank.slug = R.interp[allowed].name;  // LinnaeanRank::Other(name)
        mutated = R.beans[idx].rank != allowed_rank;
    }
    fold_beans(S, bean_level, beans);
    std::vector<size_t> F;
    for (size_t k = 0; k < R.beans.size(); k++)
        if (identity >= R.interp[k].cut) F.push_back(k);
    if (!(single && beans.size() == 1) && F.size() > idx + 1) F.resize(idx + 1);
    std::string tax;
    for (size_t k = 0; k < F.size(); k++) {
        if (k) tax += ";";
        tax += R.beans[F[k]].bean_key;
    }
    const Bean& last = F.empty() ? R.beans[idx] : R.beans[F.back()];
    emit_taxon(o, last.rank, allowed >= 0 ? &allowed_rank : nullptr, last.ident, rr.pident, rr.bits, tax, mutated, false, beans);
}

struct Output {
    std::string jsonl;
    uint64_t n_queries = 0, n_rows = 0;
};

static void run(const Handle& H, const char* text, size_t n, int threads, const std::vector<std::string>& headers, Output& out) {
    if (n == 0) throw DataError("empty blast output");
    if (memchr(text, '"', n) || memchr(text, '\r', n)) throw DataError("quote or CR byte in blast output");
    threads = std::max(1, threads);
    // 1. chunk at newlines, count rows, parse rows in parallel
    std::vector<size_t> cuts(threads + 1);
    cuts[0] = 0;
    cuts[threads] = n;
    for (int t = 1; t < threads; t++) {
        size_t c = std::max(cuts[t - 1], n * t / threads);
        const char* nl = c < n ? (const char*)memchr(text + c, '\n', n - c) : nullptr;
        cuts[t] = nl ? (size_t)(nl - text) + 1 : n;
    }
    std::vector<std::vector<Row>> parts(threads);
    std::vector<std::string> errs(threads);
    auto par = [&](auto fn) {
        std::vector<std::thread> th;
        for (int t = 1; t < threads; t++) th.emplace_back(fn, t);
        fn(0);
        for (auto& x : th) x.join();
    };
    par([&](int t) {
        try {
            const char* p = text + cuts[t];
            const char* e = text + cuts[t + 1];
            auto& v = parts[t];
            v.reserve((e - p) / 60 + 16);
            while (p < e) {
                const char* nl = (const char*)memchr(p, '\n', e - p);
                const char* le = nl ? nl : e;
                if (le > p) {
                    v.emplace_back();
                    parse_row(p, le, v.back());
                }
                p = le + 1;
            }
        } catch (const DataError& ex) {
            errs[t] = ex.what();
        }
    });
    for (auto& e : errs)
        if (!e.empty()) throw DataError(e);
    std::vector<size_t> base(threads + 1, 0);
    for (int t = 0; t < threads; t++) base[t + 1] = base[t] + parts[t].size();
    const size_t R = base[threads];
    if (R == 0) throw DataError("no rows");
    std::vector<Row> rows(R);
    par([&](int t) { std::copy(parts[t].begin(), parts[t].end(), rows.begin() + base[t]); });
    parts.clear();
    // 2. group by query (fold_results_by_query mod.rs:134-221): runs of equal qseqid, merged by a map of run heads
    std::vector<uint32_t> run_start;
    for (size_t i = 0; i < R; i++)
        if (i == 0 || rows[i].query != rows[i - 1].query) run_start.push_back((uint32_t)i);
    run_start.push_back((uint32_t)R);
    std::unordered_map<std::string_view, uint32_t> qidx;
    qidx.reserve(run_start.size() * 2);
    std::vector<std::vector<uint32_t>> qruns;
    std::vector<std::string_view> qname;
    for (size_t k = 0; k + 1 < run_start.size(); k++) {
        auto q = rows[run_start[k]].query;
        auto it = qidx.find(q);
        if (it == qidx.end()) {
            qidx.emplace(q, (uint32_t)qruns.size());
            qruns.push_back({(uint32_t)k});
            qname.push_back(q);
        } else
            qruns[it->second].push_back((uint32_t)k);
    }
    const size_t Q = qruns.size();
    // 3. per-query consensus in parallel (mod.rs:104-128)
    std::vector<std::string> js(Q);
    std::atomic<size_t> next{0};
    par([&](int t) {
        try {
            std::vector<const Row*> qr;
            while (true) {
                size_t b = next.fetch_add(256);
                if (b >= Q) break;
                for (size_t qi = b; qi < std::min(Q, b + 256); qi++) {
                    qr.clear();
                    for (uint32_t k : qruns[qi])
                        for (uint32_t i = run_start[k]; i < run_start[k + 1]; i++) qr.push_back(&rows[i]);
                    std::string& o = js[qi];
                    o += "{\"query\":";
                    jstr(o, qname[qi]);
                    o += ",\"taxon\":";
                    consensus_for_query(H, qr, o);
                    o += "}\n";
                }
            }
        } catch (const DataError& ex) {
            errs[t] = ex.what();
        }
    });
    for (auto& e : errs)
        if (!e.empty()) throw DataError(e);
    // 4. hit-less headers (mod.rs:84-102) + writer's sort by query (write_blutils_output.rs:111)
    std::vector<std::pair<std::string_view, const std::string*>> order;
    order.reserve(Q + headers.size());
    for (size_t i = 0; i < Q; i++) order.push_back({qname[i], &js[i]});
    std::vector<std::string> extra;
    extra.reserve(headers.size());
    for (auto& h : headers)
        if (!qidx.count(h)) {
            std::string o = "{\"query\":";
            jstr(o, h);
            o += ",\"taxon\":null}\n";
            extra.push_back(std::move(o));
            order.push_back({std::string_view(h), &extra.back()});
        }
    std::stable_sort(order.begin(), order.end(), [](auto& a, auto& b) { return a.first < b.first; });
    size_t tot = 0;
    for (auto& x : order) tot += x.second->size();
    out.jsonl.reserve(tot);
    for (auto& x : order) out.jsonl += *x.second;
    out.n_queries = order.size();
    out.n_rows = R;
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// C ABI for ctypes (tests/ and bench.py only)
// ---------------------------------------------------------------------------------------------
extern "C" {

// lineages: `n` strings concatenated in `blob`, string i = blob[off[i] .. off[i+1])
void* blu_oracle_create(const int64_t* taxids, const uint64_t* off, const char* blob, uint64_t n, int taxon, const int* custom8,
                        int strategy, int threads, char* err, int errlen) {
    try {
        auto H = std::make_unique<Handle>();
        H->bb = backbone_for(taxon, custom8);
        H->strategy = strategy;
        H->lineages.resize(n);
        H->by_taxid.reserve(n * 2);
        for (uint64_t i = 0; i < n; i++) {
            if (!H->by_taxid.emplace(taxids[i], (uint32_t)i).second) throw DataError("duplicate taxid in taxonomy");
        }
        threads = std::max(1, threads);
        std::vector<std::thread> th;
        auto work = [&](int t) {
            for (uint64_t i = t; i < n; i += threads) parse_lineage(std::string(blob + off[i], off[i + 1] - off[i]), H->bb, H->lineages[i]);
        };
        for (int t = 1; t < threads; t++) th.emplace_back(work, t);
        work(0);
        for (auto& x : th) x.join();
        {
            std::unordered_map<std::string_view, uint32_t> ord;
            for (auto& L : H->lineages)
                for (auto& bean : L.beans) bean.key_ord = ord.emplace(bean.bean_key, (uint32_t)ord.size()).first->second;
        }
        return H.release();
    } catch (const std::exception& e) {
        snprintf(err, errlen, "%s", e.what());
        return nullptr;
    }
}

void blu_oracle_destroy(void* h) { delete (Handle*)h; }

// returns 0 ok, 2 data error (reference aborts / unpinned), 1 other
int blu_oracle_run(void* h, const char* text, uint64_t n, int threads, const char* headers_nl, uint64_t headers_len, char** out,
                   uint64_t* out_len, uint64_t* n_queries, uint64_t* n_rows, char* err, int errlen) {
    try {
        std::vector<std::string> headers;
        if (headers_nl) {
            const char* p = headers_nl;
            const char* e = headers_nl + headers_len;
            while (p < e) {
                const char* nl = (const char*)memchr(p, '\n', e - p);
                const char* le = nl ? nl : e;
                headers.emplace_back(p, le - p);
                p = le + 1;
            }
        }
        Output o;
        run(*(Handle*)h, text, n, threads, headers, o);
        *out = (char*)malloc(o.jsonl.size() + 1);
        memcpy(*out, o.jsonl.data(), o.jsonl.size());
        (*out)[o.jsonl.size()] = 0;
        *out_len = o.jsonl.size();
        *n_queries = o.n_queries;
        *n_rows = o.n_rows;
        return 0;
    } catch (const DataError& e) {
        snprintf(err, errlen, "%s", e.what());
        return 2;
    } catch (const std::exception& e) {
        snprintf(err, errlen, "%s", e.what());
        return 1;
    }
}

void blu_oracle_free(char* p) { free(p); }

// order-independent checksum of a JSONL buffer: sum over lines of FNV-1a-64(line without '\n')
// (same definition as blu_result_checksum in include/blu_consensus.h)
uint64_t blu_oracle_checksum_jsonl(const char* js, uint64_t n) {
    uint64_t total = 0;
    const char* p = js;
    const char* e = js + n;
    while (p < e) {
        const char* nl = (const char*)memchr(p, '\n', e - p);
        const char* le = nl ? nl : e;
        uint64_t h = 0xcbf29ce484222325ull;
        for (const char* c = p; c < le; c++) {
            h ^= (unsigned char)*c;
            h *= 0x100000001b3ull;
        }
        total += h;
        p = le + 1;
    }
    return total;
}

// interpolation only, for known-answer tests: ranks = '\n'-separated rank strings
int blu_oracle_interpolate(const char* ranks_nl, int taxon, const int* custom8, double* out, int cap) {
    try {
        auto bb = backbone_for(taxon, custom8);
        std::vector<Rank> ranks;
        std::vector<std::string> parts;
        split(ranks_nl, "\n", parts);
        for (auto& p : parts) ranks.push_back(rank_from_str(p));
        auto v = interpolate(ranks, bb);
        if ((int)v.size() > cap) return -1;
        for (size_t i = 0; i < v.size(); i++) out[i] = v[i].cut;
        return (int)v.size();
    } catch (...) {
        return -2;
    }
}
}
