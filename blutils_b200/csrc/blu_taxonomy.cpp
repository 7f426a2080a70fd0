// blu_taxonomy.cpp -- see blu_taxonomy.h.  Host-only; runs once per context.
#include "blu_taxonomy.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <map>
#include <numeric>

#include <sys/stat.h>
#include <unistd.h>

#include "blu_json.h"

namespace blu {

static const char* kRankFull[9] = {"undefined", "domain", "kingdom", "phylum", "class", "order", "family", "genus", "species"};

static bool is_rust_ws(unsigned char c) { return c == ' ' || (c >= 0x09 && c <= 0x0D); }

std::string slugify_ascii(std::string_view in) {
    // slugify 0.1.0: unidecode -> lowercase -> trim -> trim sep -> ' ' => sep -> keep [a-z0-9], collapse the rest
    std::string s;
    s.reserve(in.size());
    for (unsigned char c : in) {
        if (c > 127) throw DataErr("non-ASCII rank name is not supported (slugify/unidecode)");
        s.push_back((char)((c >= 'A' && c <= 'Z') ? c + 32 : c));
    }
    size_t a = 0, b = s.size();
    while (a < b && is_rust_ws((unsigned char)s[a])) a++;
    while (b > a && is_rust_ws((unsigned char)s[b - 1])) b--;
    while (a < b && s[a] == '-') a++;
    while (b > a && s[b - 1] == '-') b--;
    std::string out;
    bool sep = true;
    for (size_t i = a; i < b; i++) {
        char c = s[i];
        if ((c >= 'a' && c <= 'z') || (c >= '0' && c <= '9')) {
            out.push_back(c);
            sep = false;
        } else if (!sep) {
            out.push_back('-');
            sep = true;
        }
    }
    if (out.empty()) throw DataErr("rank name slugifies to an empty string (the reference panics)");
    if (out.back() == '-') out.pop_back();
    return out;
}

RankInfo rank_from_str(std::string_view in) {
    std::string t;
    t.reserve(in.size());
    for (unsigned char c : in) t.push_back((char)((c >= 'A' && c <= 'Z') ? c + 32 : c));
    size_t a = 0, b = t.size();
    while (a < b && is_rust_ws((unsigned char)t[a])) a++;
    while (b > a && is_rust_ws((unsigned char)t[b - 1])) b--;
    std::string_view v(t.data() + a, b - a);
    RankInfo r;
    for (int i = 0; i < 9; i++) {
        if (v == kRankFull[i] || (v.size() == 1 && v[0] == kRankFull[i][0])) {
            r.def = i;
            r.display = std::string(1, kRankFull[i][0]);
            r.full = kRankFull[i];
            return r;
        }
    }
    r.def = -1;
    r.slug = slugify_ascii(v);
    r.display = r.slug;
    r.full = r.slug;
    return r;
}

std::vector<BackboneEntry> make_backbone(const Cutoffs& c) {
    std::vector<BackboneEntry> bb;
    if (c.taxon == BLU_TAXON_CUSTOM) {
        if (!c.has_custom) throw DataErr("Custom taxon values are required");
        if (c.custom[0] == BLU_CUTOFF_ABSENT || c.custom[7] == BLU_CUTOFF_ABSENT)
            throw DataErr("custom cutoffs: `domain` and `species` are mandatory");
        for (int i = 0; i < 8; i++) bb.push_back({i + 1, c.custom[i] == BLU_CUTOFF_ABSENT ? 0.0 : (double)c.custom[i]});
        return bb;
    }
    static const double bact[7] = {99, 97, 92, 85, 80, 75, 60};
    static const double fung[7] = {97, 95, 90, 85, 80, 75, 60};
    static const int order[7] = {8, 7, 6, 5, 4, 3, 1};
    const double* v = c.taxon == BLU_TAXON_BACTERIA ? bact : fung;  // Fungi == Eukaryotes tables
    for (int i = 0; i < 7; i++) bb.push_back({order[i], v[i]});
    return bb;
}

std::vector<double> interpolate_cutoffs(const std::vector<const RankInfo*>& ranks, const std::vector<BackboneEntry>& bb) {
    const int n = (int)ranks.size();
    std::vector<int> def(n, -1);  // backbone index when the rank is a default rank present in the backbone
    std::vector<double> cut(n, 0.0);
    bool all_default = true;
    for (int j = 0; j < n; j++) {
        if (ranks[j]->def >= 0)
            for (size_t b = 0; b < bb.size(); b++)
                if (bb[b].def == ranks[j]->def) {
                    def[j] = (int)b;
                    cut[j] = bb[b].cut;
                    break;
                }
        if (def[j] < 0) all_default = false;
    }
    if (all_default) return cut;
    // first index holding an element equal to element j (enum equality of RankedLinnaeanIdentity)
    auto first_equal = [&](int j) {
        for (int i = 0; i < n; i++) {
            if (def[j] >= 0 ? def[i] == def[j] : (def[i] < 0 && ranks[i]->display == ranks[j]->display)) return i;
        }
        return j;
    };
    std::vector<double> out = cut;
    for (int x = 0; x < n; x++) {
        if (def[x] >= 0) continue;
        int prev = 0;
        for (int i = x - 1; i >= 0; i--)
            if (def[i] >= 0) {
                prev = i;
                break;
            }
        int next = n - 1;
        for (int i = x; i < n; i++)
            if (def[i] >= 0) {
                next = i;
                break;
            }
        const int p = first_equal(prev);
        const int q = first_equal(next);
        const int wlen = std::min(q + 1, n - p);  // window = m[p..][..q+1]
        const int w_last = p + wlen - 1;
        const double first = def[p] >= 0 ? cut[p] : bb[0].cut;
        const double last = def[w_last] >= 0 ? cut[w_last] : 100.0;
        const double weight = last - first;
        const double size = (double)(wlen - 1);
        const double t = (double)(x - p);
        const double v = first + (t * (weight / size));
        out[x] = std::round(v * 1000.0) / 1000.0;  // domain/utils/mod.rs:1-4
    }
    return out;
}

void HostTaxonomy::append_bean(std::string& o, uint32_t pos) const {
    o += ranks[pos_rank[pos]].display;
    o += "__";
    o += idents[pos_ident[pos]];
}

void HostTaxonomy::append_lineage(std::string& o, uint32_t lin) const {
    for (uint32_t p = lin_off[lin]; p < lin_off[lin + 1]; p++) {
        if (p != lin_off[lin]) o.push_back(';');
        append_bean(o, p);
    }
}

size_t HostTaxonomy::device_bytes() const {
    return lin_off.size() * 4 + lin_ok.size() + lvl_key.size() * (4 + 4 + 4 + 8 + 2 + 2) + slots.size() * sizeof(HashSlot);
}

namespace {

// string -> dense id in first-appearance order.  Open addressing over one byte arena: no allocation per key and one
// probable cache miss per lookup (the std::unordered_map<std::string, ..> this replaces spent 15 s on 2 M lineages).
class StrInterner {
    std::vector<uint64_t> slots_;  // (hash & ~0xFFFFFFFF) | (id + 1); 0 = empty
    std::vector<uint64_t> off_;    // key id -> arena offset (n + 1 entries)
    std::string arena_;
    uint64_t mask_;

    static uint64_t hash(std::string_view s) {
        uint64_t h = 0x9E3779B97F4A7C15ull ^ (uint64_t)s.size();
        const char* p = s.data();
        size_t n = s.size();
        for (; n >= 8; p += 8, n -= 8) {
            uint64_t w;
            memcpy(&w, p, 8);
            h = mix64(h ^ w);
        }
        if (n) {
            uint64_t w = 0;
            memcpy(&w, p, n);
            h = mix64(h ^ w ^ 0xA5A5A5A5A5A5A5A5ull);
        }
        return mix64(h);
    }
    void grow() {
        std::vector<uint64_t> old;
        old.swap(slots_);
        slots_.assign(old.size() * 2, 0);
        mask_ = slots_.size() - 1;
        for (uint64_t e : old)
            if (e) {
                uint64_t i = (e >> 32) & mask_;
                while (slots_[i]) i = (i + 1) & mask_;
                slots_[i] = e;
            }
    }

   public:
    explicit StrInterner(size_t expect = 64) {
        size_t cap = 64;
        while (cap < 2 * expect) cap <<= 1;
        slots_.assign(cap, 0);
        mask_ = cap - 1;
        off_.push_back(0);
    }
    size_t size() const { return off_.size() - 1; }
    std::string_view at(uint32_t id) const { return std::string_view(arena_.data() + off_[id], off_[id + 1] - off_[id]); }
    // id of `s`; *fresh = true when this call added it
    uint32_t intern(std::string_view s, bool* fresh = nullptr) {
        const uint64_t h = hash(s), tag = h & ~0xFFFFFFFFull;
        uint64_t i = (h >> 32) & mask_;
        for (uint64_t e; (e = slots_[i]) != 0; i = (i + 1) & mask_)
            if ((e & ~0xFFFFFFFFull) == tag) {
                const uint32_t id = (uint32_t)e - 1;
                if (at(id) == s) {
                    if (fresh) *fresh = false;
                    return id;
                }
            }
        const uint32_t id = (uint32_t)size();
        if (id >= 0xFFFFFFF0u) throw std::invalid_argument("more than 2^32 distinct strings in the taxonomy");
        arena_.append(s.data(), s.size());
        off_.push_back(arena_.size());
        slots_[i] = tag | (uint64_t)(id + 1);
        if (fresh) *fresh = true;
        if (2 * (size_t)(id + 1) > slots_.size()) grow();
        return id;
    }
};

// (rank id, identifier id) -> dense pair id, same scheme for 64-bit keys
class PairInterner {
    std::vector<uint64_t> keys_;  // key + 1 (0 = empty)
    std::vector<uint32_t> vals_;
    uint64_t mask_;
    uint32_t n_ = 0;

    void grow() {
        std::vector<uint64_t> ok;
        std::vector<uint32_t> ov;
        ok.swap(keys_);
        ov.swap(vals_);
        keys_.assign(ok.size() * 2, 0);
        vals_.assign(ok.size() * 2, 0);
        mask_ = keys_.size() - 1;
        for (size_t j = 0; j < ok.size(); j++)
            if (ok[j]) {
                uint64_t i = mix64(ok[j]) & mask_;
                while (keys_[i]) i = (i + 1) & mask_;
                keys_[i] = ok[j];
                vals_[i] = ov[j];
            }
    }

   public:
    explicit PairInterner(size_t expect = 64) {
        size_t cap = 64;
        while (cap < 2 * expect) cap <<= 1;
        keys_.assign(cap, 0);
        vals_.assign(cap, 0);
        mask_ = cap - 1;
    }
    uint32_t intern(uint64_t key, bool* fresh) {
        const uint64_t k1 = key + 1;
        uint64_t i = mix64(k1) & mask_;
        for (; keys_[i]; i = (i + 1) & mask_)
            if (keys_[i] == k1) {
                *fresh = false;
                return vals_[i];
            }
        keys_[i] = k1;
        vals_[i] = n_;
        *fresh = true;
        const uint32_t id = n_++;
        if (2 * (size_t)n_ > keys_.size()) grow();
        return id;
    }
};

}  // namespace

void build_taxonomy(const int64_t* taxids, const uint64_t* off, const char* blob, uint64_t n, const Cutoffs& cutoffs, HostTaxonomy& T) {
    T = HostTaxonomy();
    const auto bb = make_backbone(cutoffs);
    StrInterner rank_raw;                                  // raw rank text -> raw id
    std::vector<int64_t> raw_rid;                          // raw id -> rank id, -1 = rank_from_str rejects it
    std::unordered_map<std::string, uint32_t> rank_canon;  // (def|slug) -> rank id
    StrInterner ident_id(n);
    PairInterner pair_id(n);  // (rank, ident) -> pair
    std::vector<std::pair<uint32_t, uint32_t>> pairs;
    std::vector<uint32_t> pos_pair;
    std::unordered_map<std::string, std::vector<double>> interp_memo;
    std::string memo_key;

    T.taxids.assign(taxids, taxids + n);
    T.lin_off.reserve(n + 1);
    T.lin_off.push_back(0);
    T.lin_ok.resize(n);
    std::vector<uint32_t> tmp_rank, tmp_ident;
    std::vector<const RankInfo*> rk;
    std::string_view memo_raw[65];  // views into `blob`
    uint32_t memo_raw_id[65];
    std::fill(memo_raw_id, memo_raw_id + 65, 0xFFFFFFFFu);  // = not set (an empty view would equal an empty rank name)
    for (uint64_t i = 0; i < n; i++) {
        std::string_view s(blob + off[i], off[i + 1] - off[i]);
        tmp_rank.clear();
        tmp_ident.clear();
        bool ok = true;
        size_t pos = 0;
        while (ok) {  // split(";")  blast_result.rs:49
            size_t semi = s.find(';', pos);
            std::string_view part = s.substr(pos, semi == std::string_view::npos ? std::string_view::npos : semi - pos);
            // split("__") must give exactly two pieces (blast_result.rs:60-67)
            size_t k = part.find("__");
            if (k == std::string_view::npos || part.find("__", k + 2) != std::string_view::npos) {
                ok = false;
                break;
            }
            const std::string_view rraw = part.substr(0, k), ident = part.substr(k + 2);
            bool fresh = false;
            // neighbouring lineages mostly carry the same rank name at the same depth: one compare instead of a lookup
            const size_t depth = std::min<size_t>(tmp_rank.size(), 64);
            uint32_t raw;
            if (memo_raw_id[depth] != 0xFFFFFFFFu && memo_raw[depth] == rraw)
                raw = memo_raw_id[depth];
            else {
                raw = rank_raw.intern(rraw, &fresh);
                memo_raw[depth] = rraw, memo_raw_id[depth] = raw;
            }
            if (fresh) {
                int64_t rid = -1;  // the reference would panic while parsing this lineage; only fatal if a top row uses it
                try {
                    RankInfo ri = rank_from_str(rraw);
                    std::string canon = ri.def >= 0 ? std::string("D") + std::to_string(ri.def) : "O" + ri.slug;
                    auto ic = rank_canon.find(canon);
                    if (ic != rank_canon.end())
                        rid = ic->second;
                    else {
                        rid = (int64_t)T.ranks.size();
                        T.ranks.push_back(ri);
                        rank_canon.emplace(canon, (uint32_t)rid);
                    }
                } catch (const DataErr&) {
                }
                raw_rid.push_back(rid);
            }
            if (raw_rid[raw] < 0) {
                ok = false;
                break;
            }
            const uint32_t iid = ident_id.intern(ident, &fresh);
            if (fresh) T.idents.emplace_back(ident);
            tmp_rank.push_back((uint32_t)raw_rid[raw]);
            tmp_ident.push_back(iid);
            if (semi == std::string_view::npos) break;
            pos = semi + 1;
        }
        if (ok && tmp_rank.size() > 64)
            throw std::invalid_argument("lineage with more than 64 ranks is not supported (taxid " + std::to_string(taxids[i]) + ")");
        T.lin_ok[i] = ok ? 1 : 0;
        if (ok) {
            for (size_t j = 0; j < tmp_rank.size(); j++) {
                T.pos_rank.push_back(tmp_rank[j]);
                T.pos_ident.push_back(tmp_ident[j]);
                bool fresh;
                const uint32_t pid = pair_id.intern(((uint64_t)tmp_rank[j] << 32) | tmp_ident[j], &fresh);
                if (fresh) pairs.push_back({tmp_rank[j], tmp_ident[j]});
                pos_pair.push_back(pid);
            }
            // cutoffs: a pure function of the rank vector -> memoised
            memo_key.assign((const char*)tmp_rank.data(), tmp_rank.size() * 4);
            auto im = interp_memo.find(memo_key);
            if (im == interp_memo.end()) {
                rk.clear();
                for (uint32_t r : tmp_rank) rk.push_back(&T.ranks[r]);
                im = interp_memo.emplace(memo_key, interpolate_cutoffs(rk, bb)).first;
            }
            T.cut.insert(T.cut.end(), im->second.begin(), im->second.end());
        }
        T.lin_off.push_back((uint32_t)T.pos_rank.size());
    }
    // string dictionaries over the distinct (rank, identifier) pairs
    StrInterner lvl_dict(pairs.size()), bean_dict(pairs.size());
    std::vector<uint32_t> pair_lvl(pairs.size()), pair_bean(pairs.size());
    std::string tmp;
    for (size_t p = 0; p < pairs.size(); p++) {
        const std::string& d = T.ranks[pairs[p].first].display;
        const std::string& id = T.idents[pairs[p].second];
        tmp.assign(d).append(id);
        pair_lvl[p] = lvl_dict.intern(tmp);  // fmtc.rs:153-157
        tmp.assign(d).append("__").append(id);
        pair_bean[p] = bean_dict.intern(tmp);  // consensus_result.rs:70-73
    }
    // identifier order (String cmp = bytewise) for the bean sort (bbci.rs:50-60)
    std::vector<uint32_t> iord(T.idents.size());
    std::iota(iord.begin(), iord.end(), 0);
    std::sort(iord.begin(), iord.end(), [&](uint32_t a, uint32_t b) { return T.idents[a] < T.idents[b]; });
    std::vector<uint32_t> irank(T.idents.size());
    for (size_t i = 0; i < iord.size(); i++) irank[iord[i]] = (uint32_t)i;
    // rank equality classes: class of the parsed LinnaeanRank, and of the rank that
    // get_rank_adjusted_by_identity would hand back for that position (bbci.rs:22-30):
    // DefaultRank(rank,_) -> rank ; NonDefaultRank(s,_) -> Other(s) with s = rank.to_string()
    std::unordered_map<std::string, uint16_t> cls;
    auto cls_of = [&](const std::string& k) {
        auto it = cls.find(k);
        if (it != cls.end()) return it->second;
        if (cls.size() >= 65535) throw std::invalid_argument("more than 65535 distinct rank names");
        uint16_t v = (uint16_t)cls.size();
        cls.emplace(k, v);
        return v;
    };
    std::vector<uint16_t> rank_c(T.ranks.size()), allowed_c(T.ranks.size());
    for (size_t r = 0; r < T.ranks.size(); r++) {
        const RankInfo& ri = T.ranks[r];
        rank_c[r] = cls_of(ri.def >= 0 ? std::string("D") + std::to_string(ri.def) : "O" + ri.slug);
        bool in_bb = false;
        if (ri.def >= 0)
            for (auto& b : bb) in_bb |= b.def == ri.def;
        allowed_c[r] = in_bb ? rank_c[r] : cls_of("O" + ri.display);
    }
    const size_t np = T.pos_rank.size();
    T.lvl_key.resize(np);
    T.bean_key.resize(np);
    T.ident_rank.resize(np);
    T.rank_cls.resize(np);
    T.allowed_cls.resize(np);
    for (size_t p = 0; p < np; p++) {
        T.lvl_key[p] = pair_lvl[pos_pair[p]];
        T.bean_key[p] = pair_bean[pos_pair[p]];
        T.ident_rank[p] = irank[T.pos_ident[p]];
        T.rank_cls[p] = rank_c[T.pos_rank[p]];
        T.allowed_cls[p] = allowed_c[T.pos_rank[p]];
    }
    // taxid -> lineage hash table (left join key, mod.rs:72-76)
    uint64_t cap = 16;
    while (cap < 2 * n) cap <<= 1;
    if (cap > (1ull << 31)) throw std::invalid_argument("taxonomy too large");
    T.slots.assign(cap, HashSlot{0, 0, 0});
    T.hash_mask = (uint32_t)(cap - 1);
    for (uint64_t i = 0; i < n; i++) {
        uint32_t h = (uint32_t)mix64((uint64_t)taxids[i]) & T.hash_mask;
        while (T.slots[h].used) {
            if (T.slots[h].key == taxids[i])
                throw DataErr("duplicate taxid " + std::to_string(taxids[i]) +
                              " in the taxonomy file (the reference's left join would duplicate hit rows; not supported)");
            h = (h + 1) & T.hash_mask;
        }
        T.slots[h] = HashSlot{taxids[i], (uint32_t)i, 1};
    }
}

// ---------------------------------------------------------------------------------------------------------------
static std::string slurp(const char* path, const char* what) {
    std::ifstream f(path, std::ios::binary);
    if (!f) throw IoErr(std::string(what) + " file not found: " + path);
    f.seekg(0, std::ios::end);
    std::streamoff n = f.tellg();
    f.seekg(0);
    std::string s((size_t)n, '\0');
    if (n && !f.read(s.data(), n)) throw IoErr(std::string("Unexpected error on read `") + what + "` file");
    return s;
}

void read_taxonomy_json(const char* path, bool use_taxid, std::vector<int64_t>& taxids, std::vector<uint64_t>& off, std::string& blob) {
    std::string buf = slurp(path, "Taxonomies");
    taxids.clear();
    off.assign(1, 0);
    blob.clear();
    try {
        JsonCursor c(buf.data(), buf.size());
        bool have_version = false, have_source = false, have_tax = false;
        std::string key, num, txt, tmp;
        c.expect('{');
        if (!c.consume('}')) {
            do {
                c.string(&key);
                c.expect(':');
                if (key == "blutilsVersion") {
                    if (have_version) c.fail("duplicate field `blutilsVersion`");
                    c.string(nullptr);
                    have_version = true;
                } else if (key == "sourceDatabase") {
                    if (have_source) c.fail("duplicate field `sourceDatabase`");
                    c.string(nullptr);
                    have_source = true;
                } else if (key == "ignoreTaxids") {
                    if (!c.consume_null()) {
                        c.expect('[');
                        if (!c.consume(']')) {
                            do c.u64();
                            while (c.consume(','));
                            c.expect(']');
                        }
                    }
                } else if (key == "replaceRank") {
                    if (!c.consume_null()) {
                        c.expect('{');
                        if (!c.consume('}')) {
                            do {
                                c.string(nullptr);
                                c.expect(':');
                                c.string(nullptr);
                            } while (c.consume(','));
                            c.expect('}');
                        }
                    }
                } else if (key == "dropNonLinnaeanTaxonomies") {
                    c.skip_value();
                } else if (key == "taxonomies") {
                    if (have_tax) c.fail("duplicate field `taxonomies`");
                    have_tax = true;
                    c.expect('[');
                    if (!c.consume(']')) {
                        do {
                            bool h_id = false, h_rank = false, h_num = false, h_txt = false, h_acc = false;
                            uint64_t tid = 0;
                            c.expect('{');
                            if (!c.consume('}')) {
                                do {
                                    c.string(&key);
                                    c.expect(':');
                                    if (key == "taxid") {
                                        tid = c.u64();
                                        h_id = true;
                                    } else if (key == "rank") {
                                        c.string(nullptr);
                                        h_rank = true;
                                    } else if (key == "numericLineage") {
                                        c.string(use_taxid ? &num : nullptr);
                                        h_num = true;
                                    } else if (key == "textLineage") {
                                        c.string(use_taxid ? nullptr : &txt);
                                        h_txt = true;
                                    } else if (key == "accessions") {
                                        h_acc = true;
                                        c.expect('[');
                                        if (!c.consume(']')) {
                                            do {
                                                bool a = false, o = false;
                                                c.expect('{');
                                                if (!c.consume('}')) {
                                                    do {
                                                        c.string(&tmp);
                                                        c.expect(':');
                                                        if (tmp == "accession") {
                                                            c.string(nullptr);
                                                            a = true;
                                                        } else if (tmp == "oid") {
                                                            c.string(nullptr);
                                                            o = true;
                                                        } else
                                                            c.skip_value();
                                                    } while (c.consume(','));
                                                    c.expect('}');
                                                }
                                                if (!a || !o) c.fail("missing field in `accessions` entry");
                                            } while (c.consume(','));
                                            c.expect(']');
                                        }
                                    } else
                                        c.skip_value();
                                } while (c.consume(','));
                                c.expect('}');
                            }
                            if (!(h_id && h_rank && h_num && h_txt && h_acc)) c.fail("missing field in `taxonomies` entry");
                            // taxid: u64 -> to_f64() -> cast Int64 (mod.rs:278,309); an out-of-range cast is a null key
                            double f = (double)tid;
                            if (f < 9223372036854775808.0) {
                                taxids.push_back((int64_t)f);
                                const std::string& s = use_taxid ? num : txt;
                                blob += s;
                                off.push_back(blob.size());
                            }
                        } while (c.consume(','));
                        c.expect(']');
                    }
                } else
                    c.skip_value();
            } while (c.consume(','));
            c.expect('}');
        }
        c.end();
        if (!have_version) c.fail("missing field `blutilsVersion`");
        if (!have_source) c.fail("missing field `sourceDatabase`");
        if (!have_tax) c.fail("missing field `taxonomies`");
    } catch (const JsonError& e) {
        throw IoErr(std::string("Unexpected error detected on parse `taxonomies` as json: ") + e.what());
    }
}

static bool ends_with(const std::string& s, const char* suf) {
    size_t n = strlen(suf);
    return s.size() >= n && !s.compare(s.size() - n, n, suf);
}

void read_custom_cutoffs(const char* path, Cutoffs& out) {
    static const char* names[8] = {"domain", "kingdom", "phylum", "class", "order", "family", "genus", "species"};
    std::string p(path);
    size_t dot = p.find_last_of('.');
    size_t slash = p.find_last_of('/');
    if (dot == std::string::npos || (slash != std::string::npos && dot < slash)) throw DataErr("File must have an extension");
    const bool yaml = ends_with(p, ".yaml"), json = ends_with(p, ".json");
    if (!yaml && !json) throw DataErr("Custom taxon file must be a YAML or JSON file");
    std::string buf;
    try {
        buf = slurp(path, "custom taxon");
    } catch (const IoErr& e) {
        throw DataErr(std::string("Could not open custom taxon file: ") + e.what());
    }
    int32_t v[8];
    bool seen[8] = {false};
    for (int i = 0; i < 8; i++) v[i] = BLU_CUTOFF_ABSENT;
    auto set = [&](const std::string& k, bool is_null, int64_t val) {
        for (int i = 0; i < 8; i++)
            if (k == names[i]) {
                if (seen[i]) throw DataErr("Could not parse custom taxon file: duplicate field `" + k + "`");
                seen[i] = true;
                if (!is_null) {
                    if (val < -32768 || val > 32767) throw DataErr("Could not parse custom taxon file: `" + k + "` is not an i16");
                    v[i] = (int32_t)val;
                }
                return;
            }
        // unknown keys are ignored (serde default)
    };
    if (json) {
        try {
            JsonCursor c(buf.data(), buf.size());
            std::string key;
            c.expect('{');
            if (!c.consume('}')) {
                do {
                    c.string(&key);
                    c.expect(':');
                    bool known = false;
                    for (auto nm : names) known |= key == nm;
                    if (!known)
                        c.skip_value();
                    else if (c.consume_null())
                        set(key, true, 0);
                    else
                        set(key, false, c.i64());
                } while (c.consume(','));
                c.expect('}');
            }
            c.end();
        } catch (const JsonError& e) {
            throw DataErr(std::string("Could not parse custom taxon file from JSON: ") + e.what());
        }
    } else {
        // YAML subset: a flat block mapping `key: integer` (the shape of assets/custom-taxon-cutoffs-*.yaml),
        // comments, blank lines, `---`, `~`/`null`.
        size_t pos = 0;
        while (pos < buf.size()) {
            size_t nl = buf.find('\n', pos);
            std::string line = buf.substr(pos, nl == std::string::npos ? std::string::npos : nl - pos);
            pos = nl == std::string::npos ? buf.size() : nl + 1;
            size_t hash = line.find('#');
            if (hash != std::string::npos) line.resize(hash);
            while (!line.empty() && is_rust_ws((unsigned char)line.back())) line.pop_back();
            size_t a = 0;
            while (a < line.size() && line[a] == ' ') a++;
            if (a == line.size() || line == "---" || line == "...") continue;
            if (a != 0) throw DataErr("Could not parse custom taxon file from YAML: nested structures are not supported");
            size_t colon = line.find(':');
            if (colon == std::string::npos) throw DataErr("Could not parse custom taxon file from YAML: expected `key: value`");
            std::string k = line.substr(0, colon);
            while (!k.empty() && k.back() == ' ') k.pop_back();
            if (k.size() >= 2 && ((k.front() == '"' && k.back() == '"') || (k.front() == '\'' && k.back() == '\''))) k = k.substr(1, k.size() - 2);
            std::string val = line.substr(colon + 1);
            size_t b = 0;
            while (b < val.size() && val[b] == ' ') b++;
            val = val.substr(b);
            if (val.empty() || val == "~" || val == "null" || val == "Null" || val == "NULL") {
                set(k, true, 0);
                continue;
            }
            size_t i = (val[0] == '-' || val[0] == '+') ? 1 : 0;
            if (i == val.size()) throw DataErr("Could not parse custom taxon file from YAML: `" + k + "` is not an integer");
            int64_t x = 0;
            for (size_t j = i; j < val.size(); j++) {
                if (val[j] < '0' || val[j] > '9' || j - i > 9) throw DataErr("Could not parse custom taxon file from YAML: `" + k + "` is not an integer");
                x = x * 10 + (val[j] - '0');
            }
            set(k, false, val[0] == '-' ? -x : x);
        }
    }
    if (v[0] == BLU_CUTOFF_ABSENT) throw DataErr("Could not parse custom taxon file: missing field `domain`");
    if (v[7] == BLU_CUTOFF_ABSENT) throw DataErr("Could not parse custom taxon file: missing field `species`");
    out.has_custom = true;
    for (int i = 0; i < 8; i++) out.custom[i] = v[i];
}


// ---------------------------------------------------------------------------------------------------------------
// Side-car cache.  Layout (native little-endian): CacheHeader, then the sections in the order of write_all() below,
// each vector as u64 count + raw elements; strings as u32 length + bytes.  `payload_hash` covers everything behind
// the header, `file_size` the whole file, so a cut-off or bit-flipped cache is rejected.
namespace {

constexpr uint64_t kCacheMagic = 0x3143584154554C42ull;  // "BLUTAXC1"
constexpr uint32_t kCacheVersion = 1;

struct CacheHeader {
    uint64_t magic;
    uint32_t version, header_bytes;
    TaxCacheKey key;
    uint64_t file_size, payload_hash;
};

// four independent multiply-xorshift lanes over 32-byte blocks (~ memory speed), folded at the end
struct StreamHash {
    uint64_t h[4] = {0x243F6A8885A308D3ull, 0x13198A2E03707344ull, 0xA4093822299F31D0ull, 0x082EFA98EC4E6C89ull};
    uint64_t total = 0;
    unsigned char tail[32];
    size_t n_tail = 0;

    static uint64_t step(uint64_t h, uint64_t w) {
        h = (h ^ w) * 0x9E3779B97F4A7C15ull;
        return h ^ (h >> 29);
    }
    void block(const unsigned char* p) {
        uint64_t w[4];
        memcpy(w, p, 32);
        for (int i = 0; i < 4; i++) h[i] = step(h[i], w[i]);
    }
    void update(const void* data, size_t n) {
        const unsigned char* p = (const unsigned char*)data;
        total += n;
        if (n_tail) {
            size_t take = std::min(n, 32 - n_tail);
            memcpy(tail + n_tail, p, take);
            n_tail += take, p += take, n -= take;
            if (n_tail < 32) return;
            block(tail);
            n_tail = 0;
        }
        for (; n >= 32; p += 32, n -= 32) block(p);
        if (n) memcpy(tail, p, n), n_tail = n;
    }
    uint64_t finish() {
        if (n_tail) {
            memset(tail + n_tail, 0, 32 - n_tail);
            block(tail);
        }
        uint64_t x = total;
        for (int i = 0; i < 4; i++) x = mix64(x ^ h[i]);
        return x;
    }
};

struct FileCloser {
    FILE* f;
    ~FileCloser() {
        if (f) fclose(f);
    }
};

struct CacheWriter {
    FILE* f;
    StreamHash hash;
    void raw(const void* p, size_t n) {
        if (n && fwrite(p, 1, n, f) != n) throw IoErr("could not write the taxonomy cache (disk full?)");
        hash.update(p, n);
    }
    template <class T>
    void vec(const std::vector<T>& v) {
        const uint64_t n = v.size();
        raw(&n, 8);
        raw(v.data(), n * sizeof(T));
    }
    void str(const std::string& s) {
        const uint32_t n = (uint32_t)s.size();
        raw(&n, 4);
        raw(s.data(), n);
    }
};

// reads straight from the file into the destination vectors, hashing as it goes; `left` = payload bytes not yet read
struct CacheReader {
    FILE* f;
    uint64_t left;
    StreamHash hash;
    void raw(void* dst, size_t n) {
        if (left < n || (n && fread(dst, 1, n, f) != n)) throw IoErr("taxonomy cache is truncated");
        left -= n;
        hash.update(dst, n);
    }
    template <class T>
    void vec(std::vector<T>& v) {
        uint64_t n;
        raw(&n, 8);
        if (n > left / sizeof(T)) throw IoErr("taxonomy cache is truncated");
        v.resize(n);
        raw(v.data(), n * sizeof(T));
    }
    void str(std::string& s) {
        uint32_t n;
        raw(&n, 4);
        if (n > left) throw IoErr("taxonomy cache is truncated");
        s.resize(n);
        raw(s.data(), n);
    }
};

bool same_key(const TaxCacheKey& a, const TaxCacheKey& b) {
    if (a.json_size != b.json_size || a.json_hash != b.json_hash || a.use_taxid != b.use_taxid || a.taxon != b.taxon || a.has_custom != b.has_custom)
        return false;
    if (a.has_custom)
        for (int i = 0; i < 8; i++)
            if (a.custom[i] != b.custom[i]) return false;
    return true;
}

}  // namespace

uint64_t hash_file(const char* path, uint64_t* size) {
    FileCloser fc{fopen(path, "rb")};
    if (!fc.f) throw IoErr(std::string("Taxonomies file not found: ") + path);
    std::vector<char> buf(4 << 20);
    StreamHash h;
    for (;;) {
        size_t n = fread(buf.data(), 1, buf.size(), fc.f);
        h.update(buf.data(), n);
        if (n < buf.size()) {
            if (ferror(fc.f)) throw IoErr("Unexpected error on read `Taxonomies` file");
            break;
        }
    }
    if (size) *size = h.total;
    return h.finish();
}

TaxCacheKey make_cache_key(const char* json_path, bool use_taxid, const Cutoffs& cut) {
    TaxCacheKey k;
    k.json_hash = hash_file(json_path, &k.json_size);
    k.use_taxid = use_taxid ? 1 : 0;
    k.taxon = cut.taxon;
    k.has_custom = cut.has_custom ? 1 : 0;
    for (int i = 0; i < 8; i++) k.custom[i] = cut.has_custom ? cut.custom[i] : 0;
    return k;
}

void save_taxonomy_cache(const char* cache_path, const TaxCacheKey& key, const HostTaxonomy& T) {
    const std::string tmp = std::string(cache_path) + ".tmp." + std::to_string((long long)getpid());
    CacheHeader hd{};
    hd.magic = kCacheMagic;
    hd.version = kCacheVersion;
    hd.header_bytes = sizeof(CacheHeader);
    hd.key = key;
    try {
        FileCloser fc{fopen(tmp.c_str(), "wb")};
        if (!fc.f) throw IoErr("could not create the taxonomy cache " + tmp);
        if (fwrite(&hd, 1, sizeof hd, fc.f) != sizeof hd) throw IoErr("could not write the taxonomy cache");
        CacheWriter w{fc.f, StreamHash{}};
        const uint64_t n_ranks = T.ranks.size();
        w.raw(&n_ranks, 8);
        for (const RankInfo& r : T.ranks) {
            const int32_t def = r.def;
            w.raw(&def, 4);
            w.str(r.slug);
            w.str(r.display);
            w.str(r.full);
        }
        // identifiers: lengths, then one blob
        std::vector<uint32_t> len(T.idents.size());
        std::vector<char> blob;
        size_t total = 0;
        for (size_t i = 0; i < T.idents.size(); i++) total += T.idents[i].size(), len[i] = (uint32_t)T.idents[i].size();
        blob.reserve(total);
        for (const std::string& s : T.idents) blob.insert(blob.end(), s.begin(), s.end());
        w.vec(len);
        w.vec(blob);
        w.vec(T.taxids), w.vec(T.lin_off), w.vec(T.lin_ok);
        w.vec(T.pos_rank), w.vec(T.pos_ident), w.vec(T.lvl_key), w.vec(T.bean_key), w.vec(T.ident_rank);
        w.vec(T.cut), w.vec(T.rank_cls), w.vec(T.allowed_cls), w.vec(T.slots);
        w.raw(&T.hash_mask, 4);
        hd.file_size = sizeof hd + w.hash.total;
        hd.payload_hash = w.hash.finish();
        if (fseek(fc.f, 0, SEEK_SET) != 0 || fwrite(&hd, 1, sizeof hd, fc.f) != sizeof hd || fflush(fc.f) != 0)
            throw IoErr("could not write the taxonomy cache");
    } catch (...) {
        remove(tmp.c_str());
        throw;
    }
    if (rename(tmp.c_str(), cache_path) != 0) {
        remove(tmp.c_str());
        throw IoErr(std::string("could not move the taxonomy cache into place: ") + cache_path);
    }
}

bool load_taxonomy_cache(const char* cache_path, const TaxCacheKey& key, HostTaxonomy& T) {
    FileCloser fc{fopen(cache_path, "rb")};
    if (!fc.f) return false;
    CacheHeader hd;
    if (fread(&hd, 1, sizeof hd, fc.f) != sizeof hd) return false;
    if (hd.magic != kCacheMagic || hd.version != kCacheVersion || hd.header_bytes != sizeof(CacheHeader) || !same_key(hd.key, key)) return false;
    struct stat st;
    if (fstat(fileno(fc.f), &st) != 0 || (uint64_t)st.st_size != hd.file_size || hd.file_size < sizeof hd) return false;
    try {
        T = HostTaxonomy();
        CacheReader r{fc.f, hd.file_size - sizeof hd, StreamHash{}};
        uint64_t n_ranks;
        r.raw(&n_ranks, 8);
        if (n_ranks > 65536) throw IoErr("bad rank count");
        T.ranks.resize(n_ranks);
        for (RankInfo& ri : T.ranks) {
            int32_t def;
            r.raw(&def, 4);
            ri.def = def;
            r.str(ri.slug), r.str(ri.display), r.str(ri.full);
        }
        std::vector<uint32_t> len;
        std::vector<char> blob;
        r.vec(len);
        r.vec(blob);
        T.idents.resize(len.size());
        uint64_t at = 0;
        for (size_t i = 0; i < len.size(); i++) {
            if (at + len[i] > blob.size()) throw IoErr("bad identifier table");
            T.idents[i].assign(blob.data() + at, len[i]);
            at += len[i];
        }
        r.vec(T.taxids), r.vec(T.lin_off), r.vec(T.lin_ok);
        r.vec(T.pos_rank), r.vec(T.pos_ident), r.vec(T.lvl_key), r.vec(T.bean_key), r.vec(T.ident_rank);
        r.vec(T.cut), r.vec(T.rank_cls), r.vec(T.allowed_cls), r.vec(T.slots);
        r.raw(&T.hash_mask, 4);
        // shape checks: the kernels index these arrays without bounds tests
        const size_t np = T.pos_rank.size(), nl = T.taxids.size();
        const bool ok = r.left == 0 && r.hash.finish() == hd.payload_hash && T.lin_off.size() == nl + 1 && T.lin_ok.size() == nl &&
                        T.lin_off.back() == np && T.pos_ident.size() == np && T.lvl_key.size() == np && T.bean_key.size() == np &&
                        T.ident_rank.size() == np && T.cut.size() == np && T.rank_cls.size() == np && T.allowed_cls.size() == np &&
                        T.slots.size() == (size_t)T.hash_mask + 1 && (T.slots.size() & (T.slots.size() - 1)) == 0 && T.slots.size() >= 2 * nl;
        if (!ok) throw IoErr("taxonomy cache failed its checks");
        return true;
    } catch (const std::exception&) {  // IoErr, or bad_alloc / length_error from a damaged count
        T = HostTaxonomy();
        return false;
    }
}

}  // namespace blu
