// blu_taxonomy.cpp -- see blu_taxonomy.h.  Host-only; runs once per context.
#include "blu_taxonomy.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <map>
#include <numeric>

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
struct SvHash {
    size_t operator()(std::string_view s) const { return std::hash<std::string_view>()(s); }
};
}  // namespace

void build_taxonomy(const int64_t* taxids, const uint64_t* off, const char* blob, uint64_t n, const Cutoffs& cutoffs, HostTaxonomy& T) {
    T = HostTaxonomy();
    const auto bb = make_backbone(cutoffs);
    std::unordered_map<std::string, uint32_t> rank_raw;    // raw rank text -> rank id
    std::unordered_map<std::string, uint32_t> rank_canon;  // (def|slug) -> rank id
    std::unordered_map<std::string, uint32_t> ident_id;
    std::unordered_map<uint64_t, uint32_t> pair_id;  // (rank, ident) -> pair
    std::vector<std::pair<uint32_t, uint32_t>> pairs;
    std::vector<uint32_t> pos_pair;
    std::unordered_map<std::string, std::vector<double>> interp_memo;

    T.taxids.assign(taxids, taxids + n);
    T.lin_off.reserve(n + 1);
    T.lin_off.push_back(0);
    T.lin_ok.resize(n);
    std::vector<uint32_t> tmp_rank, tmp_ident;
    std::vector<const RankInfo*> rk;
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
            std::string rraw(part.substr(0, k));
            std::string ident(part.substr(k + 2));
            uint32_t rid;
            auto it = rank_raw.find(rraw);
            if (it != rank_raw.end())
                rid = it->second;
            else {
                RankInfo ri;
                try {
                    ri = rank_from_str(rraw);
                } catch (const DataErr&) {
                    ok = false;  // the reference would panic while parsing this lineage; only fatal if a top row uses it
                    break;
                }
                std::string canon = ri.def >= 0 ? std::string("D") + std::to_string(ri.def) : "O" + ri.slug;
                auto ic = rank_canon.find(canon);
                if (ic != rank_canon.end())
                    rid = ic->second;
                else {
                    rid = (uint32_t)T.ranks.size();
                    T.ranks.push_back(ri);
                    rank_canon.emplace(canon, rid);
                }
                rank_raw.emplace(rraw, rid);
            }
            uint32_t iid;
            auto ii = ident_id.find(ident);
            if (ii != ident_id.end())
                iid = ii->second;
            else {
                iid = (uint32_t)T.idents.size();
                T.idents.push_back(ident);
                ident_id.emplace(std::move(ident), iid);
            }
            tmp_rank.push_back(rid);
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
                uint64_t pk = ((uint64_t)tmp_rank[j] << 32) | tmp_ident[j];
                auto ip = pair_id.find(pk);
                uint32_t pid;
                if (ip != pair_id.end())
                    pid = ip->second;
                else {
                    pid = (uint32_t)pairs.size();
                    pairs.push_back({tmp_rank[j], tmp_ident[j]});
                    pair_id.emplace(pk, pid);
                }
                pos_pair.push_back(pid);
            }
            // cutoffs: a pure function of the rank vector -> memoised
            std::string key((const char*)tmp_rank.data(), tmp_rank.size() * 4);
            auto im = interp_memo.find(key);
            if (im == interp_memo.end()) {
                rk.clear();
                for (uint32_t r : tmp_rank) rk.push_back(&T.ranks[r]);
                im = interp_memo.emplace(key, interpolate_cutoffs(rk, bb)).first;
            }
            T.cut.insert(T.cut.end(), im->second.begin(), im->second.end());
        }
        T.lin_off.push_back((uint32_t)T.pos_rank.size());
    }
    // string dictionaries over the distinct (rank, identifier) pairs
    std::unordered_map<std::string, uint32_t> lvl_dict, bean_dict;
    std::vector<uint32_t> pair_lvl(pairs.size()), pair_bean(pairs.size());
    for (size_t p = 0; p < pairs.size(); p++) {
        const std::string& d = T.ranks[pairs[p].first].display;
        const std::string& id = T.idents[pairs[p].second];
        pair_lvl[p] = lvl_dict.emplace(d + id, (uint32_t)lvl_dict.size()).first->second;            // fmtc.rs:153-157
        pair_bean[p] = bean_dict.emplace(d + "__" + id, (uint32_t)bean_dict.size()).first->second;  // consensus_result.rs:70-73
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

}  // namespace blu
