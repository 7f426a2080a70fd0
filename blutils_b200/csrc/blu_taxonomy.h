// blu_taxonomy.h -- host-side taxonomy model: `.blutils.json` -> dictionary-encoded per-rank integer IDs,
// per-lineage interpolated cutoffs and the taxid hash table that the kernels read from HBM.
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <string_view>
#include <unordered_map>
#include <vector>

#include "blu_core.cuh"

namespace blu {

struct DataErr : std::runtime_error {
    using std::runtime_error::runtime_error;
};
struct IoErr : std::runtime_error {
    using std::runtime_error::runtime_error;
};

// LinnaeanRank (reference core/src/domain/dtos/linnaean_ranks.rs:14-107)
struct RankInfo {
    int def;              // 0..8 = Undefined, Domain .. Species; -1 = Other(slug)
    std::string slug;     // Other(slug)
    std::string display;  // Display / to_string(): one letter or the slug      (:74-89)
    std::string full;     // serde / as_full_rank_string(): full name or slug   (:14-29,:92-106)
};

struct Cutoffs {
    int taxon = BLU_TAXON_BACTERIA;
    bool has_custom = false;
    int32_t custom[8] = {0, 0, 0, 0, 0, 0, 0, 0};  // domain-first; BLU_CUTOFF_ABSENT for None
};

// One entry of the backbone (Taxon::get_taxon_cutoff, taxon.rs:105-185)
struct BackboneEntry {
    int def;  // 1..8 (Domain..Species)
    double cut;
};
std::vector<BackboneEntry> make_backbone(const Cutoffs& c);

// slugify 0.1.0 `slugify!(s)`; ASCII only (throws DataErr otherwise or on an empty slug)
std::string slugify_ascii(std::string_view s);
// LinnaeanRank::from_str (linnaean_ranks.rs:55-71)
RankInfo rank_from_str(std::string_view s);

struct HostTaxonomy {
    // dictionaries
    std::vector<RankInfo> ranks;          // distinct ranks
    std::vector<std::string> idents;      // distinct identifiers
    // per lineage
    std::vector<int64_t> taxids;
    std::vector<uint32_t> lin_off;        // n_lin + 1
    std::vector<uint8_t> lin_ok;
    // per position (SoA; uploaded as-is)
    std::vector<uint32_t> pos_rank;       // index into ranks   (host decode only)
    std::vector<uint32_t> pos_ident;      // index into idents  (host decode only)
    std::vector<uint32_t> lvl_key, bean_key, ident_rank;
    std::vector<double> cut;
    std::vector<uint16_t> rank_cls, allowed_cls;
    // taxid hash table
    std::vector<HashSlot> slots;
    uint32_t hash_mask = 0;

    size_t n_lin() const { return taxids.size(); }
    int lin_len(uint32_t l) const { return (int)(lin_off[l + 1] - lin_off[l]); }
    // "{rank}__{identifier}" of one position (TaxonomyBean::taxonomy_to_string, taxonomy_bean.rs:19-27)
    void append_bean(std::string& o, uint32_t pos) const;
    // Taxonomy::taxonomy_beans_to_string of the whole lineage (taxonomy_bean.rs:36-45)
    void append_lineage(std::string& o, uint32_t lin) const;
    size_t device_bytes() const;
};

// Builds everything from (taxid, lineage string) pairs.  Throws DataErr on duplicate taxids or lineages
// longer than 64 positions.
void build_taxonomy(const int64_t* taxids, const uint64_t* off, const char* blob, uint64_t n, const Cutoffs& cut, HostTaxonomy& out);

// get_taxonomies_dataframe (mod.rs:246-327): reads + validates the TaxonomiesMap JSON, picks numericLineage or
// textLineage, applies the u64 -> f64 -> i64 key cast.  Throws IoErr (maps to Err(MappedErrors)).
void read_taxonomy_json(const char* path, bool use_taxid, std::vector<int64_t>& taxids, std::vector<uint64_t>& off, std::string& blob);

// --- binary side-car cache of a built HostTaxonomy -------------------------------------
// The reference re-parses the taxonomy JSON on every run (mod.rs:254-265).  The cache stores every array that
// build_taxonomy() produces, keyed by the content hash of the JSON file and by everything else the arrays depend
// on (use_taxid, cutoff backbone).  A missing, stale, truncated or corrupted cache reads as a miss, never an error.
struct TaxCacheKey {
    uint64_t json_size = 0, json_hash = 0;
    int32_t use_taxid = 0, taxon = 0, has_custom = 0;
    int32_t custom[8] = {0, 0, 0, 0, 0, 0, 0, 0};
};
// 64-bit content hash + size of a file.  Throws IoErr if it cannot be read.
uint64_t hash_file(const char* path, uint64_t* size);
TaxCacheKey make_cache_key(const char* json_path, bool use_taxid, const Cutoffs& cut);
bool load_taxonomy_cache(const char* cache_path, const TaxCacheKey& key, HostTaxonomy& out);
// Written to `<cache_path>.tmp.<pid>` and renamed into place.  Throws IoErr.
void save_taxonomy_cache(const char* cache_path, const TaxCacheKey& key, const HostTaxonomy& T);

// CustomTaxon::from_file (taxon.rs:28-65)
void read_custom_cutoffs(const char* path, Cutoffs& out);

// InterpolatedIdentity::interpolate_identities (linnaean_ranks.rs:220-383); exposed for tests.
std::vector<double> interpolate_cutoffs(const std::vector<const RankInfo*>& ranks, const std::vector<BackboneEntry>& bb);

}  // namespace blu
