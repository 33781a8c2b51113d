/* matchy_b200.h — C ABI of the B200 log-scan engine (libmatchy_b200.so).
 *
 * This is the boundary a host-language binding (the reference's Rust FFI crate, ctypes, cgo …) binds.
 * Plain pointers and sizes only.  Three groups:
 *
 *   mgpu_*   the device path that replaces the body of the reference's
 *            processing::Worker::process_bytes   (crates/matchy/src/processing/mod.rs:353-448)
 *            = Extractor::extract_from_chunk     (crates/matchy-extractor/src/lib.rs:409-488)
 *            + Database::lookup_extracted        (crates/matchy/src/database.rs:889-901)
 *              → SearchTree::lookup              (crates/matchy-format/src/mmdb/tree.rs:46-125)
 *              → LiteralHash::lookup             (crates/matchy-literal-hash/src/lib.rs:467-575)
 *              → Paraglob::find_all              (crates/matchy-paraglob/src/paraglob_offset.rs:1028-1182)
 *   mxyr_*   host-side helpers for matched records only: MMDB data decode + `matchy match` NDJSON line
 *            (crates/matchy/src/bin/match_processor/parallel.rs:297-369, bin/cli_utils.rs:107-201)
 *   mxyb_*   `.mxy` writer = DatabaseBuilder      (crates/matchy-format/src/mmdb_builder.rs:392-760),
 *            what matchy_builder_new/add/save wrap (crates/matchy/include/matchy/matchy.h)
 *   mgen_*   deterministic synthetic DB / log generators for the BASELINE.json configs (bench + tests only)
 *
 * Error convention: int functions return 0 (MGPU_OK) or a negative MGPU_E_* code; mgpu_last_error() has text.
 * There is NO CPU fallback: without a CUDA device mgpu_create fails.
 */
#ifndef MATCHY_B200_H
#define MATCHY_B200_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define MGPU_OK 0
#define MGPU_E_CUDA (-1)      /* a CUDA call failed */
#define MGPU_E_FORMAT (-2)    /* not a valid / supported .mxy database */
#define MGPU_E_PARAM (-3)     /* bad argument */
#define MGPU_E_OVERFLOW (-4)  /* result buffers exhausted even after re-chunking */
#define MGPU_E_NODB (-5)      /* scan before mgpu_db_upload */

/* extractor enable bits == MATCHY_EXTRACT_* (crates/matchy/include/matchy/matchy.h:188-228) */
#define MGPU_X_DOMAINS (1u << 0)
#define MGPU_X_EMAILS (1u << 1)
#define MGPU_X_IPV4 (1u << 2)
#define MGPU_X_IPV6 (1u << 3)
#define MGPU_X_HASHES (1u << 4)
#define MGPU_X_BITCOIN (1u << 5)  /* Base58Check ('1…', '3…') and bech32/bech32m ("bc1…") words, lib.rs:1269-1319 */
#define MGPU_X_ETHEREUM (1u << 6) /* "0x" + 40 hex digits with the EIP-55 rule, lib.rs:1328-1361 */
#define MGPU_X_MONERO (1u << 7)   /* lib.rs:1367-1409 */
#define MGPU_X_CRYPTO (MGPU_X_BITCOIN | MGPU_X_ETHEREUM | MGPU_X_MONERO)
#define MGPU_X_SUPPORTED 0xFFu

/* item types == MATCHY_ITEM_TYPE_* (matchy.h:233-288) */
#define MGPU_T_DOMAIN 0
#define MGPU_T_EMAIL 1
#define MGPU_T_IPV4 2
#define MGPU_T_IPV6 3
#define MGPU_T_MD5 4
#define MGPU_T_SHA1 5
#define MGPU_T_SHA256 6
#define MGPU_T_SHA384 7
#define MGPU_T_SHA512 8
#define MGPU_T_BITCOIN 9
#define MGPU_T_ETHEREUM 10
#define MGPU_T_MONERO 11

#define MGPU_KIND_IP 1      /* QueryResult::Ip      */
#define MGPU_KIND_PATTERN 2 /* QueryResult::Pattern */
#define MGPU_NO_DATA 0xFFFFFFFFu

/* One match == one reference MatchResult (processing/mod.rs:131-145), flattened. */
typedef struct mgpu_match {
  uint64_t offset;      /* absolute byte offset of the token (base + span.0) */
  uint32_t len;         /* token length; matched_text = log[offset .. offset+len) */
  uint8_t item_type;    /* MGPU_T_* */
  uint8_t kind;         /* MGPU_KIND_* */
  uint8_t prefix_len;   /* kind==IP: tree depth at which the record sits (tree.rs:76-84,117-120) */
  uint8_t reserved;
  uint32_t n_ids;       /* kind==PATTERN: number of (pattern_id, data_offset) pairs */
  uint32_t ids_index;   /* first pair in the ids array; literal id first, then ascending glob ids (database.rs:916-965) */
  uint32_t data_offset; /* kind==IP: offset into the data section; else MGPU_NO_DATA */
  uint32_t pad;
} mgpu_match;

typedef struct mgpu_id_pair { uint32_t pattern_id, data_offset; } mgpu_id_pair;

/* WorkerStats counters (processing/mod.rs:86-128): lines, bytes, candidates, matches, by item type [12] */
typedef struct mgpu_counters { uint64_t lines, bytes, candidates, matches, by_type[12]; } mgpu_counters;

/* per-kernel device time of the last scan (CUDA events on the scan stream), milliseconds */
#define MGPU_K_TOKENIZE 0 /* tokenize_kernel: log bytes -> candidate words / anchors */
#define MGPU_K_TOKEN 1    /* token_kernel: candidates -> typed tokens (+ constant-time string filters on the fast path) */
#define MGPU_K_IPTRIE 2   /* iptrie_kernel */
#define MGPU_K_LITHASH 3  /* lithash_kernel (generic string path only) */
#define MGPU_K_STRINGS 4  /* exact_kernel (fast string path) or acglob_kernel (generic string path) */
#define MGPU_K_COUNT 5
typedef struct mgpu_timing {
  float kernel_ms[MGPU_K_COUNT];    /* summed over the chunks of the scan */
  uint32_t launches[MGPU_K_COUNT];
  float total_ms;                   /* sum of per-chunk kernel spans */
  uint32_t chunks;
  float scan_ms;                    /* first kernel of the scan to last, inter-chunk gaps included */
  uint32_t aux_launches;            /* helper kernels (newline cut search) */
} mgpu_timing;

typedef struct mgpu_db_info {
  uint32_t node_count, record_bits, ip_version, match_mode;
  uint32_t has_ip, has_literal, has_glob;
  uint32_t literal_count, glob_count, ac_node_count;
  uint64_t tree_bytes, literal_bytes, paraglob_bytes, file_bytes;
} mgpu_db_info;

typedef struct mgpu_ctx mgpu_ctx;

/* ---- device engine ------------------------------------------------------------------------------ */
mgpu_ctx* mgpu_create(int device, size_t chunk_bytes /* 0 = default */);
void mgpu_destroy(mgpu_ctx*);
const char* mgpu_last_error(void);
int mgpu_set_psl(mgpu_ctx*, const uint8_t* psl_text, size_t len);  /* Public Suffix List text (extractor lib.rs:1546-1563) */
int mgpu_db_upload(mgpu_ctx*, const uint8_t* mxy, size_t len);     /* replaces Database::from_storage (database.rs:649-713) */
int mgpu_db_info_get(mgpu_ctx*, mgpu_db_info* out);
uint32_t mgpu_default_flags(mgpu_ctx*);                            /* match_cmd.rs:277-303, crypto extractors off */

/* Scan one newline-aligned buffer in HOST memory (pinned or pageable): H2D on double-buffered streams,
 * kernels, D2H of records.  == Worker::process_bytes over every chunk FileReader::next_batch would cut. */
int mgpu_scan(mgpu_ctx*, const uint8_t* host, size_t len, uint64_t base, uint32_t flags);
/* Same, for a buffer already resident in this device's HBM.  `dev` must be 16-byte aligned and readable up to
 * len rounded up to a multiple of 1024, plus 1024 (the tokenizer reads whole 1 KiB tiles, token readers up to 20 bytes past a
 * token); mgpu_dev_alloc() allocates with that slack. */
int mgpu_scan_device(mgpu_ctx*, const uint8_t* dev, size_t len, uint64_t base, uint32_t flags);
/* Results of the last scan, sorted by (offset, item_type, len); valid until the next scan. */
int mgpu_results(mgpu_ctx*, const mgpu_match** recs, size_t* n_recs, const mgpu_id_pair** ids, size_t* n_ids);
int mgpu_counters_get(mgpu_ctx*, mgpu_counters* out);
int mgpu_timing_get(mgpu_ctx*, mgpu_timing* out);
void mgpu_set_keep_results(mgpu_ctx*, int keep); /* 0: count matches only (bench), 1: default */
/* String lookups: 0 = automatic (constant-time filters + sparse exact lookups when the database allows it, anchored
 * walks when provably identical), 1 = generic kernels with the reference's Aho-Corasick goto/failure walk, 2 = generic
 * kernels with anchored walks.  Results are identical in every mode; the switch exists for tests. */
void mgpu_set_ac_mode(mgpu_ctx*, int mode);
/* Test / debug switches, never needed for results: "tok_reserve" (slots per token-list reservation), "cap_str" / "cap_ip" /
 * "cap_rec" / "cap_ids" (pretend the work buffers are this small: drives the overflow -> split-and-redo path; 0 restores),
 * "verify_tokens" (audit every piece's IP token list on the device; mgpu_debug_get returns the sums: slots still poisoned,
 * padding slots, IPv4 tokens, IPv6 tokens, lookup hits, slots of unknown type, slots audited, 0; then the IP-trie kernel's own
 * sums: hits per thread, per warp ballot, per block) ; "variant": experiment switches inside kernels, 0 in production;
 * "device_sort_min" (records a piece must produce for its records to be sorted on the device before they are copied out;
 * default 131072, 1 = always; results are identical either way — mgpu_debug_get()[59] tells whether the last scan's records
 * arrived in order). */
int mgpu_set_option(mgpu_ctx*, const char* key, uint64_t value);
int mgpu_debug_get(mgpu_ctx*, uint64_t out[64]);

/* Extraction only (== Extractor::extract_from_chunk): triples (item_type, start, end) as uint64, sorted by
 * (start, item_type).  Returns the number of items (may exceed cap; only cap are written). */
int64_t mgpu_extract(mgpu_ctx*, const uint8_t* host, size_t len, uint32_t flags, uint64_t* out, size_t cap);

/* Single-string / single-IP lookups through the device tables (== Database::lookup, what matchy_query wraps,
 * c_api/matchy.rs:1099-1169).  Returns number of pairs for strings; 1/0 for IPs. */
int mgpu_lookup_string(mgpu_ctx*, const uint8_t* q, size_t len, mgpu_id_pair* out, size_t cap);
int mgpu_lookup_ip(mgpu_ctx*, const uint8_t ip16[16], int is_v6, uint32_t* data_offset, uint8_t* prefix_len);

/* device memory helpers for HBM-resident inputs (bench) */
void* mgpu_dev_alloc(mgpu_ctx*, size_t bytes);
void mgpu_dev_free(mgpu_ctx*, void*);
int mgpu_dev_upload(mgpu_ctx*, void* dst, const void* src, size_t bytes);
int mgpu_dev_download(mgpu_ctx*, void* dst, const void* src, size_t bytes);
void* mgpu_host_alloc_pinned(size_t bytes);
void mgpu_host_free_pinned(void*);
int mgpu_flush_l2(mgpu_ctx*);

/* ---- host-side record formatting (matched records only) ----------------------------------------- */
typedef struct mxyr_db mxyr_db;
mxyr_db* mxyr_open(const uint8_t* mxy, size_t len); /* keeps the pointer; NULL on bad format */
void mxyr_close(mxyr_db*);
/* JSON text of the MMDB value at data_offset (bin/cli_utils.rs:177-201); returns length, text in *out (valid until next call) */
size_t mxyr_data_json(mxyr_db*, uint32_t data_offset, const char** out);
/* `matchy match` NDJSON lines (parallel mode) for recs[0..n); log/base locate matched_text. */
size_t mxyr_ndjson(mxyr_db*, const mgpu_match* recs, size_t n, const mgpu_id_pair* ids, const uint8_t* log, uint64_t base,
                   const char* source, const char** out);
/* The same lines as `matchy match --threads 1` / --follow print them (bin/match_processor/sequential.rs:205-390): `timestamp` is
 * the caller's text (wall clock, "%.3f"), matched_text and cidr use the canonical text of an address (Ipv6Addr Display). */
size_t mxyr_ndjson_sequential(mxyr_db*, const mgpu_match* recs, size_t n, const mgpu_id_pair* ids, const uint8_t* log, uint64_t base,
                              const char* source, const char* timestamp, const char** out);

/* ---- .mxy writer -------------------------------------------------------------------------------- */
typedef struct mxyb_builder mxyb_builder;
mxyb_builder* mxyb_new(int case_insensitive);
void mxyb_free(mxyb_builder*);
const char* mxyb_error(mxyb_builder*);
void mxyb_data_begin(mxyb_builder*);
void mxyb_data_str(mxyb_builder*, const char* key, const char* val, size_t vlen);
void mxyb_data_i32(mxyb_builder*, const char* key, int32_t v);
void mxyb_data_u16(mxyb_builder*, const char* key, uint16_t v);
void mxyb_data_u32(mxyb_builder*, const char* key, uint32_t v);
void mxyb_data_u64(mxyb_builder*, const char* key, uint64_t v);
void mxyb_data_f64(mxyb_builder*, const char* key, double v);
void mxyb_data_bool(mxyb_builder*, const char* key, int v);
uint32_t mxyb_data_commit(mxyb_builder*);
int mxyb_add(mxyb_builder*, int kind /*0 auto,1 ip,2 literal,3 glob*/, const char* key, size_t klen, uint32_t data_offset);
void mxyb_set_epoch(mxyb_builder*, uint64_t epoch);
void mxyb_set_type(mxyb_builder*, const char* database_type);
void mxyb_set_description(mxyb_builder*, const char* lang, const char* text);
int mxyb_build(mxyb_builder*);
const uint8_t* mxyb_bytes(mxyb_builder*, size_t* len);
void mxyb_counts(mxyb_builder*, uint64_t* out3);
int mxyb_save(mxyb_builder*, const char* path);
uint64_t mxyb_xxh64(const uint8_t* p, size_t n);

/* ---- synthetic workloads (BASELINE.json configs 1-5) -------------------------------------------- */
/* Builds the config's database into a fresh builder (caller frees with mxyb_free). scale in (0,1] shrinks counts. */
mxyb_builder* mgen_db(int config, double scale);
/* Fills out[0..len) with the config's log lines for byte range [offset, offset+len) of the infinite
 * deterministic stream (block-addressable); the last line is cut at len and padded with '\n'. */
int mgen_log(int config, double scale, uint64_t offset, uint8_t* out, size_t len, int threads);
/* The same bytes, generated in the HBM of `device` (offset and len multiples of 64 KiB): config 5's 100 / 50 / 25 GB shards
 * never cross PCIe.  The same integer code runs on both sides (csrc/synth_gen.h). */
int mgen_log_device(int device, int config, double scale, uint64_t offset, void* dev_out, size_t len);

#ifdef __cplusplus
}
#endif
#endif
