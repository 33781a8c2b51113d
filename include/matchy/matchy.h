/* matchy/matchy.h — the reference's PUBLIC C ABI, served by the B200 engine (libmatchy_b200.so).
 *
 * Written from scratch for this repository: the declarations below reproduce the names, argument order, struct
 * layouts and constant values of matchy v1.2.2's cbindgen header (crates/matchy/include/matchy/matchy.h, cited per
 * item as "ref:LINE") so that a C/C++ program compiled against the reference's header links and runs against this
 * library unchanged.  Behaviour follows crates/matchy/src/c_api/matchy.rs (cited in matchy_capi.cpp).  What happens
 * behind the calls is different: matchy_open uploads the .mxy sections into HBM, matchy_query runs the lookup
 * kernels, matchy_extractor_extract_chunk runs the tokenizer + token kernels.  There is no CPU fallback: without a
 * CUDA device matchy_open / matchy_extractor_create return NULL.
 *
 * Environment: MATCHY_B200_DEVICE (CUDA ordinal, default 0), MATCHY_B200_PSL (path of public_suffix_list.dat;
 * default: ../data/ next to the shared object).
 */
#ifndef MATCHY_H
#define MATCHY_H

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

/* value of one matchy_entry_data_t (ref:24-36) */
typedef union matchy_entry_data_value_u {
  uint32_t pointer;
  const char *utf8_string;
  double double_value;
  const uint8_t *bytes;
  uint16_t uint16;
  uint32_t uint32;
  int32_t int32;
  uint64_t uint64;
  uint8_t uint128[16];
  bool boolean;
  float float_value;
} matchy_entry_data_value_u;

/* status codes (ref:46-86, 163-173) */
#define MATCHY_SUCCESS 0
#define MATCHY_ERROR_FILE_NOT_FOUND -1
#define MATCHY_ERROR_INVALID_FORMAT -2
#define MATCHY_ERROR_CORRUPT_DATA -3
#define MATCHY_ERROR_OUT_OF_MEMORY -4
#define MATCHY_ERROR_INVALID_PARAM -5
#define MATCHY_ERROR_IO -6
#define MATCHY_ERROR_SCHEMA_VALIDATION -7
#define MATCHY_ERROR_UNKNOWN_SCHEMA -8
#define MATCHY_ERROR_LOOKUP_PATH_INVALID -7
#define MATCHY_ERROR_NO_DATA -8
#define MATCHY_ERROR_DATA_PARSE -9

/* MMDB type codes reported in matchy_entry_data_t.type_ (ref:92-157) */
#define MATCHY_DATA_TYPE_EXTENDED 0
#define MATCHY_DATA_TYPE_POINTER 1
#define MATCHY_DATA_TYPE_UTF8_STRING 2
#define MATCHY_DATA_TYPE_DOUBLE 3
#define MATCHY_DATA_TYPE_BYTES 4
#define MATCHY_DATA_TYPE_UINT16 5
#define MATCHY_DATA_TYPE_UINT32 6
#define MATCHY_DATA_TYPE_MAP 7
#define MATCHY_DATA_TYPE_INT32 8
#define MATCHY_DATA_TYPE_UINT64 9
#define MATCHY_DATA_TYPE_UINT128 10
#define MATCHY_DATA_TYPE_ARRAY 11
#define MATCHY_DATA_TYPE_BOOLEAN 14
#define MATCHY_DATA_TYPE_FLOAT 15

#define MATCHY_VALIDATION_STANDARD 0 /* ref:178 */
#define MATCHY_VALIDATION_STRICT 1   /* ref:183 */

/* extractor selection bits for matchy_extractor_create (ref:188-228) */
#define MATCHY_EXTRACT_DOMAINS (1 << 0)
#define MATCHY_EXTRACT_EMAILS (1 << 1)
#define MATCHY_EXTRACT_IPV4 (1 << 2)
#define MATCHY_EXTRACT_IPV6 (1 << 3)
#define MATCHY_EXTRACT_HASHES (1 << 4)
#define MATCHY_EXTRACT_BITCOIN (1 << 5)
#define MATCHY_EXTRACT_ETHEREUM (1 << 6)
#define MATCHY_EXTRACT_MONERO (1 << 7)
#define MATCHY_EXTRACT_ALL 255

/* matchy_match_t.item_type (ref:233-288) */
#define MATCHY_ITEM_TYPE_DOMAIN 0
#define MATCHY_ITEM_TYPE_EMAIL 1
#define MATCHY_ITEM_TYPE_IPV4 2
#define MATCHY_ITEM_TYPE_IPV6 3
#define MATCHY_ITEM_TYPE_MD5 4
#define MATCHY_ITEM_TYPE_SHA1 5
#define MATCHY_ITEM_TYPE_SHA256 6
#define MATCHY_ITEM_TYPE_SHA384 7
#define MATCHY_ITEM_TYPE_SHA512 8
#define MATCHY_ITEM_TYPE_BITCOIN 9
#define MATCHY_ITEM_TYPE_ETHEREUM 10
#define MATCHY_ITEM_TYPE_MONERO 11

#ifdef __cplusplus
namespace matchy {
extern "C" {
#endif

typedef struct matchy_builder_t matchy_builder_t;     /* opaque (ref:293) */
typedef struct matchy_t matchy_t;                     /* opaque (ref:396) */
typedef struct matchy_extractor_t matchy_extractor_t; /* opaque (ref:513) */

typedef struct matchy_reload_event_t { /* ref:302-323 */
  const char *path;
  bool success;
  const char *error;
  uint64_t generation;
} matchy_reload_event_t;
typedef void (*matchy_reload_callback_t)(const struct matchy_reload_event_t *event, void *user_data); /* ref:354 */

typedef struct matchy_open_options_t { /* ref:361-391; defaults: 10000, false, NULL, NULL */
  uint32_t cache_capacity;
  bool auto_reload;
  matchy_reload_callback_t reload_callback;
  void *reload_callback_user_data;
} matchy_open_options_t;

typedef struct matchy_stats_t { /* ref:403-432 */
  uint64_t total_queries;
  uint64_t queries_with_match;
  uint64_t queries_without_match;
  uint64_t cache_hits;
  uint64_t cache_misses;
  uint64_t ip_queries;
  uint64_t string_queries;
} matchy_stats_t;

typedef struct matchy_result_t { /* ref:437-454 */
  bool found;
  uint8_t prefix_len;
  void *_data_cache;
  const struct matchy_t *_db_ref;
} matchy_result_t;

typedef struct matchy_entry_s { /* ref:459-468 */
  const struct matchy_t *db;
  const void *data_ptr;
} matchy_entry_s;

typedef struct matchy_entry_data_t { /* ref:473-494 */
  bool has_data;
  uint32_t type_;
  union matchy_entry_data_value_u value;
  uint32_t data_size;
  uint32_t offset;
} matchy_entry_data_t;

typedef struct matchy_entry_data_list_t { /* ref:499-508 */
  struct matchy_entry_data_t entry_data;
  struct matchy_entry_data_list_t *next;
} matchy_entry_data_list_t;

typedef struct matchy_match_t { /* ref:520-538 */
  uint8_t item_type;
  const char *value;
  uintptr_t start;
  uintptr_t end;
} matchy_match_t;

typedef struct matchy_matches_t { /* ref:543-556 */
  const struct matchy_match_t *items;
  uintptr_t count;
  void *_internal;
} matchy_matches_t;

/* ---- builder (ref:578-753): backed by the from-scratch .mxy writer (csrc/mxy_builder.cpp) ---- */
struct matchy_builder_t *matchy_builder_new(void);
int32_t matchy_builder_set_case_insensitive(struct matchy_builder_t *builder, bool case_insensitive);
int32_t matchy_builder_set_schema(struct matchy_builder_t *builder, const char *schema_name); /* always UNKNOWN_SCHEMA: schema validation is out of scope */
int32_t matchy_builder_add(struct matchy_builder_t *builder, const char *key, const char *json_data);
int32_t matchy_builder_set_description(struct matchy_builder_t *builder, const char *description);
int32_t matchy_builder_save(struct matchy_builder_t *builder, const char *filename);
int32_t matchy_builder_build(struct matchy_builder_t *builder, uint8_t **buffer, uintptr_t *size); /* malloc'd; caller frees */
void matchy_builder_free(struct matchy_builder_t *builder);

/* ---- database (ref:777-1190) ---- */
void matchy_init_open_options(struct matchy_open_options_t *options);
struct matchy_t *matchy_open_with_options(const char *filename, const struct matchy_open_options_t *options);
struct matchy_t *matchy_open(const char *filename);
struct matchy_t *matchy_open_buffer(const uint8_t *buffer, uintptr_t size);
void matchy_get_stats(const struct matchy_t *db, struct matchy_stats_t *stats);
void matchy_clear_cache(const struct matchy_t *db);
void matchy_close(struct matchy_t *db);
struct matchy_result_t matchy_query(const struct matchy_t *db, const char *query);
void matchy_query_into(const struct matchy_t *db, const char *query, struct matchy_result_t *result);
void matchy_free_result(struct matchy_result_t *result);
void matchy_free_string(char *string);
const char *matchy_version(void);
const char *matchy_format(const struct matchy_t *db);
bool matchy_has_ip_data(const struct matchy_t *db);
bool matchy_has_string_data(const struct matchy_t *db);
bool matchy_has_literal_data(const struct matchy_t *db);
bool matchy_has_glob_data(const struct matchy_t *db);
bool matchy_has_pattern_data(const struct matchy_t *db);
char *matchy_metadata(const struct matchy_t *db);
char *matchy_get_pattern_string(const struct matchy_t *db, uint32_t pattern_id);
uintptr_t matchy_pattern_count(const struct matchy_t *db);

/* ---- structured access to a result's data (ref:1220-1323) ---- */
int32_t matchy_result_get_entry(const struct matchy_result_t *result, struct matchy_entry_s *entry);
int32_t matchy_aget_value(const struct matchy_entry_s *entry, struct matchy_entry_data_t *entry_data, const char *const *path);
int32_t matchy_get_entry_data_list(const struct matchy_entry_s *entry, struct matchy_entry_data_list_t **entry_data_list);
void matchy_free_entry_data_list(struct matchy_entry_data_list_t *list);
char *matchy_result_to_json(const struct matchy_result_t *result);

int32_t matchy_validate(const char *filename, int32_t level, char **error_message); /* ref:1357 (structural checks only) */

/* ---- extractor (ref:1380-1460) ---- */
struct matchy_extractor_t *matchy_extractor_create(uint32_t flags);
int32_t matchy_extractor_extract_chunk(const struct matchy_extractor_t *extractor, const uint8_t *data, uintptr_t len, struct matchy_matches_t *matches);
void matchy_matches_free(struct matchy_matches_t *matches);
void matchy_extractor_free(struct matchy_extractor_t *extractor);
const char *matchy_item_type_name(uint8_t item_type);

#ifdef __cplusplus
} /* extern "C" */
} /* namespace matchy */
#endif

#endif /* MATCHY_H */
