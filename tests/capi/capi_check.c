/* capi_check.c — a C program written against include/matchy/matchy.h only (the reference's public C ABI), linked with
 * libmatchy_b200.so.  Same flow as the reference's own C smoke test (crates/matchy/tests/test_c_api.c: build three globs,
 * save, open, pattern_count, query, open_with_options x3), then the parts of the ABI that test does not reach: IP and
 * literal entries with data, result JSON, structured access, stats, the extractor, NULL arguments.
 * usage: capi_check <tmpfile>     prints "capi ok" and exits 0, or prints the failed check and exits 1. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "matchy/matchy.h"

#define CHECK(c) do { if (!(c)) { fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); return 1; } } while (0)

static int add(matchy_builder_t* b, const char* k, const char* j) { return matchy_builder_add(b, k, j); }

int main(int argc, char** argv) {
  const char* path = argc > 1 ? argv[1] : "/tmp/matchy_b200_capi.mxy";

  /* ---- the reference's smoke flow ---- */
  matchy_builder_t* b = matchy_builder_new();
  CHECK(b != NULL);
  CHECK(add(b, "*.txt", "{}") == MATCHY_SUCCESS);
  CHECK(add(b, "*.log", "{}") == MATCHY_SUCCESS);
  CHECK(add(b, "test_*", "{}") == MATCHY_SUCCESS);
  CHECK(matchy_builder_save(b, path) == MATCHY_SUCCESS);
  matchy_builder_free(b);

  matchy_t* db = matchy_open(path);
  CHECK(db != NULL);
  CHECK(matchy_pattern_count(db) == 3);
  matchy_result_t r = matchy_query(db, "test_file.txt");
  CHECK(r.found);
  matchy_free_result(&r);
  r = matchy_query(db, "nothing.bin");
  CHECK(!r.found);
  matchy_free_result(&r);
  CHECK(matchy_has_glob_data(db) && matchy_has_string_data(db) && matchy_has_pattern_data(db) && !matchy_has_literal_data(db));
  char* ps = matchy_get_pattern_string(db, 0);
  CHECK(ps != NULL && (strcmp(ps, "*.txt") == 0 || strcmp(ps, "*.log") == 0 || strcmp(ps, "test_*") == 0));
  matchy_free_string(ps);
  CHECK(matchy_get_pattern_string(db, 3) == NULL);
  matchy_close(db);

  matchy_open_options_t opts;
  matchy_init_open_options(&opts);
  CHECK(opts.cache_capacity == 10000 && !opts.auto_reload && opts.reload_callback == NULL && opts.reload_callback_user_data == NULL);
  unsigned caps[3] = {10000, 0, 100};
  for (int k = 0; k < 3; k++) {
    matchy_init_open_options(&opts);
    opts.cache_capacity = caps[k];
    matchy_t* d2 = matchy_open_with_options(path, &opts);
    CHECK(d2 != NULL);
    for (int i = 0; i < 5; i++) {
      r = matchy_query(d2, "test_file.txt");
      CHECK(r.found);
      matchy_free_result(&r);
    }
    matchy_stats_t st;
    matchy_get_stats(d2, &st);
    CHECK(st.total_queries == 5 && st.queries_with_match == 5 && st.string_queries == 5);
    CHECK(st.cache_hits == (caps[k] ? 4u : 0u) && st.cache_misses == (caps[k] ? 1u : 0u));
    matchy_close(d2);
  }

  /* ---- a combined database with data, through the in-memory path ---- */
  b = matchy_builder_new();
  CHECK(add(b, "1.2.3.0/24", "{\"threat_level\":\"high\",\"score\":7,\"tags\":[\"a\",\"b\"],\"big\":70000,\"neg\":-5,\"pi\":1.5,\"ok\":true}") == MATCHY_SUCCESS);
  CHECK(add(b, "2001:db8::/32", "{\"v\":6}") == MATCHY_SUCCESS);
  CHECK(add(b, "evil.com", "{\"category\":\"malware\"}") == MATCHY_SUCCESS);
  CHECK(add(b, "*.bad.org", "\"scalar\"") == MATCHY_SUCCESS);            /* wrapped as {"value":"scalar"} */
  CHECK(add(b, "x.com", "null") == MATCHY_ERROR_INVALID_FORMAT);          /* no null in the data model */
  CHECK(add(b, "x.com", "{bad json") == MATCHY_ERROR_INVALID_FORMAT);
  CHECK(add(b, NULL, "{}") == MATCHY_ERROR_INVALID_PARAM);
  CHECK(matchy_builder_set_description(b, "capi check") == MATCHY_SUCCESS);
  CHECK(matchy_builder_set_schema(b, "no-such-schema") == MATCHY_ERROR_UNKNOWN_SCHEMA);
  uint8_t* buf = NULL;
  uintptr_t size = 0;
  CHECK(matchy_builder_build(b, &buf, &size) == MATCHY_SUCCESS && buf != NULL && size > 64);
  matchy_builder_free(b);
  db = matchy_open_buffer(buf, size);
  free(buf);
  CHECK(db != NULL);
  CHECK(matchy_has_ip_data(db) && matchy_has_literal_data(db) && matchy_has_glob_data(db));
  CHECK(strcmp(matchy_format(db), "Combined IP+Pattern database") == 0);
  char* meta = matchy_metadata(db);
  CHECK(meta != NULL && strstr(meta, "\"node_count\"") && strstr(meta, "capi check"));
  matchy_free_string(meta);

  r = matchy_query(db, "1.2.3.4");
  CHECK(r.found && r.prefix_len == 24 && r._db_ref == db);
  char* js = matchy_result_to_json(&r);
  CHECK(js != NULL);
  CHECK(strcmp(js, "{\"big\":70000,\"neg\":-5,\"ok\":true,\"pi\":1.5,\"score\":7,\"tags\":[\"a\",\"b\"],\"threat_level\":\"high\"}") == 0);
  matchy_free_string(js);
  matchy_entry_s entry;
  CHECK(matchy_result_get_entry(&r, &entry) == MATCHY_SUCCESS);
  matchy_entry_data_t ed;
  const char* p1[] = {"threat_level", NULL};
  CHECK(matchy_aget_value(&entry, &ed, p1) == MATCHY_SUCCESS && ed.has_data && ed.type_ == MATCHY_DATA_TYPE_UTF8_STRING);
  CHECK(ed.data_size == 4 && strcmp(ed.value.utf8_string, "high") == 0);
  const char* p2[] = {"tags", "1", NULL};
  CHECK(matchy_aget_value(&entry, &ed, p2) == MATCHY_SUCCESS && ed.type_ == MATCHY_DATA_TYPE_UTF8_STRING && strcmp(ed.value.utf8_string, "b") == 0);
  const char* p3[] = {"score", NULL};
  CHECK(matchy_aget_value(&entry, &ed, p3) == MATCHY_SUCCESS && ed.type_ == MATCHY_DATA_TYPE_UINT16 && ed.value.uint16 == 7 && ed.data_size == 2);
  const char* p4[] = {"big", NULL};
  CHECK(matchy_aget_value(&entry, &ed, p4) == MATCHY_SUCCESS && ed.type_ == MATCHY_DATA_TYPE_UINT32 && ed.value.uint32 == 70000);
  const char* p5[] = {"neg", NULL};
  CHECK(matchy_aget_value(&entry, &ed, p5) == MATCHY_SUCCESS && ed.type_ == MATCHY_DATA_TYPE_INT32 && ed.value.int32 == -5);
  const char* p6[] = {"pi", NULL};
  CHECK(matchy_aget_value(&entry, &ed, p6) == MATCHY_SUCCESS && ed.type_ == MATCHY_DATA_TYPE_DOUBLE && ed.value.double_value == 1.5);
  const char* p7[] = {"ok", NULL};
  CHECK(matchy_aget_value(&entry, &ed, p7) == MATCHY_SUCCESS && ed.type_ == MATCHY_DATA_TYPE_BOOLEAN && ed.value.boolean);
  const char* p8[] = {"tags", "2", NULL};
  CHECK(matchy_aget_value(&entry, &ed, p8) == MATCHY_ERROR_LOOKUP_PATH_INVALID && !ed.has_data);
  const char* p9[] = {"missing", NULL};
  CHECK(matchy_aget_value(&entry, &ed, p9) == MATCHY_ERROR_LOOKUP_PATH_INVALID);
  const char* p0[] = {NULL};
  CHECK(matchy_aget_value(&entry, &ed, p0) == MATCHY_SUCCESS && ed.type_ == MATCHY_DATA_TYPE_MAP && ed.data_size == 7);
  matchy_entry_data_list_t* list = NULL;
  CHECK(matchy_get_entry_data_list(&entry, &list) == MATCHY_SUCCESS && list != NULL);
  int nodes = 0;
  for (matchy_entry_data_list_t* n = list; n; n = n->next) nodes++;
  CHECK(nodes == 1 + 7 + 2); /* the map, its seven values, the two array items */
  CHECK(list->entry_data.type_ == MATCHY_DATA_TYPE_MAP);
  matchy_free_entry_data_list(list);
  matchy_free_result(&r);
  CHECK(r._data_cache == NULL);
  CHECK(matchy_result_to_json(&r) == NULL);

  r = matchy_query(db, "1.2.4.4");
  CHECK(!r.found);
  r = matchy_query(db, "2001:db8:1::5");
  CHECK(r.found && r.prefix_len == 32);
  js = matchy_result_to_json(&r);
  CHECK(js && strcmp(js, "{\"v\":6}") == 0);
  matchy_free_string(js);
  matchy_free_result(&r);
  matchy_query_into(db, "evil.com", &r);
  CHECK(r.found && r.prefix_len == 0);
  js = matchy_result_to_json(&r);
  CHECK(js && strcmp(js, "{\"category\":\"malware\"}") == 0);
  matchy_free_string(js);
  matchy_free_result(&r);
  r = matchy_query(db, "sub.bad.org");
  CHECK(r.found);
  js = matchy_result_to_json(&r);
  CHECK(js && strcmp(js, "{\"value\":\"scalar\"}") == 0);
  matchy_free_string(js);
  matchy_free_result(&r);
  r = matchy_query(db, "good.com");
  CHECK(!r.found && r._data_cache == NULL);
  CHECK(matchy_result_get_entry(&r, &entry) == MATCHY_ERROR_NO_DATA);
  r = matchy_query(db, "\xff\xfe");   /* not UTF-8 */
  CHECK(!r.found);
  r = matchy_query(NULL, "x");
  CHECK(!r.found);
  r = matchy_query(db, NULL);
  CHECK(!r.found);
  matchy_stats_t st;
  matchy_get_stats(db, &st);
  CHECK(st.total_queries == 6 && st.queries_with_match == 4 && st.queries_without_match == 2);
  CHECK(st.ip_queries == 2 && st.string_queries == 4);   /* (sic) the IP miss counts as a string query, as in the reference */
  matchy_clear_cache(db);
  matchy_close(db);
  CHECK(matchy_open("/nonexistent/db.mxy") == NULL);
  CHECK(matchy_open(NULL) == NULL);
  CHECK(matchy_open_buffer(NULL, 10) == NULL);
  CHECK(matchy_open_buffer((const uint8_t*)"garbage-garbage-garbage", 23) == NULL);
  char* err = NULL;
  CHECK(matchy_validate(path, MATCHY_VALIDATION_STANDARD, &err) == MATCHY_SUCCESS && err == NULL);
  CHECK(matchy_validate("/nonexistent/db.mxy", MATCHY_VALIDATION_STRICT, &err) == MATCHY_ERROR_FILE_NOT_FOUND && err != NULL);
  matchy_free_string(err);

  /* ---- extractor ---- */
  matchy_extractor_t* ex = matchy_extractor_create(MATCHY_EXTRACT_ALL);
  CHECK(ex != NULL);
  const char* text = "GET http://evil.com/x from 192.168.001.1 and 10.0.0.7 mail bob@example.org v6 2001:0DB8::0001 "
                     "md5=5d41402abc4b2a76b9719d911017c592 done\n";
  matchy_matches_t ms;
  CHECK(matchy_extractor_extract_chunk(ex, (const uint8_t*)text, strlen(text), &ms) == MATCHY_SUCCESS);
  /* reference order: IPv6, IPv4, e-mail, domains, hashes */
  const char* want_v[] = {"2001:db8::1", "10.0.0.7", "bob@example.org", "evil.com", "example.org", "5d41402abc4b2a76b9719d911017c592"};
  const int want_t[] = {MATCHY_ITEM_TYPE_IPV6, MATCHY_ITEM_TYPE_IPV4, MATCHY_ITEM_TYPE_EMAIL, MATCHY_ITEM_TYPE_DOMAIN, MATCHY_ITEM_TYPE_DOMAIN, MATCHY_ITEM_TYPE_MD5};
  CHECK(ms.count == 6);
  for (int k = 0; k < 6; k++) {
    CHECK(ms.items[k].item_type == want_t[k]);
    CHECK(strcmp(ms.items[k].value, want_v[k]) == 0);
    CHECK(ms.items[k].end > ms.items[k].start && ms.items[k].end <= strlen(text));
  }
  CHECK(strncmp(text + ms.items[0].start, "2001:0DB8::0001", 15) == 0);  /* spans index the raw input */
  matchy_matches_free(&ms);
  CHECK(ms.items == NULL && ms.count == 0 && ms._internal == NULL);
  CHECK(matchy_extractor_extract_chunk(ex, NULL, 0, &ms) == MATCHY_ERROR_INVALID_PARAM);
  CHECK(matchy_extractor_extract_chunk(NULL, (const uint8_t*)text, 1, &ms) == MATCHY_ERROR_INVALID_PARAM);
  matchy_extractor_free(ex);
  matchy_extractor_t* ex2 = matchy_extractor_create(MATCHY_EXTRACT_IPV4);
  CHECK(ex2 != NULL);
  CHECK(matchy_extractor_extract_chunk(ex2, (const uint8_t*)text, strlen(text), &ms) == MATCHY_SUCCESS);
  CHECK(ms.count == 1 && ms.items[0].item_type == MATCHY_ITEM_TYPE_IPV4);
  matchy_matches_free(&ms);
  matchy_extractor_free(ex2);
  CHECK(strcmp(matchy_item_type_name(MATCHY_ITEM_TYPE_SHA256), "SHA256") == 0 && strcmp(matchy_item_type_name(200), "Unknown") == 0);
  CHECK(matchy_version() != NULL);
  printf("capi ok\n");
  return 0;
}
