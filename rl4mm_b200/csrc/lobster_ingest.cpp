// lobster_ingest.cpp -- fast host-side reader of LOBSTER CSV files for the packer (rl4mm_b200/packing.py).
//
// Replaces, for the hot path's data feed, the pandas / SQLAlchemy ingest of rl4mm/database/populate_database.py:38-95 and
// database_population_helpers.py:45-193: the message file is parsed column-wise in one pass
// over a memory map (time as exact integer nanoseconds, no float round trip), and of the orderbook file -- which has
// one row per message -- only the rows actually needed for the per-second snapshots are parsed.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

namespace {
struct Map {
  const char* p = nullptr; size_t n = 0; int fd = -1;
  bool open(const char* path) {
    fd = ::open(path, O_RDONLY);
    if (fd < 0) return false;
    struct stat st;
    if (fstat(fd, &st) != 0) return false;
    n = (size_t)st.st_size;
    if (n == 0) { p = ""; return true; }
    void* m = mmap(nullptr, n, PROT_READ, MAP_PRIVATE, fd, 0);
    if (m == MAP_FAILED) return false;
    p = (const char*)m;
    madvise(m, n, MADV_SEQUENTIAL);
    return true;
  }
  ~Map() { if (p && n) munmap((void*)p, n); if (fd >= 0) close(fd); }
};

inline const char* parse_i64(const char* s, const char* e, int64_t* out) {
  bool neg = false;
  if (s < e && (*s == '-' || *s == '+')) { neg = *s == '-'; s++; }
  int64_t v = 0;
  while (s < e && *s >= '0' && *s <= '9') { v = v * 10 + (*s - '0'); s++; }
  *out = neg ? -v : v;
  return s;
}
// "seconds[.fraction]" -> integer nanoseconds (fraction truncated to 9 digits)
inline const char* parse_time_ns(const char* s, const char* e, int64_t* out) {
  int64_t sec = 0;
  while (s < e && *s >= '0' && *s <= '9') { sec = sec * 10 + (*s - '0'); s++; }
  int64_t frac = 0; int digits = 0;
  if (s < e && *s == '.') {
    s++;
    while (s < e && *s >= '0' && *s <= '9') { if (digits < 9) { frac = frac * 10 + (*s - '0'); digits++; } s++; }
  }
  while (digits < 9) { frac *= 10; digits++; }
  *out = sec * 1000000000LL + frac;
  return s;
}
}  // namespace

extern "C" {

int64_t lobingest_count_lines(const char* path) {
  Map m;
  if (!m.open(path)) return -1;
  int64_t n = 0;
  const char* s = m.p; const char* e = m.p + m.n;
  while (s < e) { const char* nl = (const char*)memchr(s, '\n', (size_t)(e - s)); if (!nl) { n++; break; } n++; s = nl + 1; }
  return n;
}

// columns 0-5 of a LOBSTER message file: time, type, order id, size, price, direction (populate_database.py:71-78)
int lobingest_parse_messages(const char* path, int64_t max_rows, int64_t* time_ns, int32_t* type, int64_t* order_id, int64_t* size,
                             int64_t* price, int32_t* direction, int64_t* n_out) {
  Map m;
  if (!m.open(path)) return -1;
  const char* s = m.p; const char* e = m.p + m.n;
  int64_t n = 0;
  while (s < e && n < max_rows) {
    const char* nl = (const char*)memchr(s, '\n', (size_t)(e - s));
    const char* le = nl ? nl : e;
    if (le > s && !(le - s == 1 && *s == '\r')) {
      int64_t v;
      const char* q = parse_time_ns(s, le, &time_ns[n]);
      if (q >= le || *q != ',') return -2;
      q = parse_i64(q + 1, le, &v); type[n] = (int32_t)v; if (q >= le || *q != ',') return -2;
      q = parse_i64(q + 1, le, &order_id[n]); if (q >= le || *q != ',') return -2;
      q = parse_i64(q + 1, le, &size[n]); if (q >= le || *q != ',') return -2;
      q = parse_i64(q + 1, le, &price[n]); if (q >= le || *q != ',') return -2;
      q = parse_i64(q + 1, le, &v); direction[n] = (int32_t)v;
      n++;
    }
    if (!nl) break;
    s = nl + 1;
  }
  *n_out = n;
  return 0;
}

// rows `row_idx` (ascending, duplicates allowed) of an orderbook file with n_cols integer columns
int lobingest_parse_book_rows(const char* path, const int64_t* row_idx, int64_t n_rows, int32_t n_cols, int64_t* out) {
  Map m;
  if (!m.open(path)) return -1;
  const char* s = m.p; const char* e = m.p + m.n;
  int64_t line = 0, k = 0;
  while (s < e && k < n_rows) {
    const char* nl = (const char*)memchr(s, '\n', (size_t)(e - s));
    const char* le = nl ? nl : e;
    if (line == row_idx[k]) {
      const char* q = s;
      for (int c = 0; c < n_cols; c++) {
        q = parse_i64(q, le, &out[k * n_cols + c]);
        if (c + 1 < n_cols) { if (q >= le || *q != ',') return -2; q++; }
      }
      k++;
      while (k < n_rows && row_idx[k] == line) { memcpy(&out[k * n_cols], &out[(k - 1) * n_cols], sizeof(int64_t) * (size_t)n_cols); k++; }
    }
    line++;
    if (!nl) break;
    s = nl + 1;
  }
  return k == n_rows ? 0 : -3;
}

}  // extern "C"
