// lobster_ingest.cpp -- fast host-side reader of LOBSTER CSV files for the packer (rl4mm_b200/packing.py).
//
// Replaces, for the hot path's data feed, the pandas / SQLAlchemy ingest of rl4mm/database/populate_database.py:38-95 and
// database_population_helpers.py:45-193: the message file is parsed column-wise in one pass
// over a memory map (time as exact integer nanoseconds, no float round trip), and of the orderbook file -- which has
// one row per message -- only the rows actually needed for the per-second snapshots are parsed.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <algorithm>
#include <cstdlib>
#include <string>
#include <thread>
#include <chrono>
#include <vector>
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

// an integer field; an empty field (no digit) is malformed: returns nullptr
inline const char* parse_i64(const char* s, const char* e, int64_t* out) {
  bool neg = false;
  if (s < e && (*s == '-' || *s == '+')) { neg = *s == '-'; s++; }
  int64_t v = 0;
  const char* d0 = s;
  while (s < e && *s >= '0' && *s <= '9') { v = v * 10 + (*s - '0'); s++; }
  if (s == d0) return nullptr;
  *out = neg ? -v : v;
  return s;
}
inline bool blank_line(const char* s, const char* le) { return le == s || (le - s == 1 && *s == '\r'); }
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
// ---- chunk-parallel scanning: the memory map is cut into `parts` byte ranges that start at line starts; every reader first counts
//      the data rows of each range (prefix sum = the row index each range starts at), then works on the ranges independently ------
int ingest_threads() {
  const char* e = getenv("LOBINGEST_THREADS");
  int t = e ? atoi(e) : (int)std::thread::hardware_concurrency();
  return t < 1 ? 1 : (t > 32 ? 32 : t);
}
template <class F>
void parallel_parts(int parts, F f) {
  if (parts <= 1) { f(0); return; }
  std::vector<std::thread> th;
  for (int i = 1; i < parts; i++) th.emplace_back([&f, i] { f(i); });
  f(0);
  for (auto& t : th) t.join();
}
struct Ranges {
  std::vector<const char*> cut;       // parts + 1 boundaries, each at a line start (or the end of the map)
  std::vector<int64_t> row0;          // parts + 1: index of the first data row of each range; row0[parts] = total
  int parts() const { return (int)cut.size() - 1; }
};
inline int64_t count_rows_range(const char* s, const char* e) {
  int64_t n = 0;
  while (s < e) {
    const char* nl = (const char*)memchr(s, '\n', (size_t)(e - s));
    const char* le = nl ? nl : e;
    if (!blank_line(s, le)) n++;
    if (!nl) break;
    s = nl + 1;
  }
  return n;
}
Ranges split_rows(const Map& m) {
  Ranges r;
  int parts = ingest_threads();
  if (m.n < (size_t)(1 << 20)) parts = 1;
  const char* b = m.p; const char* e = m.p + m.n;
  r.cut.push_back(b);
  for (int i = 1; i < parts; i++) {
    const char* c = b + m.n / (size_t)parts * (size_t)i;
    if (c <= r.cut.back()) continue;
    const char* nl = (const char*)memchr(c, '\n', (size_t)(e - c));
    c = nl ? nl + 1 : e;
    if (c > r.cut.back() && c < e) r.cut.push_back(c);
  }
  r.cut.push_back(e);
  const int np = r.parts();
  std::vector<int64_t> cnt((size_t)np);
  parallel_parts(np, [&](int i) { cnt[(size_t)i] = count_rows_range(r.cut[(size_t)i], r.cut[(size_t)i + 1]); });
  r.row0.assign((size_t)np + 1, 0);
  for (int i = 0; i < np; i++) r.row0[(size_t)i + 1] = r.row0[(size_t)i] + cnt[(size_t)i];
  return r;
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

// data rows = lines that are not blank (both readers below skip blank and "\r"-only lines WITHOUT counting them, so row k of the
// message file and row k of the orderbook file stay aligned)
int64_t lobingest_count_rows(const char* path) {
  Map m;
  if (!m.open(path)) return -1;
  return split_rows(m).row0.back();
}

// columns 0-5 of a LOBSTER message file: time, type, order id, size, price, direction (populate_database.py:71-78)
static int parse_messages_ranges(const Ranges& r, int64_t max_rows, int64_t* time_ns, int32_t* type, int64_t* order_id, int64_t* size,
                                 int64_t* price, int32_t* direction, int64_t* n_out) {
  std::vector<int> rcs((size_t)r.parts(), 0);
  parallel_parts(r.parts(), [&](int part) {
    const char* s = r.cut[(size_t)part]; const char* e = r.cut[(size_t)part + 1];
    int64_t n = r.row0[(size_t)part];
    while (s < e && n < max_rows) {
      const char* nl = (const char*)memchr(s, '\n', (size_t)(e - s));
      const char* le = nl ? nl : e;
      if (!blank_line(s, le)) {
        int64_t v;
        const char* q = parse_time_ns(s, le, &time_ns[n]);
        if (q == s || q >= le || *q != ',') { rcs[(size_t)part] = -2; return; }
        q = parse_i64(q + 1, le, &v); if (!q || q >= le || *q != ',') { rcs[(size_t)part] = -2; return; } type[n] = (int32_t)v;
        q = parse_i64(q + 1, le, &order_id[n]); if (!q || q >= le || *q != ',') { rcs[(size_t)part] = -2; return; }
        q = parse_i64(q + 1, le, &size[n]); if (!q || q >= le || *q != ',') { rcs[(size_t)part] = -2; return; }
        q = parse_i64(q + 1, le, &price[n]); if (!q || q >= le || *q != ',') { rcs[(size_t)part] = -2; return; }
        q = parse_i64(q + 1, le, &v); if (!q) { rcs[(size_t)part] = -2; return; } direction[n] = (int32_t)v;
        n++;
      }
      if (!nl) break;
      s = nl + 1;
    }
  });
  for (int rc : rcs) if (rc) return rc;
  *n_out = r.row0.back() < max_rows ? r.row0.back() : max_rows;
  return 0;
}
int lobingest_parse_messages(const char* path, int64_t max_rows, int64_t* time_ns, int32_t* type, int64_t* order_id, int64_t* size,
                             int64_t* price, int32_t* direction, int64_t* n_out) {
  Map m;
  if (!m.open(path)) return -1;
  return parse_messages_ranges(split_rows(m), max_rows, time_ns, type, order_id, size, price, direction, n_out);
}

// rows `row_idx` (ascending, duplicates allowed) of an orderbook file with n_cols integer columns
static int parse_book_rows_ranges(const Ranges& r, const int64_t* row_idx, int64_t n_rows, int32_t n_cols, int64_t* out) {
  if (n_rows > 0 && row_idx[n_rows - 1] >= r.row0.back()) return -3;
  std::vector<int> rcs((size_t)r.parts(), 0);
  parallel_parts(r.parts(), [&](int part) {
    const char* s = r.cut[(size_t)part]; const char* e = r.cut[(size_t)part + 1];
    int64_t line = r.row0[(size_t)part];
    const int64_t line_end = r.row0[(size_t)part + 1];
    int64_t k = (int64_t)(std::lower_bound(row_idx, row_idx + n_rows, line) - row_idx);
    while (s < e && k < n_rows && row_idx[k] < line_end) {
      const char* nl = (const char*)memchr(s, '\n', (size_t)(e - s));
      const char* le = nl ? nl : e;
      if (blank_line(s, le)) { if (!nl) break; s = nl + 1; continue; }   // not a data row (same rule as the message reader)
      if (line == row_idx[k]) {
        const char* q = s;
        for (int c = 0; c < n_cols; c++) {
          q = parse_i64(q, le, &out[k * n_cols + c]);
          if (!q) { rcs[(size_t)part] = -2; return; }
          if (c + 1 < n_cols) { if (q >= le || *q != ',') { rcs[(size_t)part] = -2; return; } q++; }
        }
        k++;
        while (k < n_rows && row_idx[k] == line) { memcpy(&out[k * n_cols], &out[(k - 1) * n_cols], sizeof(int64_t) * (size_t)n_cols); k++; }
      }
      line++;
      if (!nl) break;
      s = nl + 1;
    }
  });
  for (int rc : rcs) if (rc) return rc;
  return 0;
}
int lobingest_parse_book_rows(const char* path, const int64_t* row_idx, int64_t n_rows, int32_t n_cols, int64_t* out) {
  Map m;
  if (!m.open(path)) return -1;
  return parse_book_rows_ranges(split_rows(m), row_idx, n_rows, n_cols, out);
}

// ====================================================================================================================
//  The packer: LOBSTER message + orderbook files -> the device-ready stream buffers, in one native pass.
//  (rl4mm/database/populate_database.py:38-95 + database_population_helpers.py:45-62,116-193 + the query semantics of
//  HistoricalDatabase.py:46-62,103-119 and HistoricalOrderGenerator.py:49-57; the numpy restatement in packing.py::pack_arrays
//  produces bit-identical buffers and is kept as the cross-check.)
//    * type map: 1 limit, 2 cancellation, 3 deletion, 4 market (visible execution), 5 market_hidden (dropped), 6 cross_trade /
//      7 trading_halt (rejected) -- database_population_helpers.py:151-160, HistoricalOrderGenerator.py:52-57;
//    * direction +1 buy / -1 sell, flipped for executions (:132-136);  timestamps truncated to microseconds (models.py:14);
//    * replay order = ORDER BY (timestamp, id) with the STRING id "..._{row}" => same-microsecond ties ordered
//      lexicographically by the decimal row number (row + (row // batch) * batch, :163-181);
//    * messages with ts <= t0 are never replayed (range query start < ts);  CSR offsets per step (t0 + k*step, t0 + (k+1)*step];
//    * snapshot at second T = orderbook row of the last message with time <= T (:45-62,139-148); LOBSTER dummy prices
//      (+-9999999999) => no level.
// ====================================================================================================================
namespace {
struct PackedMsg { int32_t price, volume; uint32_t ref, meta; };
struct Pack {
  std::vector<PackedMsg> msgs;
  std::vector<uint32_t> step_off;
  std::vector<int32_t> snapshots;      // [n_seconds + 1][2][L][2]
  std::vector<uint8_t> snap_valid;
  std::vector<int64_t> ext_ids;        // ref -> original order id (ext_ids[0] = 0: the snapshot aggregate)
  int64_t t0_us = 0, n_grid = 0, n_seconds = 0, n_rows = 0;
  std::string err;
};
const int64_t kDummy = 9999999999LL;
const int32_t kNoPrice = INT32_MIN;

inline int dec_digits(int64_t v, char* buf) { int n = 0; char t[24]; do { t[n++] = (char)('0' + v % 10); v /= 10; } while (v); for (int i = 0; i < n; i++) buf[i] = t[n - 1 - i]; return n; }
inline bool dec_less(int64_t a, int64_t b) {   // str(a) < str(b)
  char sa[24], sb[24];
  const int na = dec_digits(a, sa), nb = dec_digits(b, sb);
  const int c = memcmp(sa, sb, (size_t)(na < nb ? na : nb));
  return c < 0 || (c == 0 && na < nb);
}
inline int64_t ceil_div(int64_t a, int64_t b) { return a >= 0 ? (a + b - 1) / b : -((-a) / b); }   // b > 0

int pack_files(Pack& P, const char* msg_csv, const char* book_csv, int n_levels, int64_t step_us, int64_t t0_us_in, int tie_reference,
               int64_t db_batch, int64_t max_rows) {
  if (n_levels <= 0 || step_us <= 0 || 1000000 % step_us) { P.err = "step_us must divide one second"; return -10; }
  if (db_batch <= 0) db_batch = 1000000;
  const bool timing = getenv("LOBINGEST_TIMING") != nullptr;
  auto t_prev = std::chrono::steady_clock::now();
  auto lap = [&](const char* what) {
    if (!timing) return;
    auto t = std::chrono::steady_clock::now();
    fprintf(stderr, "lobingest: %-28s %.3f s\n", what, std::chrono::duration<double>(t - t_prev).count());
    t_prev = t;
  };
  Map mm, mb;                                                   // each file is mapped and split into row ranges ONCE
  if (!mm.open(msg_csv)) { P.err = std::string("cannot open ") + msg_csv; return -1; }
  if (!mb.open(book_csv)) { P.err = std::string("cannot open ") + book_csv; return -1; }
  const Ranges rm = split_rows(mm), rb = split_rows(mb);
  lap("map + split rows (both files)");
  int64_t n = rm.row0.back();
  const bool truncated = max_rows >= 0 && max_rows < n;
  if (truncated) n = max_rows;
  if (n == 0) { P.err = "empty message file"; return -11; }
  std::vector<int64_t> time_ns((size_t)n), oid((size_t)n), size((size_t)n), price((size_t)n);
  std::vector<int32_t> type((size_t)n), dir((size_t)n);
  int64_t got = 0;
  int rc = parse_messages_ranges(rm, n, time_ns.data(), type.data(), oid.data(), size.data(), price.data(), dir.data(), &got);
  if (rc != 0 || got != n) { P.err = "malformed LOBSTER message file"; return -2; }
  lap("parse messages");
  if (!truncated && rb.row0.back() != n) { P.err = "message file and orderbook file have different row counts"; return -12; }
  P.n_rows = n;
  for (int64_t i = 1; i < n; i++) if (time_ns[(size_t)i] < time_ns[(size_t)i - 1]) { P.err = "LOBSTER messages must be time ordered"; return -13; }
  auto ts_us = [&](int64_t i) { return time_ns[(size_t)i] / 1000; };
  int64_t t0 = t0_us_in >= 0 ? t0_us_in : ts_us(0) / 1000000 * 1000000;
  if (t0 % 1000000) { P.err = "t0_us must be a whole second"; return -14; }
  P.t0_us = t0;
  const int64_t last_us = ts_us(n - 1);
  int64_t n_grid = std::max<int64_t>(1, ceil_div(last_us - t0, step_us));
  const int64_t n_seconds = ceil_div(n_grid * step_us, 1000000);
  n_grid = n_seconds * (1000000 / step_us);
  P.n_grid = n_grid; P.n_seconds = n_seconds;

  // ---- replay order: stable by microsecond; ties by the decimal string of the database row id --------------------------------
  std::vector<int64_t> order((size_t)n);
  for (int64_t i = 0; i < n; i++) order[(size_t)i] = i;
  if (tie_reference) {
    auto row_id = [&](int64_t r) { return r + (r / db_batch) * db_batch; };
    for (int64_t a = 0; a < n;) {
      int64_t b = a + 1;
      const int64_t t = ts_us(a);
      while (b < n && ts_us(b) == t) b++;
      if (b - a > 1) std::stable_sort(order.begin() + a, order.begin() + b, [&](int64_t x, int64_t y) { return dec_less(row_id(x), row_id(y)); });
      a = b;
    }
  }
  lap("tie order");
  // ---- keep: inside the replay range, not hidden; reject what the reference rejects ----------------------------------------------
  std::vector<int64_t> keep;
  keep.reserve((size_t)n);
  for (int64_t k = 0; k < n; k++) {
    const int64_t i = order[(size_t)k];
    if (ts_us(i) > t0 && type[(size_t)i] != 5) keep.push_back(i);
  }
  for (int64_t i : keep) {
    const int t = type[(size_t)i];
    if (t == 6 || t == 7) { P.err = "cross_trade / trading_halt messages inside the replay range"; return -15; }
    if (t < 1 || t > 4) { P.err = "unknown LOBSTER message type"; return -16; }
    if (price[(size_t)i] >= (1LL << 31) || size[(size_t)i] >= (1LL << 31)) { P.err = "price / size does not fit int32"; return -17; }
  }
  lap("filter + checks");
  // ---- dense order references: rank among the distinct external ids (np.unique order), + 1 -------------------------------------
  // (threads: chunk sorts + pairwise merges for the distinct ids, then the per-message rank lookups and record fills in parallel)
  const int nthr = keep.size() > (size_t)(1 << 18) ? ingest_threads() : 1;
  auto chunk = [&](int part, size_t total) { return std::pair<size_t, size_t>(total * (size_t)part / (size_t)nthr, total * (size_t)(part + 1) / (size_t)nthr); };
  std::vector<int64_t> uniq(keep.size());
  parallel_parts(nthr, [&](int part) {
    const auto [a, b] = chunk(part, keep.size());
    for (size_t k = a; k < b; k++) uniq[k] = oid[(size_t)keep[k]];
    std::sort(uniq.begin() + (ptrdiff_t)a, uniq.begin() + (ptrdiff_t)b);
  });
  for (int width = 1; width < nthr; width *= 2) {                    // merge sorted runs [part, part + width) pairwise
    const int pairs = (nthr + 2 * width - 1) / (2 * width);
    parallel_parts(pairs, [&](int pr) {
      const int lo = pr * 2 * width, mid = lo + width, hi = std::min(nthr, lo + 2 * width);
      if (mid >= hi) return;
      std::inplace_merge(uniq.begin() + (ptrdiff_t)chunk(lo, keep.size()).first, uniq.begin() + (ptrdiff_t)chunk(mid, keep.size()).first,
                         uniq.begin() + (ptrdiff_t)chunk(hi - 1, keep.size()).second);
    });
  }
  uniq.erase(std::unique(uniq.begin(), uniq.end()), uniq.end());
  if ((int64_t)uniq.size() + 1 >= (1LL << 31)) { P.err = "too many distinct order ids"; return -18; }
  P.ext_ids.assign(1, 0);
  P.ext_ids.insert(P.ext_ids.end(), uniq.begin(), uniq.end());
  P.msgs.resize(keep.size());
  P.step_off.assign((size_t)n_grid + 1, 0u);
  std::vector<int> bad((size_t)nthr, 0);
  std::vector<int64_t> steps(keep.size());
  parallel_parts(nthr, [&](int part) {
    const auto [a, b] = chunk(part, keep.size());
    for (size_t k = a; k < b; k++) {
      const int64_t i = keep[k];
      const int t = type[(size_t)i];
      int side = dir[(size_t)i] == 1 ? 0 : 1;          // +1 buy / otherwise sell
      if (t == 4) side ^= 1;                             // executions carry the direction of the RESTING order: flip to the aggressor's
      PackedMsg& m = P.msgs[k];
      m.price = (int32_t)price[(size_t)i];
      m.volume = (int32_t)size[(size_t)i];
      m.ref = (uint32_t)(std::lower_bound(uniq.begin(), uniq.end(), oid[(size_t)i]) - uniq.begin()) + 1u;
      m.meta = (uint32_t)t | ((uint32_t)side << 3);
      const int64_t step = ceil_div(ts_us(i) - t0, step_us) - 1;       // ts in (t0 + k*step, t0 + (k+1)*step]  =>  k
      if (step < 0 || step >= n_grid) { bad[(size_t)part] = 1; return; }
      steps[k] = step;
    }
  });
  for (int b : bad) if (b) { P.err = "internal: step index out of range"; return -19; }
  for (size_t k = 0; k < keep.size(); k++) P.step_off[(size_t)steps[k] + 1] += 1u;
  for (int64_t k = 0; k < n_grid; k++) P.step_off[(size_t)k + 1] += P.step_off[(size_t)k];

  lap("refs + records + offsets");
  // ---- per-second snapshots: the orderbook row of the last message with time <= T ------------------------------------------------
  const int64_t ns = n_seconds + 1;
  std::vector<int64_t> need((size_t)ns);
  P.snap_valid.assign((size_t)ns, 0);
  for (int64_t k = 0; k < ns; k++) {
    const int64_t sec_ns = (t0 + k * 1000000) * 1000;
    const int64_t idx = (int64_t)(std::upper_bound(time_ns.begin(), time_ns.end(), sec_ns) - time_ns.begin()) - 1;
    P.snap_valid[(size_t)k] = idx >= 0 ? 1 : 0;
    need[(size_t)k] = idx >= 0 ? idx : 0;
  }
  const int ncol = 4 * n_levels;
  std::vector<int64_t> rows((size_t)ns * ncol);
  rc = parse_book_rows_ranges(rb, need.data(), ns, ncol, rows.data());
  if (rc != 0) { P.err = "could not read the snapshot rows of the orderbook file"; return -3; }
  lap("snapshot rows");
  P.snapshots.assign((size_t)ns * 2 * n_levels * 2, 0);
  for (int64_t k = 0; k < ns; k++)
    for (int l = 0; l < n_levels; l++) {
      const int64_t* r = &rows[((size_t)k * n_levels + l) * 4];   // ask price, ask size, bid price, bid size (rl4mm/orderbook/helpers.py:52-55)
      for (int s = 0; s < 2; s++) {
        const int64_t p = s ? r[0] : r[2], v = s ? r[1] : r[3];    // side 1 = sell = ask columns
        const bool dummy = (p < 0 ? -p : p) >= kDummy;
        int32_t* o = &P.snapshots[(((size_t)k * 2 + s) * n_levels + l) * 2];
        o[0] = dummy ? kNoPrice : (int32_t)p;
        o[1] = dummy ? 0 : (int32_t)v;
      }
    }
  return 0;
}
}  // namespace

// open = parse + pack; then read the sizes, allocate, copy, close.  Returns null on failure with the reason in *rc / err_out.
void* lobingest_pack_open(const char* msg_csv, const char* book_csv, int32_t n_levels, int64_t step_us, int64_t t0_us, int32_t tie_reference,
                          int64_t db_batch_size, int64_t max_rows, int32_t* rc_out, char* err_out, int32_t err_cap) {
  Pack* P = new Pack();
  const int rc = pack_files(*P, msg_csv, book_csv, n_levels, step_us, t0_us, tie_reference, db_batch_size, max_rows);
  if (rc_out) *rc_out = rc;
  if (rc != 0) {
    if (err_out && err_cap > 0) { strncpy(err_out, P->err.c_str(), (size_t)err_cap - 1); err_out[err_cap - 1] = 0; }
    delete P;
    return nullptr;
  }
  return P;
}
// sizes[6] = {n_msgs, n_grid_steps, n_seconds, n_ext_ids, t0_us, n_rows}
void lobingest_pack_sizes(void* h, int64_t* sizes) {
  Pack* P = (Pack*)h;
  sizes[0] = (int64_t)P->msgs.size(); sizes[1] = P->n_grid; sizes[2] = P->n_seconds; sizes[3] = (int64_t)P->ext_ids.size();
  sizes[4] = P->t0_us; sizes[5] = P->n_rows;
}
void lobingest_pack_copy(void* h, void* msgs, uint32_t* step_off, int32_t* snapshots, uint8_t* snap_valid, int64_t* ext_ids) {
  Pack* P = (Pack*)h;
  if (!P->msgs.empty()) memcpy(msgs, P->msgs.data(), P->msgs.size() * sizeof(PackedMsg));
  memcpy(step_off, P->step_off.data(), P->step_off.size() * sizeof(uint32_t));
  memcpy(snapshots, P->snapshots.data(), P->snapshots.size() * sizeof(int32_t));
  memcpy(snap_valid, P->snap_valid.data(), P->snap_valid.size());
  memcpy(ext_ids, P->ext_ids.data(), P->ext_ids.size() * sizeof(int64_t));
}
void lobingest_pack_close(void* h) { delete (Pack*)h; }

}  // extern "C"
