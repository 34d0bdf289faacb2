/* lobingest.h -- C ABI of libingest.so (rl4mm_b200/csrc/lobster_ingest.cpp): LOBSTER message / orderbook CSV files -> the packed,
 * device-ready stream buffers that lobsim_load_stream (lobsim.h) takes.  Host code only (no CUDA).
 *
 * Replaces, for the hot path's data feed, the pandas / SQLAlchemy / Postgres ingest of the reference:
 *   rl4mm/database/populate_database.py:38-95             (file pair -> rows; column layout :71-78)
 *   rl4mm/database/database_population_helpers.py:45-62   (per-second book snapshots), :116-160 (type map, direction flip,
 *                                                          timestamps), :163-181 (row id string = the tie order)
 *   rl4mm/database/HistoricalDatabase.py:46-62,103-119    (range query `start < ts <= end ORDER BY timestamp, id`, last snapshot)
 *   rl4mm/simulation/HistoricalOrderGenerator.py:49-57    (hidden executions dropped, cross trades rejected)
 * The numpy restatement rl4mm_b200/packing.py::pack_arrays produces bit-identical buffers (tests/test_abi_cpu.py).
 *
 * Return codes: 0 ok; -1 cannot open a file; -2 malformed row (wrong field count, empty / non-numeric field); -3 an orderbook row
 * that the snapshots need is missing; -10..-19 semantic errors (see lobingest_pack_open's err_out text).
 * LOBINGEST_THREADS overrides the number of scanning threads (default: hardware concurrency, at most 32).               */
#ifndef LOBINGEST_H
#define LOBINGEST_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* physical lines / data rows (lines that are not blank or "\r"-only: both readers skip those without counting them) */
int64_t lobingest_count_lines(const char* path);
int64_t lobingest_count_rows(const char* path);

/* columns 0-5 of a message file (time as exact integer nanoseconds, type, order id, size, price, direction); at most max_rows */
int lobingest_parse_messages(const char* path, int64_t max_rows, int64_t* time_ns, int32_t* type, int64_t* order_id, int64_t* size,
                             int64_t* price, int32_t* direction, int64_t* n_out);

/* rows `row_idx` (ascending, duplicates allowed) of an orderbook file with n_cols integer columns -> out[n_rows][n_cols] */
int lobingest_parse_book_rows(const char* path, const int64_t* row_idx, int64_t n_rows, int32_t n_cols, int64_t* out);

/* The whole packer.  open = parse + pack (t0_us < 0: the whole second at or before the first message; max_rows < 0: all rows;
 * tie_reference != 0: same-microsecond ties in the reference's lexicographic row-id order, else file order); returns NULL on
 * failure with the code in *rc_out and a message in err_out.  Then: sizes[6] = {n_msgs, n_grid_steps, n_seconds, n_ext_ids,
 * t0_us, n_rows}; copy into caller buffers msgs[n_msgs] (lobsim_msg_t), step_off[n_grid_steps + 1], snapshots[n_seconds + 1][2][L][2],
 * snap_valid[n_seconds + 1], ext_ids[n_ext_ids] (ref -> original order id); close.                                          */
void* lobingest_pack_open(const char* msg_csv, const char* book_csv, int32_t n_levels, int64_t step_us, int64_t t0_us,
                          int32_t tie_reference, int64_t db_batch_size, int64_t max_rows, int32_t* rc_out, char* err_out, int32_t err_cap);
void lobingest_pack_sizes(void* h, int64_t* sizes);
void lobingest_pack_copy(void* h, void* msgs, uint32_t* step_off, int32_t* snapshots, uint8_t* snap_valid, int64_t* ext_ids);
void lobingest_pack_close(void* h);

#ifdef __cplusplus
}
#endif
#endif
