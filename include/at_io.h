/* at_io.h — C ABI of libat_io.so: host-side input pipeline of the covid model.
 *
 * Replaces the text parsing inside
 *   agent_torch/models/covid/substeps/utils.py:129-148  network_from_file
 *     (pd.read_csv(header=None) -> header detection -> first two columns -> to_numeric(coerce) -> dropna -> int64)
 * with a memory-mapped, multi-threaded parser that writes the (2, E) int64 `edge_index` directly: a population-scale
 * `mobility_networks/0.csv` (168 M edges, 2.5 GB of text) takes minutes in pandas and seconds here.  Host code only
 * (plain C, pthreads; no CUDA): the edge list then goes to the GPU CSR builder (agenttorch_b200/graph.py).
 *
 * Row semantics (same result as the reference on numeric edge lists):
 *   - rows end with '\n' (an optional '\r' before it is ignored); empty rows are skipped;
 *   - fields are separated by ','; only the first two are used, blanks around a field are ignored;
 *   - a field is numeric if it is a decimal integer or a floating-point literal that is finite (a pair of double quotes around
 *     it is ignored); floats are truncated towards zero like `astype("int64")`; empty fields and pandas' NA tokens ("nan", "NA",
 *     "null", ...) are missing values; anything else is text;
 *   - a row whose first or second field is missing or text is dropped (`to_numeric(errors="coerce")` + `dropna`);
 *   - if ANY field of the file is text — the classic header row, but also a stray token in a later row or column, which makes
 *     pandas read that whole column as strings — the first non-empty row is dropped
 *     (`df.iloc[0].apply(lambda v: isinstance(v, str)).any()` -> `header=0`);
 *   - "inf" tokens are treated as missing (the reference raises IntCastingNaNError on them);
 *   - row order is preserved (the kernels accumulate in edge order).
 * Every function returns 0 on success, non-zero on error (text from at_io_last_error()), never aborts.
 */
#ifndef AT_IO_H
#define AT_IO_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AT_IO_ABI_VERSION 1

int at_io_abi_version(void);
const char* at_io_last_error(void);

/* Pass 1: number of edges the file holds (rows that survive the rules above).  n_threads <= 0: one per online core. */
int at_io_edge_list_count(const char* path, int32_t n_threads, int64_t* n_edges);

/* Pass 2: fill src[0..n) and dst[0..n) (e.g. the two rows of a (2, E) int64 tensor).  `capacity` must be >= the count of
 * pass 1; *n_edges receives the number written.  Also reports the largest id seen (-1 for an empty list) so that the caller
 * can choose int32 storage. */
int at_io_edge_list_parse(const char* path, int32_t n_threads, int64_t capacity, int64_t* src, int64_t* dst, int64_t* n_edges,
                          int64_t* max_id);

#ifdef __cplusplus
}
#endif
#endif
