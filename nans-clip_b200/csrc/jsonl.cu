// jsonl.cu — host-side parser for the reference's feature files.
//
// cn_clip/eval/extract_features.py:179-181 / 200-202 write one JSON object per line,
//     {"image_id": 123, "feature": [0.0123, -4.5e-05, ...]}
// and make_topk_predictions.py:57-65 parses them back with json.loads — a few thousand lines per
// second, minutes for a 1M-row gallery in front of a millisecond kernel.  This is the same parse in
// C++: std::from_chars (correctly rounded, like Python's float()) to double, then the float32 cast
// numpy does, so the resulting array is bit-identical to the reference's; lines are independent, so
// they are parsed by a pool of threads.  No GPU involved; nothing here touches CUDA.
#include <charconv>
#include <cstring>
#include <thread>
#include <vector>

#include "common.cuh"

namespace nans {
namespace {

inline const char* skip_ws(const char* p, const char* e) {
  while (p < e && (*p == ' ' || *p == '\t' || *p == '\r')) ++p;
  return p;
}

// first occurrence of "key" (with the quotes) followed by ':' in [p, e); returns the position after ':'
const char* find_key(const char* p, const char* e, const char* key, size_t klen) {
  while (p + klen + 2 < e) {
    const char* q = static_cast<const char*>(memchr(p, '"', static_cast<size_t>(e - p)));
    if (!q || q + klen + 2 > e) return nullptr;
    if (memcmp(q + 1, key, klen) == 0 && q[klen + 1] == '"') {
      const char* r = skip_ws(q + klen + 2, e);
      if (r < e && *r == ':') return r + 1;
    }
    p = q + 1;
  }
  return nullptr;
}

// A JSON number as json.loads accepts it: -?(0|[1-9][0-9]*)(\.[0-9]+)?([eE][+-]?[0-9]+)?.  from_chars is more
// liberal (nan, inf, infinity in any case, "012", ".5", "5."): those tokens are refused here so that the
// caller falls back to json.loads, which either parses them its own way (NaN, Infinity) or raises as the
// reference would.
inline bool json_number_start(const char* p, const char* e) {
  if (p < e && *p == '-') ++p;
  if (p >= e || *p < '0' || *p > '9') return false;
  if (*p == '0' && p + 1 < e && p[1] >= '0' && p[1] <= '9') return false;
  return true;
}
inline bool json_number_tail_ok(const char* b, const char* e) {  // [b, e) = the token from_chars consumed
  for (const char* q = b; q < e; ++q)
    if (*q == '.' && (q + 1 >= e || q[1] < '0' || q[1] > '9')) return false;
  return true;
}

// 0 ok, else a small positive code; *nvals = numbers found in the feature list.  Accepts exactly the
// shape extract_features.py writes — {"<id_key>": int, "feature": [numbers]} — and refuses everything
// else (non-integer ids, non-JSON number tokens, trailing text) so that the json.loads fallback decides.
int parse_line(const char* p, const char* e, const char* id_key, size_t id_klen, int64_t* id, float* feat,
               int64_t cap, int64_t* nvals) {
  if (p >= e || *p != '{') return 9;
  const char* v = find_key(p, e, id_key, id_klen);
  if (!v) return 1;
  v = skip_ws(v, e);
  if (!json_number_start(v, e)) return 2;
  auto ri = std::from_chars(v, e, *id);
  if (ri.ec != std::errc()) return 2;
  {
    // an integer token must end here: 12.5, 12.0 or 1e3 are floats for json.loads, not the int 12 / 1
    const char* a = skip_ws(ri.ptr, e);
    if (a >= e || (*a != ',' && *a != '}')) return 10;
  }
  if (cap < 0) {  // id only (rows outside this rank's gallery shard)
    *nvals = -1;
    return 0;
  }
  const char* f = find_key(p, e, "feature", 7);
  if (!f) return 3;
  f = skip_ws(f, e);
  if (f >= e || *f != '[') return 4;
  ++f;
  int64_t n = 0;
  f = skip_ws(f, e);
  auto closes = [&](const char* q) {  // after ']': either more members (',') or '}' and nothing else
    q = skip_ws(q + 1, e);
    if (q >= e) return false;
    if (*q == ',') return true;
    return *q == '}' && skip_ws(q + 1, e) == e;
  };
  if (f < e && *f == ']') {
    *nvals = 0;
    return closes(f) ? 0 : 11;
  }
  for (;;) {
    f = skip_ws(f, e);
    if (!json_number_start(f, e)) return 5;
    double d;
    auto rf = std::from_chars(f, e, d);  // JSON numbers are a subset of what from_chars accepts
    if (rf.ec != std::errc() || !json_number_tail_ok(f, rf.ptr)) return 5;
    if (n < cap) feat[n] = static_cast<float>(d);
    ++n;
    f = skip_ws(rf.ptr, e);
    if (f >= e) return 6;
    if (*f == ',') {
      ++f;
      continue;
    }
    if (*f == ']') break;
    return 7;
  }
  if (!closes(f)) return 11;
  *nvals = n;
  return 0;
}

}  // namespace
}  // namespace nans

using namespace nans;

// Counts the non-blank lines of a buffer (rows of the feature matrix) and the length of the first
// line's feature list.
extern "C" int nans_jsonl_scan(const char* buf, int64_t len, const char* id_key, int64_t* rows, int64_t* D) {
  if (!buf || len < 0 || !id_key || !rows || !D) {
    set_error("jsonl_scan: null pointer");
    return NANS_ERR_ARG;
  }
  int64_t n = 0, d = -1;
  const char* p = buf;
  const char* end = buf + len;
  const size_t klen = strlen(id_key);
  while (p < end) {
    const char* nl = static_cast<const char*>(memchr(p, '\n', static_cast<size_t>(end - p)));
    const char* e = nl ? nl : end;
    const char* q = skip_ws(p, e);
    if (q < e) {
      if (d < 0) {
        int64_t id, nv = 0;
        const int rc = parse_line(q, e, id_key, klen, &id, nullptr, 0, &nv);
        if (rc != 0) {
          set_error("jsonl_scan: line 1 is not {\"%s\": int, \"feature\": [...]} (code %d)", id_key, rc);
          return NANS_ERR_ARG;
        }
        d = nv;
      }
      ++n;
    }
    p = e + 1;
  }
  *rows = n;
  *D = d < 0 ? 0 : d;
  return NANS_OK;
}

// Parses every non-blank line into ids[rows] / feats[rows, D] (row-major float32).  rows and D must be
// what nans_jsonl_scan reported; a line with a different feature length or a malformed line fails
// (the caller falls back to a general JSON parser).  n_threads <= 0: hardware concurrency.
extern "C" int nans_jsonl_parse(const char* buf, int64_t len, const char* id_key, int64_t rows, int64_t D,
                                int64_t* ids, float* feats, int n_threads) {
  return nans_jsonl_parse_rows(buf, len, id_key, rows, D, 0, rows, ids, feats, n_threads);
}

// The same, with the feature lists parsed only for lines [row_begin, row_begin + row_count) — one rank's
// gallery shard — into feats[row_count, D]; ids[rows] still covers every line (rank 0 writes the ids of
// all shards).  The other lines cost a quote scan and one integer.
extern "C" int nans_jsonl_parse_rows(const char* buf, int64_t len, const char* id_key, int64_t rows, int64_t D,
                                     int64_t row_begin, int64_t row_count, int64_t* ids, float* feats,
                                     int n_threads) {
  if (!buf || len < 0 || !id_key || rows < 0 || D < 0 || row_begin < 0 || row_count < 0 ||
      row_begin + row_count > rows || (rows > 0 && !ids) || (row_count > 0 && D > 0 && !feats)) {
    set_error("jsonl_parse: bad arguments");
    return NANS_ERR_ARG;
  }
  std::vector<std::pair<const char*, const char*>> lines;
  lines.reserve(static_cast<size_t>(rows));
  const char* p = buf;
  const char* end = buf + len;
  while (p < end) {
    const char* nl = static_cast<const char*>(memchr(p, '\n', static_cast<size_t>(end - p)));
    const char* e = nl ? nl : end;
    const char* q = skip_ws(p, e);
    if (q < e) lines.emplace_back(q, e);
    p = e + 1;
  }
  if (static_cast<int64_t>(lines.size()) != rows) {
    set_error("jsonl_parse: %zu lines, expected %lld", lines.size(), (long long)rows);
    return NANS_ERR_ARG;
  }
  int nt = n_threads > 0 ? n_threads : static_cast<int>(std::thread::hardware_concurrency());
  if (nt < 1) nt = 1;
  if (nt > 64) nt = 64;
  if (static_cast<int64_t>(nt) > rows) nt = rows > 0 ? static_cast<int>(rows) : 1;
  const size_t klen = strlen(id_key);
  std::vector<int64_t> bad(static_cast<size_t>(nt), -1);
  std::vector<int> code(static_cast<size_t>(nt), 0);
  auto work = [&](int t) {
    const int64_t lo = rows * t / nt, hi = rows * (t + 1) / nt;
    for (int64_t r = lo; r < hi; ++r) {
      int64_t nv = 0;
      const bool mine = r >= row_begin && r < row_begin + row_count;
      const int rc = parse_line(lines[static_cast<size_t>(r)].first, lines[static_cast<size_t>(r)].second, id_key, klen,
                                ids + r, mine ? feats + (r - row_begin) * D : nullptr, mine ? D : -1, &nv);
      if (rc != 0 || (mine && nv != D)) {
        bad[static_cast<size_t>(t)] = r;
        code[static_cast<size_t>(t)] = rc != 0 ? rc : 8;
        return;
      }
    }
  };
  std::vector<std::thread> pool;
  for (int t = 1; t < nt; ++t) pool.emplace_back(work, t);
  work(0);
  for (auto& th : pool) th.join();
  for (int t = 0; t < nt; ++t)
    if (bad[static_cast<size_t>(t)] >= 0) {
      set_error("jsonl_parse: line %lld is malformed or has a different feature length (code %d)",
                (long long)bad[static_cast<size_t>(t)] + 1, code[static_cast<size_t>(t)]);
      return NANS_ERR_ARG;
    }
  return NANS_OK;
}
