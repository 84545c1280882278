// encoder.cc -- batch query-line encoder (scope row f-1): the host-side step that feeds X_test to the GPU.
//
// Replaces the per-line Python loop of Estimator.predict ("TODO :: parallel encoding",
// neuroestimator/estimator/estimator.py:43-50) with a multi-threaded C++ parser that produces the same
// float64 feature rows, bit for bit, as the reference's
//   NNGPEncoder.parse_line_without_card_then_encode / parse_line / transform_to_1d_array / join_encoding
//                                                    (neuroestimator/estimator/encoder.py:187-250)
//   Table.parse_predicates / predicate_encoding / _factorized_encoding   (encoder.py:58-112)
//   GeneralQuerySampler.parse_line / transform_to_1d_array               (QuerySampler.py:157-221)
// Wire format: neuroestimator/README.md:35-48.  Host C++ only -- no CUDA in this file.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "../../include/nngp_b200.h"

namespace {

struct Column {
  std::string name;
  bool categorical = false;
  double lo = 0.0, denom = 1.0;  // numerical: (v - lo) / denom * 1000
  int64_t ncat = 0;              // categorical: number of codes
  int start = 0, end = 0;        // address inside the table's feature slice
};
struct Table {
  std::string name;
  std::vector<Column> cols;
  std::unordered_map<std::string, int> col_index;
  int feat_dim = 0, offset = 0;  // slice [offset, offset + feat_dim) of the row
};
struct JoinTriple {
  int t1, t2;
  std::string col;
};

thread_local std::string g_enc_error;

struct SV {  // string view
  const char* p;
  size_t n;
};
inline SV strip(SV s) {
  while (s.n && (s.p[0] == ' ' || s.p[0] == '\t' || s.p[0] == '\r' || s.p[0] == '\n')) { ++s.p; --s.n; }
  while (s.n && (s.p[s.n - 1] == ' ' || s.p[s.n - 1] == '\t' || s.p[s.n - 1] == '\r' || s.p[s.n - 1] == '\n')) --s.n;
  return s;
}
inline void split(SV s, char sep, std::vector<SV>& out) {
  out.clear();
  size_t b = 0;
  for (size_t i = 0; i <= s.n; ++i)
    if (i == s.n || s.p[i] == sep) { out.push_back({s.p + b, i - b}); b = i + 1; }
}
inline std::string str(SV s) { return std::string(s.p, s.n); }
inline bool to_double(SV s, double* v) {  // Python float(str): strtod on the stripped token
  s = strip(s);
  if (!s.n || s.n > 63) return false;
  char buf[64];
  memcpy(buf, s.p, s.n);
  buf[s.n] = 0;
  char* end = nullptr;
  *v = strtod(buf, &end);
  return end == buf + s.n;
}
inline bool to_int(SV s, long long* v) {
  s = strip(s);
  if (!s.n || s.n > 31) return false;
  char buf[32];
  memcpy(buf, s.p, s.n);
  buf[s.n] = 0;
  char* end = nullptr;
  *v = strtoll(buf, &end, 10);
  return end == buf + s.n;
}

}  // namespace

struct nngp_encoder {
  int chunk = 64;
  std::vector<Table> tables;
  std::unordered_map<std::string, int> table_index;
  std::vector<JoinTriple> joins;
  int join_offset = 0, dim = 0;
  std::vector<double> defaults;  // the row of a query without predicates and joins
  std::string err;

  bool encode_preds(const Table& t, SV preds, double* row, std::string& e, std::vector<SV>& a, std::vector<SV>& b) const {
    preds = strip(preds);
    if (!preds.n) return true;  // Table.parse_predicates: empty string -> no predicates
    split(preds, '#', a);
    for (SV pred : a) {
      split(pred, ',', b);
      auto it = t.col_index.find(str(strip(b[0])));
      if (it == t.col_index.end()) { e = "unknown column '" + str(strip(b[0])) + "' in table '" + t.name + "'"; return false; }
      const Column& c = t.cols[it->second];
      double* x = row + t.offset;
      if (c.categorical) {  // _factorized_encoding: bit `cat` of an (end-start)*chunk bit string, chunk-wise int(code, 2)
        const int ndim = c.end - c.start;
        std::vector<uint64_t> words(ndim, 0);
        for (size_t i = 1; i < b.size(); ++i) {
          long long cat;
          if (!to_int(b[i], &cat) || cat < 0 || cat >= (long long)ndim * chunk) { e = "bad category in '" + str(pred) + "'"; return false; }
          words[cat / chunk] |= 1ull << (chunk - 1 - (cat % chunk));  // position 0 of a chunk is its MSB
        }
        for (int d = 0; d < ndim; ++d) x[c.start + d] = (double)words[d];
      } else {
        double up, lo;
        if (b.size() < 3 || !to_double(b[1], &up) || !to_double(b[2], &lo)) { e = "bad numerical predicate '" + str(pred) + "'"; return false; }
        x[c.start] = (up - c.lo) / c.denom * 1000;
        x[c.start + 1] = (lo - c.lo) / c.denom * 1000;
      }
    }
    return true;
  }

  // format 0: "t1,t2@preds1@preds2@joins"   1: same + "@card"   2: single table "preds@card"
  bool encode_line(SV line, int format, double* row, double* card, std::string& e) const {
    std::vector<SV> terms, a, b, names;
    memcpy(row, defaults.data(), sizeof(double) * dim);
    line = strip(line);
    split(line, '@', terms);
    if (format == 2) {
      if (terms.size() != 2) { e = "expected 'preds@card'"; return false; }
      long long c;
      if (!to_int(terms[1], &c)) { e = "bad cardinality"; return false; }
      if (card) *card = (double)c;
      return encode_preds(tables[0], terms[0], row, e, a, b);
    }
    const size_t extra = format == 1 ? 3 : 2;
    split(strip(terms[0]), ',', names);
    if (names.size() + extra != terms.size()) { e = "Query Format Error!"; return false; }
    for (size_t i = 0; i < names.size(); ++i) {
      auto it = table_index.find(str(names[i]));
      if (it == table_index.end()) { e = "unknown table '" + str(names[i]) + "'"; return false; }
      if (!encode_preds(tables[it->second], terms[1 + i], row, e, a, b)) return false;
    }
    SV joins_s = strip(terms[names.size() + 1]);
    if (joins_s.n) {
      split(joins_s, '#', a);
      for (SV j : a) {
        split(j, ',', b);
        if (b.size() < 3) { e = "bad join '" + str(j) + "'"; return false; }
        auto i1 = table_index.find(str(strip(b[0]))), i2 = table_index.find(str(strip(b[1])));
        if (i1 == table_index.end() || i2 == table_index.end()) { e = "unknown table in join '" + str(j) + "'"; return false; }
        const std::string col = str(strip(b[2]));
        if (!tables[i1->second].col_index.count(col)) { e = "unknown join column '" + col + "'"; return false; }
        const int t1 = std::min(i1->second, i2->second), t2 = std::max(i1->second, i2->second);
        int idx = -1;
        for (size_t k = 0; k < joins.size(); ++k)
          if (joins[k].t1 == t1 && joins[k].t2 == t2 && joins[k].col == col) { idx = (int)k; break; }
        if (idx < 0) { e = "join '" + str(j) + "' is not in the schema's join list"; return false; }
        row[join_offset + idx * 3 + 2] = 1.0;  // join_encoding: op is always '=', join_ops_dict['='] == 2
      }
    }
    if (format == 1) {
      long long c;
      if (!to_int(terms.back(), &c)) { e = "bad cardinality"; return false; }
      if (card) *card = (double)c;
    }
    return true;
  }
};

extern "C" {

const char* nngp_encoder_last_error(const nngp_encoder* enc) { return enc ? enc->err.c_str() : g_enc_error.c_str(); }

int nngp_encoder_create(const char* schema_text, nngp_encoder** out) {
  if (!schema_text || !out) { g_enc_error = "nngp_encoder_create: null argument"; return NNGP_EINVAL; }
  *out = nullptr;
  nngp_encoder* enc = new nngp_encoder();
  std::vector<SV> lines, tok;
  split({schema_text, strlen(schema_text)}, '\n', lines);
  auto bad = [&](const std::string& m) { g_enc_error = "nngp_encoder_create: " + m; delete enc; return (int)NNGP_EINVAL; };
  for (SV ln : lines) {
    ln = strip(ln);
    if (!ln.n || ln.p[0] == '#') continue;
    split(ln, ' ', tok);
    const std::string kind = str(tok[0]);
    if (kind == "chunk_size" && tok.size() == 2) {
      long long v;
      if (!to_int(tok[1], &v) || v < 1 || v > 64) return bad("chunk_size must be in 1..64");
      enc->chunk = (int)v;
    } else if (kind == "table" && tok.size() == 2) {
      Table t;
      t.name = str(tok[1]);
      enc->table_index[t.name] = (int)enc->tables.size();
      enc->tables.push_back(t);
    } else if (kind == "col" && tok.size() >= 4 && !enc->tables.empty()) {
      Table& t = enc->tables.back();
      Column c;
      c.name = str(tok[1]);
      const std::string ty = str(tok[2]);
      if (ty == "num" && tok.size() == 5) {
        if (!to_double(tok[3], &c.lo) || !to_double(tok[4], &c.denom)) return bad("bad numerical column '" + str(ln) + "'");
        c.start = t.feat_dim; c.end = t.feat_dim + 2;
      } else if (ty == "cat" && tok.size() == 4) {
        long long n;
        if (!to_int(tok[3], &n) || n < 0) return bad("bad categorical column '" + str(ln) + "'");
        c.categorical = true; c.ncat = n;
        c.start = t.feat_dim;
        c.end = t.feat_dim + (int)((n + enc->chunk - 1) / enc->chunk);  // math.ceil(num_cat / chunk_size)
      } else {
        return bad("bad column line '" + str(ln) + "'");
      }
      t.feat_dim = c.end;
      t.col_index[c.name] = (int)t.cols.size();
      t.cols.push_back(c);
    } else if (kind == "join" && tok.size() == 4) {
      long long a, b;
      if (!to_int(tok[1], &a) || !to_int(tok[2], &b)) return bad("bad join line '" + str(ln) + "'");
      enc->joins.push_back({(int)a, (int)b, str(tok[3])});
    } else {
      return bad("unrecognised schema line '" + str(ln) + "'");
    }
  }
  if (enc->tables.empty()) return bad("schema has no table");
  int off = 0;
  for (Table& t : enc->tables) { t.offset = off; off += t.feat_dim; }
  enc->join_offset = off;
  enc->dim = off + 3 * (int)enc->joins.size();  // join_feat_dim = total_num_joins * len({'>','<','='})
  enc->defaults.assign(enc->dim, 0.0);
  for (const Table& t : enc->tables)
    for (const Column& c : t.cols)
      if (!c.categorical) enc->defaults[t.offset + c.start + 1] = 1000.0;  // predicate_encoding: no predicate -> (0, 1000)
  *out = enc;
  return NNGP_OK;
}

void nngp_encoder_destroy(nngp_encoder* enc) { delete enc; }

int nngp_encoder_dim(const nngp_encoder* enc) { return enc ? enc->dim : NNGP_EINVAL; }

int nngp_encode_lines(nngp_encoder* enc, const char* blob, int64_t blob_len, int64_t n_lines, int32_t format,
                      double* x_out, double* card_out_or_null, int32_t n_threads) {
  if (!enc) return NNGP_EINVAL;
  if (!blob || !x_out || n_lines < 0 || blob_len < 0 || format < 0 || format > 2) { enc->err = "nngp_encode_lines: bad argument"; return NNGP_EINVAL; }
  if (format == 2 && enc->tables.size() != 1) { enc->err = "nngp_encode_lines: format 2 needs a single-table schema"; return NNGP_EINVAL; }
  std::vector<int64_t> starts;
  starts.reserve(n_lines + 1);
  int64_t pos = 0;
  while (pos <= blob_len && (int64_t)starts.size() < n_lines) {
    starts.push_back(pos);
    const void* nl = pos < blob_len ? memchr(blob + pos, '\n', blob_len - pos) : nullptr;
    pos = nl ? (const char*)nl - blob + 1 : blob_len + 1;
  }
  if ((int64_t)starts.size() != n_lines) { enc->err = "nngp_encode_lines: fewer lines in the buffer than n_lines"; return NNGP_EINVAL; }
  starts.push_back(pos);
  int nt = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
  if (nt < 1) nt = 1;
  if (n_lines < 256) nt = 1;
  nt = (int)std::min<int64_t>(nt, std::max<int64_t>(1, n_lines));
  std::vector<std::string> errs(nt);
  std::vector<int64_t> bad_line(nt, -1);
  auto work = [&](int tid) {
    const int64_t lo = n_lines * tid / nt, hi = n_lines * (tid + 1) / nt;
    std::string e;
    for (int64_t i = lo; i < hi; ++i) {
      int64_t len = starts[i + 1] - starts[i] - 1;
      if (len < 0) len = 0;
      if (!enc->encode_line({blob + starts[i], (size_t)len}, format, x_out + i * enc->dim,
                            card_out_or_null ? card_out_or_null + i : nullptr, e)) {
        errs[tid] = e;
        bad_line[tid] = i;
        return;
      }
    }
  };
  if (nt == 1) {
    work(0);
  } else {
    std::vector<std::thread> th;
    for (int t = 0; t < nt; ++t) th.emplace_back(work, t);
    for (auto& t : th) t.join();
  }
  for (int t = 0; t < nt; ++t)
    if (bad_line[t] >= 0) {
      enc->err = "nngp_encode_lines: line " + std::to_string(bad_line[t]) + ": " + errs[t];
      return NNGP_EINVAL;
    }
  return NNGP_OK;
}

}  // extern "C"
