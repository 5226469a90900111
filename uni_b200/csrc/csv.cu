// csv.cu -- the on-disk format either side of the hot path: `path.loadMatD` /
// `loadSmart` (src/main/scala/uni/io/FileOps.scala:139-188, cell grammar `big(str)` src/main/scala/uni/data/Big.scala:59-64,
// row reader src/main/scala/uni/io/FastCsv.scala:21,35-41) and `m.saveCSV` (src/main/scala/uni/data/Mat.scala:90-116).
// Text is parsed / formatted on the host (all cores) and moved with one copy; the matrix itself lives on the device.
//
// Scope of the reader: one-byte delimiter (given, or the most frequent of , TAB ; | on the first content line --
// the reference's statistical detector io/Delimiter.scala is NOT reproduced), RFC-4180 quotes with "" escapes, fields
// trimmed, rows without content dropped, rows padded to the widest, header = row 0 all non-numeric and row 1 has a
// numeric cell.  A cell is `big(s).toDouble`: trim, drop every ',' and '$', optional trailing '%' (value / 100), then
// java.math.BigDecimal's grammar [+-]digits[.digits][(e|E)[+-]digits]; anything else (including "NaN", "Inf") is NaN.
#include <charconv>
#include <cmath>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "common.cuh"

namespace um {
namespace {

struct CsvTable {
  int64_t rows = 0, cols = 0;
  bool has_header = false;
  std::vector<std::string> headers;
  std::vector<double> data;
};

std::mutex g_csv_mu;
std::unordered_map<uint64_t, CsvTable*> g_csv;
uint64_t g_csv_next = 1;

struct Field {
  size_t off, len;
};

inline bool is_space(unsigned char c) { return c <= ' '; }  // String.trim: code points <= U+0020

// big(str).toDouble (Big.scala:59-64)
double parse_cell(const char* p, size_t n) {
  while (n && is_space((unsigned char)p[0])) { p++; n--; }
  while (n && is_space((unsigned char)p[n - 1])) n--;
  if (n == 0) return NAN;
  {
    // fast path: a plain literal (no ',' '$' '%'): from_chars reads java.math.BigDecimal's grammar once a leading '+' is
    // dropped and a non-numeric first character ("inf", "nan") is refused; it must consume the whole field
    bool plain = true;
    for (size_t i = 0; i < n; i++)
      if (p[i] == ',' || p[i] == '$' || p[i] == '%') { plain = false; break; }
    if (plain) {
      const char* q = p;
      size_t m = n;
      if (*q == '+') {
        q++; m--;
        if (m == 0 || *q == '+' || *q == '-') return NAN;
      }
      const char first = (*q == '-' && m > 1) ? q[1] : *q;
      if (!((first >= '0' && first <= '9') || first == '.')) return NAN;
      double v = 0.0;
      const auto r = std::from_chars(q, q + m, v);
      if (r.ec == std::errc() && r.ptr == q + m) return v == 0.0 ? 0.0 : v;
      if (r.ec != std::errc::result_out_of_range) return NAN;
      // overflow / underflow: the general routine below decides between Infinity and 0
    }
  }
  char buf[96];
  size_t m = 0;
  for (size_t i = 0; i < n; i++) {
    if (p[i] == ',' || p[i] == '$') continue;
    if (m >= sizeof(buf) - 8) return NAN;  // longer than any double literal worth reading
    buf[m++] = p[i];
  }
  bool percent = m && buf[m - 1] == '%';
  if (percent) m--;
  // java.math.BigDecimal(String) grammar
  size_t i = 0;
  bool neg = false;
  if (i < m && (buf[i] == '+' || buf[i] == '-')) { neg = buf[i] == '-'; i++; }
  const size_t num0 = i;
  size_t nd = 0;
  while (i < m && buf[i] >= '0' && buf[i] <= '9') { i++; nd++; }
  if (i < m && buf[i] == '.') {
    i++;
    while (i < m && buf[i] >= '0' && buf[i] <= '9') { i++; nd++; }
  }
  if (nd == 0) return NAN;
  const size_t mant_end = i;
  long long exp10 = 0;
  if (i < m && (buf[i] == 'e' || buf[i] == 'E')) {
    i++;
    bool eneg = false;
    if (i < m && (buf[i] == '+' || buf[i] == '-')) { eneg = buf[i] == '-'; i++; }
    size_t ed = 0;
    while (i < m && buf[i] >= '0' && buf[i] <= '9') {
      if (exp10 < 100000000) exp10 = exp10 * 10 + (buf[i] - '0');
      i++; ed++;
    }
    if (ed == 0 || ed > 10) return NAN;
    if (eneg) exp10 = -exp10;
  }
  if (i != m) return NAN;
  if (percent) exp10 -= 2;  // BigDecimal(t.init) / 100: a decimal shift, exact
  char lit[128];
  size_t k = mant_end - num0;
  memcpy(lit, buf + num0, k);
  k += (size_t)snprintf(lit + k, sizeof(lit) - k, "e%lld", exp10);
  double v = 0.0;
  auto r = std::from_chars(lit, lit + k, v);  // correctly rounded, like BigDecimal.doubleValue
  if (r.ec == std::errc::result_out_of_range) {
    // from_chars leaves v untouched: overflow -> Infinity, underflow -> 0 as doubleValue does
    bool all_zero = true;
    for (size_t q = 0; q < mant_end - num0; q++)
      if (lit[q] >= '1' && lit[q] <= '9') all_zero = false;
    v = all_zero ? 0.0 : (exp10 > 0 ? INFINITY : 0.0);
  } else if (r.ec != std::errc()) {
    return NAN;
  }
  if (v == 0.0) return 0.0;  // a BigDecimal has no negative zero
  return neg ? -v : v;
}

// Split into rows of fields (quotes honoured); fields are NOT yet trimmed.  Unquoted copies live in `store`.
void split_rows(const std::string& text, char delim, std::string& store, std::vector<Field>& fields,
                std::vector<size_t>& row_start) {
  const size_t n = text.size();
  size_t i = 0;
  if (n >= 3 && (unsigned char)text[0] == 0xEF && (unsigned char)text[1] == 0xBB && (unsigned char)text[2] == 0xBF) i = 3;
  store.reserve(n + 1);
  row_start.push_back(0);
  size_t fstart = store.size();
  bool in_quotes = false, row_has_content = false;
  auto end_field = [&]() {
    fields.push_back(Field{fstart, store.size() - fstart});
    fstart = store.size();
  };
  auto end_row = [&]() {
    end_field();
    if (row_has_content) row_start.push_back(fields.size());
    else fields.resize(row_start.back());  // FastCsv.isContentRow: a row without content is dropped
    row_has_content = false;
  };
  for (; i < n; i++) {
    const char c = text[i];
    if (in_quotes) {
      if (c == '"') {
        if (i + 1 < n && text[i + 1] == '"') { store.push_back('"'); i++; }
        else in_quotes = false;
      } else {
        store.push_back(c);
        if (!is_space((unsigned char)c)) row_has_content = true;
      }
    } else if (c == '"') {
      in_quotes = true;
    } else if (c == delim) {
      end_field();
    } else if (c == '\n') {
      end_row();
    } else if (c == '\r') {
      if (!(i + 1 < n && text[i + 1] == '\n')) end_row();
    } else {
      store.push_back(c);
      if (!is_space((unsigned char)c)) row_has_content = true;
    }
  }
  if (store.size() != fstart || fields.size() != row_start.back()) end_row();
  if (row_start.back() != fields.size()) fields.resize(row_start.back());
}

// Quote-free text (every numeric matrix file): rows are independent, so the text is cut at line ends into one byte range per
// thread.  Pass A counts content rows and the widest row of each range, pass B converts the cells straight into the table.
// '\r' and '\n' both end a row (the empty row inside "\r\n" has no content and is dropped like any blank row).
struct PlainRange {
  size_t a = 0, b = 0;
  int64_t rows = 0;
  size_t width = 0;
};
// on_row returns false to stop.  A row has content when some field is non-blank (FastCsv.isContentRow): a line of
// delimiters and blanks only is dropped.
template <class RowFn>
void plain_rows(const char* s, size_t a, size_t b, char delim, RowFn&& on_row) {
  size_t ls = a;
  while (ls < b) {
    size_t le = ls;
    bool content = false;
    while (le < b && s[le] != '\n' && s[le] != '\r') {
      if (!is_space((unsigned char)s[le]) && s[le] != delim) content = true;
      le++;
    }
    if (content && !on_row(ls, le)) return;
    ls = le + 1;
  }
}
inline size_t plain_width(const char* s, size_t ls, size_t le, char delim) {
  size_t w = 1;
  for (size_t i = ls; i < le; i++)
    if (s[i] == delim) w++;
  return w;
}
void open_plain(const std::string& text, char delim, CsvTable* t) {
  const char* s = text.data();
  const size_t n = text.size();
  size_t i0 = 0;
  if (n >= 3 && (unsigned char)s[0] == 0xEF && (unsigned char)s[1] == 0xBB && (unsigned char)s[2] == 0xBF) i0 = 3;
  const int nthreads = (int)std::min<size_t>(std::max(1u, std::thread::hardware_concurrency()), (n - i0) / (1u << 18) + 1);
  std::vector<PlainRange> rg((size_t)nthreads);
  size_t prev = i0;
  for (int k = 0; k < nthreads; k++) {
    size_t e = k + 1 == nthreads ? n : i0 + (n - i0) / (size_t)nthreads * (size_t)(k + 1);
    if (e < prev) e = prev;
    while (e < n && s[e] != '\n' && s[e] != '\r') e++;  // end the range at a line end
    if (e < n) e++;
    rg[(size_t)k].a = prev;
    rg[(size_t)k].b = e;
    prev = e;
  }
  auto run = [&](auto&& fn) {
    if (nthreads == 1) { fn(0); return; }
    std::vector<std::thread> pool;
    for (int k = 0; k < nthreads; k++) pool.emplace_back(fn, k);
    for (auto& th : pool) th.join();
  };
  run([&](int k) {
    PlainRange& r = rg[(size_t)k];
    plain_rows(s, r.a, r.b, delim, [&](size_t ls, size_t le) {
      r.rows++;
      r.width = std::max(r.width, plain_width(s, ls, le, delim));
      return true;
    });
  });
  int64_t nrows_all = 0;
  size_t width = 0;
  for (auto& r : rg) { nrows_all += r.rows; width = std::max(width, r.width); }
  if (nrows_all == 0) return;
  // the first two content rows decide the header (FileOps.scala:154-166)
  std::vector<std::pair<size_t, size_t>> head;
  plain_rows(s, i0, n, delim, [&](size_t ls, size_t le) {
    head.emplace_back(ls, le);
    return head.size() < 2;
  });
  auto row_cells = [&](size_t ls, size_t le, auto&& on_cell) {
    size_t c = 0, fs = ls;
    for (size_t i = ls; i <= le; i++)
      if (i == le || s[i] == delim) { on_cell(c++, fs, i); fs = i + 1; }
  };
  int64_t header = 0;
  if (nrows_all > 1) {
    bool row0_text = true, row1_data = false;
    row_cells(head[0].first, head[0].second, [&](size_t, size_t a, size_t b) { if (!std::isnan(parse_cell(s + a, b - a))) row0_text = false; });
    row_cells(head[1].first, head[1].second, [&](size_t, size_t a, size_t b) { if (!std::isnan(parse_cell(s + a, b - a))) row1_data = true; });
    if (row0_text && row1_data) header = 1;
  }
  if (header) {  // namedHeaders (FileOps.scala:117-121)
    t->headers.assign(width, std::string());
    row_cells(head[0].first, head[0].second, [&](size_t c, size_t a, size_t b) {
      while (a < b && is_space((unsigned char)s[a])) a++;
      while (b > a && is_space((unsigned char)s[b - 1])) b--;
      t->headers[c].assign(s + a, b - a);
    });
    for (size_t c = 0; c < width; c++)
      if (t->headers[c].empty()) t->headers[c] = "col" + std::to_string(c + 1);
  }
  t->has_header = header != 0;
  t->rows = nrows_all - header;
  t->cols = (int64_t)width;
  t->data.assign((size_t)t->rows * width, NAN);  // cells a short row does not have stay NaN (FastCsv.rectangular)
  std::vector<int64_t> first((size_t)nthreads + 1, 0);
  for (int k = 0; k < nthreads; k++) first[(size_t)k + 1] = first[(size_t)k] + rg[(size_t)k].rows;
  run([&](int k) {
    int64_t r = first[(size_t)k];
    plain_rows(s, rg[(size_t)k].a, rg[(size_t)k].b, delim, [&](size_t ls, size_t le) {
      const int64_t out_r = r++ - header;
      if (out_r < 0) return true;
      double* dst = t->data.data() + (size_t)out_r * width;
      row_cells(ls, le, [&](size_t c, size_t a, size_t b) { dst[c] = parse_cell(s + a, b - a); });
      return true;
    });
  });
}

char detect_delimiter(const std::string& text) {
  size_t i = 0;
  while (i < text.size()) {  // first line with content
    size_t e = text.find('\n', i);
    if (e == std::string::npos) e = text.size();
    bool content = false;
    for (size_t q = i; q < e; q++)
      if (!is_space((unsigned char)text[q])) content = true;
    if (content) {
      const char cand[4] = {',', '\t', ';', '|'};
      int best = 0, bestn = 0;
      for (int k = 0; k < 4; k++) {
        int cnt = 0;
        bool q = false;
        for (size_t z = i; z < e; z++) {
          if (text[z] == '"') q = !q;
          else if (!q && text[z] == cand[k]) cnt++;
        }
        if (cnt > bestn) { bestn = cnt; best = k; }
      }
      return cand[best];
    }
    i = e + 1;
  }
  return ',';
}

// java.lang.Double.toString / Float.toString layout (JDK 19+: shortest digits that round-trip): plain decimal for
// 1e-3 <= |d| < 1e7, otherwise d.dddE[-]n; always at least one digit after the point.
template <class F>
void java_to_string(F d, std::string& out) {
  if (d == 0) { out += std::signbit(d) ? "-0.0" : "0.0"; return; }
  char sci[64];
  auto r = std::to_chars(sci, sci + sizeof(sci) - 1, d, std::chars_format::scientific);  // [-]d[.ddd]e[+-]XX, shortest
  *r.ptr = 0;
  const char* p = sci;
  if (*p == '-') { out.push_back('-'); p++; }
  std::string digits;
  const char* e = p;
  while (e < r.ptr && *e != 'e') { if (*e != '.') digits.push_back(*e); e++; }
  int exp10 = atoi(e + 1);
  if (digits.size() == 1) {
    // Java renders at least two digits and, among the two-digit decimals that read back as d, takes the closest
    // (Double.MIN_VALUE is "4.9E-324", not "5.0E-324"); for every normal value that is just "d.0"
    char two[64];
    snprintf(two, sizeof(two), "%.1e", (double)(d < 0 ? -d : d));
    F back = (F)strtod(two, nullptr);
    if (sizeof(F) == 4) back = strtof(two, nullptr);
    if (back == (d < 0 ? -d : d)) {
      digits.assign(1, two[0]);
      digits.push_back(two[2]);
      exp10 = atoi(strchr(two, 'e') + 1);
      if (digits[1] == '0') digits.pop_back();
    }
  }
  const int nd = (int)digits.size();
  if (exp10 >= -3 && exp10 < 7) {
    if (exp10 >= 0) {
      for (int i = 0; i <= exp10; i++) out.push_back(i < nd ? digits[i] : '0');
      out.push_back('.');
      if (nd > exp10 + 1) out.append(digits, exp10 + 1, std::string::npos);
      else out.push_back('0');
    } else {
      out += "0.";
      for (int i = 0; i < -exp10 - 1; i++) out.push_back('0');
      out += digits;
    }
  } else {
    out.push_back(digits[0]);
    out.push_back('.');
    if (nd > 1) out.append(digits, 1, std::string::npos);
    else out.push_back('0');
    out.push_back('E');
    out += std::to_string(exp10);
  }
}

template <class F>
void format_float_cell(F v, const char* nan_as, std::string& out) {  // Mat.scala:95-101
  if (v != v) out += nan_as;
  else if (std::isinf(v)) out += v > 0 ? "Inf" : "-Inf";
  else java_to_string(v, out);
}

}  // namespace
}  // namespace um

using namespace um;

extern "C" {

int32_t um_csv_open(const char* path, int32_t delimiter, uint64_t* handle, int64_t* rows, int64_t* cols, int32_t* has_header) {
  if (!path || !handle || !rows || !cols) return fail(UM_ERR_ARG, "um_csv_open: null argument");
  FILE* f = fopen(path, "rb");
  if (!f) return fail(UM_ERR_ARG, "um_csv_open: cannot read %s", path);
  std::string text;
  char chunk[1 << 16];
  size_t got;
  while ((got = fread(chunk, 1, sizeof(chunk), f)) > 0) text.append(chunk, got);
  fclose(f);
  const char delim = delimiter > 0 ? (char)delimiter : detect_delimiter(text);
  if (memchr(text.data(), '"', text.size()) == nullptr) {  // no quoted field anywhere: rows can be cut apart in parallel
    CsvTable* tp = new CsvTable();
    open_plain(text, delim, tp);
    {
      std::lock_guard<std::mutex> lk(g_csv_mu);
      *handle = g_csv_next++;
      g_csv[*handle] = tp;
    }
    *rows = tp->rows;
    *cols = tp->cols;
    if (has_header) *has_header = tp->has_header ? 1 : 0;
    return UM_OK;
  }
  std::string store;
  std::vector<Field> fields;
  std::vector<size_t> row_start;
  split_rows(text, delim, store, fields, row_start);
  const int64_t nrows_all = (int64_t)row_start.size() - 1;
  CsvTable* t = new CsvTable();
  if (nrows_all > 0) {
    size_t width = 0;
    for (int64_t r = 0; r < nrows_all; r++) width = std::max(width, row_start[r + 1] - row_start[r]);
    auto cell = [&](int64_t r, size_t c) -> double {
      if (c >= row_start[r + 1] - row_start[r]) return NAN;  // padded "" (FastCsv.rectangular) -> BigNaN
      const Field& fd = fields[row_start[r] + c];
      return parse_cell(store.data() + fd.off, fd.len);
    };
    // header detection (FileOps.scala:154-166)
    int64_t header = 0;
    if (nrows_all > 1) {
      bool row0_text = true, row1_data = false;
      for (size_t c = 0; c < width; c++) {
        if (!std::isnan(cell(0, c))) row0_text = false;
        if (!std::isnan(cell(1, c))) row1_data = true;
      }
      if (row0_text && row1_data) header = 1;
    }
    if (header) {  // namedHeaders (FileOps.scala:117-121)
      for (size_t c = 0; c < width; c++) {
        std::string h;
        if (c < row_start[1] - row_start[0]) {
          const Field& fd = fields[row_start[0] + c];
          size_t a = fd.off, b = fd.off + fd.len;
          while (a < b && is_space((unsigned char)store[a])) a++;
          while (b > a && is_space((unsigned char)store[b - 1])) b--;
          h.assign(store, a, b - a);
        }
        if (h.empty()) h = "col" + std::to_string(c + 1);
        t->headers.push_back(h);
      }
    }
    t->has_header = header != 0;
    t->rows = nrows_all - header;
    t->cols = (int64_t)width;
    t->data.resize((size_t)t->rows * width);
    const int nthreads = (int)std::min<int64_t>(std::max(1u, std::thread::hardware_concurrency()), std::max<int64_t>(1, t->rows / 256));
    auto work = [&](int64_t r0, int64_t r1) {
      for (int64_t r = r0; r < r1; r++)
        for (size_t c = 0; c < width; c++) t->data[(size_t)r * width + c] = cell(r + header, c);
    };
    if (nthreads <= 1) {
      work(0, t->rows);
    } else {
      std::vector<std::thread> pool;
      const int64_t per = (t->rows + nthreads - 1) / nthreads;
      for (int k = 0; k < nthreads; k++) pool.emplace_back(work, std::min<int64_t>(k * per, t->rows), std::min<int64_t>((k + 1) * per, t->rows));
      for (auto& th : pool) th.join();
    }
  }
  {
    std::lock_guard<std::mutex> lk(g_csv_mu);
    *handle = g_csv_next++;
    g_csv[*handle] = t;
  }
  *rows = t->rows;
  *cols = t->cols;
  if (has_header) *has_header = t->has_header ? 1 : 0;
  return UM_OK;
}

static CsvTable* csv_of(uint64_t h) {
  std::lock_guard<std::mutex> lk(g_csv_mu);
  auto it = g_csv.find(h);
  return it == g_csv.end() ? nullptr : it->second;
}

int32_t um_csv_header(uint64_t handle, int64_t col, char* buf, int64_t buflen) {
  CsvTable* t = csv_of(handle);
  if (!t || !buf || buflen <= 0) return fail(UM_ERR_ARG, "um_csv_header: bad handle or buffer");
  if (!t->has_header || col < 0 || col >= (int64_t)t->headers.size()) return fail(UM_ERR_ARG, "um_csv_header: no header %lld", (long long)col);
  snprintf(buf, (size_t)buflen, "%s", t->headers[(size_t)col].c_str());
  return UM_OK;
}

int32_t um_csv_read_host(uint64_t handle, double* dst, int64_t count) {
  CsvTable* t = csv_of(handle);
  if (!t || (!dst && count > 0)) return fail(UM_ERR_ARG, "um_csv_read_host: bad handle or buffer");
  if (count != t->rows * t->cols) return fail(UM_ERR_ARG, "um_csv_read_host: %lld elements, table has %lld", (long long)count, (long long)(t->rows * t->cols));
  if (count) memcpy(dst, t->data.data(), (size_t)count * sizeof(double));
  return UM_OK;
}

int32_t um_csv_upload(uint64_t handle, const um_view* dst) {
  CsvTable* t = csv_of(handle);
  if (!t) return fail(UM_ERR_ARG, "um_csv_upload: bad handle");
  UM_CTX();
  UM_VIEW(dst);
  if (dst->rows != t->rows || dst->cols != t->cols) return fail(UM_ERR_ARG, "um_csv_upload: destination %lldx%lld, table %lldx%lld", (long long)dst->rows, (long long)dst->cols, (long long)t->rows, (long long)t->cols);
  const bool contiguous = dst->cs == 1 && (dst->rows <= 1 || dst->rs == dst->cols);
  if (!contiguous && t->rows * t->cols > 0) return fail(UM_ERR_ARG, "um_csv_upload: destination must be contiguous row-major");
  const size_t n = (size_t)(t->rows * t->cols);
  if (n == 0) return UM_OK;
  if (dst->dtype == UM_F64) {
    UM_CUDA(cudaMemcpyAsync(static_cast<double*>(dst->data) + dst->offset, t->data.data(), n * sizeof(double), cudaMemcpyHostToDevice, ctx()->stream));
    UM_CUDA(cudaStreamSynchronize(ctx()->stream));
  } else if (dst->dtype == UM_F32) {  // loadMatF: _.toDouble.toFloat
    std::vector<float> tmp(n);
    for (size_t i = 0; i < n; i++) tmp[i] = (float)t->data[i];
    UM_CUDA(cudaMemcpyAsync(static_cast<float*>(dst->data) + dst->offset, tmp.data(), n * sizeof(float), cudaMemcpyHostToDevice, ctx()->stream));
    UM_CUDA(cudaStreamSynchronize(ctx()->stream));
  } else {
    return fail(UM_ERR_ARG, "um_csv_upload: destination must be f64 or f32");
  }
  return UM_OK;
}

int32_t um_csv_close(uint64_t handle) {
  std::lock_guard<std::mutex> lk(g_csv_mu);
  auto it = g_csv.find(handle);
  if (it == g_csv.end()) return fail(UM_ERR_ARG, "um_csv_close: bad handle");
  delete it->second;
  g_csv.erase(it);
  return UM_OK;
}

/* formats one value the way saveCSV writes it (host only; exposed for the binding's toString and for tests) */
int32_t um_csv_format_f64(double v, const char* nan_as, char* buf, int64_t buflen) {
  if (!buf || buflen <= 0) return fail(UM_ERR_ARG, "um_csv_format_f64: bad buffer");
  std::string s;
  format_float_cell(v, nan_as ? nan_as : "NaN", s);
  snprintf(buf, (size_t)buflen, "%s", s.c_str());
  return UM_OK;
}
int32_t um_csv_format_f32(float v, const char* nan_as, char* buf, int64_t buflen) {
  if (!buf || buflen <= 0) return fail(UM_ERR_ARG, "um_csv_format_f32: bad buffer");
  std::string s;
  format_float_cell(v, nan_as ? nan_as : "NaN", s);
  snprintf(buf, (size_t)buflen, "%s", s.c_str());
  return UM_OK;
}
int32_t um_csv_parse_cell(const char* text, double* out) {
  if (!text || !out) return fail(UM_ERR_ARG, "um_csv_parse_cell: null argument");
  *out = parse_cell(text, strlen(text));
  return UM_OK;
}

int32_t um_csv_save(const um_view* m, const char* path, const char* sep, const char* nan_as) {
  if (!path) return fail(UM_ERR_ARG, "um_csv_save: null path");
  UM_CTX();
  UM_VIEW(m);
  const char* sp = sep ? sep : ",";
  const char* na = nan_as ? nan_as : "NaN";
  const int64_t R = m->rows, Cn = m->cols;
  const bool contiguous = m->cs == 1 && (R <= 1 || m->rs == Cn);
  if (!contiguous && R * Cn > 0) return fail(UM_ERR_ARG, "um_csv_save: materialise the view first (contiguous row-major expected)");
  const size_t n = (size_t)(R * Cn), esz = dtype_size(m->dtype);
  std::vector<unsigned char> host(n * esz);
  if (n) {
    UM_CUDA(cudaMemcpyAsync(host.data(), static_cast<const unsigned char*>(m->data) + (size_t)m->offset * esz, n * esz, cudaMemcpyDeviceToHost, ctx()->stream));
    UM_CUDA(cudaStreamSynchronize(ctx()->stream));
  }
  const int nthreads = (int)std::min<int64_t>(std::max(1u, std::thread::hardware_concurrency()), std::max<int64_t>(1, R / 256));
  std::vector<std::string> parts((size_t)nthreads);
  auto work = [&](int k, int64_t r0, int64_t r1) {
    std::string& out = parts[(size_t)k];
    out.reserve((size_t)(r1 - r0) * (size_t)Cn * 20);
    for (int64_t r = r0; r < r1; r++) {
      for (int64_t c = 0; c < Cn; c++) {
        if (c) out += sp;
        const size_t i = (size_t)(r * Cn + c);
        switch (m->dtype) {
          case UM_F64: format_float_cell(reinterpret_cast<const double*>(host.data())[i], na, out); break;
          case UM_F32: format_float_cell(reinterpret_cast<const float*>(host.data())[i], na, out); break;
          case UM_I32: out += std::to_string(reinterpret_cast<const int32_t*>(host.data())[i]); break;
          default: out += host[i] ? "true" : "false"; break;  // String.valueOf(Boolean)
        }
      }
      out.push_back('\n');
    }
  };
  if (nthreads <= 1) {
    work(0, 0, R);
  } else {
    std::vector<std::thread> pool;
    const int64_t per = (R + nthreads - 1) / nthreads;
    for (int k = 0; k < nthreads; k++) pool.emplace_back(work, k, std::min<int64_t>(k * per, R), std::min<int64_t>((k + 1) * per, R));
    for (auto& th : pool) th.join();
  }
  FILE* f = fopen(path, "wb");
  if (!f) return fail(UM_ERR_ARG, "um_csv_save: cannot write %s", path);
  for (auto& s : parts) fwrite(s.data(), 1, s.size(), f);
  fclose(f);
  return UM_OK;
}

}  // extern "C"
