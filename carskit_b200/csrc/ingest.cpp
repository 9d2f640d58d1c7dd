// ingest.cpp -- native ingest, k-fold split and columnar storage of the rating data.  Host code.
//
// The reference loads ratings into boxed Guava tables (DataDAO.java:174, 342): at 10^8 ratings that is a JVM memory
// wall before the first SGD step.  This file restates the id / layout rules of DataDAO.readData
// (src/carskit/data/processor/DataDAO.java:199-354) over flat arrays, writes them as a columnar file that loads with
// one read per column, and replicates DataSplitter's k-fold assignment (DataSplitter.java:102-133) on a
// java.util.Random clone -- so that cars_desc can be filled without the JVM holding the data at all.
//
//   header  `User, Item, Rating, dim:cond, ...` split on [tab ,]+ (:201); condition id = column - 3 (:209); dimension
//           id = first appearance of the text before ':' (:207-208); "...:na" conditions are the empty ones (:214-215)
//   lines   split on ',' (:226); user / item ids by first appearance (:238-242); the user-item PAIR id by first
//           appearance of "row,col" (:266-268); the context id by first appearance of the comma-joined ids of the
//           conditions whose flag is 1 (:279-330); a repeated (pair, context) overwrites the rating (Table.put, :342)
//   order   `for (MatrixEntry me : trainMatrix)`: pair id ascending, context id ascending; 0.0 ratings are not stored
//   mean    SparseMatrix.getGlobalAvg: sequential sum / count (data/structure/SparseMatrix.java:49-56)
#include "../../include/carskit_b200.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <string>
#include <unordered_map>
#include <vector>

struct cars_dataset {
  int32_t num_users = 0, num_items = 0, num_pairs = 0, num_contexts = 0, num_conditions = 0, num_context_dims = 0;
  std::vector<int32_t> u, j, ctx, pair;  // one entry per rating, CRS order
  std::vector<double> r;
  std::vector<int32_t> ctx_ptr, ctx_cond;  // context -> condition ids (getContextConditionsList, :1035)
  std::vector<int32_t> empty_conds;        // EmptyContextConditions
  std::vector<double> rating_scale;        // getRatingScale(): sorted distinct ratings of the WHOLE file
  std::vector<std::string> user_names, item_names, cond_names;
  double global_mean = 0.0;
};

static thread_local std::string g_ds_error;
static int ds_fail(int code, const std::string& msg) {
  g_ds_error = msg;
  return code;
}

static double sequential_mean(const std::vector<double>& r) {
  double s = 0.0;
  for (double v : r) s += v;
  return r.empty() ? 0.0 : s / (double)r.size();
}

static void split(const std::string& s, const char* seps, bool collapse, std::vector<std::string>* out) {
  out->clear();
  size_t i = 0;
  while (i <= s.size()) {
    size_t k = s.find_first_of(seps, i);
    if (k == std::string::npos) k = s.size();
    if (!(collapse && k == i)) out->push_back(s.substr(i, k - i));
    i = k + 1;
  }
}
static std::string trim(const std::string& s) {
  size_t a = 0, b = s.size();
  while (a < b && (s[a] == ' ' || s[a] == '\t' || s[a] == '\r' || s[a] == '\n')) a++;
  while (b > a && (s[b - 1] == ' ' || s[b - 1] == '\t' || s[b - 1] == '\r' || s[b - 1] == '\n')) b--;
  return s.substr(a, b - a);
}

template <typename Map>
static int32_t intern(Map& m, std::vector<std::string>* names, const std::string& key) {
  auto it = m.find(key);
  if (it != m.end()) return it->second;
  const int32_t id = (int32_t)m.size();
  m.emplace(key, id);
  if (names) names->push_back(key);
  return id;
}

extern "C" int cars_dataset_read_binary_csv(const char* path, cars_dataset** out) {
  if (!path || !out) return ds_fail(CARS_E_INVALID, "NULL argument");
  *out = nullptr;
  FILE* f = fopen(path, "rb");
  if (!f) return ds_fail(CARS_E_INVALID, std::string("cannot open ") + path);
  std::string text;
  {
    char buf[1 << 16];
    size_t n;
    while ((n = fread(buf, 1, sizeof buf, f)) > 0) text.append(buf, n);
    fclose(f);
  }
  cars_dataset* d = new (std::nothrow) cars_dataset();
  if (!d) return ds_fail(CARS_E_OOM, "host allocation failed");
  try {
    size_t pos = 0;
    auto next_line = [&](std::string* line) {
      if (pos >= text.size()) return false;
      size_t e = text.find('\n', pos);
      if (e == std::string::npos) e = text.size();
      *line = text.substr(pos, e - pos);
      pos = e + 1;
      return true;
    };
    std::string line;
    std::vector<std::string> tok;
    if (!next_line(&line)) { delete d; return ds_fail(CARS_E_INVALID, "empty file"); }
    split(trim(line), "\t,", true, &tok);
    std::unordered_map<std::string, int32_t> dims;
    for (size_t i = 3; i < tok.size(); i++) {
      const std::string context = trim(tok[i]);
      const std::string dim = trim(context.substr(0, context.find(':')));
      intern(dims, nullptr, dim);
      d->cond_names.push_back(context);
      if (context.size() >= 3 && context.compare(context.size() - 3, 3, ":na") == 0) d->empty_conds.push_back((int32_t)i - 3);
    }
    d->num_conditions = (int32_t)d->cond_names.size();
    d->num_context_dims = (int32_t)dims.size();
    std::unordered_map<std::string, int32_t> users, items, ctxs;
    std::unordered_map<uint64_t, int32_t> pairs;
    std::vector<int32_t> pair_user, pair_item;
    std::unordered_map<uint64_t, double> table;  // (pair << 32 | context) -> rating, last one wins
    std::vector<double> scale;
    std::string ctxkey;
    std::vector<int32_t> conds;
    int64_t lineno = 1;
    // Data lines are parsed in place (no per-token strings: a 10^8-line file has 3.5 * 10^9 tokens).
    const char* base = text.data();
    const size_t total = text.size();
    std::string ukey, ikey;
    while (pos < total) {
      lineno++;
      size_t eol = text.find('\n', pos);
      if (eol == std::string::npos) eol = total;
      size_t a = pos, b = eol;
      pos = eol + 1;
      while (a < b && (base[a] == ' ' || base[a] == '\t' || base[a] == '\r')) a++;   // line.strip()
      while (b > a && (base[b - 1] == ' ' || base[b - 1] == '\t' || base[b - 1] == '\r')) b--;
      if (a == b) continue;
      // the first three fields: user, item, rating (split on ',' exactly, DataDAO.java:226)
      size_t c1 = a;
      while (c1 < b && base[c1] != ',') c1++;
      size_t c2 = c1 + 1;
      while (c2 < b && base[c2] != ',') c2++;
      size_t c3 = c2 + 1;
      while (c3 < b && base[c3] != ',') c3++;
      if (c1 >= b || c2 >= b) { delete d; return ds_fail(CARS_E_INVALID, "line " + std::to_string(lineno) + ": fewer than 3 fields"); }
      ukey.assign(base + a, c1 - a);
      ikey.assign(base + c1 + 1, c2 - c1 - 1);
      char* endp = nullptr;
      const std::string rtxt(base + c2 + 1, (c3 < b ? c3 : b) - c2 - 1);
      const double rate = strtod(rtxt.c_str(), &endp);
      if (endp == rtxt.c_str()) { delete d; return ds_fail(CARS_E_INVALID, "line " + std::to_string(lineno) + ": rating is not a number"); }
      scale.push_back(rate);
      const int32_t row = intern(users, &d->user_names, ukey);
      const int32_t col = intern(items, &d->item_names, ikey);
      const uint64_t pk = ((uint64_t)(uint32_t)row << 32) | (uint32_t)col;
      auto pit = pairs.find(pk);
      int32_t pid;
      if (pit == pairs.end()) {
        pid = (int32_t)pairs.size();
        pairs.emplace(pk, pid);
        pair_user.push_back(row);
        pair_item.push_back(col);
      } else {
        pid = pit->second;
      }
      // the condition flags: field k >= 3 with integer value 1 switches condition k - 3 on (:279-290)
      conds.clear();
      ctxkey.clear();
      int32_t col_idx = 0;
      size_t p = c3 + 1;
      while (c3 < b && p <= b) {
        size_t q = p;
        while (q < b && base[q] != ',') q++;
        size_t x = p, y = q;
        while (x < y && (base[x] == ' ' || base[x] == '\t')) x++;
        while (y > x && (base[y - 1] == ' ' || base[y - 1] == '\t')) y--;
        bool one = (y - x == 1 && base[x] == '1');
        if (!one && y > x) one = atoi(std::string(base + x, y - x).c_str()) == 1;  // "01", "+1": rare, Integer.parseInt semantics
        if (one) {
          conds.push_back(col_idx);
          if (!ctxkey.empty()) ctxkey.push_back(',');
          ctxkey += std::to_string(col_idx);
        }
        col_idx++;
        p = q + 1;
      }
      auto cit = ctxs.find(ctxkey);
      int32_t cid;
      if (cit == ctxs.end()) {
        cid = (int32_t)ctxs.size();
        ctxs.emplace(ctxkey, cid);
        if (d->ctx_ptr.empty()) d->ctx_ptr.push_back(0);
        d->ctx_cond.insert(d->ctx_cond.end(), conds.begin(), conds.end());
        d->ctx_ptr.push_back((int32_t)d->ctx_cond.size());
      } else {
        cid = cit->second;
      }
      table[((uint64_t)(uint32_t)pid << 32) | (uint32_t)cid] = rate;
    }
    (void)tok;
    if (d->ctx_ptr.empty()) d->ctx_ptr.push_back(0);
    d->num_users = (int32_t)users.size();
    d->num_items = (int32_t)items.size();
    d->num_pairs = (int32_t)pairs.size();
    d->num_contexts = (int32_t)ctxs.size();
    std::sort(scale.begin(), scale.end());
    scale.erase(std::unique(scale.begin(), scale.end()), scale.end());
    d->rating_scale = scale;
    std::vector<uint64_t> keys;
    keys.reserve(table.size());
    for (const auto& kv : table)
      if (kv.second != 0.0) keys.push_back(kv.first);  // zero entries are not stored in the sparse matrix
    std::sort(keys.begin(), keys.end());                 // CRS order: pair id, then context id
    const size_t n = keys.size();
    d->u.resize(n); d->j.resize(n); d->ctx.resize(n); d->pair.resize(n); d->r.resize(n);
    for (size_t k = 0; k < n; k++) {
      const int32_t pid = (int32_t)(keys[k] >> 32);
      d->pair[k] = pid;
      d->u[k] = pair_user[(size_t)pid];
      d->j[k] = pair_item[(size_t)pid];
      d->ctx[k] = (int32_t)(keys[k] & 0xffffffffu);
      d->r[k] = table[keys[k]];
    }
    d->global_mean = sequential_mean(d->r);
  } catch (const std::bad_alloc&) {
    delete d;
    return ds_fail(CARS_E_OOM, "host allocation failed while reading the ratings");
  }
  *out = d;
  return CARS_OK;
}

extern "C" int cars_dataset_from_arrays(int32_t num_users, int32_t num_items, int32_t num_conditions, int32_t num_contexts,
                                        int32_t num_context_dims, int64_t nnz, const int32_t* u, const int32_t* j,
                                        const int32_t* ctx, const double* r, const int32_t* ctx_ptr, const int32_t* ctx_cond,
                                        cars_dataset** out) {
  if (!out || nnz < 0 || (nnz > 0 && (!u || !j || !r))) return ds_fail(CARS_E_INVALID, "bad arguments");
  cars_dataset* d = new (std::nothrow) cars_dataset();
  if (!d) return ds_fail(CARS_E_OOM, "host allocation failed");
  try {
    d->num_users = num_users; d->num_items = num_items; d->num_conditions = num_conditions; d->num_contexts = num_contexts;
    d->num_context_dims = num_context_dims;
    d->u.assign(u, u + nnz); d->j.assign(j, j + nnz); d->r.assign(r, r + nnz);
    if (ctx) d->ctx.assign(ctx, ctx + nnz);
    if (ctx_ptr) {
      d->ctx_ptr.assign(ctx_ptr, ctx_ptr + num_contexts + 1);
      d->ctx_cond.assign(ctx_cond, ctx_cond + ctx_ptr[num_contexts]);
    } else {
      d->ctx_ptr.assign(1, 0);
    }
    d->pair.resize((size_t)nnz);  // pair ids: consecutive entries of one (user, item) share a CRS row
    int32_t pid = -1;
    for (int64_t n = 0; n < nnz; n++) {
      if (n == 0 || u[n] != u[n - 1] || j[n] != j[n - 1]) pid++;
      d->pair[(size_t)n] = pid;
    }
    d->num_pairs = pid + 1;
    std::vector<double> scale(d->r);
    std::sort(scale.begin(), scale.end());
    scale.erase(std::unique(scale.begin(), scale.end()), scale.end());
    d->rating_scale = scale;
    d->global_mean = sequential_mean(d->r);
  } catch (const std::bad_alloc&) {
    delete d;
    return ds_fail(CARS_E_OOM, "host allocation failed");
  }
  *out = d;
  return CARS_OK;
}

// ---- columnar file: magic, 16 int64 header words, then the columns back to back --------------------------------------
static const char kMagic[8] = {'C', 'A', 'R', 'S', 'C', 'O', 'L', '1'};

template <typename T>
static bool put(FILE* f, const std::vector<T>& v) { return v.empty() || fwrite(v.data(), sizeof(T), v.size(), f) == v.size(); }
template <typename T>
static bool get(FILE* f, std::vector<T>* v, int64_t n) {
  v->resize((size_t)n);
  return n == 0 || fread(v->data(), sizeof(T), (size_t)n, f) == (size_t)n;
}
static std::string join(const std::vector<std::string>& v) {
  std::string s;
  for (const auto& x : v) { s += x; s.push_back('\n'); }
  return s;
}
static void unjoin(const std::string& s, std::vector<std::string>* v) {
  v->clear();
  size_t i = 0;
  while (i < s.size()) {
    size_t k = s.find('\n', i);
    v->push_back(s.substr(i, k - i));
    i = k + 1;
  }
}

extern "C" int cars_dataset_save(const cars_dataset* d, const char* path) {
  if (!d || !path) return ds_fail(CARS_E_INVALID, "NULL argument");
  FILE* f = fopen(path, "wb");
  if (!f) return ds_fail(CARS_E_INVALID, std::string("cannot create ") + path);
  const std::string un = join(d->user_names), in = join(d->item_names), cn = join(d->cond_names);
  int64_t hdr[16] = {d->num_users, d->num_items, d->num_pairs, d->num_contexts, d->num_conditions, d->num_context_dims,
                     (int64_t)d->r.size(), (int64_t)d->ctx.size(), (int64_t)d->ctx_cond.size(), (int64_t)d->empty_conds.size(),
                     (int64_t)d->rating_scale.size(), (int64_t)un.size(), (int64_t)in.size(), (int64_t)cn.size(), 0, 0};
  memcpy(&hdr[14], &d->global_mean, 8);
  bool ok = fwrite(kMagic, 1, 8, f) == 8 && fwrite(hdr, 8, 16, f) == 16 && put(f, d->u) && put(f, d->j) && put(f, d->ctx) &&
            put(f, d->pair) && put(f, d->r) && put(f, d->ctx_ptr) && put(f, d->ctx_cond) && put(f, d->empty_conds) &&
            put(f, d->rating_scale) && fwrite(un.data(), 1, un.size(), f) == un.size() &&
            fwrite(in.data(), 1, in.size(), f) == in.size() && fwrite(cn.data(), 1, cn.size(), f) == cn.size();
  ok = (fclose(f) == 0) && ok;
  return ok ? CARS_OK : ds_fail(CARS_E_INVALID, std::string("short write to ") + path);
}

extern "C" int cars_dataset_load(const char* path, cars_dataset** out) {
  if (!path || !out) return ds_fail(CARS_E_INVALID, "NULL argument");
  *out = nullptr;
  FILE* f = fopen(path, "rb");
  if (!f) return ds_fail(CARS_E_INVALID, std::string("cannot open ") + path);
  char magic[8];
  int64_t hdr[16];
  if (fread(magic, 1, 8, f) != 8 || memcmp(magic, kMagic, 8) != 0 || fread(hdr, 8, 16, f) != 16) {
    fclose(f);
    return ds_fail(CARS_E_INVALID, std::string(path) + " is not a carskit_b200 columnar file");
  }
  cars_dataset* d = new (std::nothrow) cars_dataset();
  if (!d) { fclose(f); return ds_fail(CARS_E_OOM, "host allocation failed"); }
  bool ok = true;
  try {
    d->num_users = (int32_t)hdr[0]; d->num_items = (int32_t)hdr[1]; d->num_pairs = (int32_t)hdr[2];
    d->num_contexts = (int32_t)hdr[3]; d->num_conditions = (int32_t)hdr[4]; d->num_context_dims = (int32_t)hdr[5];
    memcpy(&d->global_mean, &hdr[14], 8);
    const int64_t n = hdr[6];
    std::string un((size_t)hdr[11], 0), in((size_t)hdr[12], 0), cn((size_t)hdr[13], 0);
    ok = get(f, &d->u, n) && get(f, &d->j, n) && get(f, &d->ctx, hdr[7]) && get(f, &d->pair, n) && get(f, &d->r, n) &&
         get(f, &d->ctx_ptr, (int64_t)d->num_contexts + 1) && get(f, &d->ctx_cond, hdr[8]) && get(f, &d->empty_conds, hdr[9]) &&
         get(f, &d->rating_scale, hdr[10]) && (un.empty() || fread(&un[0], 1, un.size(), f) == un.size()) &&
         (in.empty() || fread(&in[0], 1, in.size(), f) == in.size()) && (cn.empty() || fread(&cn[0], 1, cn.size(), f) == cn.size());
    unjoin(un, &d->user_names); unjoin(in, &d->item_names); unjoin(cn, &d->cond_names);
  } catch (const std::bad_alloc&) {
    ok = false;
  }
  fclose(f);
  if (!ok) { delete d; return ds_fail(CARS_E_INVALID, std::string(path) + ": truncated"); }
  *out = d;
  return CARS_OK;
}

extern "C" int cars_dataset_get_view(const cars_dataset* d, cars_dataset_view* v) {
  if (!d || !v) return ds_fail(CARS_E_INVALID, "NULL argument");
  v->num_users = d->num_users; v->num_items = d->num_items; v->num_pairs = d->num_pairs; v->num_contexts = d->num_contexts;
  v->num_conditions = d->num_conditions; v->num_context_dims = d->num_context_dims;
  v->nnz = (int64_t)d->r.size();
  v->u = d->u.data(); v->j = d->j.data(); v->ctx = d->ctx.empty() ? nullptr : d->ctx.data(); v->pair = d->pair.data();
  v->r = d->r.data(); v->ctx_ptr = d->ctx_ptr.data(); v->ctx_cond = d->ctx_cond.data();
  v->global_mean = d->global_mean;
  v->min_rate = d->rating_scale.empty() ? 0.0 : d->rating_scale.front();
  v->max_rate = d->rating_scale.empty() ? 0.0 : d->rating_scale.back();
  v->num_empty_conditions = (int32_t)d->empty_conds.size();
  v->empty_conditions = d->empty_conds.data();
  return CARS_OK;
}

// ---- DataSplitter (DataSplitter.java:102-133): fold labels from java.util.Random(seed) -------------------------------
static inline uint64_t jr_next(uint64_t* s, int bits) {
  *s = (*s * 0x5DEECE66DULL + 0xBULL) & ((1ULL << 48) - 1);
  return *s >> (48 - bits);
}

extern "C" int cars_dataset_kfold(const cars_dataset* d, int32_t kfold, int64_t seed, int32_t fold, cars_dataset** train,
                                  cars_dataset** test) {
  if (!d || !train || !test || kfold < 1) return ds_fail(CARS_E_INVALID, "bad arguments");
  *train = *test = nullptr;
  const int64_t n = (int64_t)d->r.size();
  const int32_t nf = n ? (int32_t)std::min<int64_t>(kfold, n) : kfold;
  if (fold < 1 || fold > nf) return ds_fail(CARS_E_INVALID, "fold outside 1..numFold");
  cars_dataset *tr = nullptr, *te = nullptr;
  try {
    // rdm[i] = Randoms.uniform() = nextDouble(); fold[i] = (int)(i / (numRates / numFold)) + 1   (:108-118)
    std::vector<double> rdm((size_t)n);
    uint64_t s = ((uint64_t)seed ^ 0x5DEECE66DULL) & ((1ULL << 48) - 1);
    for (int64_t i = 0; i < n; i++) {
      const uint64_t hi = jr_next(&s, 26), lo = jr_next(&s, 27);
      rdm[(size_t)i] = (double)((hi << 27) + lo) * (1.0 / 9007199254740992.0);
    }
    const double indv = ((double)n + 0.0) / nf;
    std::vector<int64_t> order((size_t)n);
    std::iota(order.begin(), order.end(), 0);
    // Sortor.quickSort(rdm, fold, ...): rdm ascending carrying fold == an argsort while no two draws are equal (:120)
    std::stable_sort(order.begin(), order.end(), [&](int64_t a, int64_t b) { return rdm[(size_t)a] < rdm[(size_t)b]; });
    std::vector<int32_t> assign((size_t)n);
    for (int64_t f = 0; f < n; f++) {
      const int64_t i = order[(size_t)f];
      if (f > 0 && rdm[(size_t)i] == rdm[(size_t)order[(size_t)f - 1]] &&
          (int32_t)((double)i / indv) != (int32_t)((double)order[(size_t)f - 1] / indv))
        return ds_fail(CARS_E_UNSUPPORTED, "two equal random draws carry different fold labels (Lomuto order not reproduced)");
      assign[(size_t)f] = (int32_t)((double)i / indv) + 1;  // the f-th CRS entry gets fold[order[f]]  (:125-132)
    }
    tr = new cars_dataset(*d);
    te = new cars_dataset(*d);
    for (cars_dataset* x : {tr, te}) { x->u.clear(); x->j.clear(); x->ctx.clear(); x->pair.clear(); x->r.clear(); }
    for (int64_t f = 0; f < n; f++) {
      cars_dataset* x = assign[(size_t)f] == fold ? te : tr;  // label k = the TEST set of fold k (:80-83)
      x->u.push_back(d->u[(size_t)f]); x->j.push_back(d->j[(size_t)f]); x->pair.push_back(d->pair[(size_t)f]);
      x->r.push_back(d->r[(size_t)f]);
      if (!d->ctx.empty()) x->ctx.push_back(d->ctx[(size_t)f]);
    }
    tr->global_mean = sequential_mean(tr->r);  // globalMean of the TRAINING matrix (Recommender.java:265)
    te->global_mean = sequential_mean(te->r);
  } catch (const std::bad_alloc&) {
    delete tr;
    delete te;
    return ds_fail(CARS_E_OOM, "host allocation failed");
  }
  *train = tr;
  *test = te;
  return CARS_OK;
}

extern "C" void cars_dataset_free(cars_dataset* d) { delete d; }
extern "C" const char* cars_dataset_last_error(void) { return g_ds_error.c_str(); }
