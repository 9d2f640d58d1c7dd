// tuning.h -- developer knobs of one handle, parsed from cars_desc.tuning ("key=value;key=value").
// The library reads no environment variable: a test or a profiling script that wants another launch shape,
// the host-side level pass or tiny FM row blocks says so in the descriptor of the handle it creates.
//   engine   shape=<n>        launch shape of the SGD kernel (see pick_*_plan in engine.cu)
//            levels=host      dependency levels by the sequential host pass instead of Kahn's algorithm on the device
//            sched_trace=1    host wall time of every schedule-build phase on stderr
//            copy_threads=<n> host threads of the staged copier (1..8)
//            fast_chunk=<n>   ratings per chunk of the FAST schedule
//            fast_hot_rows=<n> fast_hot_flush=<n>   FAST: hot item rows per CTA (0 = off), updates between flushes
//            tagged=1 tagged_ctas=<n>   EXACT by tagged rows (K1t) instead of completion counters
//            pool=0           cudaMalloc / cudaFree instead of the stream-ordered pool
//   FM       fm_block_rows=<n>      rows per item block of the internal row order (0 = keep the caller's order)
//            fm_dense_min_rows=<n>  smallest input whose context field is reduced by streaming
//            fm_runs=0              users by the gathering piece reduce instead of fm_run_reduce_kernel
//            fm_run_users=<32..512> users per CTA of fm_run_reduce_kernel; fm_run_min_blocks=<n>; fm_run_fuse_update=0
//            fm_lanes_per_piece, fm_ppg_short, fm_ppg_long   shapes of fm_piece_reduce_kernel
//            fm_prepare_tiled=0|1   pre-pass over factor-major V (small inputs) / over a coordinate-major scratch copy
#pragma once
#include <cstdlib>
#include <string>
#include <utility>
#include <vector>

namespace cars {

class Tuning {
 public:
  Tuning() = default;
  explicit Tuning(const char* s) {
    if (!s) return;
    std::string cur;
    auto flush = [&]() {
      if (cur.empty()) return;
      const size_t eq = cur.find('=');
      if (eq == std::string::npos) kv_.emplace_back(cur, "1");
      else kv_.emplace_back(cur.substr(0, eq), cur.substr(eq + 1));
      cur.clear();
    };
    for (const char* p = s; *p; p++) {
      if (*p == ';' || *p == ',' || *p == ' ') flush();
      else cur.push_back(*p);
    }
    flush();
  }
  const char* get(const char* key) const {
    for (const auto& e : kv_)
      if (e.first == key) return e.second.c_str();
    return nullptr;
  }
  long long get_ll(const char* key, long long dflt) const {
    const char* v = get(key);
    return v ? atoll(v) : dflt;
  }
  bool is(const char* key, const char* value) const {
    const char* v = get(key);
    return v && std::string(v) == value;
  }

 private:
  std::vector<std::pair<std::string, std::string>> kv_;
};

}  // namespace cars
