// ref_bench.cc -- the reference's OWN CPU kernels timed on the host cores.
//
// TEST / MEASUREMENT INFRASTRUCTURE ONLY (lives under oracle/): bench.py's `--impl reference` arm and
// its cpu_baseline leg execute this binary; nothing under arrow_b200/ does.
//
// It links the installed reference binaries (pyarrow 24.0.0 wheel: libarrow.so.2400 +
// libarrow_compute.so.2400 + libarrow_acero.so.2400 = the same kernels as /root/reference for this
// path; the reference source itself cannot be configured offline) and calls
// arrow::compute::CallFunction exactly as a user of the reference would:
//
//   pipeline (BASELINE.json configs[1]): add(cast(take(values, indices), float32), other)
//     values  float64 uniform [0, 1e6), null_probability 0.1      
//     indices int64 uniform [0, rows), no nulls
//     other   float32 uniform [0, 1e6), null_probability 0.1
//   filter   (configs[0]): filter(int64 values null_p 0.1, boolean mask s = 0.5)
//   groupby  (configs[2]): Acero aggregate hash_sum + hash_count over int64 key / int64 value
//   sort     (configs[3]): sort_indices(int64 with validity)
//
// CallFunction is single-threaded by design (compute/exec.h:85-91), so "all cores" = `threads`
// row-range slices of the index / other columns in flight on std::threads, each slice one
// CallFunction chain over a zero-copy Slice (values stay shared: any slice may gather any row).
// Timing: steady_clock around the calls, `warmup` untimed + `steps` timed passes, best and mean
// reported; input generation is excluded, output allocation is included (the reference does it per call).
//
// usage: ref_bench <op> <rows> <steps> <warmup> <threads> [groups]
#include <arrow/acero/exec_plan.h>
#include <arrow/acero/options.h>
#include <arrow/api.h>
#include <arrow/compute/api.h>
#include <arrow/compute/initialize.h>
#include <arrow/util/thread_pool.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <random>
#include <string>
#include <thread>
#include <vector>

namespace cp = arrow::compute;
namespace ac = arrow::acero;
using arrow::Datum;

static constexpr uint64_t kSeed = 0x0ff1ce;  // the reference benchmarks' seed (vector_selection_benchmark.cc:37)

#define CHECK_OK(expr)                                                        \
  do {                                                                        \
    auto _st = (expr);                                                        \
    if (!_st.ok()) {                                                          \
      fprintf(stderr, "ref_bench: %s\n", _st.ToString().c_str());             \
      exit(2);                                                                \
    }                                                                         \
  } while (0)

template <typename T>
static T Unwrap(arrow::Result<T> r) {
  CHECK_OK(r.status());
  return std::move(r).ValueUnsafe();
}

// splitmix64: cheap counter-based generator so every thread fills its slice independently
static inline uint64_t Mix(uint64_t x) {
  x += 0x9e3779b97f4a7c15ull;
  x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull;
  x = (x ^ (x >> 27)) * 0x94d049bb133111ebull;
  return x ^ (x >> 31);
}

template <typename Fn>
static void ParallelFor(int64_t n, int threads, Fn fn) {
  std::vector<std::thread> ts;
  for (int t = 0; t < threads; ++t) {
    int64_t lo = n * t / threads / 64 * 64, hi = t == threads - 1 ? n : n * (t + 1) / threads / 64 * 64;
    ts.emplace_back([=] { fn(lo, hi, t); });
  }
  for (auto& t : ts) t.join();
}

// validity bitmap with Bernoulli(1 - null_p) bits; returns the null count
static std::shared_ptr<arrow::Buffer> MakeValidity(int64_t n, double null_p, uint64_t stream, int threads, int64_t* nulls) {
  auto buf = Unwrap(arrow::AllocateBuffer((n + 7) / 8 + 8));
  memset(buf->mutable_data(), 0, buf->size());
  const uint64_t thresh = static_cast<uint64_t>(null_p * 18446744073709551615.0);
  std::vector<int64_t> local(threads, 0);
  ParallelFor(n, threads, [&](int64_t lo, int64_t hi, int t) {
    uint8_t* bits = buf->mutable_data();
    int64_t nn = 0;
    for (int64_t i = lo; i < hi; ++i) {
      const bool valid = Mix(stream * 0x100000001b3ull + i) >= thresh;
      if (valid) bits[i >> 3] |= uint8_t(1u << (i & 7));
      else ++nn;
    }
    // single writer per 64-row aligned range, so the |= never races
    local[t] += nn;
  });
  *nulls = 0;
  for (auto v : local) *nulls += v;
  return buf;
}

template <typename T, typename Gen>
static std::shared_ptr<arrow::Array> MakeColumn(std::shared_ptr<arrow::DataType> type, int64_t n, double null_p, uint64_t stream,
                                                int threads, Gen gen) {
  auto data = Unwrap(arrow::AllocateBuffer(n * sizeof(T)));
  T* out = reinterpret_cast<T*>(data->mutable_data());
  ParallelFor(n, threads, [&](int64_t lo, int64_t hi, int) {
    for (int64_t i = lo; i < hi; ++i) out[i] = gen(Mix(stream * 0x9e3779b1ull + i));
  });
  std::shared_ptr<arrow::Buffer> validity;
  int64_t nulls = 0;
  if (null_p > 0) validity = MakeValidity(n, null_p, stream + 101, threads, &nulls);
  return arrow::MakeArray(arrow::ArrayData::Make(type, n, {validity, std::move(data)}, nulls));
}

static std::string CpuModel() {
  std::ifstream f("/proc/cpuinfo");
  std::string line;
  while (std::getline(f, line))
    if (line.rfind("model name", 0) == 0) return line.substr(line.find(':') + 2);
  return "unknown";
}

struct Timing {
  double best_s, mean_s;
};

template <typename Fn>
static Timing TimeIt(int steps, int warmup, Fn fn) {
  for (int i = 0; i < warmup; ++i) fn();
  double best = 1e30, total = 0;
  for (int i = 0; i < steps; ++i) {
    auto t0 = std::chrono::steady_clock::now();
    fn();
    double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    best = std::min(best, s);
    total += s;
  }
  return {best, total / steps};
}

int main(int argc, char** argv) {
  if (argc < 6) {
    fprintf(stderr, "usage: ref_bench <pipeline|filter|groupby|sort> <rows> <steps> <warmup> <threads> [groups]\n");
    return 1;
  }
  const std::string op = argv[1];
  const int64_t n = atoll(argv[2]);
  const int steps = atoi(argv[3]), warmup = atoi(argv[4]);
  int threads = atoi(argv[5]);
  const int hw = (int)std::thread::hardware_concurrency();
  if (threads <= 0) threads = hw;
  const int64_t groups = argc > 6 ? atoll(argv[6]) : 10000000;
  CHECK_OK(cp::Initialize());
  const int gen_threads = std::max(1, hw);
  int64_t checksum = 0;
  Timing t{};

  if (op == "pipeline") {
    CHECK_OK(arrow::SetCpuThreadPoolCapacity(1));
    auto values = MakeColumn<double>(arrow::float64(), n, 0.1, 1, gen_threads, [](uint64_t r) { return (r >> 11) * (1e6 / 9007199254740992.0); });
    auto indices = MakeColumn<int64_t>(arrow::int64(), n, 0.0, 2, gen_threads, [n](uint64_t r) { return (int64_t)(r % (uint64_t)n); });
    auto other = MakeColumn<float>(arrow::float32(), n, 0.1, 3, gen_threads, [](uint64_t r) { return (float)((r >> 40) * (1e6 / 16777216.0)); });
    std::vector<int64_t> nulls(threads, 0);
    auto pass = [&] {
      ParallelFor(n, threads, [&](int64_t lo, int64_t hi, int t) {
        if (hi <= lo) return;
        Datum taken = Unwrap(cp::CallFunction("take", {values, indices->Slice(lo, hi - lo)}));
        cp::CastOptions co = cp::CastOptions::Unsafe(arrow::float32());
        Datum casted = Unwrap(cp::CallFunction("cast", {taken}, &co));
        Datum sum = Unwrap(cp::CallFunction("add", {casted, other->Slice(lo, hi - lo)}));
        nulls[t] = sum.null_count();
      });
    };
    t = TimeIt(steps, warmup, pass);
    for (auto v : nulls) checksum += v;
  } else if (op == "filter") {
    CHECK_OK(arrow::SetCpuThreadPoolCapacity(1));
    auto values = MakeColumn<int64_t>(arrow::int64(), n, 0.1, 4, gen_threads, [](uint64_t r) { return (int64_t)(r % 201) - 100; });
    int64_t unused = 0;
    auto mask_bits = MakeValidity(n, 0.5, 5, gen_threads, &unused);
    auto mask = arrow::MakeArray(arrow::ArrayData::Make(arrow::boolean(), n, {nullptr, mask_bits}, 0));
    std::vector<int64_t> lens(threads, 0);
    auto pass = [&] {
      ParallelFor(n, threads, [&](int64_t lo, int64_t hi, int t) {
        if (hi <= lo) return;
        Datum out = Unwrap(cp::CallFunction("filter", {values->Slice(lo, hi - lo), mask->Slice(lo, hi - lo)}));
        lens[t] = out.length();
      });
    };
    t = TimeIt(steps, warmup, pass);
    for (auto v : lens) checksum += v;
  } else if (op == "sort") {
    CHECK_OK(arrow::SetCpuThreadPoolCapacity(threads));
    auto keys = MakeColumn<int64_t>(arrow::int64(), n, 0.1, 6, gen_threads, [](uint64_t r) { return (int64_t)(r >> 1) - (int64_t)(1ull << 62); });
    // sort_indices on one array is a single CallFunction (single-threaded std::stable_sort, vector_array_sort.cc:144-178);
    // with threads > 1 the column is presented as `threads` chunks, which the reference sorts chunk by chunk and merges
    // (vector_sort.cc:47-225)
    std::shared_ptr<arrow::ChunkedArray> chunked;
    {
      std::vector<std::shared_ptr<arrow::Array>> chunks;
      for (int c = 0; c < threads; ++c) {
        int64_t lo = n * c / threads, hi = n * (c + 1) / threads;
        if (hi > lo) chunks.push_back(keys->Slice(lo, hi - lo));
      }
      chunked = std::make_shared<arrow::ChunkedArray>(chunks);
    }
    auto pass = [&] {
      Datum out = threads > 1 ? Unwrap(cp::CallFunction("sort_indices", {chunked})) : Unwrap(cp::CallFunction("sort_indices", {keys}));
      checksum = out.length();
    };
    t = TimeIt(steps, warmup, pass);
  } else if (op == "groupby") {
    CHECK_OK(arrow::SetCpuThreadPoolCapacity(threads));
    auto keys = MakeColumn<int64_t>(arrow::int64(), n, 0.0, 7, gen_threads, [groups](uint64_t r) { return (int64_t)(r % (uint64_t)groups); });
    auto vals = MakeColumn<int64_t>(arrow::int64(), n, 0.1, 8, gen_threads, [](uint64_t r) { return (int64_t)(r % 201) - 100; });
    auto table = arrow::Table::Make(arrow::schema({arrow::field("k", arrow::int64()), arrow::field("v", arrow::int64())}), {keys, vals});
    auto pass = [&] {
      ac::Declaration plan = ac::Declaration::Sequence(
          {{"table_source", ac::TableSourceNodeOptions(table)},
           {"aggregate", ac::AggregateNodeOptions({{"hash_sum", nullptr, "v", "sum"}, {"hash_count", nullptr, "v", "count"}}, {"k"})}});
      auto out = Unwrap(ac::DeclarationToTable(std::move(plan), /*use_threads=*/threads > 1));
      checksum = out->num_rows();
    };
    t = TimeIt(steps, warmup, pass);
  } else {
    fprintf(stderr, "unknown op %s\n", op.c_str());
    return 1;
  }
  printf("{\"op\": \"%s\", \"rows\": %lld, \"steps\": %d, \"warmup\": %d, \"threads\": %d, \"hardware_threads\": %d, "
         "\"best_s\": %.6f, \"mean_s\": %.6f, \"rows_per_s_best\": %.1f, \"rows_per_s_mean\": %.1f, \"checksum\": %lld, "
         "\"cpu_model\": \"%s\", \"arrow_version\": \"%s\"}\n",
         op.c_str(), (long long)n, steps, warmup, threads, hw, t.best_s, t.mean_s, n / t.best_s, n / t.mean_s, (long long)checksum,
         CpuModel().c_str(), arrow::GetBuildInfo().version_string.c_str());
  return 0;
}
