// pybind11 module `_H`: host-side (CPU) natives of the input pipeline.
// No torch / CUDA dependency, so it loads on any host.
#include <pybind11/functional.h>
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <thread>

#include "car_ops.h"
#include "records.h"
#include "text_ops.h"

namespace py = pybind11;
using namespace lbh;   // NOLINT

namespace {

// ------------------------------------------------------------ BucketAdjuster ----
// Re-derives the bucket upper bounds (all but the last) from a histogram of the observed
// bucket keys so that the total padding cost  Σ_buckets count(bucket) · bound(bucket)  is
// minimal (reference `BucketAdjuster`, record_batcher.h:72 / record_batcher.cc:387-390).
// Exact dynamic programme over the distinct observed keys (thinned to ≤ 2048 candidates).
class BucketAdjuster {
 public:
  BucketAdjuster(int64_t max_bucket_key, int64_t num_buckets)
      : max_key_(max_bucket_key), num_buckets_(num_buckets),
        hist_(static_cast<size_t>(max_bucket_key) + 1, 0) {}

  void IncrementHistogram(int64_t key) {
    if (key < 0 || key > max_key_) return;
    std::lock_guard<std::mutex> l(mu_);
    ++hist_[static_cast<size_t>(key)];
  }

  // Rewrites bounds[0 .. n-2]; bounds[n-1] (the hard cap) is kept.
  void AdjustBuckets(std::vector<int64_t>* bounds) {
    std::vector<std::pair<int64_t, int64_t>> pts;          // (key, count), ascending
    {
      std::lock_guard<std::mutex> l(mu_);
      for (int64_t k = 0; k <= max_key_; ++k)
        if (hist_[static_cast<size_t>(k)] > 0) pts.emplace_back(k, hist_[static_cast<size_t>(k)]);
    }
    const int64_t nb = static_cast<int64_t>(bounds->size());
    if (nb < 2 || pts.size() < 2) return;
    // thin very long supports: merge neighbours (counts move up to the larger key)
    const size_t kMax = 2048;
    if (pts.size() > kMax) {
      std::vector<std::pair<int64_t, int64_t>> thin;
      const size_t step = (pts.size() + kMax - 1) / kMax;
      for (size_t i = 0; i < pts.size(); i += step) {
        int64_t c = 0;
        const size_t hi = std::min(pts.size(), i + step);
        for (size_t j = i; j < hi; ++j) c += pts[j].second;
        thin.emplace_back(pts[hi - 1].first, c);
      }
      pts.swap(thin);
    }
    const int64_t m = static_cast<int64_t>(pts.size());
    std::vector<double> pre(m + 1, 0.0);
    for (int64_t i = 0; i < m; ++i) pre[i + 1] = pre[i] + static_cast<double>(pts[i].second);
    // cost[b][i]: min cost of covering points [0, i) with b buckets, the b-th ending at
    // point i-1 (bound = pts[i-1].first). The final bucket always ends at the hard cap.
    const int64_t free = std::min<int64_t>(nb - 1, m);      // adjustable buckets in use
    const double kInf = 1e300;
    std::vector<std::vector<double>> cost(free + 1, std::vector<double>(m + 1, kInf));
    std::vector<std::vector<int64_t>> arg(free + 1, std::vector<int64_t>(m + 1, 0));
    cost[0][0] = 0.0;
    for (int64_t b = 1; b <= free; ++b)
      for (int64_t i = b; i <= m; ++i)
        for (int64_t j = b - 1; j < i; ++j) {
          if (cost[b - 1][j] >= kInf) continue;
          const double c = cost[b - 1][j] + (pre[i] - pre[j]) * static_cast<double>(pts[i - 1].first);
          if (c < cost[b][i]) { cost[b][i] = c; arg[b][i] = j; }
        }
    // choose how many points the adjustable buckets cover; the rest pays the hard cap
    const double cap = static_cast<double>(bounds->back());
    double best = kInf;
    int64_t best_i = 0, best_b = 0;
    for (int64_t b = 0; b <= free; ++b)
      for (int64_t i = b; i <= m; ++i) {
        if (cost[b][i] >= kInf) continue;
        const double c = cost[b][i] + (pre[m] - pre[i]) * cap;
        if (c < best) { best = c; best_i = i; best_b = b; }
      }
    std::vector<int64_t> chosen;
    for (int64_t b = best_b, i = best_i; b > 0; --b) {
      chosen.push_back(pts[i - 1].first);
      i = arg[b][i];
    }
    std::sort(chosen.begin(), chosen.end());
    // keep exactly nb-1 adjustable bounds: unused ones collapse onto their neighbour
    std::vector<int64_t> out;
    for (int64_t v : chosen)
      if (v < bounds->back()) out.push_back(v);
    while (static_cast<int64_t>(out.size()) < nb - 1)
      out.push_back(out.empty() ? bounds->back() : out.back());
    std::sort(out.begin(), out.end());
    for (int64_t i = 0; i < nb - 1; ++i) (*bounds)[static_cast<size_t>(i)] = out[static_cast<size_t>(i)];
  }

 private:
  const int64_t max_key_, num_buckets_;
  std::mutex mu_;
  std::vector<int64_t> hist_;
};

// ------------------------------------------------------------- RecordBatcher ----
// N worker threads pull records from a Yielder (no GIL), call the Python
// `processor(record: bytes, source_id: int)` (GIL) which returns None (example
// filtered out) or `(bucket_key: int, [np.ndarray, ...])`, and file the sample
// under the first bucket whose upper bound >= key (over-range samples are skipped
// and counted). When a bucket reaches its batch limit it is queued for the
// consumer; `GetNext` pads every tensor slot to the per-batch max shape and stacks.
// `flush_every_n > 0` flushes all partial buckets every n processed records (exact
// evaluation); at end of data all partial buckets are flushed, then StopIteration.
class RecordBatcher {
 public:
  // `bucket_adjust_every_n > 0`: re-optimise the bounds (all but the last) every n records.
  // `fatal_errors`: None ⇒ every processor exception aborts the pipeline; a list ⇒ only
  // exceptions whose message contains one of the strings abort, all others skip the
  // record and are counted (`records_failed`) — the reference's `fatal_errors` option.
  // `pad_values`: per tensor slot padding constant (reference `dynamic_padding_constants`).
  RecordBatcher(std::shared_ptr<Yielder> yielder, py::function processor,
                std::vector<int64_t> bucket_upper_bound, std::vector<int64_t> bucket_batch_limit,
                int num_threads, int64_t flush_every_n, int64_t bucket_adjust_every_n,
                py::object fatal_errors, std::vector<double> pad_values)
      : yielder_(std::move(yielder)), processor_(std::move(processor)),
        bounds_(std::move(bucket_upper_bound)), limits_(std::move(bucket_batch_limit)),
        flush_every_n_(flush_every_n), adjust_every_n_(bucket_adjust_every_n),
        pad_values_(std::move(pad_values)), buckets_(bounds_.size()) {
    if (bounds_.empty() || bounds_.size() != limits_.size())
      throw std::runtime_error("RecordBatcher: bucket_upper_bound / bucket_batch_limit mismatch");
    if (!fatal_errors.is_none()) {
      lenient_ = true;
      for (auto item : fatal_errors) fatal_errors_.push_back(item.cast<std::string>());
    }
    if (adjust_every_n_ > 0)
      adjuster_ = std::make_unique<BucketAdjuster>(bounds_.back(), static_cast<int64_t>(bounds_.size()));
    if (num_threads < 1) num_threads = 1;
    live_workers_ = num_threads;
    for (int i = 0; i < num_threads; ++i) workers_.emplace_back([this] { Work(); });
  }

  ~RecordBatcher() { Close(); }

  void Close() {
    if (closed_.exchange(true)) return;
    stop_.store(true);
    yielder_->Close();
    {
      std::lock_guard<std::mutex> l(mu_);
      cv_ready_.notify_all();
      cv_space_.notify_all();
    }
    py::gil_scoped_release rel;
    for (auto& t : workers_)
      if (t.joinable()) t.join();
    workers_.clear();
  }

  py::tuple GetNext() {
    Batch batch;
    {
      py::gil_scoped_release rel;
      std::unique_lock<std::mutex> l(mu_);
      {
        // time the consumer spends starved: the input pipeline is the bottleneck
        const auto t0 = std::chrono::steady_clock::now();
        cv_ready_.wait(l, [&] { return stop_ || !ready_.empty() || live_workers_ == 0; });
        consumer_wait_us_ += std::chrono::duration_cast<std::chrono::microseconds>(
                                 std::chrono::steady_clock::now() - t0).count();
      }
      if (!error_.empty()) {
        std::string e = error_;
        l.unlock();
        py::gil_scoped_acquire acq;
        throw std::runtime_error(e);
      }
      if (ready_.empty()) {
        // end of data: flush whatever is left, one bucket per call
        for (auto& b : buckets_) {
          if (!b.empty()) {
            batch = std::move(b);
            b.clear();
            break;
          }
        }
      } else {
        batch = std::move(ready_.front());
        ready_.pop_front();
        cv_space_.notify_all();
      }
    }
    if (batch.empty()) throw py::stop_iteration();
    return Merge(batch);
  }

  int64_t records_skipped() const { return skipped_.load(); }
  int64_t records_failed() const { return failed_.load(); }
  std::vector<int64_t> bucket_upper_bound() {
    std::lock_guard<std::mutex> l(mu_);
    return bounds_;
  }
  int64_t records_processed() const { return processed_.load(); }
  // Wait-time diagnostics (the reference logs these from `record_debug.cc`).
  py::dict Stats() {
    std::lock_guard<std::mutex> l(mu_);
    py::dict d;
    d["records_processed"] = processed_.load();
    d["records_skipped"] = skipped_.load();
    d["records_failed"] = failed_.load();
    d["batches_ready"] = static_cast<int64_t>(ready_.size());
    d["consumer_wait_s"] = consumer_wait_us_ * 1e-6;
    d["producer_wait_s"] = producer_wait_us_ * 1e-6;
    d["hint"] = consumer_wait_us_ > 2 * producer_wait_us_
                    ? "consumer starved: raise num_threads / file parallelism or speed up the processor"
                    : "producers blocked: batches are consumed slower than they are made";
    return d;
  }

 private:
  struct Sample {
    int64_t key;
    std::vector<py::array> tensors;
  };
  using Batch = std::vector<Sample>;

  void Work() {
    Record rec;
    while (!stop_) {
      if (!yielder_->Yield(&rec)) break;
      Sample s;
      bool keep = false;
      {
        py::gil_scoped_acquire acq;
        try {
          py::object ret = processor_(py::bytes(rec.value), rec.source_id);
          if (!ret.is_none()) {
            auto tup = ret.cast<py::tuple>();
            s.key = tup[0].cast<int64_t>();
            for (auto item : tup[1]) s.tensors.push_back(py::array::ensure(item));
            keep = true;
          }
        } catch (const std::exception& e) {
          const std::string what = e.what();
          bool fatal = !lenient_;
          for (const auto& f : fatal_errors_)
            if (what.find(f) != std::string::npos) fatal = true;
          if (fatal) {
            std::lock_guard<std::mutex> l(mu_);
            error_ = std::string("RecordBatcher processor failed: ") + what;
            stop_.store(true);
            cv_ready_.notify_all();
          } else {
            failed_.fetch_add(1);
          }
        }
        if (!keep) s.tensors.clear();   // drop references while holding the GIL
      }
      if (!keep) continue;
      const int64_t n = processed_.fetch_add(1) + 1;
      if (adjuster_) adjuster_->IncrementHistogram(s.key);
      std::unique_lock<std::mutex> l(mu_);
      if (adjuster_ && n % adjust_every_n_ == 0) {
        // Samples already filed keep their bucket; a bucket is only ever *larger* than the
        // keys it holds if its bound shrank, so flush everything before re-bucketing.
        for (auto& b : buckets_) {
          if (!b.empty()) {
            ready_.push_back(std::move(b));
            b.clear();
          }
        }
        cv_ready_.notify_all();
        adjuster_->AdjustBuckets(&bounds_);
      }
      const auto it = std::lower_bound(bounds_.begin(), bounds_.end(), s.key);
      if (it == bounds_.end()) {
        skipped_.fetch_add(1);
        l.unlock();
        py::gil_scoped_acquire acq;
        s.tensors.clear();
        continue;
      }
      const size_t bi = static_cast<size_t>(it - bounds_.begin());
      {
        // time producers spend blocked on a full queue: the trainer is the bottleneck
        const auto t0 = std::chrono::steady_clock::now();
        cv_space_.wait(l, [&] { return stop_ || ready_.size() < 4; });
        producer_wait_us_ += std::chrono::duration_cast<std::chrono::microseconds>(
                                 std::chrono::steady_clock::now() - t0).count();
      }
      if (stop_) {
        // `s` owns numpy arrays: they must be released under the GIL, not by this thread's
        // stack unwinding.
        l.unlock();
        py::gil_scoped_acquire acq;
        s.tensors.clear();
        break;
      }
      buckets_[bi].push_back(std::move(s));
      if (static_cast<int64_t>(buckets_[bi].size()) >= limits_[bi]) {
        ready_.push_back(std::move(buckets_[bi]));
        buckets_[bi].clear();
        cv_ready_.notify_one();
      }
      if (flush_every_n_ > 0 && n % flush_every_n_ == 0) {
        for (auto& b : buckets_) {
          if (!b.empty()) {
            ready_.push_back(std::move(b));
            b.clear();
          }
        }
        cv_ready_.notify_all();
      }
    }
    std::lock_guard<std::mutex> l(mu_);
    --live_workers_;
    cv_ready_.notify_all();
  }

  static void CopyInto(char* dst, const std::vector<py::ssize_t>& dst_strides, const char* src,
                       const py::buffer_info& info, int dim) {
    if (dim == info.ndim) {
      memcpy(dst, src, info.itemsize);
      return;
    }
    if (dim == info.ndim - 1 && info.strides[dim] == info.itemsize) {
      memcpy(dst, src, info.itemsize * info.shape[dim]);
      return;
    }
    for (py::ssize_t i = 0; i < info.shape[dim]; ++i)
      CopyInto(dst + i * dst_strides[dim], dst_strides, src + i * info.strides[dim], info, dim + 1);
  }

  py::tuple Merge(Batch& batch) {
    const size_t n = batch.size();
    const size_t slots = batch[0].tensors.size();
    py::array_t<int64_t> keys(static_cast<py::ssize_t>(n));
    for (size_t i = 0; i < n; ++i) keys.mutable_at(i) = batch[i].key;
    py::list outs;
    for (size_t s = 0; s < slots; ++s) {
      const py::array& first = batch[0].tensors[s];
      const int nd = static_cast<int>(first.ndim());
      std::vector<py::ssize_t> shape(nd + 1, 0);
      shape[0] = static_cast<py::ssize_t>(n);
      for (const auto& smp : batch) {
        if (smp.tensors.size() != slots || smp.tensors[s].ndim() != nd ||
            !smp.tensors[s].dtype().is(first.dtype()))
          throw std::runtime_error("RecordBatcher: samples disagree on tensor rank/dtype");
        for (int d = 0; d < nd; ++d) shape[d + 1] = std::max(shape[d + 1], smp.tensors[s].shape(d));
      }
      py::array out(first.dtype(), shape);
      memset(out.mutable_data(), 0, static_cast<size_t>(out.nbytes()));
      if (s < pad_values_.size() && pad_values_[s] != 0.0) {
        // dynamic padding constant of this slot (numpy fills with the proper dtype)
        out.attr("fill")(pad_values_[s]);
      }
      std::vector<py::ssize_t> inner_strides(out.strides() + 1, out.strides() + 1 + nd);
      for (size_t i = 0; i < n; ++i) {
        py::buffer_info info = batch[i].tensors[s].request();
        if (info.size == 0) continue;
        CopyInto(static_cast<char*>(out.mutable_data()) + i * out.strides(0), inner_strides,
                 static_cast<const char*>(info.ptr), info, 0);
      }
      outs.append(out);
    }
    batch.clear();
    return py::make_tuple(keys, outs);
  }

  std::shared_ptr<Yielder> yielder_;
  py::function processor_;
  std::vector<int64_t> bounds_, limits_;
  int64_t flush_every_n_;
  int64_t adjust_every_n_ = 0;
  std::unique_ptr<BucketAdjuster> adjuster_;
  bool lenient_ = false;
  std::vector<std::string> fatal_errors_;
  std::vector<double> pad_values_;
  std::atomic<int64_t> failed_{0};
  std::vector<Batch> buckets_;
  std::deque<Batch> ready_;
  std::mutex mu_;
  std::condition_variable cv_ready_, cv_space_;
  std::vector<std::thread> workers_;
  int live_workers_ = 0;
  std::atomic<bool> stop_{false}, closed_{false};
  std::atomic<int64_t> skipped_{0}, processed_{0};
  int64_t consumer_wait_us_ = 0, producer_wait_us_ = 0;   // guarded by mu_
  std::string error_;
};

py::array_t<int32_t> ToArray2D(const std::vector<int32_t>& v, int rows, int cols) {
  py::array_t<int32_t> a({rows, cols});
  memcpy(a.mutable_data(), v.data(), v.size() * sizeof(int32_t));
  return a;
}

std::vector<int32_t> ToVec(const py::array_t<int32_t, py::array::c_style | py::array::forcecast>& a) {
  return std::vector<int32_t>(a.data(), a.data() + a.size());
}

}  // namespace

PYBIND11_MODULE(_H, m) {
  m.doc() = "lingvo_b200 host-side natives (records, batching, tokenizers, packing)";
  m.def("crc32c", [](py::bytes b) { std::string s = b; return Crc32c(s.data(), s.size()); });
  // Zero-copy CRC over any contiguous buffer (numpy arrays of checkpoint tensors); the
  // GIL is released so an async checkpoint thread does not stall the training loop.
  m.def("crc32c_buffer", [](py::buffer b, uint32_t crc) {
    py::buffer_info info = b.request();
    const char* ptr = static_cast<const char*>(info.ptr);
    const size_t n = static_cast<size_t>(info.size) * static_cast<size_t>(info.itemsize);
    uint32_t out;
    {
      py::gil_scoped_release release;
      out = Crc32c(ptr, n, crc);
    }
    return out;
  }, py::arg("buffer"), py::arg("crc") = 0);
  m.def("masked_crc32c", [](py::bytes b) { std::string s = b; return MaskCrc(Crc32c(s.data(), s.size())); });
  m.def("glob_files", &GlobFiles);

  py::class_<TFRecordWriter>(m, "TFRecordWriter")
      .def(py::init<const std::string&>())
      .def("write", [](TFRecordWriter& w, py::bytes b) { w.Write(b); })
      .def("close", &TFRecordWriter::Close);

  py::class_<Yielder, std::shared_ptr<Yielder>>(m, "Yielder")
      .def("next", [](Yielder& y) -> py::object {
        Record r;
        bool ok;
        {
          py::gil_scoped_release rel;
          ok = y.Yield(&r);
        }
        if (!ok) return py::none();
        return py::make_tuple(py::bytes(r.value), r.source_id);
      })
      .def("close", [](Yielder& y) { py::gil_scoped_release rel; y.Close(); })
      .def_property_readonly("current_epoch", &Yielder::current_epoch);

  m.def("basic_record_yielder",
        [](const std::string& file_pattern, uint64_t seed, int64_t bufsize, int parallelism,
           int64_t num_epochs, int source_id, int num_input_replicas, int input_replica_id,
           double bufsize_in_seconds) -> std::shared_ptr<Yielder> {
          BasicYielderOptions o;
          o.file_pattern = file_pattern; o.seed = seed; o.bufsize = bufsize;
          o.parallelism = parallelism; o.num_epochs = num_epochs; o.source_id = source_id;
          o.num_input_replicas = num_input_replicas; o.input_replica_id = input_replica_id;
          o.bufsize_in_seconds = bufsize_in_seconds;
          return std::make_shared<BasicRecordYielder>(o);
        },
        py::arg("file_pattern"), py::arg("seed") = 0, py::arg("bufsize") = 16384,
        py::arg("parallelism") = 4, py::arg("num_epochs") = 0, py::arg("source_id") = 0,
        py::arg("num_input_replicas") = 1, py::arg("input_replica_id") = 0,
        py::arg("bufsize_in_seconds") = 0.0);
  m.def("sequential_record_yielder",
        [](const std::string& fp, int64_t repeat, int source_id, int num_input_replicas,
           int input_replica_id) -> std::shared_ptr<Yielder> {
          return std::make_shared<SequentialRecordYielder>(fp, repeat, source_id,
                                                           num_input_replicas, input_replica_id);
        },
        py::arg("file_pattern"), py::arg("repeat_count") = 1, py::arg("source_id") = 0,
        py::arg("num_input_replicas") = 1, py::arg("input_replica_id") = 0);
  py::class_<BucketAdjuster>(m, "BucketAdjuster")
      .def(py::init<int64_t, int64_t>(), py::arg("max_bucket_key"), py::arg("num_buckets"))
      .def("increment_histogram", &BucketAdjuster::IncrementHistogram)
      .def("adjust_buckets", [](BucketAdjuster& a, std::vector<int64_t> bounds) {
        a.AdjustBuckets(&bounds);
        return bounds;
      });
  m.def("weighted_mix_record_yielder",
        [](std::vector<std::shared_ptr<Yielder>> kids, std::vector<double> w,
           uint64_t seed) -> std::shared_ptr<Yielder> {
          return std::make_shared<WeightedMixRecordYielder>(std::move(kids), std::move(w), seed);
        },
        py::arg("children"), py::arg("weights"), py::arg("seed") = 0);

  py::class_<RecordBatcher>(m, "RecordBatcher")
      .def(py::init<std::shared_ptr<Yielder>, py::function, std::vector<int64_t>,
                    std::vector<int64_t>, int, int64_t, int64_t, py::object,
                    std::vector<double>>(),
           py::arg("yielder"), py::arg("processor"), py::arg("bucket_upper_bound"),
           py::arg("bucket_batch_limit"), py::arg("num_threads") = 4,
           py::arg("flush_every_n") = 0, py::arg("bucket_adjust_every_n") = 0,
           py::arg("fatal_errors") = py::none(), py::arg("pad_values") = std::vector<double>())
      .def("get_next", &RecordBatcher::GetNext)
      .def("close", &RecordBatcher::Close)
      .def_property_readonly("records_skipped", &RecordBatcher::records_skipped)
      .def_property_readonly("records_failed", &RecordBatcher::records_failed)
      .def_property_readonly("bucket_upper_bound", &RecordBatcher::bucket_upper_bound)
      .def_property_readonly("records_processed", &RecordBatcher::records_processed)
      .def("stats", &RecordBatcher::Stats);

  // ---- tokenizers ----
  m.def("ascii_to_ids", [](const std::string& s) { return AsciiTokenizer::Get().StringToIds(s); });
  m.def("ascii_to_string",
        [](const std::vector<int32_t>& ids) { return AsciiTokenizer::Get().IdsToString(ids); });
  m.def("ascii_num_tokens", []() { return AsciiTokenizer::Get().NumTokens(); });
  py::class_<VocabTokenizer>(m, "VocabTokenizer")
      .def(py::init<const std::string&, bool>(), py::arg("vocab_path"),
           py::arg("load_token_ids_from_vocab") = false)
      .def("to_ids", &VocabTokenizer::StringToIds)
      .def("to_string", &VocabTokenizer::IdsToString)
      .def("token_to_id", &VocabTokenizer::TokenToId)
      .def("id_to_token", &VocabTokenizer::IdToToken)
      .def_property_readonly("unk_id", &VocabTokenizer::unk_id)
      .def_property_readonly("sos_id", &VocabTokenizer::sos_id)
      .def_property_readonly("eos_id", &VocabTokenizer::eos_id)
      .def("join_ids", &VocabTokenizer::JoinIds, py::arg("ids"), py::arg("separator") = "")
      .def("__contains__", &VocabTokenizer::Contains)
      .def("__len__", &VocabTokenizer::size);
  py::class_<MlPerfSubword>(m, "MlPerfSubword")
      .def(py::init<const std::string&>(), py::arg("vocab_path"))
      .def(py::init<const std::vector<std::string>&>(), py::arg("lines"))
      .def("decode", &MlPerfSubword::Decode)
      .def("__len__", &MlPerfSubword::size);
  py::class_<StaticMap<int64_t, int64_t>>(m, "StaticMapIntInt")
      .def(py::init<const std::vector<int64_t>&, const std::vector<int64_t>&, int64_t>(),
           py::arg("keys"), py::arg("vals") = std::vector<int64_t>(), py::arg("unk") = -1)
      .def("lookup", &StaticMap<int64_t, int64_t>::Lookup)
      .def("__len__", &StaticMap<int64_t, int64_t>::size);
  py::class_<StaticMap<int64_t, std::string>>(m, "StaticMapIntString")
      .def(py::init<const std::vector<int64_t>&, const std::vector<std::string>&, std::string>(),
           py::arg("keys"), py::arg("vals"), py::arg("unk") = "")
      .def("lookup", &StaticMap<int64_t, std::string>::Lookup)
      .def("__len__", &StaticMap<int64_t, std::string>::size);
  py::class_<StaticMap<std::string, int64_t>>(m, "StaticMapStringInt")
      .def(py::init<const std::vector<std::string>&, const std::vector<int64_t>&, int64_t>(),
           py::arg("keys"), py::arg("vals") = std::vector<int64_t>(), py::arg("unk") = -1)
      .def("lookup", &StaticMap<std::string, int64_t>::Lookup)
      .def("__len__", &StaticMap<std::string, int64_t>::size);
  py::class_<BpeTokenizer>(m, "BpeTokenizer")
      .def(py::init<const std::string&, const std::string&>(), py::arg("codes_path"),
           py::arg("vocab_path"))
      .def("to_ids", &BpeTokenizer::StringToIds)
      .def("to_string", &BpeTokenizer::IdsToString)
      .def("encode_word", &BpeTokenizer::EncodeWord);

  // ---- packing ----
  m.def("pack_sequences",
        [](py::array_t<int32_t, py::array::c_style | py::array::forcecast> src_lens,
           py::array_t<int32_t, py::array::c_style | py::array::forcecast> tgt_lens,
           int packed_batch_size, int packed_src_seq_len, int packed_tgt_seq_len, uint64_t seed) {
          PackResult r = PackSequences(ToVec(src_lens), ToVec(tgt_lens), packed_batch_size,
                                       packed_src_seq_len, packed_tgt_seq_len, seed);
          return py::make_tuple(ToArray2D(r.src_segment_ids, r.rows, r.src_len),
                                ToArray2D(r.src_segment_pos, r.rows, r.src_len),
                                ToArray2D(r.src_indices_in_input, r.rows, r.src_len),
                                ToArray2D(r.tgt_segment_ids, r.rows, r.tgt_len),
                                ToArray2D(r.tgt_segment_pos, r.rows, r.tgt_len),
                                ToArray2D(r.tgt_indices_in_input, r.rows, r.tgt_len));
        },
        py::arg("src_actual_seq_len"), py::arg("tgt_actual_seq_len"), py::arg("packed_batch_size"),
        py::arg("packed_src_seq_len"), py::arg("packed_tgt_seq_len"), py::arg("seed") = 0);
  // Gathers [N, T, …] rows into the packed [B, L, …] layout (any dtype: rows move as bytes).
  m.def("apply_packing",
        [](py::array inputs, py::array padding,
           py::array_t<int32_t, py::array::c_style | py::array::forcecast> seg,
           py::array_t<int32_t, py::array::c_style | py::array::forcecast> idx) {
          if (inputs.ndim() < 2) throw std::invalid_argument("apply_packing: inputs must be [N, T, ...]");
          if (!(inputs.flags() & py::array::c_style)) throw std::invalid_argument("apply_packing: inputs must be C-contiguous");
          if (seg.ndim() != 2 || idx.ndim() != 2 || seg.shape(0) != idx.shape(0) || seg.shape(1) != idx.shape(1))
            throw std::invalid_argument("apply_packing: segment_ids / indices_in_input must both be [B, L]");
          const int64_t n = inputs.shape(0), t = inputs.shape(1), b = seg.shape(0), l = seg.shape(1);
          size_t inner = static_cast<size_t>(inputs.itemsize());
          std::vector<py::ssize_t> shape{b, l};
          for (int d = 2; d < inputs.ndim(); ++d) {
            inner *= static_cast<size_t>(inputs.shape(d));
            shape.push_back(inputs.shape(d));
          }
          if (static_cast<size_t>(padding.nbytes()) != static_cast<size_t>(inputs.itemsize()))
            throw std::invalid_argument("apply_packing: padding must be one element of the input dtype");
          py::array out(inputs.dtype(), shape);
          char* o = static_cast<char*>(out.mutable_data());
          const char* in = static_cast<const char*>(inputs.data());
          const char* pad = static_cast<const char*>(padding.data());
          const size_t isz = static_cast<size_t>(inputs.itemsize());
          const int32_t* sp = seg.data();
          const int32_t* ip = idx.data();
          std::string err;
          {
            py::gil_scoped_release rel;
            for (int64_t r = 0; r < b && err.empty(); ++r) {
              int32_t prev_seg = 0, prev_idx = -1;
              int64_t run = 0;
              for (int64_t c = 0; c < l; ++c) {
                char* dst = o + (r * l + c) * inner;
                const int32_t sg = sp[r * l + c], ix = ip[r * l + c];
                if (sg <= 0) {
                  for (size_t k = 0; k < inner; k += isz) memcpy(dst + k, pad, isz);
                  prev_seg = 0;
                  continue;
                }
                run = (sg == prev_seg && ix == prev_idx) ? run + 1 : 0;
                prev_seg = sg;
                prev_idx = ix;
                if (ix < 0 || ix >= n || run >= t) {
                  err = "apply_packing: index out of range at row " + std::to_string(r);
                  break;
                }
                memcpy(dst, in + (static_cast<size_t>(ix) * t + run) * inner, inner);
              }
            }
          }
          if (!err.empty()) throw std::out_of_range(err);
          return out;
        },
        py::arg("inputs"), py::arg("padding"), py::arg("segment_ids"), py::arg("indices_in_input"));
  m.def("pack_single_sequence",
        [](py::array_t<int32_t, py::array::c_style | py::array::forcecast> lens, int cap,
           bool sequential) { return PackSingleSequence(ToVec(lens), cap, sequential); },
        py::arg("input_lengths"), py::arg("max_packed_length"),
        py::arg("require_sequential_order") = false);

  m.def("mass",
        [](py::array_t<int32_t, py::array::c_style | py::array::forcecast> ids,
           py::array_t<float, py::array::c_style | py::array::forcecast> weights,
           py::array_t<int32_t, py::array::c_style | py::array::forcecast> lens, int32_t mask_id,
           float mask_ratio, int mask_minlen, int span_len, float random_start_prob,
           float keep_prob, float rand_prob, float mask_prob, bool mask_target, int vocab_size,
           int first_unreserved_id, uint64_t seed) {
          const int batch = static_cast<int>(ids.shape(0)), max_len = static_cast<int>(ids.shape(1));
          MassOptions o;
          o.mask_id = mask_id; o.mask_ratio = mask_ratio; o.mask_minlen = mask_minlen;
          o.span_len = span_len; o.random_start_prob = random_start_prob; o.keep_prob = keep_prob;
          o.rand_prob = rand_prob; o.mask_prob = mask_prob; o.mask_target = mask_target;
          o.vocab_size = vocab_size; o.first_unreserved_id = first_unreserved_id;
          MassResult r = Mass(ToVec(ids), std::vector<float>(weights.data(), weights.data() + weights.size()),
                              ToVec(lens), batch, max_len, o, seed);
          py::array_t<float> w({batch, max_len});
          memcpy(w.mutable_data(), r.tgt_weights.data(), r.tgt_weights.size() * sizeof(float));
          return py::make_tuple(ToArray2D(r.src_ids, batch, max_len), ToArray2D(r.tgt_ids, batch, max_len),
                                ToArray2D(r.tgt_labels, batch, max_len), w);
        },
        py::arg("ids"), py::arg("weights"), py::arg("actual_seq_len"), py::arg("mask_id") = 3,
        py::arg("mask_ratio") = 0.5f, py::arg("mask_minlen") = 0, py::arg("span_len") = 100000,
        py::arg("random_start_prob") = 0.6f, py::arg("keep_prob") = 0.1f, py::arg("rand_prob") = 0.1f,
        py::arg("mask_prob") = 0.8f, py::arg("mask_target") = true, py::arg("vocab_size") = 0,
        py::arg("first_unreserved_id") = 4, py::arg("seed") = 0);

  m.def("best_step", &BestStep, py::arg("hist_file"), py::arg("tol") = 0.0,
        py::arg("minimize") = true);

  // ---- 3-D detection geometry ----
  using FArr = py::array_t<float, py::array::c_style | py::array::forcecast>;
  m.def("pairwise_iou_3d", [](FArr a, FArr b) {
    const int n = static_cast<int>(a.shape(0)), k = static_cast<int>(b.shape(0));
    py::array_t<float> out({n, k});
    {
      py::gil_scoped_release rel;
      PairwiseIou3D(a.data(), n, b.data(), k, out.mutable_data());
    }
    return out;
  });
  m.def("nms_3d",
        [](FArr boxes, FArr scores, std::vector<float> nms_iou, std::vector<float> score_thresh,
           int max_boxes) {
          const int n = static_cast<int>(boxes.shape(0));
          const int c = scores.ndim() == 2 ? static_cast<int>(scores.shape(1)) : 1;
          auto idx = NonMaxSuppression3D(boxes.data(), scores.data(), n, c, nms_iou, score_thresh,
                                         max_boxes);
          py::array_t<int32_t> out({c, max_boxes});
          memcpy(out.mutable_data(), idx.data(), idx.size() * sizeof(int32_t));
          return out;
        },
        py::arg("boxes"), py::arg("scores"), py::arg("nms_iou_threshold"),
        py::arg("score_threshold"), py::arg("max_boxes_per_class"));
  m.def("average_precision_3d",
        [](float iou_threshold, FArr gt_bbox, py::array_t<int32_t, py::array::c_style | py::array::forcecast> gt_imageid,
           py::array_t<int32_t, py::array::c_style | py::array::forcecast> gt_ignore, FArr pd_bbox,
           py::array_t<int32_t, py::array::c_style | py::array::forcecast> pd_imageid,
           py::array_t<int32_t, py::array::c_style | py::array::forcecast> pd_ignore, FArr pd_score,
           int num_recall_points, const std::string& algorithm) {
          const int n = static_cast<int>(gt_bbox.shape(0)), k = static_cast<int>(pd_bbox.shape(0));
          ApResult r;
          {
            py::gil_scoped_release rel;
            r = AveragePrecision3D(iou_threshold, gt_bbox.data(), gt_imageid.data(), gt_ignore.data(),
                                   n, pd_bbox.data(), pd_imageid.data(), pd_ignore.data(),
                                   pd_score.data(), k, num_recall_points, algorithm == "KITTI");
          }
          py::array_t<float> pr({num_recall_points, 2});
          memcpy(pr.mutable_data(), r.precision_recall.data(), r.precision_recall.size() * 4);
          py::array_t<float> sh({k, 2});
          if (k) memcpy(sh.mutable_data(), r.score_and_hit.data(), r.score_and_hit.size() * 4);
          return py::make_tuple(r.average_precision, pr, sh);
        },
        py::arg("iou_threshold"), py::arg("groundtruth_bbox"), py::arg("groundtruth_imageid"),
        py::arg("groundtruth_ignore"), py::arg("prediction_bbox"), py::arg("prediction_imageid"),
        py::arg("prediction_ignore"), py::arg("prediction_score"),
        py::arg("num_recall_points") = 1, py::arg("algorithm") = "KITTI");
  m.def("average_precision_2d",
        [](float iou_threshold, FArr gt_bbox, py::array_t<int32_t, py::array::c_style | py::array::forcecast> gt_imageid,
           py::array_t<int32_t, py::array::c_style | py::array::forcecast> gt_ignore, FArr pd_bbox,
           py::array_t<int32_t, py::array::c_style | py::array::forcecast> pd_imageid,
           py::array_t<int32_t, py::array::c_style | py::array::forcecast> pd_ignore, FArr pd_score,
           int num_recall_points, const std::string& algorithm) {
          const int n = static_cast<int>(gt_bbox.shape(0)), k = static_cast<int>(pd_bbox.shape(0));
          ApResult r;
          {
            py::gil_scoped_release rel;
            r = AveragePrecision2D(iou_threshold, gt_bbox.data(), gt_imageid.data(), gt_ignore.data(),
                                   n, pd_bbox.data(), pd_imageid.data(), pd_ignore.data(),
                                   pd_score.data(), k, num_recall_points, algorithm == "KITTI");
          }
          py::array_t<float> pr({num_recall_points, 2});
          memcpy(pr.mutable_data(), r.precision_recall.data(), r.precision_recall.size() * 4);
          py::array_t<float> sh({k, 2});
          if (k) memcpy(sh.mutable_data(), r.score_and_hit.data(), r.score_and_hit.size() * 4);
          return py::make_tuple(r.average_precision, pr, sh);
        },
        py::arg("iou_threshold"), py::arg("groundtruth_bbox"), py::arg("groundtruth_imageid"),
        py::arg("groundtruth_ignore"), py::arg("prediction_bbox"), py::arg("prediction_imageid"),
        py::arg("prediction_ignore"), py::arg("prediction_score"),
        py::arg("num_recall_points") = 1, py::arg("algorithm") = "VOC");
  // Batched point sampling: scenes are independent, so they are spread over threads.
  m.def("sample_points",
        [](FArr points, FArr padding, int num_seeded, const std::string& center_selector,
           const std::string& neighbor_sampler, const std::string& neighbor_algorithm,
           int num_centers, float center_z_min, float center_z_max, int num_neighbors,
           float max_distance, int64_t random_seed) {
          if (points.ndim() != 3 || padding.ndim() != 2)
            throw std::invalid_argument("sample_points: points [B,N,D>=3], padding [B,N]");
          const int b = static_cast<int>(points.shape(0)), n = static_cast<int>(points.shape(1)),
                    d = static_cast<int>(points.shape(2));
          if (d < 3) throw std::invalid_argument("sample_points: need xyz");
          if (center_selector != "farthest" && center_selector != "uniform")
            throw std::invalid_argument("center_selector must be farthest|uniform");
          if (neighbor_sampler != "closest" && neighbor_sampler != "uniform")
            throw std::invalid_argument("neighbor_sampler must be closest|uniform");
          if (neighbor_algorithm != "auto" && neighbor_algorithm != "hash")
            throw std::invalid_argument("neighbor_algorithm must be auto|hash");
          SampleOptions o;
          o.farthest = center_selector == "farthest";
          o.closest = neighbor_sampler == "closest";
          o.use_hash = neighbor_algorithm == "hash";
          o.num_centers = num_centers;
          o.num_neighbors = num_neighbors;
          o.center_z_min = center_z_min;
          o.center_z_max = center_z_max;
          o.max_dist = max_distance;
          py::array_t<int32_t> center({b, num_centers}), idx({b, num_centers, num_neighbors});
          py::array_t<float> cpad({b, num_centers}), ipad({b, num_centers, num_neighbors});
          const float* pp = points.data();
          const float* pd = padding.data();
          int32_t* c_out = center.mutable_data();
          int32_t* i_out = idx.mutable_data();
          float* cp_out = cpad.mutable_data();
          float* ip_out = ipad.mutable_data();
          {
            py::gil_scoped_release rel;
            std::atomic<int> next{0};
            auto work = [&] {
              for (int i = next.fetch_add(1); i < b; i = next.fetch_add(1)) {
                SampleOptions oi = o;
                oi.seed = random_seed >= 0 ? random_seed + i : -1;
                SampleResult r = SamplePoints(pp + static_cast<size_t>(i) * n * d,
                                              pd + static_cast<size_t>(i) * n, n, d, num_seeded, oi);
                const size_t mk = static_cast<size_t>(num_centers) * num_neighbors;
                memcpy(c_out + static_cast<size_t>(i) * num_centers, r.center.data(), num_centers * 4);
                memcpy(cp_out + static_cast<size_t>(i) * num_centers, r.center_padding.data(), num_centers * 4);
                memcpy(i_out + i * mk, r.indices.data(), mk * 4);
                memcpy(ip_out + i * mk, r.indices_padding.data(), mk * 4);
              }
            };
            const int nt = std::max(1, std::min<int>(b, static_cast<int>(std::thread::hardware_concurrency())));
            std::vector<std::thread> ts;
            for (int t = 1; t < nt; ++t) ts.emplace_back(work);
            work();
            for (auto& t : ts) t.join();
          }
          return py::make_tuple(center, cpad, idx, ipad);
        },
        py::arg("points"), py::arg("points_padding"), py::arg("num_seeded_points") = 0,
        py::arg("center_selector") = "farthest", py::arg("neighbor_sampler") = "closest",
        py::arg("neighbor_algorithm") = "auto", py::arg("num_centers") = 1,
        py::arg("center_z_min") = -3.4e38f, py::arg("center_z_max") = 3.4e38f,
        py::arg("num_neighbors") = 1, py::arg("max_distance") = 3.4e38f,
        py::arg("random_seed") = -1);
  m.def("points_to_pillars",
        [](FArr points, float x0, float x1, float y0, float y1, int nx, int ny, int max_pillars,
           int points_per_pillar) {
          const int n = static_cast<int>(points.shape(0)), d = static_cast<int>(points.shape(1));
          py::array_t<float> pp({max_pillars, points_per_pillar, d});
          py::array_t<int32_t> xy({max_pillars, 2}), cnt({max_pillars});
          memset(pp.mutable_data(), 0, static_cast<size_t>(pp.nbytes()));
          memset(xy.mutable_data(), 0, static_cast<size_t>(xy.nbytes()));
          memset(cnt.mutable_data(), 0, static_cast<size_t>(cnt.nbytes()));
          const int used = PointsToPillars(points.data(), n, d, x0, x1, y0, y1, nx, ny, max_pillars,
                                           points_per_pillar, pp.mutable_data(), xy.mutable_data(),
                                           cnt.mutable_data());
          return py::make_tuple(pp, xy, cnt, used);
        });
  m.def("farthest_point_sample", [](FArr xyz, int k) {
    return FarthestPointSample(xyz.data(), static_cast<int>(xyz.shape(0)), k);
  });

  py::class_<RandomPermutationSequence>(m, "RandomPermutationSequence")
      .def(py::init<int64_t, int64_t, bool, uint64_t>(), py::arg("num"), py::arg("batch"),
           py::arg("repeat"), py::arg("seed") = 0)
      .def("next", &RandomPermutationSequence::Next);
}
