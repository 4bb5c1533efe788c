// Host-side record IO: CRC32C, TFRecord framing, record iterators and yielders.
// Native re-design of the reference's `record_yielder.{h,cc}`,
// `sequential_record_yielder`, `weighted_mix_record_yielder`.
#pragma once

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdint>
#include <functional>
#include <memory>
#include <mutex>
#include <random>
#include <string>
#include <thread>
#include <vector>

namespace lbh {

uint32_t Crc32c(const char* data, size_t n, uint32_t crc = 0);
inline uint32_t MaskCrc(uint32_t crc) { return ((crc >> 15) | (crc << 17)) + 0xa282ead8u; }

// Expands "a*,b?" (comma separated globs) into a sorted file list.
std::vector<std::string> GlobFiles(const std::string& pattern);

struct Record {
  std::string value;
  int source_id = 0;
};

// Reads one file of a given type ("tfrecord", "text", "iota").
class RecordIterator {
 public:
  virtual ~RecordIterator() = default;
  virtual bool Next(std::string* out) = 0;
  static std::unique_ptr<RecordIterator> Create(const std::string& type, const std::string& file);
};

class TFRecordWriter {
 public:
  explicit TFRecordWriter(const std::string& path);
  ~TFRecordWriter();
  void Write(const std::string& rec);
  void Close();

 private:
  FILE* f_ = nullptr;
};

class Yielder {
 public:
  virtual ~Yielder() = default;
  // Blocks until a record is available; returns false once the data is exhausted.
  virtual bool Yield(Record* out) = 0;
  virtual void Close() {}
  virtual int64_t current_epoch() const { return 0; }
};

struct BasicYielderOptions {
  std::string file_pattern;   // "type:glob[,glob...]"
  uint64_t seed = 0;          // 0: non-deterministic
  int64_t bufsize = 16384;    // shuffle-buffer records
  int parallelism = 4;        // reader threads
  int64_t num_epochs = 0;     // 0: forever
  int source_id = 0;
  // Data-parallel input sharding (reference record_yielder.h:84-85): replica
  // `input_replica_id` of `num_input_replicas` reads every num_input_replicas-th file of
  // the sorted file list; with fewer files than replicas every replica reads all files
  // and keeps every num_input_replicas-th record instead.
  int num_input_replicas = 1;
  int input_replica_id = 0;
  // > 0: size the shuffle buffer to this many seconds of consumption (measured yield
  // rate), within [1024, bufsize] (reference `bufsize_in_seconds`).
  double bufsize_in_seconds = 0.0;
};

// Shuffling yielder: per epoch the file list is shuffled and dealt round-robin to
// `parallelism` reader threads which insert records into a shared buffer at random
// positions (random-swap insertion); consumers pop once the buffer is half full
// (or the epoch is draining).
class BasicRecordYielder : public Yielder {
 public:
  explicit BasicRecordYielder(const BasicYielderOptions& opts);
  ~BasicRecordYielder() override;
  bool Yield(Record* out) override;
  void Close() override;
  int64_t current_epoch() const override { return epoch_.load(); }

 private:
  void MainLoop();
  void ReadShard(const std::vector<std::string>& files, uint64_t seed, bool shard_records);
  void Add(std::vector<std::string>* chunk, std::mt19937_64* rng);

  BasicYielderOptions opts_;
  std::string type_, glob_;
  std::mutex mu_;
  std::condition_variable cv_not_empty_, cv_not_full_;
  std::vector<std::string> buf_;
  std::mt19937_64 pop_rng_;
  bool epoch_draining_ = false;
  bool finished_ = false;
  std::atomic<bool> stop_{false};
  std::atomic<int64_t> epoch_{0};
  std::atomic<int64_t> limit_;            // current shuffle-buffer capacity
  std::atomic<int64_t> yielded_{0};
  std::chrono::steady_clock::time_point t_first_yield_;
  std::thread main_;
};

// In-order yielder (evaluation): files sorted, records in file order.
class SequentialRecordYielder : public Yielder {
 public:
  SequentialRecordYielder(const std::string& file_pattern, int64_t repeat_count, int source_id,
                          int num_input_replicas = 1, int input_replica_id = 0);
  bool Yield(Record* out) override;
  int64_t current_epoch() const override { return epoch_; }

 private:
  std::string type_;
  std::vector<std::string> files_;
  int64_t repeat_, epoch_ = 0;
  size_t file_idx_ = 0;
  int source_id_;
  int replicas_ = 1, replica_id_ = 0;
  int64_t rec_idx_ = 0;
  std::unique_ptr<RecordIterator> it_;
  std::mutex mu_;
};

// Picks child i with probability weights[i] / sum(weights) for every record.
class WeightedMixRecordYielder : public Yielder {
 public:
  WeightedMixRecordYielder(std::vector<std::shared_ptr<Yielder>> children,
                           std::vector<double> weights, uint64_t seed);
  bool Yield(Record* out) override;
  void Close() override;

 private:
  std::vector<std::shared_ptr<Yielder>> children_;
  std::discrete_distribution<int> dist_;
  std::mt19937_64 rng_;
  std::mutex mu_;
};

}  // namespace lbh
