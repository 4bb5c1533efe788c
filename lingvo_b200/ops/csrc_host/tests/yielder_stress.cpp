// Concurrency stress of the record yielders, meant to run under ThreadSanitizer
// (tools/tsan_host.py builds it with -fsanitize=thread). No Python involved: the same
// translation units that back `_H.so` are linked into a plain executable.
//
//   yielder_stress <tmpdir>
//
// Writes a few TFRecord shards, then hammers a shuffling BasicRecordYielder (4 reader
// threads) and a WeightedMix over two of them with 6 concurrent consumers across several
// epochs, checks that every epoch delivers every record exactly once, and closes a
// yielder while consumers are still blocked in Yield().
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../records.h"

using namespace lbh;  // NOLINT

static int Fail(const char* msg) {
  fprintf(stderr, "FAIL: %s\n", msg);
  return 1;
}

int main(int argc, char** argv) {
  if (argc < 2) return Fail("usage: yielder_stress <tmpdir>");
  const std::string dir = argv[1];
  const int kShards = 6, kPerShard = 500, kEpochs = 3;
  for (int s = 0; s < kShards; ++s) {
    TFRecordWriter w(dir + "/data-" + std::to_string(s));
    for (int i = 0; i < kPerShard; ++i) w.Write("r" + std::to_string(s * kPerShard + i));
    w.Close();
  }
  {  // exactly-once per epoch under concurrent consumers
    BasicYielderOptions o;
    o.file_pattern = "tfrecord:" + dir + "/data-*";
    o.seed = 7;
    o.bufsize = 256;
    o.parallelism = 4;
    o.num_epochs = kEpochs;
    BasicRecordYielder y(o);
    std::mutex mu;
    std::map<std::string, int> seen;
    std::vector<std::thread> consumers;
    for (int c = 0; c < 6; ++c)
      consumers.emplace_back([&] {
        Record r;
        while (y.Yield(&r)) {
          std::lock_guard<std::mutex> l(mu);
          ++seen[r.value];
        }
      });
    for (auto& t : consumers) t.join();
    if (seen.size() != static_cast<size_t>(kShards * kPerShard)) return Fail("missing records");
    for (auto& kv : seen)
      if (kv.second != kEpochs) return Fail("record not seen once per epoch");
  }
  {  // weighted mix of two infinite yielders + Close() while consumers are blocked
    BasicYielderOptions a;
    a.file_pattern = "tfrecord:" + dir + "/data-0";
    a.seed = 1;
    a.bufsize = 64;
    a.parallelism = 2;
    BasicYielderOptions b = a;
    b.file_pattern = "tfrecord:" + dir + "/data-1";
    b.source_id = 1;
    std::vector<std::shared_ptr<Yielder>> kids{std::make_shared<BasicRecordYielder>(a),
                                               std::make_shared<BasicRecordYielder>(b)};
    WeightedMixRecordYielder mix(kids, {0.75, 0.25}, 3);
    std::atomic<int> got{0}, from_b{0};
    std::vector<std::thread> consumers;
    for (int c = 0; c < 4; ++c)
      consumers.emplace_back([&] {
        Record r;
        while (mix.Yield(&r)) {
          if (r.source_id == 1) ++from_b;
          if (++got >= 20000) break;
        }
      });
    while (got.load() < 20000) std::this_thread::yield();
    mix.Close();
    for (auto& t : consumers) t.join();
    const double frac = static_cast<double>(from_b.load()) / got.load();
    if (frac < 0.2 || frac > 0.3) return Fail("mix ratio off");
  }
  {  // sequential yielder shared by several readers: in-order, exactly once
    SequentialRecordYielder y("tfrecord:" + dir + "/data-*", 1, 0);
    std::atomic<int> n{0};
    std::vector<std::thread> consumers;
    for (int c = 0; c < 4; ++c)
      consumers.emplace_back([&] {
        Record r;
        while (y.Yield(&r)) ++n;
      });
    for (auto& t : consumers) t.join();
    if (n.load() != kShards * kPerShard) return Fail("sequential count");
  }
  printf("YIELDER_STRESS_OK\n");
  return 0;
}
