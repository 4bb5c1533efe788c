#include "records.h"

#include <zlib.h>

#include <glob.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>
#include <stdexcept>

namespace lbh {

// ------------------------------------------------------------------ crc32c ----
namespace {
struct CrcTable {
  uint32_t t[8][256];
  CrcTable() {
    for (uint32_t i = 0; i < 256; ++i) {
      uint32_t c = i;
      for (int k = 0; k < 8; ++k) c = (c & 1) ? (c >> 1) ^ 0x82f63b78u : c >> 1;
      t[0][i] = c;
    }
    for (uint32_t i = 0; i < 256; ++i)
      for (int s = 1; s < 8; ++s) t[s][i] = (t[s - 1][i] >> 8) ^ t[0][t[s - 1][i] & 0xff];
  }
};
const CrcTable& Table() {
  static const CrcTable tab;
  return tab;
}
}  // namespace

#if defined(__x86_64__)
// Hardware CRC32C (SSE4.2 `crc32` instruction): three independent streams hide the
// 3-cycle latency, their CRCs are merged with precomputed "shift by N zero bytes"
// operators. ~15-20 GB/s per core vs ~1.5 GB/s for slicing-by-8 — checkpoint bundles
// checksum tens of GB.
__attribute__((target("sse4.2"))) static uint32_t HwCrcBlock(const uint8_t* p, size_t n,
                                                               uint32_t c) {
  while (n >= 8) {
    uint64_t v;
    memcpy(&v, p, 8);
    c = static_cast<uint32_t>(__builtin_ia32_crc32di(c, v));
    p += 8;
    n -= 8;
  }
  while (n--) c = __builtin_ia32_crc32qi(c, *p++);
  return c;
}

// GF(2) 32x32 matrix helpers (zlib's crc32_combine construction, CRC32C polynomial).
static uint32_t GfTimes(const uint32_t* mat, uint32_t vec) {
  uint32_t sum = 0;
  while (vec) {
    if (vec & 1) sum ^= *mat;
    vec >>= 1;
    ++mat;
  }
  return sum;
}
static void GfSquare(uint32_t* sq, const uint32_t* mat) {
  for (int i = 0; i < 32; ++i) sq[i] = GfTimes(mat, mat[i]);
}
// Operator that advances a (raw, un-inverted) CRC register over `len` zero bytes.
struct ZeroShift {
  uint32_t m[32];
  explicit ZeroShift(size_t len) {
    uint32_t even[32], odd[32];
    odd[0] = 0x82f63b78u;
    uint32_t row = 1;
    for (int i = 1; i < 32; ++i) {
      odd[i] = row;
      row <<= 1;
    }
    GfSquare(even, odd);
    GfSquare(odd, even);
    // identity
    for (int i = 0; i < 32; ++i) m[i] = 1u << i;
    uint32_t tmp[32];
    do {
      GfSquare(even, odd);
      if (len & 1) {
        for (int i = 0; i < 32; ++i) tmp[i] = GfTimes(even, m[i]);
        memcpy(m, tmp, sizeof(m));
      }
      len >>= 1;
      if (!len) break;
      GfSquare(odd, even);
      if (len & 1) {
        for (int i = 0; i < 32; ++i) tmp[i] = GfTimes(odd, m[i]);
        memcpy(m, tmp, sizeof(m));
      }
      len >>= 1;
    } while (len);
  }
  uint32_t Apply(uint32_t c) const { return GfTimes(m, c); }
};

__attribute__((target("sse4.2"))) static uint32_t HwCrc32c(const uint8_t* p, size_t n,
                                                             uint32_t crc) {
  constexpr size_t kLane = 8192;                 // bytes per stream per round
  static const ZeroShift shift(kLane);
  uint32_t c = ~crc;
  while (n >= 3 * kLane) {
    uint32_t c0 = c, c1 = 0, c2 = 0;
    const uint8_t* a = p;
    const uint8_t* b = p + kLane;
    const uint8_t* d = p + 2 * kLane;
    for (size_t i = 0; i < kLane; i += 8) {
      uint64_t v0, v1, v2;
      memcpy(&v0, a + i, 8);
      memcpy(&v1, b + i, 8);
      memcpy(&v2, d + i, 8);
      c0 = static_cast<uint32_t>(__builtin_ia32_crc32di(c0, v0));
      c1 = static_cast<uint32_t>(__builtin_ia32_crc32di(c1, v1));
      c2 = static_cast<uint32_t>(__builtin_ia32_crc32di(c2, v2));
    }
    c = shift.Apply(shift.Apply(c0) ^ c1) ^ c2;
    p += 3 * kLane;
    n -= 3 * kLane;
  }
  return ~HwCrcBlock(p, n, c);
}
#endif

uint32_t Crc32c(const char* data, size_t n, uint32_t crc) {
  const uint8_t* p = reinterpret_cast<const uint8_t*>(data);
#if defined(__x86_64__)
  static const bool hw = __builtin_cpu_supports("sse4.2");
  if (hw) return HwCrc32c(p, n, crc);
#endif
  const CrcTable& tb = Table();
  uint32_t c = ~crc;
  while (n >= 8) {   // slicing-by-8
    uint32_t lo, hi;
    memcpy(&lo, p, 4);
    memcpy(&hi, p + 4, 4);
    lo ^= c;
    c = tb.t[7][lo & 0xff] ^ tb.t[6][(lo >> 8) & 0xff] ^ tb.t[5][(lo >> 16) & 0xff] ^
        tb.t[4][lo >> 24] ^ tb.t[3][hi & 0xff] ^ tb.t[2][(hi >> 8) & 0xff] ^
        tb.t[1][(hi >> 16) & 0xff] ^ tb.t[0][hi >> 24];
    p += 8;
    n -= 8;
  }
  while (n--) c = tb.t[0][(c ^ *p++) & 0xff] ^ (c >> 8);
  return ~c;
}

// "base@N" names N shards "base-?????-of-0000N"; "base@*" any shard count.
std::string ExpandShardSpec(const std::string& item) {
  const size_t at = item.rfind('@');
  if (at == std::string::npos || at + 1 >= item.size()) return item;
  const std::string n = item.substr(at + 1);
  if (n == "*") return item.substr(0, at) + "-?????-of-?????";
  if (n.find_first_not_of("0123456789") != std::string::npos) return item;
  char buf[16];
  snprintf(buf, sizeof(buf), "%05d", std::stoi(n));
  return item.substr(0, at) + "-?????-of-" + buf;
}

std::vector<std::string> GlobFiles(const std::string& pattern) {
  std::vector<std::string> out;
  std::stringstream ss(pattern);
  std::string item;
  while (std::getline(ss, item, ',')) {
    if (item.empty()) continue;
    item = ExpandShardSpec(item);
    glob_t g;
    if (glob(item.c_str(), 0, nullptr, &g) == 0) {
      for (size_t i = 0; i < g.gl_pathc; ++i) out.emplace_back(g.gl_pathv[i]);
    }
    globfree(&g);
  }
  std::sort(out.begin(), out.end());
  return out;
}

// --------------------------------------------------------------- iterators ----
namespace {

class TFRecordIterator : public RecordIterator {
 public:
  explicit TFRecordIterator(const std::string& file) : f_(fopen(file.c_str(), "rb")), name_(file) {
    if (!f_) throw std::runtime_error("cannot open " + file);
    setvbuf(f_, nullptr, _IOFBF, 1 << 20);
  }
  ~TFRecordIterator() override {
    if (f_) fclose(f_);
  }
  bool Next(std::string* out) override {
    char hdr[12];
    const size_t got = fread(hdr, 1, 12, f_);
    if (got == 0) return false;
    if (got != 12) throw std::runtime_error("truncated tfrecord header in " + name_);
    uint64_t len;
    uint32_t len_crc;
    memcpy(&len, hdr, 8);
    memcpy(&len_crc, hdr + 8, 4);
    if (MaskCrc(Crc32c(hdr, 8)) != len_crc) throw std::runtime_error("corrupt length crc in " + name_);
    out->resize(len);
    char foot[4];
    if (fread(out->data(), 1, len, f_) != len || fread(foot, 1, 4, f_) != 4)
      throw std::runtime_error("truncated tfrecord payload in " + name_);
    uint32_t data_crc;
    memcpy(&data_crc, foot, 4);
    if (MaskCrc(Crc32c(out->data(), len)) != data_crc)
      throw std::runtime_error("corrupt data crc in " + name_);
    return true;
  }

 private:
  FILE* f_;
  std::string name_;
};

// TFRecord framing inside a gzip stream ("tfrecord_gzip:").
class GzTFRecordIterator : public RecordIterator {
 public:
  explicit GzTFRecordIterator(const std::string& file) : f_(gzopen(file.c_str(), "rb")), name_(file) {
    if (!f_) throw std::runtime_error("cannot open " + file);
    gzbuffer(f_, 1 << 20);
  }
  ~GzTFRecordIterator() override {
    if (f_) gzclose(f_);
  }
  bool Next(std::string* out) override {
    char hdr[12];
    const int got = gzread(f_, hdr, 12);
    if (got == 0) return false;
    if (got != 12) throw std::runtime_error("truncated tfrecord header in " + name_);
    uint64_t len;
    uint32_t len_crc;
    memcpy(&len, hdr, 8);
    memcpy(&len_crc, hdr + 8, 4);
    if (MaskCrc(Crc32c(hdr, 8)) != len_crc) throw std::runtime_error("corrupt length crc in " + name_);
    out->resize(len);
    size_t done = 0;
    while (done < len) {                               // gzread takes an unsigned int
      const unsigned want = static_cast<unsigned>(std::min<uint64_t>(len - done, 1u << 30));
      const int r = gzread(f_, out->data() + done, want);
      if (r <= 0) throw std::runtime_error("truncated tfrecord payload in " + name_);
      done += static_cast<size_t>(r);
    }
    char foot[4];
    if (gzread(f_, foot, 4) != 4) throw std::runtime_error("truncated tfrecord payload in " + name_);
    uint32_t data_crc;
    memcpy(&data_crc, foot, 4);
    if (MaskCrc(Crc32c(out->data(), len)) != data_crc)
      throw std::runtime_error("corrupt data crc in " + name_);
    return true;
  }

 private:
  gzFile f_;
  std::string name_;
};

class TextLineIterator : public RecordIterator {
 public:
  explicit TextLineIterator(const std::string& file) : in_(file) {
    if (!in_) throw std::runtime_error("cannot open " + file);
  }
  bool Next(std::string* out) override { return static_cast<bool>(std::getline(in_, *out)); }

 private:
  std::ifstream in_;
};

class IotaIterator : public RecordIterator {
 public:
  explicit IotaIterator(const std::string& n) : n_(std::stoll(n)) {}
  bool Next(std::string* out) override {
    if (i_ >= n_) return false;
    *out = std::to_string(i_++);
    return true;
  }

 private:
  int64_t n_, i_ = 0;
};

void SplitPattern(const std::string& fp, std::string* type, std::string* glob) {
  const size_t c = fp.find(':');
  if (c == std::string::npos) {
    *type = "tfrecord";
    *glob = fp;
  } else {
    *type = fp.substr(0, c);
    *glob = fp.substr(c + 1);
  }
}

// "text_indirect:<dir>/ckpt": the file is a text-format VersionedFileSet
//   current { file_pattern: "rel/a@8" file_pattern: "rel/b*" create_timestamp: … } history { … }
// and the data are the `current` patterns, relative to the directory of the ckpt file.
std::vector<std::string> IndirectPatterns(const std::string& ckpt) {
  std::ifstream in(ckpt);
  if (!in) throw std::runtime_error("cannot open fileset " + ckpt);
  std::stringstream buf;
  buf << in.rdbuf();
  const std::string text = buf.str();
  const size_t cur = text.find("current");
  if (cur == std::string::npos) throw std::runtime_error("no `current` fileset in " + ckpt);
  const size_t open = text.find('{', cur);
  if (open == std::string::npos) throw std::runtime_error("malformed fileset " + ckpt);
  int depth = 0;
  size_t close = open;
  for (; close < text.size(); ++close) {
    if (text[close] == '{') ++depth;
    if (text[close] == '}' && --depth == 0) break;
  }
  const std::string body = text.substr(open + 1, close - open - 1);
  const size_t slash = ckpt.rfind('/');
  const std::string dir = slash == std::string::npos ? std::string() : ckpt.substr(0, slash + 1);
  std::vector<std::string> out;
  size_t pos = 0;
  while ((pos = body.find("file_pattern", pos)) != std::string::npos) {
    const size_t q0 = body.find_first_of("\"'", pos);
    if (q0 == std::string::npos) break;
    const size_t q1 = body.find(body[q0], q0 + 1);
    if (q1 == std::string::npos) break;
    const std::string rel = body.substr(q0 + 1, q1 - q0 - 1);
    out.push_back(!rel.empty() && rel[0] == '/' ? rel : dir + rel);
    pos = q1 + 1;
  }
  if (out.empty()) throw std::runtime_error("fileset " + ckpt + " lists no file_pattern");
  return out;
}

std::vector<std::string> ExpandFiles(const std::string& type, const std::string& glob) {
  if (type == "iota") return {glob};
  if (type == "text_indirect") {
    std::vector<std::string> files;
    for (const std::string& pat : IndirectPatterns(glob)) {
      auto more = GlobFiles(pat);
      if (more.empty()) throw std::runtime_error("no files match " + pat);
      files.insert(files.end(), more.begin(), more.end());
    }
    return files;
  }
  auto files = GlobFiles(glob);
  if (files.empty()) throw std::runtime_error("no files match " + glob);
  return files;
}

}  // namespace

std::unique_ptr<RecordIterator> RecordIterator::Create(const std::string& type,
                                                       const std::string& file) {
  if (type == "tfrecord") return std::make_unique<TFRecordIterator>(file);
  if (type == "tfrecord_gzip") return std::make_unique<GzTFRecordIterator>(file);
  if (type == "text" || type == "text_indirect") return std::make_unique<TextLineIterator>(file);
  if (type == "iota") return std::make_unique<IotaIterator>(file);
  throw std::runtime_error("unknown record type '" + type + "'");
}

TFRecordWriter::TFRecordWriter(const std::string& path) : f_(fopen(path.c_str(), "wb")) {
  if (!f_) throw std::runtime_error("cannot create " + path);
}
TFRecordWriter::~TFRecordWriter() { Close(); }
void TFRecordWriter::Close() {
  if (f_) fclose(f_);
  f_ = nullptr;
}
void TFRecordWriter::Write(const std::string& rec) {
  char hdr[12];
  const uint64_t len = rec.size();
  memcpy(hdr, &len, 8);
  const uint32_t lc = MaskCrc(Crc32c(hdr, 8));
  memcpy(hdr + 8, &lc, 4);
  const uint32_t dc = MaskCrc(Crc32c(rec.data(), rec.size()));
  fwrite(hdr, 1, 12, f_);
  fwrite(rec.data(), 1, rec.size(), f_);
  fwrite(&dc, 1, 4, f_);
}

// ----------------------------------------------------------- basic yielder ----
BasicRecordYielder::BasicRecordYielder(const BasicYielderOptions& opts)
    : opts_(opts), pop_rng_(opts.seed ? opts.seed * 7919 + 1 : std::random_device{}()) {
  SplitPattern(opts_.file_pattern, &type_, &glob_);
  if (opts_.parallelism < 1) opts_.parallelism = 1;
  if (opts_.bufsize < 1) opts_.bufsize = 1;
  if (opts_.num_input_replicas < 1) opts_.num_input_replicas = 1;
  if (opts_.input_replica_id < 0 || opts_.input_replica_id >= opts_.num_input_replicas)
    throw std::runtime_error("BasicRecordYielder: input_replica_id out of range");
  limit_.store(opts_.bufsize);
  ExpandFiles(type_, glob_);   // fail fast on a bad pattern
  main_ = std::thread([this] { MainLoop(); });
}

BasicRecordYielder::~BasicRecordYielder() { Close(); }

void BasicRecordYielder::Close() {
  stop_.store(true);
  {
    std::lock_guard<std::mutex> l(mu_);
    cv_not_empty_.notify_all();
    cv_not_full_.notify_all();
  }
  if (main_.joinable()) main_.join();
}

void BasicRecordYielder::Add(std::vector<std::string>* chunk, std::mt19937_64* rng) {
  std::unique_lock<std::mutex> l(mu_);
  for (auto& rec : *chunk) {
    cv_not_full_.wait(l, [&] { return stop_ || static_cast<int64_t>(buf_.size()) < limit_.load(); });
    if (stop_) return;
    buf_.emplace_back(std::move(rec));
    // random-swap insertion keeps the buffer uniformly shuffled
    const size_t j = (*rng)() % buf_.size();
    std::swap(buf_[j], buf_.back());
    if (static_cast<int64_t>(buf_.size()) * 2 >= limit_.load()) cv_not_empty_.notify_one();
  }
  chunk->clear();
}

void BasicRecordYielder::ReadShard(const std::vector<std::string>& files, uint64_t seed,
                                   bool shard_records) {
  std::mt19937_64 rng(seed);
  std::vector<std::string> chunk;
  for (const auto& f : files) {
    if (stop_) return;
    auto it = RecordIterator::Create(type_, f);
    std::string rec;
    int64_t idx = 0;
    while (!stop_ && it->Next(&rec)) {
      // fewer files than replicas: shard by record position inside each file
      if (shard_records && (idx++ % opts_.num_input_replicas) != opts_.input_replica_id) {
        rec.clear();
        continue;
      }
      chunk.emplace_back(std::move(rec));
      rec.clear();
      if (chunk.size() >= 64) Add(&chunk, &rng);
    }
  }
  Add(&chunk, &rng);
}

void BasicRecordYielder::MainLoop() {
  for (int64_t epoch = 0; !stop_ && (opts_.num_epochs == 0 || epoch < opts_.num_epochs); ++epoch) {
    epoch_.store(epoch);
    auto files = ExpandFiles(type_, glob_);     // sorted ⇒ identical on every replica
    bool shard_records = false;
    if (opts_.num_input_replicas > 1) {
      if (static_cast<int>(files.size()) >= opts_.num_input_replicas) {
        std::vector<std::string> mine;
        for (size_t i = 0; i < files.size(); ++i)
          if (static_cast<int>(i % opts_.num_input_replicas) == opts_.input_replica_id)
            mine.push_back(files[i]);
        files.swap(mine);
      } else {
        shard_records = true;
      }
    }
    const uint64_t eseed =
        opts_.seed ? (opts_.seed * 0x9e3779b97f4a7c15ull) ^ static_cast<uint64_t>(epoch + 1)
                   : std::random_device{}();
    std::mt19937_64 rng(eseed);
    std::shuffle(files.begin(), files.end(), rng);
    const int shards = std::min<int>(opts_.parallelism, static_cast<int>(files.size()));
    std::vector<std::vector<std::string>> per(shards);
    for (size_t i = 0; i < files.size(); ++i) per[i % shards].push_back(files[i]);
    {
      std::lock_guard<std::mutex> l(mu_);
      epoch_draining_ = false;
    }
    std::vector<std::thread> readers;
    for (int s = 0; s < shards; ++s)
      readers.emplace_back([this, &per, s, eseed, shard_records] {
        ReadShard(per[s], eseed + 1 + s, shard_records);
      });
    for (auto& t : readers) t.join();
    // Everything of this epoch is in the buffer: let consumers drain it fully before
    // the next epoch starts (so epochs never mix).
    std::unique_lock<std::mutex> l(mu_);
    epoch_draining_ = true;
    cv_not_empty_.notify_all();
    cv_not_full_.wait(l, [&] { return stop_ || buf_.empty(); });
  }
  std::lock_guard<std::mutex> l(mu_);
  finished_ = true;
  cv_not_empty_.notify_all();
}

bool BasicRecordYielder::Yield(Record* out) {
  std::unique_lock<std::mutex> l(mu_);
  cv_not_empty_.wait(l, [&] {
    return stop_ || finished_ || (!buf_.empty() && (epoch_draining_ ||
                                                    static_cast<int64_t>(buf_.size()) * 2 >= limit_.load()));
  });
  if (buf_.empty()) return false;
  if (opts_.bufsize_in_seconds > 0) {
    // adapt the buffer to `bufsize_in_seconds` of the measured consumption rate
    const int64_t n = yielded_.fetch_add(1) + 1;
    if (n == 1) t_first_yield_ = std::chrono::steady_clock::now();
    if (n % 1024 == 0) {
      const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() -
                                                        t_first_yield_).count();
      if (secs > 0) {
        int64_t want = static_cast<int64_t>(n / secs * opts_.bufsize_in_seconds);
        want = std::max<int64_t>(1024, std::min<int64_t>(want, opts_.bufsize));
        limit_.store(want);
      }
    }
  }
  if (opts_.seed == 0) {   // extra extraction randomness for unseeded runs
    const size_t j = pop_rng_() % buf_.size();
    std::swap(buf_[j], buf_.back());
  }
  out->value = std::move(buf_.back());
  out->source_id = opts_.source_id;
  buf_.pop_back();
  cv_not_full_.notify_all();
  return true;
}

// ------------------------------------------------------ sequential yielder ----
SequentialRecordYielder::SequentialRecordYielder(const std::string& file_pattern,
                                                 int64_t repeat_count, int source_id,
                                                 int num_input_replicas, int input_replica_id)
    : repeat_(repeat_count), source_id_(source_id),
      replicas_(num_input_replicas < 1 ? 1 : num_input_replicas), replica_id_(input_replica_id) {
  std::string glob;
  SplitPattern(file_pattern, &type_, &glob);
  files_ = ExpandFiles(type_, glob);
}

bool SequentialRecordYielder::Yield(Record* out) {
  std::lock_guard<std::mutex> l(mu_);
  while (true) {
    if (!it_) {
      if (file_idx_ >= files_.size()) {
        ++epoch_;
        if (repeat_ >= 0 && epoch_ >= std::max<int64_t>(repeat_, 1)) return false;
        file_idx_ = 0;
      }
      it_ = RecordIterator::Create(type_, files_[file_idx_++]);
    }
    if (it_->Next(&out->value)) {
      // in-order evaluation data: round-robin records over the replicas
      if (replicas_ > 1 && (rec_idx_++ % replicas_) != replica_id_) continue;
      out->source_id = source_id_;
      return true;
    }
    it_.reset();
  }
}

// ---------------------------------------------------- weighted-mix yielder ----
WeightedMixRecordYielder::WeightedMixRecordYielder(std::vector<std::shared_ptr<Yielder>> children,
                                                   std::vector<double> weights, uint64_t seed)
    : children_(std::move(children)), dist_(weights.begin(), weights.end()),
      rng_(seed ? seed : std::random_device{}()) {
  if (children_.size() != weights.size() || children_.empty())
    throw std::runtime_error("WeightedMixRecordYielder: children/weights mismatch");
}

bool WeightedMixRecordYielder::Yield(Record* out) {
  int pick;
  {
    std::lock_guard<std::mutex> l(mu_);
    pick = dist_(rng_);
  }
  if (!children_[pick]->Yield(out)) return false;
  out->source_id = pick;
  return true;
}

void WeightedMixRecordYielder::Close() {
  for (auto& c : children_) c->Close();
}

}  // namespace lbh
