// Host-side text / packing algorithms: tokenizers, sequence packing,
// MASS masking, early-stop bookkeeping.  Native re-designs of the reference's
// `tokenizer_ops_kernels.cc`, `ascii_tokenizer.cc`, `simple_vocab.cc`,
// `pack_ops.cc`, `text_packing.cc`, `mass_op.cc`, `best_step_op_kernels.cc`,
// `random_ops_kernels.cc`.
#pragma once

#include <cstdint>
#include <map>
#include <random>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <unordered_map>
#include <utility>
#include <vector>

namespace lbh {

// ---- character tokenizer (76 ids; 0 <unk>, 1 <s>, 2 </s>, 3 ' ', 4 <noise> …) ----
class AsciiTokenizer {
 public:
  static const AsciiTokenizer& Get();
  int NumTokens() const { return static_cast<int>(id_to_tok_.size()); }
  std::vector<int32_t> StringToIds(const std::string& text) const;
  std::string IdsToString(const std::vector<int32_t>& ids) const;

 private:
  AsciiTokenizer();
  std::vector<std::string> id_to_tok_;
  std::unordered_map<std::string, int32_t> tok_to_id_;
  std::vector<std::pair<std::string, int32_t>> specials_;
};

// ---- whitespace tokenizer over a vocab file ("token" or "token<TAB>id" per line) ----
class VocabTokenizer {
 public:
  VocabTokenizer(const std::string& vocab_path, bool load_token_ids_from_vocab);
  int32_t TokenToId(const std::string& tok) const;
  const std::string& IdToToken(int32_t id) const;
  std::vector<int32_t> StringToIds(const std::string& text) const;
  std::string IdsToString(const std::vector<int32_t>& ids) const;
  // N-gram vocabularies: concatenates the tokens with `separator` (ref `NgramIdToToken`).
  std::string JoinIds(const std::vector<int32_t>& ids, const std::string& separator) const;
  bool Contains(const std::string& tok) const { return tok_to_id_.count(tok) > 0; }
  int32_t unk_id() const { return unk_id_; }
  int32_t sos_id() const { return sos_id_; }
  int32_t eos_id() const { return eos_id_; }
  size_t size() const { return id_to_tok_.size(); }

 private:
  std::unordered_map<std::string, int32_t> tok_to_id_;
  std::map<int32_t, std::string> id_to_tok_;
  int32_t unk_id_ = -1, sos_id_ = -1, eos_id_ = -1;
  std::string unk_ = "<unk>";
};

// ---- BPE: merge rules ("a b" per line, priority = line number) + vocab ----
class BpeTokenizer {
 public:
  BpeTokenizer(const std::string& codes_path, const std::string& vocab_path);
  // Words are split on whitespace; every word is encoded independently.
  std::vector<int32_t> StringToIds(const std::string& text) const;
  std::string IdsToString(const std::vector<int32_t>& ids) const;
  std::vector<std::string> EncodeWord(const std::string& word) const;

 private:
  std::map<std::pair<std::string, std::string>, int> rank_;
  std::unordered_map<std::string, int32_t> tok_to_id_;
  std::vector<std::string> id_to_tok_;
  int32_t unk_id_ = 0;
};

// ---- MLPerf transformer sub-word vocabulary: one quoted sub-token per line; '_' ends a
// token. Decoding joins the sub-tokens, splits on '_' and re-inserts a blank between two
// neighbouring tokens that both start with a letter or digit (ref `ml_perf_subword_op.cc`).
class MlPerfSubword {
 public:
  explicit MlPerfSubword(const std::string& vocab_path);
  explicit MlPerfSubword(const std::vector<std::string>& lines) { LoadLines(lines); }
  std::string Decode(const std::vector<int32_t>& ids) const;
  size_t size() const { return id_to_tok_.size(); }

 private:
  void LoadLines(const std::vector<std::string>& lines);
  std::vector<std::string> id_to_tok_;
};

// ---- immutable lookup tables with a default (ref `static_map_op.cc`) ----
template <class K, class V>
class StaticMap {
 public:
  StaticMap(const std::vector<K>& keys, const std::vector<V>& vals, V unk) : unk_(std::move(unk)) {
    if (!vals.empty() && vals.size() != keys.size())
      throw std::invalid_argument("StaticMap: keys / vals size mismatch");
    map_.reserve(keys.size());
    for (size_t i = 0; i < keys.size(); ++i)
      if (!map_.emplace(keys[i], vals.empty() ? Default(i) : vals[i]).second)
        throw std::invalid_argument("StaticMap: duplicate key");
  }
  std::vector<V> Lookup(const std::vector<K>& xs) const {
    std::vector<V> out;
    out.reserve(xs.size());
    for (const K& x : xs) {
      auto it = map_.find(x);
      out.push_back(it == map_.end() ? unk_ : it->second);
    }
    return out;
  }
  size_t size() const { return map_.size(); }

 private:
  static V Default(size_t i) {
    if constexpr (std::is_integral<V>::value) return static_cast<V>(i);
    else throw std::invalid_argument("StaticMap: vals required for non-integer values");
  }
  std::unordered_map<K, V> map_;
  V unk_;
};

// ---- packing ----
struct PackResult {
  int rows = 0, src_len = 0, tgt_len = 0;
  // [rows, len] row-major; 0 = empty slot (segment ids are 1-based)
  std::vector<int32_t> src_segment_ids, src_segment_pos, src_indices_in_input;
  std::vector<int32_t> tgt_segment_ids, tgt_segment_pos, tgt_indices_in_input;
};

// First-fit packing of (src_len[i], tgt_len[i]) pairs into rows of capacity
// (packed_src_seq_len, packed_tgt_seq_len); inputs that fit nowhere are dropped.
// If more than `packed_batch_size` rows result, a uniform reservoir sample of
// the rows is kept (seed 0: non-deterministic); fewer rows are padded with empties.
PackResult PackSequences(const std::vector<int32_t>& src_lens, const std::vector<int32_t>& tgt_lens,
                         int packed_batch_size, int packed_src_seq_len, int packed_tgt_seq_len,
                         uint64_t seed);

// Assigns every sequence a packed-group id (-1: longer than the limit).
// require_sequential_order: next-fit in input order; else best-fit decreasing.
std::vector<int32_t> PackSingleSequence(const std::vector<int32_t>& lens, int max_packed_len,
                                        bool require_sequential_order);

// ---- MASS (masked seq2seq pre-training) ----
struct MassOptions {
  int32_t mask_id = 3;
  float mask_ratio = 0.5f;
  int mask_minlen = 0;
  int span_len = 100000;
  float random_start_prob = 0.6f;
  float keep_prob = 0.1f, rand_prob = 0.1f, mask_prob = 0.8f;
  bool mask_target = true;
  int vocab_size = 0;
  int first_unreserved_id = 4;
};
struct MassResult {
  std::vector<int32_t> src_ids, tgt_ids, tgt_labels;
  std::vector<float> tgt_weights;
};
// ids / weights are [batch, max_len] row-major; lens[b] = actual length.
MassResult Mass(const std::vector<int32_t>& ids, const std::vector<float>& weights,
                const std::vector<int32_t>& lens, int batch, int max_len, const MassOptions& o,
                uint64_t seed);

// ---- early stop bookkeeping: best step of a "step<TAB/space>value" history file ----
std::pair<int64_t, int64_t> BestStep(const std::string& hist_file, double tol, bool minimize);

// ---- epoch-wise random permutation batches of [0, num) ----
class RandomPermutationSequence {
 public:
  RandomPermutationSequence(int64_t num, int64_t batch, bool repeat, uint64_t seed);
  std::vector<int64_t> Next();   // empty when exhausted (repeat == false)

 private:
  void Refill();
  int64_t num_, batch_;
  bool repeat_;
  std::mt19937_64 rng_;
  std::vector<int64_t> order_;
  size_t pos_ = 0;
};

}  // namespace lbh
