#include "text_ops.h"

#include <algorithm>
#include <cctype>
#include <climits>
#include <fstream>
#include <set>
#include <sstream>
#include <stdexcept>

namespace lbh {

// -------------------------------------------------------------------- ascii ----
AsciiTokenizer::AsciiTokenizer() {
  // id layout: 5 specials, a-z, 8 punctuation marks, 0-9, 24 symbols, 3 specials.
  id_to_tok_ = {"<unk>", "<s>", "</s>", " ", "<noise>"};
  for (char c = 'a'; c <= 'z'; ++c) id_to_tok_.emplace_back(1, c);
  for (char c : std::string(".'-:!~`;")) id_to_tok_.emplace_back(1, c);
  for (char c = '0'; c <= '9'; ++c) id_to_tok_.emplace_back(1, c);
  for (char c : std::string("\"#$%&()*+,/<=>?@[\\]^_{|}")) id_to_tok_.emplace_back(1, c);
  id_to_tok_.insert(id_to_tok_.end(), {"<epsilon>", "<text_only>", "<sorw>"});
  for (size_t i = 0; i < id_to_tok_.size(); ++i) tok_to_id_[id_to_tok_[i]] = static_cast<int32_t>(i);
  for (const char* s : {"<unk>", "<noise>", "<s>", "</s>", "<epsilon>", "<text_only>", "<sorw>"})
    specials_.emplace_back(s, tok_to_id_[s]);
}

const AsciiTokenizer& AsciiTokenizer::Get() {
  static const AsciiTokenizer t;
  return t;
}

std::vector<int32_t> AsciiTokenizer::StringToIds(const std::string& text) const {
  std::string s = text;
  std::transform(s.begin(), s.end(), s.begin(), [](unsigned char c) { return std::tolower(c); });
  std::vector<int32_t> ids;
  for (size_t i = 0; i < s.size();) {
    bool hit = false;
    if (s[i] == '<') {
      for (const auto& sp : specials_) {
        if (s.compare(i, sp.first.size(), sp.first) == 0) {
          ids.push_back(sp.second);
          i += sp.first.size();
          hit = true;
          break;
        }
      }
    }
    if (hit) continue;
    auto it = tok_to_id_.find(std::string(1, s[i]));
    ids.push_back(it == tok_to_id_.end() ? 0 : it->second);
    ++i;
  }
  return ids;
}

std::string AsciiTokenizer::IdsToString(const std::vector<int32_t>& ids) const {
  std::string out;
  for (int32_t id : ids)
    out += (id >= 0 && id < NumTokens()) ? id_to_tok_[id] : id_to_tok_[0];
  return out;
}

// -------------------------------------------------------------------- vocab ----
VocabTokenizer::VocabTokenizer(const std::string& vocab_path, bool ids_from_vocab) {
  std::ifstream in(vocab_path);
  if (!in) throw std::runtime_error("cannot open vocab " + vocab_path);
  std::string line;
  int32_t next = 0;
  while (std::getline(in, line)) {
    if (line.empty()) continue;
    std::string tok = line;
    int32_t id = next;
    const size_t tab = line.find('\t');
    if (tab != std::string::npos) {
      tok = line.substr(0, tab);
      if (ids_from_vocab) id = std::stoi(line.substr(tab + 1));
    }
    tok_to_id_[tok] = id;
    id_to_tok_[id] = tok;
    next = id + 1;
  }
  auto find = [&](const char* t) {
    auto it = tok_to_id_.find(t);
    return it == tok_to_id_.end() ? -1 : it->second;
  };
  unk_id_ = find("<unk>");
  if (unk_id_ < 0) unk_id_ = find("<UNK>");
  sos_id_ = find("<s>");
  if (sos_id_ < 0) sos_id_ = find("<S>");
  eos_id_ = find("</s>");
  if (eos_id_ < 0) eos_id_ = find("</S>");
}

int32_t VocabTokenizer::TokenToId(const std::string& tok) const {
  auto it = tok_to_id_.find(tok);
  return it == tok_to_id_.end() ? unk_id_ : it->second;
}

const std::string& VocabTokenizer::IdToToken(int32_t id) const {
  auto it = id_to_tok_.find(id);
  return it == id_to_tok_.end() ? unk_ : it->second;
}

std::vector<int32_t> VocabTokenizer::StringToIds(const std::string& text) const {
  std::vector<int32_t> ids;
  std::istringstream ss(text);
  std::string tok;
  while (ss >> tok) ids.push_back(TokenToId(tok));
  return ids;
}

std::string VocabTokenizer::IdsToString(const std::vector<int32_t>& ids) const {
  std::string out;
  for (size_t i = 0; i < ids.size(); ++i) {
    if (i) out += ' ';
    out += IdToToken(ids[i]);
  }
  return out;
}

// ---------------------------------------------------------------------- bpe ----
BpeTokenizer::BpeTokenizer(const std::string& codes_path, const std::string& vocab_path) {
  std::ifstream codes(codes_path);
  if (!codes) throw std::runtime_error("cannot open BPE codes " + codes_path);
  std::string line;
  int r = 0;
  while (std::getline(codes, line)) {
    if (line.empty() || line[0] == '#') continue;
    std::istringstream ss(line);
    std::string a, b;
    if (ss >> a >> b) rank_.emplace(std::make_pair(a, b), r++);
  }
  std::ifstream vocab(vocab_path);
  if (!vocab) throw std::runtime_error("cannot open BPE vocab " + vocab_path);
  while (std::getline(vocab, line)) {
    if (line.empty()) continue;
    const size_t cut = line.find_first_of("\t ");
    std::string tok = cut == std::string::npos ? line : line.substr(0, cut);
    tok_to_id_[tok] = static_cast<int32_t>(id_to_tok_.size());
    id_to_tok_.push_back(tok);
  }
  auto it = tok_to_id_.find("<unk>");
  unk_id_ = it == tok_to_id_.end() ? 0 : it->second;
}

std::vector<std::string> BpeTokenizer::EncodeWord(const std::string& word) const {
  // symbols = utf-8 code points; the last one carries the end-of-word marker.
  std::vector<std::string> sym;
  for (size_t i = 0; i < word.size();) {
    size_t n = 1;
    const unsigned char c = word[i];
    if (c >= 0xf0) n = 4; else if (c >= 0xe0) n = 3; else if (c >= 0xc0) n = 2;
    sym.push_back(word.substr(i, n));
    i += n;
  }
  if (sym.empty()) return sym;
  sym.back() += "</w>";
  while (sym.size() > 1) {
    int best = INT_MAX;
    size_t at = 0;
    for (size_t i = 0; i + 1 < sym.size(); ++i) {
      auto it = rank_.find({sym[i], sym[i + 1]});
      if (it != rank_.end() && it->second < best) {
        best = it->second;
        at = i;
      }
    }
    if (best == INT_MAX) break;
    sym[at] += sym[at + 1];
    sym.erase(sym.begin() + at + 1);
  }
  return sym;
}

std::vector<int32_t> BpeTokenizer::StringToIds(const std::string& text) const {
  std::vector<int32_t> ids;
  std::istringstream ss(text);
  std::string w;
  while (ss >> w) {
    for (const auto& piece : EncodeWord(w)) {
      auto it = tok_to_id_.find(piece);
      ids.push_back(it == tok_to_id_.end() ? unk_id_ : it->second);
    }
  }
  return ids;
}

std::string BpeTokenizer::IdsToString(const std::vector<int32_t>& ids) const {
  std::string out;
  for (int32_t id : ids) {
    std::string t = (id >= 0 && id < static_cast<int32_t>(id_to_tok_.size())) ? id_to_tok_[id] : "<unk>";
    const size_t m = t.rfind("</w>");
    if (m != std::string::npos && m + 4 == t.size()) {
      out += t.substr(0, m);
      out += ' ';
    } else {
      out += t;
    }
  }
  if (!out.empty() && out.back() == ' ') out.pop_back();
  return out;
}

// ------------------------------------------------------------------ packing ----
PackResult PackSequences(const std::vector<int32_t>& src_lens, const std::vector<int32_t>& tgt_lens,
                         int packed_batch_size, int src_cap, int tgt_cap, uint64_t seed) {
  if (src_lens.size() != tgt_lens.size()) throw std::runtime_error("PackSequences: size mismatch");
  struct Row {
    int src_used = 0, tgt_used = 0;
    std::vector<int> items;
  };
  std::vector<Row> rows;
  for (size_t i = 0; i < src_lens.size(); ++i) {
    const int s = src_lens[i], t = tgt_lens[i];
    if (s > src_cap || t > tgt_cap || (s <= 0 && t <= 0)) continue;   // dropped
    Row* dst = nullptr;
    for (auto& r : rows) {   // first fit
      if (r.src_used + s <= src_cap && r.tgt_used + t <= tgt_cap) {
        dst = &r;
        break;
      }
    }
    if (!dst) {
      rows.emplace_back();
      dst = &rows.back();
    }
    dst->src_used += s;
    dst->tgt_used += t;
    dst->items.push_back(static_cast<int>(i));
  }
  // packed_batch_size == 0: output as many rows as the packing needs (no dropping).
  if (packed_batch_size <= 0) packed_batch_size = static_cast<int>(rows.size());
  // reservoir sample rows down to packed_batch_size
  std::vector<int> keep;
  std::mt19937_64 rng(seed ? seed : std::random_device{}());
  for (int r = 0; r < static_cast<int>(rows.size()); ++r) {
    if (static_cast<int>(keep.size()) < packed_batch_size) {
      keep.push_back(r);
    } else {
      const uint64_t j = rng() % static_cast<uint64_t>(r + 1);
      if (j < static_cast<uint64_t>(packed_batch_size)) keep[j] = r;
    }
  }
  PackResult out;
  out.rows = packed_batch_size;
  out.src_len = src_cap;
  out.tgt_len = tgt_cap;
  auto alloc = [&](std::vector<int32_t>* v, int len) { v->assign(static_cast<size_t>(packed_batch_size) * len, 0); };
  alloc(&out.src_segment_ids, src_cap);
  alloc(&out.src_segment_pos, src_cap);
  alloc(&out.src_indices_in_input, src_cap);
  alloc(&out.tgt_segment_ids, tgt_cap);
  alloc(&out.tgt_segment_pos, tgt_cap);
  alloc(&out.tgt_indices_in_input, tgt_cap);
  for (size_t k = 0; k < keep.size(); ++k) {
    const Row& row = rows[keep[k]];
    int so = 0, to = 0;
    for (size_t seg = 0; seg < row.items.size(); ++seg) {
      const int i = row.items[seg];
      for (int p = 0; p < src_lens[i]; ++p, ++so) {
        const size_t at = k * src_cap + so;
        out.src_segment_ids[at] = static_cast<int32_t>(seg + 1);
        out.src_segment_pos[at] = p;
        out.src_indices_in_input[at] = i;
      }
      for (int p = 0; p < tgt_lens[i]; ++p, ++to) {
        const size_t at = k * tgt_cap + to;
        out.tgt_segment_ids[at] = static_cast<int32_t>(seg + 1);
        out.tgt_segment_pos[at] = p;
        out.tgt_indices_in_input[at] = i;
      }
    }
  }
  return out;
}

std::vector<int32_t> PackSingleSequence(const std::vector<int32_t>& lens, int cap, bool sequential) {
  std::vector<int32_t> group(lens.size(), -1);
  if (sequential) {
    int g = -1, used = cap + 1;
    for (size_t i = 0; i < lens.size(); ++i) {
      if (lens[i] > cap) continue;
      if (used + lens[i] > cap) {
        ++g;
        used = 0;
      }
      used += lens[i];
      group[i] = g;
    }
    return group;
  }
  // best-fit decreasing over an ordered multiset of (remaining space, group id)
  std::vector<int> order(lens.size());
  for (size_t i = 0; i < order.size(); ++i) order[i] = static_cast<int>(i);
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return lens[a] > lens[b]; });
  std::set<std::pair<int, int>> space;
  int groups = 0;
  for (int i : order) {
    if (lens[i] > cap) continue;
    auto it = space.lower_bound({lens[i], -1});
    int g, rem;
    if (it == space.end()) {
      g = groups++;
      rem = cap;
    } else {
      g = it->second;
      rem = it->first;
      space.erase(it);
    }
    group[i] = g;
    space.insert({rem - lens[i], g});
  }
  return group;
}

// --------------------------------------------------------------------- MASS ----
MassResult Mass(const std::vector<int32_t>& ids, const std::vector<float>& weights,
                const std::vector<int32_t>& lens, int batch, int max_len, const MassOptions& o,
                uint64_t seed) {
  std::mt19937_64 rng(seed ? seed : std::random_device{}());
  std::uniform_real_distribution<float> uni(0.f, 1.f);
  MassResult r;
  r.src_ids = ids;
  r.tgt_labels = ids;
  r.tgt_ids.assign(ids.size(), 0);
  r.tgt_weights.assign(ids.size(), 0.f);
  for (int b = 0; b < batch; ++b) {
    const int len = std::min<int>(lens[b], max_len);
    if (len <= 0) continue;
    int mask_len = std::max<int>(o.mask_minlen, static_cast<int>(len * o.mask_ratio));
    mask_len = std::min(mask_len, len);
    // split the masked budget into spans of at most span_len tokens
    std::vector<std::pair<int, int>> spans;   // [begin, end)
    int remaining = mask_len;
    std::vector<char> masked(len, 0);
    while (remaining > 0) {
      const int cur = std::min(remaining, o.span_len);
      int start;
      const float u = uni(rng);
      if (u < o.random_start_prob) start = static_cast<int>(rng() % static_cast<uint64_t>(len - cur + 1));
      else if (u < o.random_start_prob + (1.f - o.random_start_prob) / 2) start = 0;
      else start = len - cur;
      int placed = 0;
      for (int t = start; t < len && placed < cur; ++t) {
        if (!masked[t]) {
          masked[t] = 1;
          ++placed;
        }
      }
      if (placed == 0) break;
      remaining -= placed;
    }
    for (int t = 0; t < len; ++t) {
      const size_t at = static_cast<size_t>(b) * max_len + t;
      const int32_t prev = t == 0 ? 0 : ids[at - 1];
      if (masked[t]) {
        const float u = uni(rng);
        if (u < o.mask_prob) r.src_ids[at] = o.mask_id;
        else if (u < o.mask_prob + o.rand_prob && o.vocab_size > o.first_unreserved_id)
          r.src_ids[at] = o.first_unreserved_id +
                          static_cast<int32_t>(rng() % static_cast<uint64_t>(o.vocab_size - o.first_unreserved_id));
        // else keep the original token
        r.tgt_ids[at] = t == 0 ? prev : prev;       // teacher forcing: previous gold token
        r.tgt_weights[at] = weights.empty() ? 1.f : weights[at];
      } else {
        r.tgt_ids[at] = o.mask_target ? o.mask_id : prev;
      }
    }
  }
  return r;
}

// ---------------------------------------------------------------- best step ----
std::pair<int64_t, int64_t> BestStep(const std::string& hist_file, double tol, bool minimize) {
  std::ifstream in(hist_file);
  int64_t best_step = 0, last_step = 0;
  double best = 0;
  bool have = false;
  std::string line;
  while (std::getline(in, line)) {
    std::replace(line.begin(), line.end(), ',', ' ');
    std::istringstream ss(line);
    long long step;
    double val;
    if (!(ss >> step >> val)) continue;
    last_step = step;
    const bool better = !have || (minimize ? val < best - tol : val > best + tol);
    if (better) {
      best = val;
      best_step = step;
      have = true;
    }
  }
  return {best_step, last_step};
}

// ------------------------------------------------------- random permutation ----
RandomPermutationSequence::RandomPermutationSequence(int64_t num, int64_t batch, bool repeat,
                                                     uint64_t seed)
    : num_(num), batch_(batch), repeat_(repeat), rng_(seed ? seed : std::random_device{}()) {
  Refill();
}

void RandomPermutationSequence::Refill() {
  order_.resize(num_);
  for (int64_t i = 0; i < num_; ++i) order_[i] = i;
  std::shuffle(order_.begin(), order_.end(), rng_);
  pos_ = 0;
}

std::vector<int64_t> RandomPermutationSequence::Next() {
  std::vector<int64_t> out;
  while (static_cast<int64_t>(out.size()) < batch_) {
    if (pos_ == order_.size()) {
      if (!repeat_) break;
      Refill();
    }
    out.push_back(order_[pos_++]);
  }
  return out;
}

// ------------------------------------------------------------ n-gram / MLPerf ----
std::string VocabTokenizer::JoinIds(const std::vector<int32_t>& ids,
                                    const std::string& separator) const {
  std::string out;
  for (size_t i = 0; i < ids.size(); ++i) {
    if (i) out += separator;
    out += IdToToken(ids[i]);
  }
  return out;
}

namespace {

// Decodes the UTF-8 code point at the start of `s` (U+FFFD on malformed input).
uint32_t FirstCodePoint(const std::string& s) {
  if (s.empty()) return 0;
  const auto b = [&](size_t i) { return static_cast<uint32_t>(static_cast<unsigned char>(s[i])); };
  const uint32_t c = b(0);
  int extra = c < 0x80 ? 0 : (c >> 5) == 0x6 ? 1 : (c >> 4) == 0xE ? 2 : (c >> 3) == 0x1E ? 3 : -1;
  if (extra < 0 || s.size() < static_cast<size_t>(extra) + 1) return 0xFFFD;
  uint32_t cp = extra == 0 ? c : c & (0x3F >> extra);
  for (int i = 1; i <= extra; ++i) {
    if ((b(i) & 0xC0) != 0x80) return 0xFFFD;
    cp = (cp << 6) | (b(i) & 0x3F);
  }
  return cp;
}

// Letter-or-digit test without ICU: ASCII exactly, the rest by Unicode block (letters of
// the living scripts, CJK, Hangul, full-width forms); punctuation / symbol blocks excluded.
bool IsAlnumCodePoint(uint32_t cp) {
  if (cp < 0x80) return std::isalnum(static_cast<int>(cp)) != 0;
  if (cp == 0xAA || cp == 0xB5 || cp == 0xBA) return true;
  if (cp >= 0xC0 && cp <= 0x24F) return cp != 0xD7 && cp != 0xF7;
  struct Range { uint32_t lo, hi; };
  static const Range kRanges[] = {
      {0x250, 0x2AF},   {0x370, 0x373},   {0x376, 0x377},   {0x37B, 0x37D},   {0x386, 0x386},
      {0x388, 0x3FF},   {0x400, 0x481},   {0x48A, 0x52F},   {0x531, 0x556},   {0x561, 0x587},
      {0x5D0, 0x5EA},   {0x620, 0x64A},   {0x660, 0x669},   {0x671, 0x6D3},   {0x6F0, 0x6FC},
      {0x904, 0x939},   {0x958, 0x961},   {0x966, 0x96F},   {0x985, 0x9B9},   {0x9E6, 0x9F1},
      {0xA05, 0xA39},   {0xA85, 0xAB9},   {0xB05, 0xB39},   {0xB85, 0xBB9},   {0xC05, 0xC39},
      {0xC85, 0xCB9},   {0xD05, 0xD3A},   {0xE01, 0xE30},   {0xE40, 0xE46},   {0xE50, 0xE59},
      {0x10A0, 0x10FF}, {0x1100, 0x11FF}, {0x1E00, 0x1FFF}, {0x3041, 0x3096}, {0x30A1, 0x30FA},
      {0x3105, 0x312F}, {0x3400, 0x4DBF}, {0x4E00, 0x9FFF}, {0xAC00, 0xD7A3}, {0xF900, 0xFAFF},
      {0xFF10, 0xFF19}, {0xFF21, 0xFF3A}, {0xFF41, 0xFF5A}, {0xFF66, 0xFFDC}, {0x20000, 0x2FA1F}};
  for (const Range& r : kRanges)
    if (cp >= r.lo && cp <= r.hi) return true;
  return false;
}

}  // namespace

MlPerfSubword::MlPerfSubword(const std::string& vocab_path) {
  std::ifstream f(vocab_path);
  if (!f) throw std::runtime_error("MlPerfSubword: cannot open " + vocab_path);
  std::vector<std::string> lines;
  for (std::string line; std::getline(f, line);) lines.push_back(line);
  LoadLines(lines);
}

void MlPerfSubword::LoadLines(const std::vector<std::string>& lines) {
  for (std::string line : lines) {
    while (!line.empty() && (line.back() == '\r' || line.back() == '\n')) line.pop_back();
    if (line.empty()) continue;
    if (line.size() < 2) throw std::runtime_error("MlPerfSubword: bad vocab line: " + line);
    id_to_tok_.push_back(line.substr(1, line.size() - 2));      // drop the quotes
  }
}

std::string MlPerfSubword::Decode(const std::vector<int32_t>& ids) const {
  std::string joined;
  for (int32_t id : ids) {
    if (id < 0 || static_cast<size_t>(id) >= id_to_tok_.size())
      throw std::out_of_range("MlPerfSubword: id out of range: " + std::to_string(id));
    joined += id_to_tok_[id];
  }
  std::string out;
  bool prev_alnum = false, first = true;
  size_t start = 0;
  while (true) {
    const size_t end = joined.find('_', start);
    const std::string tok = joined.substr(start, end == std::string::npos ? end : end - start);
    const bool alnum = IsAlnumCodePoint(FirstCodePoint(tok));
    if (!first && prev_alnum && alnum) out += ' ';
    out += tok;
    prev_alnum = alnum;
    first = false;
    if (end == std::string::npos) break;
    start = end + 1;
  }
  return out;
}

}  // namespace lbh
