// Native `.msgpack.l.gz` shard decoder + graph tensoriser (host only; C ABI in include/buglab_shards.h).
//
// What it restates, and where the behaviour is pinned:
//   * wire format: gzip stream of back-to-back msgpack objects   (reference buglab/utils/msgpackutils.py:11-14)
//   * sample schema + open-vocabulary subtoken nodes              (reference buglab/representations/data.py:14-20, 97-167)
//   * node-label tokenisation                                     (split_identifier_into_parts + vocabulary)
// The host-language implementation of the same steps (buglab/representations/data.py, ptgnn/.../strelementrepresentationmodel.py
// in this repo, themselves pinned by tests/golden) is the checker: tests/test_shards_cpu.py demands bit-identical arrays.
//
// Rule of the file: never approximate.  Whatever cannot be reproduced exactly is reported as BL_SAMPLE_NEEDS_HOST.
#include "../../../include/buglab_shards.h"

#include <zlib.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <deque>
#include <string>
#include <string_view>
#include <unordered_map>
#include <vector>

namespace {

// ------------------------------------------------------------------------------------------------ UTF-8
// Strict validation (what CPython's utf-8 codec accepts with errors="strict", which msgpack raw=False uses).
// Calls `on_cp(code point)` for every non-ASCII code point; returns false on malformed input.
template <class F>
bool scan_utf8(const uint8_t* p, size_t n, F&& on_cp) {
  size_t i = 0;
  while (i < n) {
    // ASCII fast path, 8 bytes at a time
    while (i + 8 <= n) {
      uint64_t w;
      memcpy(&w, p + i, 8);
      if (w & 0x8080808080808080ull) break;
      i += 8;
    }
    if (i >= n) break;
    uint8_t c = p[i];
    if (c < 0x80) { ++i; continue; }
    uint32_t cp;
    int extra;
    if (c >= 0xC2 && c <= 0xDF) { cp = c & 0x1F; extra = 1; }
    else if (c >= 0xE0 && c <= 0xEF) { cp = c & 0x0F; extra = 2; }
    else if (c >= 0xF0 && c <= 0xF4) { cp = c & 0x07; extra = 3; }
    else return false;
    if (i + (size_t)extra >= n) return false;  // sequence runs past the end
    for (int k = 1; k <= extra; ++k) {
      uint8_t cc = p[i + k];
      if ((cc & 0xC0) != 0x80) return false;
      cp = (cp << 6) | (cc & 0x3F);
    }
    if (extra == 2 && (cp < 0x800 || (cp >= 0xD800 && cp <= 0xDFFF))) return false;
    if (extra == 3 && (cp < 0x10000 || cp > 0x10FFFF)) return false;
    on_cp(cp);
    i += extra + 1;
  }
  return true;
}

inline bool valid_utf8(const uint8_t* p, size_t n) {
  return scan_utf8(p, n, [](uint32_t) {});
}

// ------------------------------------------------------------------------------------------------ msgpack
struct ParseError {};

constexpr int kMaxDepth = 512;  // nesting limit of msgpack-python's C unpacker

struct Cursor {
  const uint8_t* p;
  const uint8_t* end;

  void need(size_t n) const {
    if ((size_t)(end - p) < n) throw ParseError();
  }
  uint8_t peek() const { need(1); return *p; }
  uint64_t be(int n) {
    need(n);
    uint64_t v = 0;
    for (int i = 0; i < n; ++i) v = (v << 8) | *p++;
    return v;
  }
  bool at_nil() const { return peek() == 0xc0; }

  // Container / string headers; return false (cursor untouched) when the next value is of another kind.
  bool array_header(uint64_t& n) {
    uint8_t b = peek();
    if ((b & 0xf0) == 0x90) { ++p; n = b & 0x0f; return true; }
    if (b == 0xdc) { ++p; n = be(2); return true; }
    if (b == 0xdd) { ++p; n = be(4); return true; }
    return false;
  }
  bool map_header(uint64_t& n) {
    uint8_t b = peek();
    if ((b & 0xf0) == 0x80) { ++p; n = b & 0x0f; return true; }
    if (b == 0xde) { ++p; n = be(2); return true; }
    if (b == 0xdf) { ++p; n = be(4); return true; }
    return false;
  }
  bool str(std::string_view& out) {
    uint8_t b = peek();
    uint64_t n;
    if ((b & 0xe0) == 0xa0) { ++p; n = b & 0x1f; }
    else if (b == 0xd9) { ++p; n = be(1); }
    else if (b == 0xda) { ++p; n = be(2); }
    else if (b == 0xdb) { ++p; n = be(4); }
    else return false;
    need(n);
    out = std::string_view(reinterpret_cast<const char*>(p), (size_t)n);
    p += n;
    return true;
  }
  // Any msgpack integer that fits int64.
  bool integer(int64_t& out) {
    uint8_t b = peek();
    if (b <= 0x7f) { ++p; out = b; return true; }
    if (b >= 0xe0) { ++p; out = (int8_t)b; return true; }
    switch (b) {
      case 0xcc: ++p; out = (int64_t)be(1); return true;
      case 0xcd: ++p; out = (int64_t)be(2); return true;
      case 0xce: ++p; out = (int64_t)be(4); return true;
      case 0xcf: {
        const uint8_t* save = p;
        ++p;
        uint64_t v = be(8);
        if (v > (uint64_t)INT64_MAX) { p = save; return false; }
        out = (int64_t)v;
        return true;
      }
      case 0xd0: ++p; out = (int8_t)be(1); return true;
      case 0xd1: ++p; out = (int16_t)be(2); return true;
      case 0xd2: ++p; out = (int32_t)be(4); return true;
      case 0xd3: ++p; out = (int64_t)be(8); return true;
      default: return false;
    }
  }

  // Skips one value, enforcing what msgpack.Unpacker(raw=False, strict_map_key=True) enforces while it builds objects:
  // well-formed type bytes, valid UTF-8 in every str, str/bin map keys, bounded nesting.
  void skip(int depth = 0) {
    if (depth > kMaxDepth) throw ParseError();
    uint8_t b = peek();
    uint64_t n;
    std::string_view s;
    if (b <= 0x7f || b >= 0xe0) { ++p; return; }
    if ((b & 0xe0) == 0xa0 || b == 0xd9 || b == 0xda || b == 0xdb) {
      str(s);
      if (!valid_utf8(reinterpret_cast<const uint8_t*>(s.data()), s.size())) throw ParseError();
      return;
    }
    if (array_header(n)) {
      for (uint64_t i = 0; i < n; ++i) skip(depth + 1);
      return;
    }
    if (map_header(n)) {
      for (uint64_t i = 0; i < n; ++i) {
        uint8_t k = peek();
        bool key_ok = (k & 0xe0) == 0xa0 || k == 0xd9 || k == 0xda || k == 0xdb || k == 0xc4 || k == 0xc5 || k == 0xc6;
        if (!key_ok) throw ParseError();
        skip(depth + 1);
        skip(depth + 1);
      }
      return;
    }
    ++p;
    switch (b) {
      case 0xc0: case 0xc2: case 0xc3: return;
      case 0xc4: n = be(1); need(n); p += n; return;
      case 0xc5: n = be(2); need(n); p += n; return;
      case 0xc6: n = be(4); need(n); p += n; return;
      case 0xc7: n = be(1); need(n + 1); p += n + 1; return;
      case 0xc8: n = be(2); need(n + 1); p += n + 1; return;
      case 0xc9: n = be(4); need(n + 1); p += n + 1; return;
      case 0xca: need(4); p += 4; return;
      case 0xcb: need(8); p += 8; return;
      case 0xcc: case 0xd0: need(1); p += 1; return;
      case 0xcd: case 0xd1: need(2); p += 2; return;
      case 0xce: case 0xd2: need(4); p += 4; return;
      case 0xcf: case 0xd3: need(8); p += 8; return;
      case 0xd4: need(2); p += 2; return;
      case 0xd5: need(3); p += 3; return;
      case 0xd6: need(5); p += 5; return;
      case 0xd7: need(9); p += 9; return;
      case 0xd8: need(17); p += 17; return;
      default: throw ParseError();  // 0xc1
    }
  }
};

// ------------------------------------------------------------------------------------------------ CPython set order
// Insertion-only model of CPython's setobject.c (3.7 .. 3.12: LINEAR_PROBES 9, PERTURB_SHIFT 5, grow at fill*5 >= mask*3
// to used*4 (used*2 beyond 50000 entries), re-insertion in table order).  Keys are non-negative ints < 2**61-1, whose hash
// is the value itself.  Pinned against the running interpreter by tests/test_shards_cpu.py and, at load time, by
// buglab_b200/shards.py.
class PySetOrder {
 public:
  PySetOrder() : table_(8, kEmpty), mask_(7), fill_(0) {}

  void add(int64_t key) {
    size_t perturb = (size_t)key;
    size_t i = (size_t)key & mask_;
    size_t slot;
    for (;;) {
      size_t e = i;
      int probes = (i + kLinearProbes <= mask_) ? kLinearProbes : 0;
      do {
        if (table_[e] == kEmpty) { slot = e; goto found_unused; }
        if (table_[e] == key) return;
        ++e;
      } while (probes--);
      perturb >>= kPerturbShift;
      i = (i * 5 + 1 + perturb) & mask_;
    }
  found_unused:
    table_[slot] = key;
    ++fill_;
    if (fill_ * 5 < mask_ * 3) return;
    resize(fill_ > 50000 ? fill_ * 2 : fill_ * 4);
  }

  template <class F>
  void for_each(F&& f) const {
    for (int64_t k : table_)
      if (k != kEmpty) f(k);
  }
  size_t size() const { return fill_; }

 private:
  static constexpr int64_t kEmpty = -1;
  static constexpr int kLinearProbes = 9;
  static constexpr int kPerturbShift = 5;

  void resize(size_t minused) {
    size_t newsize = 8;
    while (newsize <= minused) newsize <<= 1;
    std::vector<int64_t> fresh(newsize, kEmpty);
    size_t mask = newsize - 1;
    for (int64_t key : table_) {
      if (key == kEmpty) continue;
      size_t perturb = (size_t)key;
      size_t i = (size_t)key & mask;
      for (;;) {
        if (fresh[i] == kEmpty) { fresh[i] = key; break; }
        bool placed = false;
        if (i + kLinearProbes <= mask) {
          for (int j = 1; j <= kLinearProbes; ++j)
            if (fresh[i + j] == kEmpty) { fresh[i + j] = key; placed = true; break; }
        }
        if (placed) break;
        perturb >>= kPerturbShift;
        i = (i * 5 + 1 + perturb) & mask;
      }
    }
    table_.swap(fresh);
    mask_ = mask;
  }

  std::vector<int64_t> table_;
  size_t mask_;
  size_t fill_;
};

// ------------------------------------------------------------------------------------------------ identifier splitting
inline bool is_upper(unsigned char c) { return c >= 'A' && c <= 'Z'; }
inline bool is_lower(unsigned char c) { return c >= 'a' && c <= 'z'; }
inline bool is_digit(unsigned char c) { return c >= '0' && c <= '9'; }

// dpu_utils split_identifier_into_parts, as restated in this repo's dpu_utils/codeutils/identifiersplitting.py:
// split on '_', then per piece the runs  [A-Z]+(?![a-z]) | [A-Z][a-z]+ | [a-z]+ | [0-9]+ | [^A-Za-z0-9]+ ; ASCII-lowercased
// (callers guarantee that non-ASCII code points in the label are invariant under str.lower()).
// Parts are appended as (offset, length) into `lowered`, a lower-cased copy of the label; a label that yields no part
// (empty or underscores only) is returned whole and untouched.
struct Parts {
  std::string lowered;
  std::vector<std::pair<uint32_t, uint32_t>> spans;
  bool whole_label = false;  // the single part is the label itself, NOT lower-cased
};

void split_identifier(std::string_view label, Parts& out) {
  out.spans.clear();
  out.whole_label = false;
  out.lowered.assign(label);
  for (char& ch : out.lowered)
    if (is_upper((unsigned char)ch)) ch = (char)(ch - 'A' + 'a');
  const size_t n = label.size();
  size_t i = 0;
  while (i < n) {
    if (label[i] == '_') { ++i; continue; }
    size_t piece_end = i;
    while (piece_end < n && label[piece_end] != '_') ++piece_end;
    size_t pos = i;
    while (pos < piece_end) {
      unsigned char c = (unsigned char)label[pos];
      size_t stop = pos;
      if (is_upper(c)) {
        while (stop < piece_end && is_upper((unsigned char)label[stop])) ++stop;
        size_t run = stop - pos;
        bool lower_follows = stop < piece_end && is_lower((unsigned char)label[stop]);
        if (lower_follows) {
          if (run >= 2) {
            stop -= 1;  // acronym gives back its last capital, which starts the next Title-case part
          } else {
            while (stop < piece_end && is_lower((unsigned char)label[stop])) ++stop;
          }
        }
      } else if (is_lower(c)) {
        while (stop < piece_end && is_lower((unsigned char)label[stop])) ++stop;
      } else if (is_digit(c)) {
        while (stop < piece_end && is_digit((unsigned char)label[stop])) ++stop;
      } else {
        while (stop < piece_end) {
          unsigned char d = (unsigned char)label[stop];
          if (is_upper(d) || is_lower(d) || is_digit(d)) break;
          ++stop;
        }
      }
      out.spans.emplace_back((uint32_t)pos, (uint32_t)(stop - pos));
      pos = stop;
    }
    i = piece_end;
  }
  if (out.spans.empty()) {
    out.whole_label = true;
    out.lowered.assign(label);
    out.spans.emplace_back(0u, (uint32_t)n);
  }
}

// ------------------------------------------------------------------------------------------------ vocabulary
inline uint64_t hash_bytes(const char* p, size_t n) {
  uint64_t h = 0xcbf29ce484222325ull;
  for (size_t i = 0; i < n; ++i) { h ^= (unsigned char)p[i]; h *= 0x100000001b3ull; }
  h ^= h >> 29;
  return h;
}

}  // namespace

struct bl_tokenizer {
  std::string blob;
  struct Slot { uint64_t hash; int64_t off; int32_t len; int32_t id; };
  std::vector<Slot> slots;   // open addressing, len < 0 = empty
  uint64_t mask = 0;
  int32_t unk_id = -1;
  int32_t kind = BL_SPLIT_SUBTOKEN;
  int32_t max_subtokens = 1;
  std::vector<int32_t> lower_variant;  // sorted

  // id of `s`, or unk_id (possibly < 0) on a miss
  int32_t lookup(const char* p, size_t n) const {
    uint64_t h = hash_bytes(p, n);
    for (uint64_t i = h & mask;; i = (i + 1) & mask) {
      const Slot& s = slots[i];
      if (s.len < 0) return unk_id;
      if (s.hash == h && (size_t)s.len == n && memcmp(blob.data() + s.off, p, n) == 0) return s.id;
    }
  }

  // true when the label can be tokenised natively with exactly the host's result
  bool label_supported(std::string_view label) const {
    bool ok = true;
    bool valid = scan_utf8(reinterpret_cast<const uint8_t*>(label.data()), label.size(), [&](uint32_t cp) {
      if (std::binary_search(lower_variant.begin(), lower_variant.end(), (int32_t)cp)) ok = false;
    });
    return valid && ok;
  }

  // Writes <= max_subtokens ids; returns the count or -1 (needs host).  `label` must satisfy label_supported().
  int32_t ids_of(std::string_view label, Parts& scratch, int32_t* out) const {
    if (kind == BL_SPLIT_TOKEN) {
      int32_t id = lookup(label.data(), label.size());
      if (id < 0) return -1;
      out[0] = id;
      return 1;
    }
    split_identifier(label, scratch);
    int32_t count = (int32_t)std::min<size_t>(scratch.spans.size(), (size_t)max_subtokens);
    for (int32_t k = 0; k < count; ++k) {
      int32_t id = lookup(scratch.lowered.data() + scratch.spans[k].first, scratch.spans[k].second);
      if (id < 0) return -1;
      out[k] = id;
    }
    return count;
  }
};

struct bl_shard {
  std::vector<uint8_t> raw;
  std::vector<int64_t> offsets;  // object i = raw[offsets[i] .. offsets[i+1])
  int32_t status = BL_SHARDS_OK;
};

namespace {

struct EdgeList {
  std::string_view name;
  std::vector<int32_t> src, tgt;
  std::vector<uint32_t> args_marked;  // positions whose third element is the string "args" (only kept for "Child")
};

struct NeedsHost {};

}  // namespace

struct bl_sample {
  // outputs
  std::vector<int32_t> node_ids, node_lens, edge_src, edge_tgt, reference_nodes, call_args;
  std::vector<int64_t> edge_offsets;
  // scratch, reused between decodes
  std::vector<std::string_view> labels;
  std::deque<std::string> new_labels;
  std::vector<EdgeList> edge_lists;
  size_t num_edge_lists = 0;
  Parts parts;
  std::unordered_map<std::string, int32_t> subtoken_node;

  EdgeList& edge_list(std::string_view name) {
    for (size_t i = 0; i < num_edge_lists; ++i)
      if (edge_lists[i].name == name) {  // duplicate key: the later value replaces the earlier, like a dict
        edge_lists[i].src.clear(); edge_lists[i].tgt.clear(); edge_lists[i].args_marked.clear();
        return edge_lists[i];
      }
    if (num_edge_lists == edge_lists.size()) edge_lists.emplace_back();
    EdgeList& e = edge_lists[num_edge_lists++];
    e.name = name;
    e.src.clear(); e.tgt.clear(); e.args_marked.clear();
    return e;
  }
  EdgeList* find_edge_list(std::string_view name) {
    for (size_t i = 0; i < num_edge_lists; ++i)
      if (edge_lists[i].name == name) return &edge_lists[i];
    return nullptr;
  }
};

// Accumulator of the model metadata a pass over the training data collects (mirror gnn.py:223-225 ->
// graphneuralnetwork.py update_metadata_from): (sub)token counts of the node labels and the set of edge-type names.
struct bl_metadata {
  std::unordered_map<std::string, int64_t> token_counts;
  std::unordered_map<std::string, int64_t> edge_types;  // name -> number of samples that have it
  int64_t num_samples = 0;
};

namespace {

int32_t as_i32(int64_t v) {
  if (v < INT32_MIN || v > INT32_MAX) throw NeedsHost();
  return (int32_t)v;
}

void parse_edges(Cursor& c, bl_sample& s) {
  uint64_t num_types;
  if (!c.map_header(num_types)) throw NeedsHost();
  for (uint64_t t = 0; t < num_types; ++t) {
    std::string_view name;
    if (!c.str(name)) throw NeedsHost();
    EdgeList& list = s.edge_list(name);
    const bool is_child = name == "Child";
    uint64_t num_edges;
    if (!c.array_header(num_edges)) throw NeedsHost();
    list.src.reserve(num_edges);
    list.tgt.reserve(num_edges);
    for (uint64_t e = 0; e < num_edges; ++e) {
      uint64_t arity;
      if (!c.array_header(arity) || arity < 2) throw NeedsHost();
      int64_t a, b;
      if (!c.integer(a) || !c.integer(b)) throw NeedsHost();
      list.src.push_back(as_i32(a));
      list.tgt.push_back(as_i32(b));
      for (uint64_t k = 2; k < arity; ++k) {
        std::string_view extra;
        if (is_child && arity == 3 && c.str(extra)) {
          if (extra == "args") list.args_marked.push_back((uint32_t)(list.src.size() - 1));
        } else {
          c.skip(1);
        }
      }
    }
  }
}

void parse_graph(Cursor& c, bl_sample& s, bool& have_nodes, bool& have_refs) {
  uint64_t n;
  if (!c.map_header(n)) throw NeedsHost();
  for (uint64_t i = 0; i < n; ++i) {
    std::string_view key;
    if (!c.str(key)) throw NeedsHost();
    if (key == "nodes") {
      uint64_t count;
      if (!c.array_header(count)) throw NeedsHost();
      s.labels.clear();
      s.labels.reserve(count);
      for (uint64_t k = 0; k < count; ++k) {
        std::string_view label;
        if (!c.str(label)) throw NeedsHost();
        s.labels.push_back(label);
      }
      have_nodes = true;
    } else if (key == "edges") {
      s.num_edge_lists = 0;
      parse_edges(c, s);
    } else if (key == "reference_nodes") {
      uint64_t count;
      if (!c.array_header(count)) throw NeedsHost();
      s.reference_nodes.clear();
      for (uint64_t k = 0; k < count; ++k) {
        int64_t v;
        if (!c.integer(v)) throw NeedsHost();
        s.reference_nodes.push_back(as_i32(v));
      }
      have_refs = true;
    } else {
      c.skip(1);
    }
  }
}

// data.py:97-121 — the identifier tokens (endpoints of NextToken, in CPython set order) get one node per distinct
// lower-cased part and a HasSubtoken edge to it.
void add_open_vocab_nodes_and_edges(bl_sample& s, const bl_tokenizer& tok) {
  EdgeList* next_token = s.find_edge_list("NextToken");
  if (next_token == nullptr) return;
  PySetOrder token_nodes;
  const int64_t num_file_nodes = (int64_t)s.labels.size();
  for (size_t e = 0; e < next_token->src.size(); ++e) {
    int64_t a = next_token->src[e], b = next_token->tgt[e];
    if (a < 0 || b < 0 || a >= num_file_nodes || b >= num_file_nodes) throw NeedsHost();  // negative index / IndexError
    token_nodes.add(a);
    token_nodes.add(b);
  }
  s.subtoken_node.clear();
  s.new_labels.clear();
  // Built aside: `labels` grows while it is read, and `HasSubtoken` may already exist in the file.
  std::vector<int32_t> hs_src, hs_tgt;
  std::vector<int64_t> order;
  order.reserve(token_nodes.size());
  token_nodes.for_each([&](int64_t k) { order.push_back(k); });
  for (int64_t token_idx : order) {
    std::string_view label = s.labels[(size_t)token_idx];
    if (label.empty()) continue;
    unsigned char c0 = (unsigned char)label[0];
    if (!(is_upper(c0) || is_lower(c0) || c0 == '_')) continue;
    split_identifier(label, s.parts);
    for (const auto& span : s.parts.spans) {
      std::string part(s.parts.lowered.data() + span.first, span.second);
      auto it = s.subtoken_node.find(part);
      int32_t part_idx;
      if (it == s.subtoken_node.end()) {
        part_idx = as_i32((int64_t)s.labels.size());
        s.new_labels.push_back(part);
        s.labels.push_back(std::string_view(s.new_labels.back()));
        s.subtoken_node.emplace(std::move(part), part_idx);
      } else {
        part_idx = it->second;
      }
      hs_src.push_back((int32_t)token_idx);
      hs_tgt.push_back(part_idx);
    }
  }
  (void)tok;
  EdgeList& has_subtoken = s.edge_list("HasSubtoken");
  has_subtoken.src.swap(hs_src);
  has_subtoken.tgt.swap(hs_tgt);
}

int32_t decode_sample(const bl_shard& shard, int64_t index, const bl_tokenizer& tok, const char* const* edge_type_names,
                      int32_t num_edge_types, bl_sample& s, bl_sample_view& v) {
  memset(&v, 0, sizeof(v));
  const uint8_t* begin = shard.raw.data() + shard.offsets[(size_t)index];
  const uint8_t* end = shard.raw.data() + shard.offsets[(size_t)index + 1];
  v.raw = begin;
  v.raw_len = end - begin;
  v.num_edge_types = num_edge_types;
  v.max_subtokens = tok.kind == BL_SPLIT_TOKEN ? 1 : tok.max_subtokens;
  Cursor c{begin, end};
  if (c.at_nil()) { v.status = BL_SAMPLE_NIL; return BL_SHARDS_OK; }
  try {
    uint64_t n;
    if (!c.map_header(n)) throw NeedsHost();
    bool have_graph = false, have_nodes = false, have_refs = false, have_target_key = false;
    s.num_edge_lists = 0;
    s.labels.clear();
    s.reference_nodes.clear();
    for (uint64_t i = 0; i < n; ++i) {
      std::string_view key;
      if (!c.str(key)) throw NeedsHost();
      const uint8_t* value_begin = c.p;
      if (key == "graph") {
        have_nodes = have_refs = false;
        parse_graph(c, s, have_nodes, have_refs);
        have_graph = true;
      } else if (key == "candidate_rewrites") {
        c.skip(1);
        v.rewrites_off = value_begin - begin; v.rewrites_len = c.p - value_begin;
      } else if (key == "candidate_rewrite_metadata") {
        c.skip(1);
        v.metadata_off = value_begin - begin; v.metadata_len = c.p - value_begin;
      } else if (key == "candidate_rewrite_logprobs") {
        c.skip(1);
        v.logprobs_off = value_begin - begin; v.logprobs_len = c.p - value_begin;
      } else if (key == "target_fix_action_idx") {
        have_target_key = true;
        if (c.at_nil()) { ++c.p; v.has_target = 0; }
        else {
          int64_t t;
          if (!c.integer(t)) throw NeedsHost();
          v.has_target = 1; v.target_fix_action_idx = t;
        }
      } else {
        c.skip(1);
      }
    }
    if (!have_graph || !have_nodes || !have_refs || !have_target_key || v.rewrites_len == 0 || v.metadata_len == 0)
      throw NeedsHost();  // KeyError territory: let the host raise it
    EdgeList* child = s.find_edge_list("Child");
    if (child == nullptr) throw NeedsHost();  // basemodel.py:84 indexes graph["edges"]["Child"]

    for (std::string_view label : s.labels)
      if (!tok.label_supported(label)) throw NeedsHost();
    v.num_file_nodes = as_i32((int64_t)s.labels.size());

    add_open_vocab_nodes_and_edges(s, tok);
    const size_t num_nodes = s.labels.size();
    v.num_nodes = as_i32((int64_t)num_nodes);

    // node labels -> ids
    const int32_t T = v.max_subtokens;
    s.node_ids.assign(num_nodes * (size_t)T, 0);
    s.node_lens.resize(num_nodes);
    for (size_t node = 0; node < num_nodes; ++node) {
      int32_t count = tok.ids_of(s.labels[node], s.parts, s.node_ids.data() + node * (size_t)T);
      if (count < 0) throw NeedsHost();
      s.node_lens[node] = count;
    }

    // edges in the model's edge-type order
    s.edge_offsets.assign((size_t)num_edge_types + 1, 0);
    s.edge_src.clear();
    s.edge_tgt.clear();
    for (int32_t k = 0; k < num_edge_types; ++k) {
      EdgeList* list = s.find_edge_list(edge_type_names[k]);
      if (list != nullptr) {
        s.edge_src.insert(s.edge_src.end(), list->src.begin(), list->src.end());
        s.edge_tgt.insert(s.edge_tgt.end(), list->tgt.begin(), list->tgt.end());
      }
      s.edge_offsets[(size_t)k + 1] = (int64_t)s.edge_src.size();
    }

    // positional arguments of Call nodes (basemodel.py:84-88)
    s.call_args.clear();
    child = s.find_edge_list("Child");
    for (uint32_t pos : child->args_marked) {
      int64_t call = child->src[pos];
      if (call < 0 || call >= (int64_t)num_nodes) throw NeedsHost();
      if (s.labels[(size_t)call] == "Call") {
        s.call_args.push_back(child->src[pos]);
        s.call_args.push_back(child->tgt[pos]);
      }
    }

    v.node_ids = s.node_ids.data();
    v.node_lens = s.node_lens.data();
    v.edge_offsets = s.edge_offsets.data();
    v.edge_src = s.edge_src.data();
    v.edge_tgt = s.edge_tgt.data();
    v.num_reference_nodes = as_i32((int64_t)s.reference_nodes.size());
    v.reference_nodes = s.reference_nodes.data();
    v.num_call_args = (int32_t)(s.call_args.size() / 2);
    v.call_args = s.call_args.data();
    v.status = BL_SAMPLE_OK;
  } catch (const NeedsHost&) {
    v.status = BL_SAMPLE_NEEDS_HOST;
  } catch (const ParseError&) {
    v.status = BL_SAMPLE_NEEDS_HOST;  // cannot happen for indexed objects; the host path will raise if it does
  }
  return BL_SHARDS_OK;
}

// One object's contribution to the metadata: what GnnBugLabModel.update_metadata_from does with it in the host chain —
// BugLabData.as_graph_data (needs graph.nodes / graph.reference_nodes / target_fix_action_idx, appends the open-vocabulary
// nodes and the HasSubtoken edges), then the node model counts the (sub)tokens of every node label and the graph model
// records the edge-type names.  Returns BL_SAMPLE_OK / BL_SAMPLE_NIL / BL_SAMPLE_NEEDS_HOST; nothing is added unless OK.
int32_t collect_metadata(const bl_shard& shard, int64_t index, const bl_tokenizer& tok, bl_sample& s, bl_metadata& md) {
  const uint8_t* begin = shard.raw.data() + shard.offsets[(size_t)index];
  const uint8_t* end = shard.raw.data() + shard.offsets[(size_t)index + 1];
  Cursor c{begin, end};
  if (c.at_nil()) return BL_SAMPLE_NIL;
  try {
    uint64_t n;
    if (!c.map_header(n)) throw NeedsHost();
    bool have_graph = false, have_nodes = false, have_refs = false, have_target_key = false, has_target = false;
    int64_t target = 0;
    s.num_edge_lists = 0;
    s.labels.clear();
    s.reference_nodes.clear();
    for (uint64_t i = 0; i < n; ++i) {
      std::string_view key;
      if (!c.str(key)) throw NeedsHost();
      if (key == "graph") {
        have_nodes = have_refs = false;
        parse_graph(c, s, have_nodes, have_refs);
        have_graph = true;
      } else if (key == "target_fix_action_idx") {
        have_target_key = true;
        if (c.at_nil()) { ++c.p; has_target = false; }
        else {
          if (!c.integer(target)) throw NeedsHost();
          has_target = true;
        }
      } else {
        c.skip(1);
      }
    }
    if (!have_graph || !have_nodes || !have_refs || !have_target_key) throw NeedsHost();  // KeyError territory
    if (has_target && (target < 0 || target >= (int64_t)s.reference_nodes.size())) throw NeedsHost();  // inverse[target]
    for (std::string_view label : s.labels)
      if (!tok.label_supported(label)) throw NeedsHost();
    add_open_vocab_nodes_and_edges(s, tok);
  } catch (const NeedsHost&) {
    return BL_SAMPLE_NEEDS_HOST;
  } catch (const ParseError&) {
    return BL_SAMPLE_NEEDS_HOST;
  }
  // from here on nothing throws NeedsHost: the sample is counted as a whole
  for (std::string_view label : s.labels) {
    if (tok.kind == BL_SPLIT_TOKEN) {
      md.token_counts[std::string(label)] += 1;
    } else {
      split_identifier(label, s.parts);
      for (const auto& span : s.parts.spans)
        md.token_counts[std::string(s.parts.lowered.data() + span.first, span.second)] += 1;
    }
  }
  for (size_t i = 0; i < s.num_edge_lists; ++i) md.edge_types[std::string(s.edge_lists[i].name)] += 1;
  md.num_samples += 1;
  return BL_SAMPLE_OK;
}

// gzip (RFC 1952) members back to back, as Python's gzip module reads them.
int32_t inflate_all(const uint8_t* gz, size_t gz_len, std::vector<uint8_t>& out) {
  z_stream zs;
  memset(&zs, 0, sizeof(zs));
  if (inflateInit2(&zs, 15 + 16) != Z_OK) return BL_SHARDS_ERR_GZIP;
  zs.next_in = const_cast<Bytef*>(gz);
  size_t in_left = gz_len;
  out.clear();
  out.resize(std::max<size_t>(gz_len * 3, 1 << 16));
  size_t produced = 0;
  int32_t status = BL_SHARDS_OK;
  bool any_member = false;
  for (;;) {
    if (produced == out.size()) out.resize(out.size() * 2);
    uInt in_chunk = (uInt)std::min<size_t>(in_left, 1u << 30);
    uInt out_chunk = (uInt)std::min<size_t>(out.size() - produced, 1u << 30);
    zs.avail_in = in_chunk;
    zs.next_out = out.data() + produced;
    zs.avail_out = out_chunk;
    int rc = inflate(&zs, Z_NO_FLUSH);
    in_left -= in_chunk - zs.avail_in;
    produced += out_chunk - zs.avail_out;
    if (rc == Z_STREAM_END) {
      any_member = true;
      // trailing zero padding is ignored by Python's gzip reader; another member restarts the inflater
      while (in_left > 0 && *zs.next_in == 0) { ++zs.next_in; --in_left; }
      if (in_left == 0) break;
      if (inflateReset(&zs) != Z_OK) { status = BL_SHARDS_ERR_GZIP; break; }
      continue;
    }
    if (rc == Z_OK) {
      if (in_left == 0 && zs.avail_out != 0) { status = BL_SHARDS_ERR_GZIP; break; }  // truncated stream
      continue;
    }
    if (rc == Z_BUF_ERROR && zs.avail_out == 0) continue;
    status = BL_SHARDS_ERR_GZIP;
    break;
  }
  inflateEnd(&zs);
  out.resize(produced);
  if (!any_member && produced == 0 && status == BL_SHARDS_OK) status = BL_SHARDS_ERR_GZIP;
  return status;
}

void index_objects(bl_shard& shard) {
  Cursor c{shard.raw.data(), shard.raw.data() + shard.raw.size()};
  shard.offsets.clear();
  shard.offsets.push_back(0);
  while (c.p < c.end) {
    try {
      c.skip(0);
    } catch (const ParseError&) {
      if (shard.status == BL_SHARDS_OK) shard.status = BL_SHARDS_ERR_MSGPACK;
      break;
    }
    shard.offsets.push_back(c.p - shard.raw.data());
  }
}

}  // namespace

extern "C" {

int32_t bl_shards_version(void) { return 1; }

const char* bl_shards_error_string(int32_t code) {
  switch (code) {
    case BL_SHARDS_OK: return "ok";
    case BL_SHARDS_ERR_IO: return "cannot read shard file";
    case BL_SHARDS_ERR_GZIP: return "corrupt or truncated gzip stream";
    case BL_SHARDS_ERR_ARG: return "bad argument";
    case BL_SHARDS_ERR_MSGPACK: return "malformed msgpack object";
    case BL_SHARDS_ERR_NOMEM: return "out of memory";
    default: return "unknown error";
  }
}

int32_t bl_shard_open_buffer(const uint8_t* gz, int64_t gz_len, bl_shard** out) {
  if (out == nullptr || (gz == nullptr && gz_len != 0) || gz_len < 0) return BL_SHARDS_ERR_ARG;
  bl_shard* shard = nullptr;
  try {
    shard = new bl_shard();
    shard->status = inflate_all(gz, (size_t)gz_len, shard->raw);
    index_objects(*shard);
  } catch (...) {  // std::bad_alloc / length_error on absurd sizes
    delete shard;
    return BL_SHARDS_ERR_NOMEM;
  }
  *out = shard;
  return BL_SHARDS_OK;
}

int32_t bl_shard_open(const char* path, bl_shard** out) {
  if (path == nullptr || out == nullptr) return BL_SHARDS_ERR_ARG;
  FILE* f = fopen(path, "rb");
  if (f == nullptr) return BL_SHARDS_ERR_IO;
  std::vector<uint8_t> gz;
  uint8_t buf[1 << 16];
  size_t got;
  bool failed = false;
  try {
    while ((got = fread(buf, 1, sizeof(buf), f)) > 0) gz.insert(gz.end(), buf, buf + got);
  } catch (...) {
    fclose(f);
    return BL_SHARDS_ERR_NOMEM;
  }
  failed = ferror(f) != 0;
  fclose(f);
  if (failed) return BL_SHARDS_ERR_IO;
  return bl_shard_open_buffer(gz.data(), (int64_t)gz.size(), out);
}

void bl_shard_close(bl_shard* shard) { delete shard; }

int64_t bl_shard_num_objects(const bl_shard* shard) { return shard ? (int64_t)shard->offsets.size() - 1 : 0; }
int64_t bl_shard_raw_bytes(const bl_shard* shard) { return shard ? (int64_t)shard->raw.size() : 0; }
int32_t bl_shard_status(const bl_shard* shard) { return shard ? shard->status : BL_SHARDS_ERR_ARG; }

int32_t bl_shard_object(const bl_shard* shard, int64_t index, const uint8_t** data, int64_t* len) {
  if (!shard || !data || !len || index < 0 || index >= bl_shard_num_objects(shard)) return BL_SHARDS_ERR_ARG;
  *data = shard->raw.data() + shard->offsets[(size_t)index];
  *len = shard->offsets[(size_t)index + 1] - shard->offsets[(size_t)index];
  return BL_SHARDS_OK;
}

int32_t bl_tokenizer_create(const uint8_t* blob, const int64_t* offsets, const int32_t* ids, int32_t num_tokens,
                            int32_t unk_id, int32_t splitting_kind, int32_t max_subtokens,
                            const int32_t* lower_variant_codepoints, int32_t num_lower_variant, bl_tokenizer** out) {
  if (!out || num_tokens < 0 || (num_tokens > 0 && (!blob || !offsets || !ids)) || max_subtokens < 1 ||
      (splitting_kind != BL_SPLIT_TOKEN && splitting_kind != BL_SPLIT_SUBTOKEN) || num_lower_variant < 0 ||
      (num_lower_variant > 0 && !lower_variant_codepoints))
    return BL_SHARDS_ERR_ARG;
  bl_tokenizer* tok = nullptr;
  try {
  tok = new bl_tokenizer();
  if (num_tokens > 0) tok->blob.assign(reinterpret_cast<const char*>(blob), (size_t)offsets[num_tokens]);
  size_t capacity = 16;
  while (capacity < (size_t)num_tokens * 2 + 2) capacity <<= 1;
  tok->slots.assign(capacity, bl_tokenizer::Slot{0, 0, -1, 0});
  tok->mask = capacity - 1;
  for (int32_t t = 0; t < num_tokens; ++t) {
    int64_t off = offsets[t];
    int32_t len = (int32_t)(offsets[t + 1] - off);
    uint64_t h = hash_bytes(tok->blob.data() + off, (size_t)len);
    uint64_t i = h & tok->mask;
    bool duplicate = false;
    while (tok->slots[i].len >= 0) {
      const auto& s = tok->slots[i];
      if (s.hash == h && s.len == len && memcmp(tok->blob.data() + s.off, tok->blob.data() + off, (size_t)len) == 0) {
        duplicate = true;
        break;
      }
      i = (i + 1) & tok->mask;
    }
    if (duplicate) tok->slots[i].id = ids[t];
    else tok->slots[i] = bl_tokenizer::Slot{h, off, len, ids[t]};
  }
  tok->unk_id = unk_id;
  tok->kind = splitting_kind;
  tok->max_subtokens = splitting_kind == BL_SPLIT_TOKEN ? 1 : max_subtokens;
  tok->lower_variant.assign(lower_variant_codepoints, lower_variant_codepoints + num_lower_variant);
  std::sort(tok->lower_variant.begin(), tok->lower_variant.end());
  } catch (...) {
    delete tok;
    return BL_SHARDS_ERR_NOMEM;
  }
  *out = tok;
  return BL_SHARDS_OK;
}

void bl_tokenizer_destroy(bl_tokenizer* tok) { delete tok; }

int32_t bl_tokenizer_ids(const bl_tokenizer* tok, const uint8_t* label, int64_t len, int32_t* ids_out) {
  if (!tok || (!label && len != 0) || len < 0 || !ids_out) return -1;
  std::string_view sv(reinterpret_cast<const char*>(label), (size_t)len);
  if (!tok->label_supported(sv)) return -1;
  Parts scratch;
  return tok->ids_of(sv, scratch, ids_out);
}

int32_t bl_sample_create(bl_sample** out) {
  if (!out) return BL_SHARDS_ERR_ARG;
  try {
    *out = new bl_sample();
  } catch (...) {
    return BL_SHARDS_ERR_NOMEM;
  }
  return BL_SHARDS_OK;
}

void bl_sample_destroy(bl_sample* sample) { delete sample; }

int32_t bl_sample_decode(const bl_shard* shard, int64_t index, const bl_tokenizer* tok,
                         const char* const* edge_type_names, int32_t num_edge_types, bl_sample* sample,
                         bl_sample_view* view) {
  if (!shard || !tok || !sample || !view || num_edge_types < 0 || (num_edge_types > 0 && !edge_type_names) || index < 0 ||
      index >= bl_shard_num_objects(shard))
    return BL_SHARDS_ERR_ARG;
  try {
    return decode_sample(*shard, index, *tok, edge_type_names, num_edge_types, *sample, *view);
  } catch (...) {  // decode_sample handles its own control-flow exceptions; what is left is allocation failure
    return BL_SHARDS_ERR_NOMEM;
  }
}

int32_t bl_sample_decode_many(const bl_shard* shard, const int64_t* indices, int32_t count, const bl_tokenizer* tok,
                              const char* const* edge_type_names, int32_t num_edge_types, bl_sample* const* samples,
                              bl_sample_view* views) {
  if (count < 0 || (count > 0 && (!indices || !samples || !views))) return BL_SHARDS_ERR_ARG;
  for (int32_t i = 0; i < count; ++i) {
    int32_t rc = bl_sample_decode(shard, indices[i], tok, edge_type_names, num_edge_types, samples[i], &views[i]);
    if (rc != BL_SHARDS_OK) return rc;
  }
  return BL_SHARDS_OK;
}

int64_t bl_shard_non_nil(const bl_shard* shard, int64_t* indices) {
  if (!shard) return -1;
  int64_t k = 0;
  const int64_t n = (int64_t)shard->offsets.size() - 1;
  for (int64_t i = 0; i < n; ++i) {
    const bool nil = shard->offsets[(size_t)i + 1] - shard->offsets[(size_t)i] == 1 && shard->raw[(size_t)shard->offsets[(size_t)i]] == 0xc0;
    if (!nil) {
      if (indices) indices[k] = i;
      ++k;
    }
  }
  return k;
}

int32_t bl_metadata_create(bl_metadata** out) {
  if (!out) return BL_SHARDS_ERR_ARG;
  try {
    *out = new bl_metadata();
  } catch (...) {
    return BL_SHARDS_ERR_NOMEM;
  }
  return BL_SHARDS_OK;
}

void bl_metadata_destroy(bl_metadata* md) { delete md; }

int32_t bl_metadata_add(bl_metadata* md, const bl_shard* shard, const int64_t* indices, int32_t count, const bl_tokenizer* tok,
                        bl_sample* scratch, int32_t* needs_host, int32_t* num_needs_host) {
  if (!md || !shard || !tok || !scratch || !num_needs_host || count < 0 || (count > 0 && (!indices || !needs_host)))
    return BL_SHARDS_ERR_ARG;
  *num_needs_host = 0;
  try {
    for (int32_t i = 0; i < count; ++i) {
      if (indices[i] < 0 || indices[i] >= bl_shard_num_objects(shard)) return BL_SHARDS_ERR_ARG;
      if (collect_metadata(*shard, indices[i], *tok, *scratch, *md) == BL_SAMPLE_NEEDS_HOST) needs_host[(*num_needs_host)++] = i;
    }
  } catch (...) {
    return BL_SHARDS_ERR_NOMEM;
  }
  return BL_SHARDS_OK;
}

int64_t bl_metadata_num_samples(const bl_metadata* md) { return md ? md->num_samples : -1; }

int64_t bl_metadata_size(const bl_metadata* md, int32_t which, int64_t* blob_bytes) {
  if (!md || (which != 0 && which != 1)) return -1;
  const auto& table = which == 0 ? md->token_counts : md->edge_types;
  int64_t bytes = 0;
  for (const auto& kv : table) bytes += (int64_t)kv.first.size();
  if (blob_bytes) *blob_bytes = bytes;
  return (int64_t)table.size();
}

int32_t bl_metadata_export(const bl_metadata* md, int32_t which, uint8_t* blob, int64_t* offsets, int64_t* counts) {
  if (!md || (which != 0 && which != 1) || !offsets) return BL_SHARDS_ERR_ARG;
  const auto& table = which == 0 ? md->token_counts : md->edge_types;
  int64_t off = 0, k = 0;
  offsets[0] = 0;
  for (const auto& kv : table) {
    if (!kv.first.empty()) {
      if (!blob) return BL_SHARDS_ERR_ARG;
      memcpy(blob + off, kv.first.data(), kv.first.size());
    }
    off += (int64_t)kv.first.size();
    if (counts) counts[k] = kv.second;
    offsets[++k] = off;
  }
  return BL_SHARDS_OK;
}

int64_t bl_pyset_iteration_order(const int64_t* values, int64_t n, int64_t* out) {
  if (n < 0 || (n > 0 && (!values || !out))) return -1;
  try {
    PySetOrder set;
    for (int64_t i = 0; i < n; ++i) {
      if (values[i] < 0 || values[i] >= ((int64_t)1 << 61) - 1) return -1;
      set.add(values[i]);
    }
    int64_t k = 0;
    set.for_each([&](int64_t v) { out[k++] = v; });
    return k;
  } catch (...) {
    return -1;
  }
}

}  // extern "C"
