// Memory-safety fuzz driver for the shard decoder: decodes every object of every file given on the command line.
//   g++ -O1 -g -std=c++17 -fsanitize=address,undefined -o /tmp/fuzz_shards scripts/fuzz_shards_driver.cpp \
//       neurips21-self-supervised-bug-detection-and-repair_b200/csrc/shards/shards.cpp -lz
//   python scripts/fuzz_shards_gen.py SEED /tmp/corpus 6000 && /tmp/fuzz_shards /tmp/corpus/*.gz
// Round 2 (decode_many, metadata pass and non-nil index added to the driver): 12 000 byte-mutated files (seeds 11-13), no report.
// Round 1: 19 500 byte-mutated files (seeds 1-4), no sanitizer report; 15 000 structure-level differential cases against the
// Python path (tests/test_shards_cpu.py::mutate_sample, seeds 1-3): identical tensors or the same exception type in all.
#include "../include/buglab_shards.h"
#include <cstdio>
#include <cstring>
#include <vector>
#include <string>
int main(int argc, char** argv) {
  const char* toks[] = {"%PAD%", "%UNK%", "foo", "bar", "call", "name", "x", "loss", "config", "step", "path"};
  std::string blob; std::vector<int64_t> offs{0}; std::vector<int32_t> ids;
  for (int i = 0; i < 11; ++i) { blob += toks[i]; offs.push_back((int64_t)blob.size()); ids.push_back(i); }
  int32_t variants[] = {0xC9, 0x391};
  bl_tokenizer* tok = nullptr;
  if (bl_tokenizer_create((const uint8_t*)blob.data(), offs.data(), ids.data(), 11, 1, BL_SPLIT_SUBTOKEN, 6, variants, 2, &tok)) return 2;
  const char* names[] = {"Child", "NextToken", "HasSubtoken", "LastMayWrite", "Nope"};
  bl_sample* sample = nullptr; bl_sample_create(&sample);
  bl_tokenizer* splitter = nullptr;  // no vocabulary: the metadata pass only splits
  if (bl_tokenizer_create(nullptr, nullptr, nullptr, 0, -1, BL_SPLIT_SUBTOKEN, 6, variants, 2, &splitter)) return 2;
  bl_metadata* md = nullptr;
  if (bl_metadata_create(&md)) return 2;
  long ok = 0, host = 0, nil = 0, objs = 0, files = 0, counted = 0, handed_back = 0; long long checksum = 0;
  for (int a = 1; a < argc; ++a) {
    bl_shard* sh = nullptr;
    if (bl_shard_open(argv[a], &sh) != 0) continue;
    ++files;
    int64_t n = bl_shard_num_objects(sh);
    for (int64_t i = 0; i < n; ++i) {
      bl_sample_view v;
      if (bl_sample_decode(sh, i, tok, names, 5, sample, &v) != 0) return 3;
      ++objs;
      if (v.status == 0) {
        ++ok;
        for (int64_t k = 0; k < (int64_t)v.num_nodes * v.max_subtokens; ++k) checksum += v.node_ids[k];
        for (int64_t k = 0; k < v.edge_offsets[v.num_edge_types]; ++k) checksum += v.edge_src[k] ^ v.edge_tgt[k];
        for (int k = 0; k < v.num_reference_nodes; ++k) checksum += v.reference_nodes[k];
        for (int k = 0; k < 2 * v.num_call_args; ++k) checksum += v.call_args[k];
        if (v.rewrites_off + v.rewrites_len > v.raw_len || v.metadata_off + v.metadata_len > v.raw_len) return 4;
      } else if (v.status == 1) ++nil; else ++host;
    }
    // the chunk entry point over all objects at once (distinct handles) must agree with the per-object calls above ...
    if (n > 0 && n <= 64) {
      std::vector<int64_t> idx((size_t)n);
      std::vector<bl_sample*> many((size_t)n, nullptr);
      std::vector<bl_sample_view> views((size_t)n);
      for (int64_t i = 0; i < n; ++i) { idx[(size_t)i] = i; bl_sample_create(&many[(size_t)i]); }
      if (bl_sample_decode_many(sh, idx.data(), (int32_t)n, tok, names, 5, many.data(), views.data()) != 0) return 5;
      for (int64_t i = 0; i < n; ++i) {
        bl_sample_view v;
        if (bl_sample_decode(sh, i, tok, names, 5, sample, &v) != 0) return 3;
        if (v.status != views[(size_t)i].status || v.num_nodes != views[(size_t)i].num_nodes) return 6;
        if (v.status == 0)
          for (int64_t k = 0; k < (int64_t)v.num_nodes * v.max_subtokens; ++k)
            if (v.node_ids[k] != views[(size_t)i].node_ids[k]) return 7;
      }
      for (auto* m : many) bl_sample_destroy(m);
      // ... and the metadata pass over the non-nil objects: every object is either counted or handed back
      std::vector<int64_t> not_nil((size_t)n);
      int64_t k = bl_shard_non_nil(sh, not_nil.data());
      if (k < 0 || k > n || bl_shard_non_nil(sh, nullptr) != k) return 8;
      std::vector<int32_t> declined((size_t)n + 1);
      int32_t num_declined = -1;
      int64_t before = bl_metadata_num_samples(md);
      if (bl_metadata_add(md, sh, not_nil.data(), (int32_t)k, splitter, sample, declined.data(), &num_declined) != 0) return 9;
      if (num_declined < 0 || num_declined > k || bl_metadata_num_samples(md) - before != k - num_declined) return 10;
      counted += k - num_declined; handed_back += num_declined;
    }
    bl_shard_close(sh);
  }
  for (int which = 0; which < 2; ++which) {
    int64_t bytes = -1, entries = bl_metadata_size(md, which, &bytes);
    if (entries < 0 || bytes < 0) return 11;
    std::vector<uint8_t> keys((size_t)bytes + 1);
    std::vector<int64_t> offsets((size_t)entries + 1), counts((size_t)entries + 1);
    if (bl_metadata_export(md, which, keys.data(), offsets.data(), counts.data()) != 0 || offsets[(size_t)entries] != bytes) return 12;
    for (int64_t e = 0; e < entries; ++e) checksum += counts[(size_t)e] + (offsets[(size_t)e + 1] - offsets[(size_t)e]);
  }
  bl_metadata_destroy(md);
  bl_sample_destroy(sample); bl_tokenizer_destroy(tok); bl_tokenizer_destroy(splitter);
  printf("files %ld objects %ld ok %ld host %ld nil %ld metadata: counted %ld handed back %ld checksum %lld\n", files, objs, ok,
         host, nil, counted, handed_back, checksum);
  return 0;
}
