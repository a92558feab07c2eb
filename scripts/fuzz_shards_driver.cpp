// Memory-safety fuzz driver for the shard decoder: decodes every object of every file given on the command line.
//   g++ -O1 -g -std=c++17 -fsanitize=address,undefined -o /tmp/fuzz_shards scripts/fuzz_shards_driver.cpp \
//       neurips21-self-supervised-bug-detection-and-repair_b200/csrc/shards/shards.cpp -lz
//   python scripts/fuzz_shards_gen.py SEED /tmp/corpus 6000 && /tmp/fuzz_shards /tmp/corpus/*.gz
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
  long ok = 0, host = 0, nil = 0, objs = 0, files = 0; long long checksum = 0;
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
    bl_shard_close(sh);
  }
  bl_sample_destroy(sample); bl_tokenizer_destroy(tok);
  printf("files %ld objects %ld ok %ld host %ld nil %ld checksum %lld\n", files, objs, ok, host, nil, checksum);
  return 0;
}
