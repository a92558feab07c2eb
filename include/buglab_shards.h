/*
 * buglab_shards.h — C ABI of the native `.msgpack.l.gz` shard decoder / graph tensoriser.
 *
 * Host-only library (C++17 + zlib, no CUDA): `libbuglab_shards.so`.  It replaces, for the training data path, the
 * per-object Python work of the reference
 *     buglab/utils/msgpackutils.py:11-14   (gzip stream -> msgpack.Unpacker -> Python dicts)
 *     buglab/representations/data.py:97-121 (add_open_vocab_nodes_and_edges: subtoken nodes + HasSubtoken edges)
 *     buglab/representations/data.py:139-167 (as_graph_data: edge lists -> int32 arrays)
 *     ptgnn StrElementRepresentationModel.tensorize (split_identifier_into_parts + vocabulary lookup, <= T ids per node)
 * and writes packed int32 arrays directly.  The wire format is unchanged (files stay byte-identical); results are
 * bit-identical to the host-language path, which remains the checker (tests/test_shards_cpu.py).
 *
 * Anything the native path cannot reproduce EXACTLY (labels with code points that change under Python's str.lower(),
 * invalid UTF-8, node ids that are not int32, negative / out-of-range token indices, missing unknown-token id ...)
 * is not guessed: the sample is reported as BL_SAMPLE_NEEDS_HOST and the caller runs the host-language path on the raw
 * msgpack bytes of that object (`raw`, `raw_len`), so behaviour including exceptions is the reference's.
 *
 * Threading: a `bl_shard` and a `bl_tokenizer` are immutable after creation and may be shared between threads; a
 * `bl_sample` belongs to one thread at a time.  No call takes a lock or touches the Python runtime, so host threads
 * decode in parallel (ctypes releases the GIL for the duration of a call).
 */
#ifndef BUGLAB_SHARDS_H_
#define BUGLAB_SHARDS_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BL_SHARDS_OK 0
#define BL_SHARDS_ERR_IO 1        /* file cannot be read */
#define BL_SHARDS_ERR_GZIP 2      /* not a gzip stream / corrupt deflate data */
#define BL_SHARDS_ERR_ARG 3       /* bad argument (null handle, index out of range) */
#define BL_SHARDS_ERR_MSGPACK 4   /* malformed msgpack inside an object */
#define BL_SHARDS_ERR_NOMEM 5     /* allocation failed (no C++ exception ever crosses this boundary) */

#define BL_SAMPLE_OK 0            /* arrays below are valid */
#define BL_SAMPLE_NIL 1           /* the object is msgpack nil: skipped by the loader (msgpackutils.py:38) */
#define BL_SAMPLE_NEEDS_HOST 2    /* decode `raw` with the host-language path instead */

#define BL_SPLIT_TOKEN 0          /* token_splitting="token": one id per node label */
#define BL_SPLIT_SUBTOKEN 1       /* token_splitting="subtoken": split_identifier_into_parts, first T parts */

typedef struct bl_shard bl_shard;
typedef struct bl_tokenizer bl_tokenizer;
typedef struct bl_sample bl_sample;

/* One decoded sample.  Every pointer aims into memory owned by the `bl_sample` (or, for `raw`, by the `bl_shard`) and
 * stays valid until the next bl_sample_decode on the same handle / bl_sample_destroy (bl_shard_close for `raw`). */
typedef struct bl_sample_view {
  int32_t status;               /* BL_SAMPLE_* */
  int32_t num_file_nodes;       /* len(graph["nodes"]) as stored */
  int32_t num_nodes;            /* after appending the open-vocabulary subtoken nodes (data.py:97-121) */
  int32_t max_subtokens;        /* T: row length of node_ids */
  const int32_t* node_ids;      /* [num_nodes, T] vocabulary ids, zero padded */
  const int32_t* node_lens;     /* [num_nodes] number of valid ids per row */
  int32_t num_edge_types;       /* = the count passed to bl_sample_decode */
  int32_t num_reference_nodes;
  const int64_t* edge_offsets;  /* [num_edge_types + 1] into edge_src / edge_tgt, in the caller's edge-type order */
  const int32_t* edge_src;
  const int32_t* edge_tgt;
  const int32_t* reference_nodes;   /* graph["reference_nodes"] as stored (one per candidate rewrite) */
  int32_t num_call_args;
  int32_t has_target;               /* 0 when target_fix_action_idx is nil */
  const int32_t* call_args;         /* [num_call_args, 2] (Call node, argument node) in "Child" edge order, the
                                       positional-argument table of basemodel.py:84-88 */
  int64_t target_fix_action_idx;
  const uint8_t* raw;               /* the whole msgpack object */
  int64_t raw_len;
  /* byte ranges inside `raw` of the small fields the host decodes itself; length 0 = key absent */
  int64_t rewrites_off, rewrites_len;           /* "candidate_rewrites" */
  int64_t metadata_off, metadata_len;           /* "candidate_rewrite_metadata" */
  int64_t logprobs_off, logprobs_len;           /* "candidate_rewrite_logprobs" */
} bl_sample_view;

int32_t bl_shards_version(void);
const char* bl_shards_error_string(int32_t code);

/* Reads and inflates a whole `.msgpack.l.gz` file (multi-member gzip accepted) and indexes its top-level msgpack
 * objects.  A stream that breaks half way keeps the objects before the break — as the reference's generator yields them
 * before raising (msgpackutils.py:44-45) — and reports the break through bl_shard_status. */
int32_t bl_shard_open(const char* path, bl_shard** out);
/* Same, from a gzip byte buffer already in memory. */
int32_t bl_shard_open_buffer(const uint8_t* gz, int64_t gz_len, bl_shard** out);
void bl_shard_close(bl_shard* shard);
int64_t bl_shard_num_objects(const bl_shard* shard);
int64_t bl_shard_raw_bytes(const bl_shard* shard);       /* inflated size */
int32_t bl_shard_status(const bl_shard* shard);          /* BL_SHARDS_OK, or the error that truncated the stream */
/* Byte range of object `index` inside the inflated stream (for hosts that want the generic decoder). */
int32_t bl_shard_object(const bl_shard* shard, int64_t index, const uint8_t** data, int64_t* len);

/* Vocabulary of the node-label embedder (ptgnn StrElementRepresentationModel): `num_tokens` UTF-8 strings stored back
 * to back in `blob`, token i = blob[offsets[i] .. offsets[i+1]) with id ids[i].  `unk_id` < 0 = vocabulary without
 * %UNK% (any miss then reports BL_SAMPLE_NEEDS_HOST, where the host raises as the reference does).
 * `lower_variant_codepoints`: sorted code points c >= 0x80 with chr(c).lower() != chr(c) in the host's Unicode tables;
 * labels containing one are routed to the host path. */
int32_t bl_tokenizer_create(const uint8_t* blob, const int64_t* offsets, const int32_t* ids, int32_t num_tokens,
                            int32_t unk_id, int32_t splitting_kind, int32_t max_subtokens,
                            const int32_t* lower_variant_codepoints, int32_t num_lower_variant, bl_tokenizer** out);
void bl_tokenizer_destroy(bl_tokenizer* tok);
/* Tokenises one label (testing hook): writes <= max_subtokens ids, returns their count, or -1 = needs host. */
int32_t bl_tokenizer_ids(const bl_tokenizer* tok, const uint8_t* label, int64_t len, int32_t* ids_out);

int32_t bl_sample_create(bl_sample** out);
void bl_sample_destroy(bl_sample* sample);
/* Decodes + tensorises object `index`.  `edge_type_names`: the model's edge types (forward kinds only, metadata order),
 * NUL-terminated UTF-8.  Returns BL_SHARDS_OK and fills `view` (view->status tells what was produced). */
int32_t bl_sample_decode(const bl_shard* shard, int64_t index, const bl_tokenizer* tok,
                         const char* const* edge_type_names, int32_t num_edge_types, bl_sample* sample,
                         bl_sample_view* view);
/* The same for `count` objects in ONE call: object indices[i] is decoded into samples[i] (distinct handles) and views[i].
 * A host loader thread decodes a whole chunk of a shard without returning to its interpreter in between (one GIL
 * hand-off per chunk instead of one per sample).  Stops at the first error and returns its code. */
int32_t bl_sample_decode_many(const bl_shard* shard, const int64_t* indices, int32_t count, const bl_tokenizer* tok,
                              const char* const* edge_type_names, int32_t num_edge_types, bl_sample* const* samples,
                              bl_sample_view* views);

/* ---- metadata pass (vocabulary + edge types) -----------------------------------------------------------------------------
 * What the reference does once before training, one Python object at a time: `model.compute_metadata(data)` ->
 * GnnBugLabModel.update_metadata_from (buglab/models/gnn.py:350-358) -> BugLabData.as_graph_data (data.py:139-167, incl. the
 * open-vocabulary nodes) -> the node model counts the (sub)tokens of every node label, the graph model records the edge-type
 * names.  Both results are order independent, so shards are counted in parallel into per-thread accumulators and merged. */
typedef struct bl_metadata bl_metadata;
int32_t bl_metadata_create(bl_metadata** out);
void bl_metadata_destroy(bl_metadata* md);
/* Writes the indices of the non-nil objects of the shard (capacity bl_shard_num_objects) and returns their number;
 * `indices` may be NULL to count only.  (The loader's element limit counts non-nil objects, msgpackutils.py:38-43.) */
int64_t bl_shard_non_nil(const bl_shard* shard, int64_t* indices);
/* Adds objects indices[0..count) to `md`.  `tok`: a tokenizer made by bl_tokenizer_create (only its splitting kind and
 * code-point table are used; it may hold no vocabulary).  `scratch`: a bl_sample of the calling thread.  Objects the native
 * path declines (anything the decoder would hand to the host, see BL_SAMPLE_NEEDS_HOST) are NOT added: their positions in
 * `indices` are written to needs_host (capacity count) and counted in *num_needs_host, for the host to add. */
int32_t bl_metadata_add(bl_metadata* md, const bl_shard* shard, const int64_t* indices, int32_t count, const bl_tokenizer* tok,
                        bl_sample* scratch, int32_t* needs_host, int32_t* num_needs_host);
int64_t bl_metadata_num_samples(const bl_metadata* md);
/* which = 0: (sub)token counts; which = 1: edge-type names (count = number of samples that have the type).
 * bl_metadata_size returns the number of entries and the total UTF-8 bytes of their keys; bl_metadata_export writes the keys
 * back to back into blob with offsets[entries + 1] and counts[entries] (unspecified order). */
int64_t bl_metadata_size(const bl_metadata* md, int32_t which, int64_t* blob_bytes);
int32_t bl_metadata_export(const bl_metadata* md, int32_t which, uint8_t* blob, int64_t* offsets, int64_t* counts);

/* Iteration order of a CPython `set` filled with the given non-negative ints in this order (testing hook for the
 * emulation that data.py:103-108 makes load-bearing: the order decides the ids of the subtoken nodes).
 * Returns the number of distinct values written to `out` (capacity n), or -1 for unsupported values. */
int64_t bl_pyset_iteration_order(const int64_t* values, int64_t n, int64_t* out);

#ifdef __cplusplus
}
#endif
#endif /* BUGLAB_SHARDS_H_ */
