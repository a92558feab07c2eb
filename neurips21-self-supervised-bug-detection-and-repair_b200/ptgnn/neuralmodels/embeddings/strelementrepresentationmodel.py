"""String element (node label) representation: vocabulary + embedding (+ subtoken max-pool on the GPU).

Reference wiring: buglab/models/modelregistry.py:57-66,79-82 builds
``StrElementRepresentationModel(embedding_size=H, token_splitting="subtoken", max_num_subtokens=6,
subtoken_combination="max", vocabulary_size=15000)``.  SURVEY.md §8a P2 states the semantics:
vocabulary of identifier subtokens (``split_identifier_into_parts``), <= ``max_num_subtokens`` ids per node,
``Embedding -> dropout -> masked max over subtokens``.  The lookup + max-pool (+dropout) is ONE kernel
(``bl_subtoken_maxpool_fwd``); the padded ``[N, T]`` id table never materialises an ``[N, T, H]`` tensor.
"""
from collections import Counter
from typing import Any, Dict, List, Optional, Tuple, Union

import numpy as np
import torch
from dpu_utils.codeutils import split_identifier_into_parts
from dpu_utils.mlutils import Vocabulary
from torch import nn

from ptgnn.baseneuralmodel import AbstractNeuralModel


class TokenUnitEmbedder(nn.Module):
    """Whole-token embedding (used for edge features when ``edge_feature_size > 0``; off by default)."""

    def __init__(self, vocabulary_size: int, embedding_size: int, dropout_rate: float, padding_idx: int = 0):
        super().__init__()
        self.__embeddings = nn.Embedding(vocabulary_size, embedding_size, padding_idx=padding_idx)
        self.__dropout_layer = nn.Dropout(p=dropout_rate)

    @property
    def embedding_layer(self) -> nn.Embedding:
        return self.__embeddings

    def forward(self, token_idxs: torch.Tensor) -> torch.Tensor:
        return self.__dropout_layer(self.__embeddings(token_idxs.long()))


class SubtokenUnitEmbedder(nn.Module):
    def __init__(self, vocabulary_size: int, embedding_size: int, dropout_rate: float,
                 subtoken_combination: str = "max", padding_idx: int = 0):
        super().__init__()
        if subtoken_combination != "max":
            raise NotImplementedError(
                "only subtoken_combination='max' (the gnn-mlp registry default, modelregistry.py:63-64) has a B200 kernel")
        self.__embeddings = nn.Embedding(vocabulary_size, embedding_size, padding_idx=padding_idx)
        self.__dropout_rate = float(dropout_rate)
        self.__subtoken_combination = subtoken_combination

    @property
    def embedding_layer(self) -> nn.Embedding:
        return self.__embeddings

    def forward(self, token_idxs: torch.Tensor, lengths: torch.Tensor) -> torch.Tensor:
        from buglab_b200 import ops

        return ops.subtoken_maxpool(self.__embeddings.weight, token_idxs, lengths, self.__dropout_rate, self.training)


class CharUnitEmbedder(nn.Module):
    def __init__(self, *args, **kwargs):
        super().__init__()
        raise NotImplementedError("character-level node embeddings are outside the gnn-mlp hot path")


class StrElementRepresentationModel(
    AbstractNeuralModel[str, np.ndarray, Union[TokenUnitEmbedder, SubtokenUnitEmbedder, CharUnitEmbedder]]
):
    def __init__(self, *, token_splitting: str, embedding_size: int = 128, dropout_rate: float = 0.2,
                 vocabulary_size: int = 10000, min_freq_threshold: int = 5, max_num_subtokens: Optional[int] = 5,
                 subtoken_combination: str = "sum", **unsupported):
        super().__init__()
        if token_splitting not in ("token", "subtoken"):
            raise NotImplementedError(f"token_splitting={token_splitting!r} is outside the gnn-mlp hot path")
        self.splitting_kind = token_splitting
        self.embedding_size = embedding_size
        self.dropout_rate = dropout_rate
        self.max_vocabulary_size = vocabulary_size
        self.min_freq_threshold = min_freq_threshold
        self.max_num_subtokens = max_num_subtokens if token_splitting == "subtoken" else 1
        self.subtoken_combination = subtoken_combination
        self.__tok_counter: Optional[Counter] = Counter()
        self.__vocabulary: Optional[Vocabulary] = None

    @property
    def vocabulary(self) -> Vocabulary:
        return self.__vocabulary

    def representation_size(self) -> int:
        return self.embedding_size

    # ---- metadata -------------------------------------------------------------------------------
    def update_metadata_from(self, datapoint: str) -> None:
        if self.splitting_kind == "token":
            self.__tok_counter[datapoint] += 1
        else:
            self.__tok_counter.update(split_identifier_into_parts(datapoint))

    def update_metadata_from_many(self, datapoints: List[str]) -> None:
        if self.splitting_kind == "token":
            self.__tok_counter.update(datapoints)
        else:
            for token, count in Counter(datapoints).items():
                for part in split_identifier_into_parts(token):
                    self.__tok_counter[part] += count

    def update_metadata_from_token_counts(self, counts: Dict[str, int]) -> None:
        """Adds already-split (sub)token counts — the result of ``update_metadata_from_many`` over many graphs, computed
        elsewhere (buglab_b200.shards.NativeMetadataPass)."""
        self.__tok_counter.update(counts)

    def finalize_metadata(self) -> None:
        self.__vocabulary = Vocabulary.create_vocabulary(
            self.__tok_counter, max_size=self.max_vocabulary_size, count_threshold=self.min_freq_threshold, add_pad=True)
        self.__tok_counter = None
        self.__id_cache: Dict[str, Tuple[int, ...]] = {}

    def build_neural_module(self):
        pad = self.__vocabulary.get_id_or_unk(Vocabulary.get_pad())
        if self.splitting_kind == "token":
            return TokenUnitEmbedder(len(self.__vocabulary), self.embedding_size, self.dropout_rate, padding_idx=pad)
        return SubtokenUnitEmbedder(len(self.__vocabulary), self.embedding_size, self.dropout_rate,
                                    self.subtoken_combination, padding_idx=pad)

    # ---- tensorisation --------------------------------------------------------------------------
    def _ids_of(self, token: str) -> Tuple[int, ...]:
        cache = self.__dict__.setdefault("_StrElementRepresentationModel__id_cache", {})
        ids = cache.get(token)
        if ids is None:
            vocab = self.__vocabulary
            if self.splitting_kind == "token":
                ids = (vocab.get_id_or_unk(token),)
            else:
                ids = tuple(vocab.get_id_or_unk(t) for t in split_identifier_into_parts(token)[: self.max_num_subtokens])
            if len(cache) < 2_000_000:
                cache[token] = ids
        return ids

    def tensorize(self, datapoint: str) -> np.ndarray:
        return np.array(self._ids_of(datapoint), dtype=np.int32)

    def tensorize_many(self, datapoints: List[str]) -> Tuple[np.ndarray, np.ndarray]:
        """Padded ``[n, T]`` id table (pad id 0) and ``[n]`` lengths for a whole graph's node labels."""
        T = self.max_num_subtokens
        n = len(datapoints)
        ids = np.zeros((n, T), dtype=np.int32)
        lens = np.empty(n, dtype=np.int32)
        for i, token in enumerate(datapoints):
            t = self._ids_of(token)
            lens[i] = len(t)
            ids[i, : len(t)] = t
        return ids, lens

    # per-element minibatch protocol (kept for API compatibility; the GNN model uses tensorize_many)
    def initialize_minibatch(self) -> Dict[str, Any]:
        return {"ids": [], "lens": []}

    def extend_minibatch_with(self, tensorized_datapoint: np.ndarray, partial_minibatch: Dict[str, Any]) -> bool:
        row = np.zeros(self.max_num_subtokens, dtype=np.int32)
        row[: len(tensorized_datapoint)] = tensorized_datapoint
        partial_minibatch["ids"].append(row)
        partial_minibatch["lens"].append(len(tensorized_datapoint))
        return True

    def finalize_minibatch(self, accumulated_minibatch_data: Dict[str, Any], device) -> Dict[str, Any]:
        ids = np.stack(accumulated_minibatch_data["ids"]) if accumulated_minibatch_data["ids"] else np.zeros(
            (0, self.max_num_subtokens), dtype=np.int32)
        lens = np.array(accumulated_minibatch_data["lens"], dtype=np.int32)
        if self.splitting_kind == "token":
            return {"token_idxs": torch.from_numpy(ids[:, 0].copy()).to(device)}
        return {"token_idxs": torch.from_numpy(ids).to(device), "lengths": torch.from_numpy(lens).to(device)}
