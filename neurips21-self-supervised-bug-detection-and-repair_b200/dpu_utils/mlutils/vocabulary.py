"""Token vocabulary with %PAD% / %UNK% specials (dpu_utils.mlutils.Vocabulary behaviour, restated).

Used by buglab/models/basemodel.py:66-70 (rewrite-operator ids), buglab/representations/data.py:158-161
(``Vocabulary.get_pad()``) and the ptgnn string embedders.  One deliberate difference: ties between
equally frequent tokens are broken by the token string instead of by insertion order, so the ids do not
depend on PYTHONHASHSEED / set iteration order — every data-parallel rank must build the same table.
"""
from collections import Counter
from typing import Dict, Iterable, List, Optional, Union


class Vocabulary:
    def __init__(self, add_unk: bool = True, add_pad: bool = False) -> None:
        self.token_to_id: Dict[str, int] = {}
        self.id_to_token: List[str] = []
        if add_pad:
            self.add_or_get_id(self.get_pad())
        if add_unk:
            self.add_or_get_id(self.get_unk())

    @staticmethod
    def get_unk() -> str:
        return "%UNK%"

    @staticmethod
    def get_pad() -> str:
        return "%PAD%"

    def add_or_get_id(self, token: str) -> int:
        idx = self.token_to_id.get(token)
        if idx is None:
            idx = len(self.id_to_token)
            self.token_to_id[token] = idx
            self.id_to_token.append(token)
        return idx

    def is_unk(self, token: str) -> bool:
        return token not in self.token_to_id

    def get_id_or_unk(self, token: str) -> int:
        idx = self.token_to_id.get(token)
        if idx is not None:
            return idx
        return self.token_to_id[self.get_unk()]  # KeyError if the vocabulary was built without %UNK%

    def get_id_or_unk_multiple(self, tokens: List[str], pad_to_size: Optional[int] = None,
                               padding_element: int = 0) -> List[int]:
        if pad_to_size is not None:
            tokens = tokens[:pad_to_size]
        ids = [self.get_id_or_unk(t) for t in tokens]
        if pad_to_size is not None and len(ids) < pad_to_size:
            ids += [padding_element] * (pad_to_size - len(ids))
        return ids

    def get_name_for_id(self, token_id: int) -> str:
        return self.id_to_token[token_id]

    def __len__(self) -> int:
        return len(self.token_to_id)

    def __contains__(self, token: str) -> bool:
        return token in self.token_to_id

    def update(self, token_counter: Counter, max_size: int, count_threshold: int = 5) -> None:
        ranked = sorted(token_counter.items(), key=lambda kv: (-kv[1], kv[0]))
        for token, count in ranked:
            if len(self) >= max_size:
                break
            if count >= count_threshold:
                self.add_or_get_id(token)

    @staticmethod
    def create_vocabulary(tokens: Union[Iterable[str], Counter], max_size: int, count_threshold: int = 5,
                          add_unk: bool = True, add_pad: bool = False) -> "Vocabulary":
        counter = tokens if isinstance(tokens, Counter) else Counter(tokens)
        vocab = Vocabulary(add_unk=add_unk, add_pad=add_pad)
        vocab.update(counter, max_size, count_threshold)
        return vocab
