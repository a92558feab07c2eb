"""``.msgpack.l.gz`` shards: a gzip stream of concatenated msgpack objects (wire format of the reference's
buglab/utils/msgpackutils.py:11-45; byte-compatible in both directions)."""
import gzip
import random
from collections import OrderedDict
from os import PathLike
from typing import Any, Iterable, Iterator, Optional

import msgpack
from dpu_utils.utils import RichPath


def load_msgpack_l_gz(filename: PathLike) -> Iterator[Any]:
    with gzip.open(filename, "rb") as stream:
        yield from msgpack.Unpacker(stream, raw=False, object_pairs_hook=OrderedDict, max_buffer_size=0)


def save_msgpack_l_gz(data: Iterable[Any], filename: PathLike) -> None:
    packer = msgpack.Packer(use_bin_type=True)
    with gzip.GzipFile(filename, "wb") as stream:
        for element in data:
            stream.write(packer.pack(element))


def select_shard_files(path: RichPath, shuffle: bool = False, take_only_first_n_files: Optional[int] = None,
                       rank: int = 0, world_size: int = 1):
    """The ``*.msgpack.l.gz`` files one rank reads: the sorted list (optionally truncated), split round-robin across
    ranks when there are enough files — otherwise every rank reads every file and keeps each ``world_size``-th element
    (second return value) — then optionally shuffled."""
    files = sorted(path.iterate_filtered_files_in_dir("*.msgpack.l.gz"))
    if take_only_first_n_files is not None:
        files = files[:take_only_first_n_files]
    if world_size > 1 and len(files) >= world_size:
        files = files[rank::world_size]
        shard_elements = False
    else:
        shard_elements = world_size > 1
    if shuffle:
        random.shuffle(files)
    return files, shard_elements


def load_all_msgpack_l_gz(path: RichPath, shuffle: bool = False, take_only_first_n_files: Optional[int] = None,
                          limit_num_yielded_elements: Optional[int] = None, rank: int = 0, world_size: int = 1) -> Iterator:
    """All non-None elements of every ``*.msgpack.l.gz`` under ``path`` (sorted, optionally shuffled file order).
    ``rank`` / ``world_size`` shard the (sorted) file list round-robin for data-parallel training."""
    files, shard_elements = select_shard_files(path, shuffle, take_only_first_n_files, rank, world_size)
    num_yielded = 0
    for file in files:
        try:
            for i, element in enumerate(load_msgpack_l_gz(file.to_local_path().path)):
                if element is None or (shard_elements and i % world_size != rank):
                    continue
                num_yielded += 1
                yield element
                if limit_num_yielded_elements is not None and num_yielded > limit_num_yielded_elements:
                    return
        except Exception as e:  # a corrupt shard is skipped, as in the reference (msgpackutils.py:44-45)
            print(f"Error loading {file}: {e}.")
