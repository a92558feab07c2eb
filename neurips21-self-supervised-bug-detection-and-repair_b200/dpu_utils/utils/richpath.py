"""Local-filesystem ``RichPath`` (the subset used by buglab/utils/msgpackutils.py:24-45 and train.py:74-90)."""
import fnmatch
import gzip
import json
import os
import pickle
from typing import Any, Iterable, Optional


class RichPath:
    def __init__(self, path: str):
        self.path = path

    @staticmethod
    def create(path: str, azure_info_path: Optional[str] = None) -> "RichPath":
        if str(path).startswith("azure://"):
            raise NotImplementedError("Azure storage paths are outside the scope of this build")
        return LocalPath(str(path))

    def __repr__(self) -> str:
        return f"{type(self).__name__}({self.path!r})"

    def __lt__(self, other: "RichPath") -> bool:
        return self.path < other.path

    def __eq__(self, other) -> bool:
        return isinstance(other, RichPath) and self.path == other.path

    def __hash__(self) -> int:
        return hash(self.path)


class LocalPath(RichPath):
    def is_dir(self) -> bool:
        return os.path.isdir(self.path)

    def is_file(self) -> bool:
        return os.path.isfile(self.path)

    def exists(self) -> bool:
        return os.path.exists(self.path)

    def make_as_dir(self) -> None:
        os.makedirs(self.path, exist_ok=True)

    def join(self, filename: str) -> "LocalPath":
        return LocalPath(os.path.join(self.path, filename))

    def basename(self) -> str:
        return os.path.basename(self.path)

    def to_local_path(self) -> "LocalPath":
        return self

    def get_size(self) -> int:
        return os.stat(self.path).st_size

    def iterate_filtered_files_in_dir(self, file_pattern: str) -> Iterable["LocalPath"]:
        if os.path.isfile(self.path):
            if fnmatch.fnmatch(os.path.basename(self.path), file_pattern):
                yield self
            return
        for root, _dirs, files in os.walk(self.path):
            for name in files:
                if fnmatch.fnmatch(name, file_pattern):
                    yield LocalPath(os.path.join(root, name))

    def get_filtered_files_in_dir(self, file_pattern: str):
        return list(self.iterate_filtered_files_in_dir(file_pattern))

    def read_as_text(self) -> str:
        with open(self.path, "r", encoding="utf-8") as f:
            return f.read()

    def read_as_json(self) -> Any:
        return json.loads(self.read_as_text())

    def read_as_pickle(self) -> Any:
        opener = gzip.open if self.path.endswith(".gz") else open
        with opener(self.path, "rb") as f:
            return pickle.load(f)

    def save_as_compressed_file(self, data: Any) -> None:
        if self.path.endswith(".json.gz"):
            with gzip.open(self.path, "wt", encoding="utf-8") as f:
                json.dump(data, f)
        elif self.path.endswith(".pkl.gz"):
            with gzip.open(self.path, "wb") as f:
                pickle.dump(data, f)
        else:
            raise ValueError(f"unsupported suffix: {self.path}")
