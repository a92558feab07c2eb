from .richpath import LocalPath, RichPath
from .debughelper import run_and_debug

__all__ = ["RichPath", "LocalPath", "run_and_debug"]
