import pdb
import sys
import traceback
from typing import Callable


def run_and_debug(func: Callable[[], None], enable_debugging: bool) -> None:
    """Run ``func``; on an exception print the trace and (if asked) drop into the post-mortem debugger."""
    try:
        func()
    except Exception:
        if enable_debugging:
            _, value, tb = sys.exc_info()
            traceback.print_exc()
            pdb.post_mortem(tb)
        else:
            raise
