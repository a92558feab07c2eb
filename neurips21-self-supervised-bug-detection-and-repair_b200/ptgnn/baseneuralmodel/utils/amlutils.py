"""Logging helpers with the signatures the reference calls (train.py:64,111-116); Azure ML runs are out of scope."""
import logging
import os
import tempfile
from typing import Any, Dict, Optional


def configure_logging(aml_ctx: Optional[Any] = None) -> str:
    log_path = os.path.join(tempfile.gettempdir(), f"buglab_b200_{os.getpid()}.log")
    root = logging.getLogger()
    if not root.handlers:
        logging.basicConfig(
            level=logging.INFO,
            format="%(asctime)s [%(name)-35.35s @ %(lineno)-4d] [%(levelname)-5.5s] %(message)s",
            handlers=[logging.FileHandler(log_path), logging.StreamHandler()],
        )
    return log_path


def log_run(aml_ctx: Optional[Any], fold_name: str, model: Any, epoch_idx: int, metrics: Dict[str, Any]) -> None:
    if aml_ctx is None:
        return
    for name, value in metrics.items():
        try:
            aml_ctx.log(f"{fold_name}-{name}", float(value))
        except (TypeError, ValueError):
            pass
