from typing import Any, Dict

from torch import nn


class ModuleWithMetrics(nn.Module):
    """``nn.Module`` with running metrics (protocol used at buglab/models/gnn.py:95-114,
    layers/localizationmodule.py:30-52, layers/fixermodules.py:19-29)."""

    def __init__(self):
        super().__init__()
        self._reset_module_metrics()

    def _reset_module_metrics(self) -> None:
        pass

    def _module_metrics(self) -> Dict[str, Any]:
        return {}

    def reset_metrics(self) -> None:
        for module in self.modules():
            if isinstance(module, ModuleWithMetrics):
                module._reset_module_metrics()

    def report_metrics(self) -> Dict[str, Any]:
        metrics: Dict[str, Any] = {}
        for module in self.modules():
            if isinstance(module, ModuleWithMetrics):
                for key, value in module._module_metrics().items():
                    if key in metrics:
                        raise ValueError(f"metric `{key}` is reported by more than one module")
                    metrics[key] = value
        return metrics
