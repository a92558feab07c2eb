from .abstractneuralmodel import AbstractNeuralModel
from .modulewithmetrics import ModuleWithMetrics
from .trainer import AbstractScheduler, ModelTrainer

__all__ = ["AbstractNeuralModel", "ModuleWithMetrics", "ModelTrainer", "AbstractScheduler"]
