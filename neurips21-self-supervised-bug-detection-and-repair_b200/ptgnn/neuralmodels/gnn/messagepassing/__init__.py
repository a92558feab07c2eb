from .abstractmessagepassing import AbstractMessagePassingLayer
from .gatedmessagepassing import GatedMessagePassingLayer
from .mlpmessagepassing import MlpMessagePassingLayer

__all__ = ["AbstractMessagePassingLayer", "GatedMessagePassingLayer", "MlpMessagePassingLayer"]
