from .structs import GnnOutput, GraphData, TensorizedGraphData
from .graphneuralnetwork import GraphNeuralNetwork, GraphNeuralNetworkModel

__all__ = ["GnnOutput", "GraphData", "GraphNeuralNetwork", "GraphNeuralNetworkModel", "TensorizedGraphData"]
