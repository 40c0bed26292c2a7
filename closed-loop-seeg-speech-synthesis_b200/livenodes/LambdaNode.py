"""Applies a function to every frame (reference: livenodes/LambdaNode.py)."""
from . import Node


class LambdaNode(Node.Node):
    def __init__(self, feature_function, name='LambdaNode'):
        super().__init__(name=name)
        self.feature_function = feature_function

    def add_data(self, data_frame, data_id=0):
        self.output_data(self.feature_function(data_frame))
