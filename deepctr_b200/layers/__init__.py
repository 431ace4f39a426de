"""deepctr.layers surface for the hot path (deepctr/layers/__init__.py:3-54 lists the reference's)."""
from ..engine import Layer, Dense, Flatten, Concatenate, Add, Lambda, Input
from .activation import Dice, Activation, activation_layer
from .core import DNN, LocalActivationUnit, PredictionLayer
from .interaction import FM, CrossNet, CIN, InteractingLayer
from .normalization import BatchNormalization, Dropout
from .sequence import SequencePoolingLayer, WeightedSequenceLayer, AttentionSequencePoolingLayer
from .utils import (NoMask, Hash, Linear, Concat, _Add, concat_func, add_func, combined_dnn_input,
                    reduce_sum, reduce_mean, reduce_max, div, softmax)

custom_objects = {
    'DNN': DNN, 'PredictionLayer': PredictionLayer, 'LocalActivationUnit': LocalActivationUnit,
    'FM': FM, 'CrossNet': CrossNet, 'CIN': CIN, 'InteractingLayer': InteractingLayer,
    'SequencePoolingLayer': SequencePoolingLayer, 'WeightedSequenceLayer': WeightedSequenceLayer,
    'AttentionSequencePoolingLayer': AttentionSequencePoolingLayer, 'Dice': Dice, 'Hash': Hash,
    'Linear': Linear, 'Concat': Concat, 'NoMask': NoMask, '_Add': _Add,
}
