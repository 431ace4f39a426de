"""BatchNormalization / Dropout used by DNN(use_bn, dropout_rate) (deepctr/layers/core.py:177-179) and
(The reference's LayerNormalization, deepctr/layers/normalization.py:18-51, is Transformer-only: out of scope.)"""
from .. import ops
from ..engine import Layer, Zeros, Ones


class BatchNormalization(Layer):
    """tf.keras.layers.BatchNormalization defaults: momentum 0.99, epsilon 1e-3, gamma/beta trainable."""

    def __init__(self, axis=-1, momentum=0.99, epsilon=1e-3, center=True, scale=True, **kwargs):
        Layer.__init__(self, **kwargs)
        self.momentum, self.epsilon, self.center, self.scale = momentum, epsilon, center, scale

    def build(self, input_shape):
        n = int(input_shape[-1])
        self.gamma = self.add_weight("gamma", (n,), Ones()) if self.scale else None
        self.beta = self.add_weight("beta", (n,), Zeros()) if self.center else None
        self.moving_mean = self.add_weight("moving_mean", (n,), Zeros(), trainable=False)
        self.moving_variance = self.add_weight("moving_variance", (n,), Ones(), trainable=False)
        self.built = True

    def call(self, inputs, training=None, **kwargs):
        return ops.batchnorm(inputs, self.gamma, self.beta, self.moving_mean, self.moving_variance,
                             self.epsilon, bool(training), self.momentum)


class Dropout(Layer):
    def __init__(self, rate, seed=None, **kwargs):
        Layer.__init__(self, **kwargs)
        self.rate, self.seed = rate, seed
        self._calls = 0

    def call(self, inputs, training=None, **kwargs):
        if not training or not self.rate:
            return inputs
        self._calls += 1
        return ops.dropout(inputs, self.rate, (self.seed or 0) * 1000003 + self._calls)
