"""Host mirror of deepctr/layers/sequence.py for the hot path: SequencePoolingLayer (:41-120),
WeightedSequenceLayer (:123-197), AttentionSequencePoolingLayer (:200-315).

When these layers hang directly off an embedding lookup of a model input (the shape
``get_varlen_pooling_list`` produces) the EmbeddingPlanner folds them into the fused gather and their
``call`` never runs; the standalone kernels below serve every other use."""
from .. import engine as E
from .. import ops
from ..engine import Layer
from .core import LocalActivationUnit


class SequencePoolingLayer(Layer):
    """sum / mean / max over the valid positions of [B,T,E] -> [B,1,E]."""

    def __init__(self, mode='mean', supports_masking=False, **kwargs):
        if mode not in ['sum', 'mean', 'max']:
            raise ValueError("mode must be sum or mean")
        self.mode = mode
        self.eps = 1e-8
        Layer.__init__(self, **kwargs)
        self.supports_masking = supports_masking

    def build(self, input_shape):
        if not self.supports_masking:
            self.seq_len_max = int(input_shape[0][1])
        self.built = True

    def call(self, seq_value_len_list, mask=None, **kwargs):
        if self.supports_masking:
            if mask is None:
                raise ValueError("When supports_masking=True,input must support masking")
            seq = seq_value_len_list
            m = mask.materialize()
            return ops.seqpool(seq, self.mode, mask_u8=m)
        seq, lengths = seq_value_len_list
        return ops.seqpool(seq, self.mode, lengths=lengths)

    def compute_output_shape(self, input_shape):
        if self.supports_masking:
            return (None, 1, input_shape[-1])
        return (None, 1, input_shape[0][-1])

    def compute_mask(self, inputs, mask=None):
        return None

    def get_config(self):
        config = {'mode': self.mode, 'supports_masking': self.supports_masking}
        base = Layer.get_config(self)
        return dict(list(base.items()) + list(config.items()))


class WeightedSequenceLayer(Layer):
    """[B,T,E] * (masked, optionally soft-maxed) per-position weights [B,T,1] -> [B,T,E]."""

    def __init__(self, weight_normalization=True, supports_masking=False, **kwargs):
        Layer.__init__(self, **kwargs)
        self.weight_normalization = weight_normalization
        self.supports_masking = supports_masking

    def build(self, input_shape):
        if not self.supports_masking:
            self.seq_len_max = int(input_shape[0][1])
        self.built = True

    def call(self, input_list, mask=None, **kwargs):
        if self.supports_masking:
            if mask is None:
                raise ValueError("When supports_masking=True,input must support masking")
            key_input, value_input = input_list
            m = mask[0].materialize()
            return ops.weighted_seq(key_input, value_input, self.weight_normalization, mask_u8=m)
        key_input, key_length_input, value_input = input_list
        return ops.weighted_seq(key_input, value_input, self.weight_normalization, lengths=key_length_input)

    def compute_output_shape(self, input_shape):
        return input_shape[0]

    def compute_mask(self, inputs, mask=None):
        if self.supports_masking:
            return mask[0] if mask is not None else None
        return None

    def get_config(self):
        config = {'weight_normalization': self.weight_normalization, 'supports_masking': self.supports_masking}
        base = Layer.get_config(self)
        return dict(list(base.items()) + list(config.items()))


class AttentionSequencePoolingLayer(Layer):
    """DIN attentional pooling, deepctr/layers/sequence.py:200-315.
    score = LocalActivationUnit([q, keys]); masked fill (0, or -2^32+1 then softmax when
    weight_normalization); out = score @ keys -> [B,1,E] (or the scores when return_score)."""

    def __init__(self, att_hidden_units=(80, 40), att_activation='sigmoid', weight_normalization=False,
                 return_score=False, supports_masking=False, **kwargs):
        self.att_hidden_units = att_hidden_units
        self.att_activation = att_activation
        self.weight_normalization = weight_normalization
        self.return_score = return_score
        Layer.__init__(self, **kwargs)
        self.supports_masking = supports_masking

    def build(self, input_shape):
        if not self.supports_masking:
            if not isinstance(input_shape, list) or len(input_shape) != 3:
                raise ValueError('A `AttentionSequencePoolingLayer` layer should be called '
                                 'on a list of 3 inputs')
            if len(input_shape[0]) != 3 or len(input_shape[1]) != 3 or len(input_shape[2]) != 2:
                raise ValueError(
                    "Unexpected inputs dimensions,the 3 tensor dimensions are %d,%d and %d , expect to be 3,3 and 2" % (
                        len(input_shape[0]), len(input_shape[1]), len(input_shape[2])))
            if input_shape[0][-1] != input_shape[1][-1] or input_shape[0][1] != 1 or input_shape[2][1] != 1:
                raise ValueError('A `AttentionSequencePoolingLayer` layer requires '
                                 'inputs of a 3 tensor with shape (None,1,embedding_size),(None,T,embedding_size) and (None,1)'
                                 'Got different shapes: %s' % (input_shape))
        self.local_att = self._track(LocalActivationUnit(
            self.att_hidden_units, self.att_activation, l2_reg=0, dropout_rate=0, use_bn=False, seed=1024,
            name=self.name + "/local_activation_unit"))
        self.local_att._maybe_build([tuple(input_shape[0]), tuple(input_shape[1])])
        self.built = True

    def call(self, inputs, mask=None, training=None, **kwargs):
        if self.supports_masking:
            if mask is None:
                raise ValueError("When supports_masking=True,input must support masking")
            queries, keys = inputs
            km = mask[-1]
            if km is None:
                raise ValueError("When supports_masking=True,input must support masking")
            key_mask = km.materialize()
        else:
            queries, keys, keys_length = inputs
            key_mask = E.KMask(lengths=keys_length.data.reshape(-1), maxlen=keys.data.shape[1]).materialize()
        from .. import kernels as K
        with K.profile_tag("din_att"):
            score = self.local_att.call([queries, keys], training=training)      # [B,T,1]
            return ops.din_attention_pool(score, keys, key_mask, self.weight_normalization, self.return_score)

    def compute_output_shape(self, input_shape):
        if self.return_score:
            return (None, 1, input_shape[1][1])
        return (None, 1, input_shape[0][-1])

    def compute_mask(self, inputs, mask=None):
        return None

    def get_config(self):
        config = {'att_hidden_units': self.att_hidden_units, 'att_activation': self.att_activation,
                  'weight_normalization': self.weight_normalization, 'return_score': self.return_score,
                  'supports_masking': self.supports_masking}
        base = Layer.get_config(self)
        return dict(list(base.items()) + list(config.items()))
