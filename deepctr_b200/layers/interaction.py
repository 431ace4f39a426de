"""Host mirror of deepctr/layers/interaction.py for the operators on the hot path:
FM (:563-607), CrossNet (:344-435), CIN (:209-341), InteractingLayer (:697-790).
Same constructor arguments, weight names, shape checks and error messages as the reference."""
from .. import engine as E
from .. import ops
from ..engine import Layer, Zeros, glorot_normal, glorot_uniform, TruncatedNormal, l2


class FM(Layer):
    """Factorization Machine second-order term, deepctr/layers/interaction.py:563-607.
    [B,F,E] -> [B,1].  When the input is the fused gather's buffer the value comes out of the gather
    kernel's epilogue (no extra pass over [B,F,E])."""

    def build(self, input_shape):
        if len(input_shape) != 3:
            raise ValueError("Unexpected inputs dimensions % d,\
                             expect to be 3 dimensions" % (len(input_shape)))
        self.built = True

    def call(self, inputs, **kwargs):
        if inputs.data.dim() != 3:
            raise ValueError("Unexpected inputs dimensions %d, expect to be 3 dimensions" % (inputs.data.dim()))
        planner = getattr(self, "_planner", None)
        if planner is not None:
            fused = planner.lookup_fm(inputs)
            if fused is not None:
                return fused
        return ops.fm(inputs)

    def compute_output_shape(self, input_shape):
        return (None, 1)


class CrossNet(Layer):
    """deepctr/layers/interaction.py:344-435.  vector: x_{l+1} = x_0 (x_l . w_l) + b_l + x_l;
    matrix: x_{l+1} = x_0 * (W_l x_l + b_l) + x_l."""

    def __init__(self, layer_num=2, parameterization='vector', l2_reg=0, seed=1024, **kwargs):
        self.layer_num = layer_num
        self.parameterization = parameterization
        self.l2_reg = l2_reg
        self.seed = seed
        print('CrossNet parameterization:', self.parameterization)
        Layer.__init__(self, **kwargs)

    def build(self, input_shape):
        if len(input_shape) != 2:
            raise ValueError("Unexpected inputs dimensions %d, expect to be 2 dimensions" % (len(input_shape),))
        dim = int(input_shape[-1])
        if self.parameterization == 'vector':
            shape = (dim, 1)
        elif self.parameterization == 'matrix':
            shape = (dim, dim)
        else:
            raise ValueError("parameterization should be 'vector' or 'matrix'")
        self.kernels = [self.add_weight(name='kernel' + str(i), shape=shape,
                                        initializer=glorot_normal(seed=self.seed),
                                        regularizer=l2(self.l2_reg), trainable=True)
                        for i in range(self.layer_num)]
        self.bias = [self.add_weight(name='bias' + str(i), shape=(dim, 1), initializer=Zeros(), trainable=True)
                     for i in range(self.layer_num)]
        self.built = True

    def call(self, inputs, **kwargs):
        if inputs.data.dim() != 2:
            raise ValueError("Unexpected inputs dimensions %d, expect to be 2 dimensions" % (inputs.data.dim()))
        x_0 = inputs
        x_l = x_0
        for i in range(self.layer_num):
            if self.parameterization == 'vector':
                x_l = ops.cross_vector(x_0, x_l, self.kernels[i], self.bias[i])
            else:
                x_l = ops.cross_matrix(x_0, x_l, self.kernels[i], self.bias[i])
        return x_l

    def get_config(self):
        config = {'layer_num': self.layer_num, 'parameterization': self.parameterization,
                  'l2_reg': self.l2_reg, 'seed': self.seed}
        base = Layer.get_config(self)
        base.update(config)
        return base

    def compute_output_shape(self, input_shape):
        return input_shape


class CIN(Layer):
    """Compressed Interaction Network, deepctr/layers/interaction.py:209-341.
    out[b,d,n] = act(sum_{i,j} X0[b,i,d] * Xk[b,j,d] * W_k[i*H_k + j, n] + bias_k[n]); the
    [B, D, m*H_k] outer product of the reference (:291-297) is never materialised."""

    def __init__(self, layer_size=(128, 128), activation='relu', split_half=True, l2_reg=1e-5, seed=1024,
                 **kwargs):
        if len(layer_size) == 0:
            raise ValueError("layer_size must be a list(tuple) of length greater than 1")
        self.layer_size = layer_size
        self.split_half = split_half
        self.activation = activation
        self.l2_reg = l2_reg
        self.seed = seed
        Layer.__init__(self, **kwargs)

    def build(self, input_shape):
        if len(input_shape) != 3:
            raise ValueError("Unexpected inputs dimensions %d, expect to be 3 dimensions" % (len(input_shape)))
        self.field_nums = [int(input_shape[1])]
        self.filters = []
        self.bias = []
        for i, size in enumerate(self.layer_size):
            self.filters.append(self.add_weight(name='filter' + str(i),
                                                shape=[1, self.field_nums[-1] * self.field_nums[0], size],
                                                initializer=glorot_uniform(seed=self.seed + i),
                                                regularizer=l2(self.l2_reg)))
            self.bias.append(self.add_weight(name='bias' + str(i), shape=[size], initializer=Zeros()))
            if self.split_half:
                if i != len(self.layer_size) - 1 and size % 2 > 0:
                    raise ValueError(
                        "layer_size must be even number except for the last layer when split_half=True")
                self.field_nums.append(size // 2)
            else:
                self.field_nums.append(size)
        from .activation import fusable_activation
        if not fusable_activation(self.activation):
            raise ValueError("CIN activation %r is not supported" % (self.activation,))
        self.built = True

    def call(self, inputs, **kwargs):
        if inputs.data.dim() != 3:
            raise ValueError("Unexpected inputs dimensions %d, expect to be 3 dimensions" % (inputs.data.dim()))
        return ops.cin(inputs, self.filters, self.bias, self.layer_size, self.activation, self.split_half)

    def compute_output_shape(self, input_shape):
        if self.split_half:
            featuremap_num = sum(self.layer_size[:-1]) // 2 + self.layer_size[-1]
        else:
            featuremap_num = sum(self.layer_size)
        return (None, featuremap_num)

    def get_config(self):
        config = {'layer_size': self.layer_size, 'split_half': self.split_half, 'activation': self.activation,
                  'seed': self.seed}
        base = Layer.get_config(self)
        base.update(config)
        return base


class InteractingLayer(Layer):
    """AutoInt multi-head self-attention over fields, deepctr/layers/interaction.py:697-790."""

    def __init__(self, att_embedding_size=8, head_num=2, use_res=True, scaling=False, seed=1024, **kwargs):
        if head_num <= 0:
            raise ValueError('head_num must be a int > 0')
        self.att_embedding_size = att_embedding_size
        self.head_num = head_num
        self.use_res = use_res
        self.seed = seed
        self.scaling = scaling
        Layer.__init__(self, **kwargs)

    def build(self, input_shape):
        if len(input_shape) != 3:
            raise ValueError("Unexpected inputs dimensions %d, expect to be 3 dimensions" % (len(input_shape)))
        embedding_size = int(input_shape[-1])
        n = self.att_embedding_size * self.head_num
        self.W_Query = self.add_weight(name='query', shape=[embedding_size, n],
                                       initializer=TruncatedNormal(seed=self.seed))
        self.W_key = self.add_weight(name='key', shape=[embedding_size, n],
                                     initializer=TruncatedNormal(seed=self.seed + 1))
        self.W_Value = self.add_weight(name='value', shape=[embedding_size, n],
                                       initializer=TruncatedNormal(seed=self.seed + 2))
        if self.use_res:
            self.W_Res = self.add_weight(name='res', shape=[embedding_size, n],
                                         initializer=TruncatedNormal(seed=self.seed))
        self.built = True

    def call(self, inputs, **kwargs):
        if inputs.data.dim() != 3:
            raise ValueError("Unexpected inputs dimensions %d, expect to be 3 dimensions" % (inputs.data.dim()))
        querys = ops.dense(inputs, self.W_Query)
        keys = ops.dense(inputs, self.W_key)
        values = ops.dense(inputs, self.W_Value)
        res = ops.dense(inputs, self.W_Res) if self.use_res else None
        return ops.interacting_attention(querys, keys, values, res, self.head_num, self.att_embedding_size,
                                         self.scaling)

    def compute_output_shape(self, input_shape):
        return (None, input_shape[1], self.att_embedding_size * self.head_num)

    def get_config(self):
        config = {'att_embedding_size': self.att_embedding_size, 'head_num': self.head_num,
                  'use_res': self.use_res, 'seed': self.seed}
        base = Layer.get_config(self)
        base.update(config)
        return base
