"""Feature-column descriptors and the functions that turn them into model inputs, embedding windows and
the linear logit - the host-side mirror of the reference's ``deepctr.feature_column`` surface
(deepctr/feature_column.py:34-233): same public names, constructor arguments, defaults, return
structures and error behaviour, implemented against this package's ``engine`` / ``inputs``.

The three descriptors are plain slotted value objects built on one small base class (`_Column`) instead
of namedtuple subclasses; they keep what callers of the reference rely on: attribute access, ``_replace``,
field-wise equality, hashing by feature name, and a constructor-style repr.
"""
from collections import OrderedDict

from .engine import Input, RandomNormal, Zeros
from .inputs import (create_embedding_matrix, embedding_lookup, get_dense_input, varlen_embedding_lookup,
                     get_varlen_pooling_list, mergeDict)

DEFAULT_GROUP_NAME = "default_group"


class _Column(object):
    """Immutable record with named fields (``_fields``); hashed by ``name`` so that a column can key a dict
    or sit in a set the way the reference's columns do (deepctr/feature_column.py:56-57,105-106,128-129)."""
    __slots__ = ()
    _fields = ()

    def _init(self, values):
        for field, value in zip(self._fields, values):
            object.__setattr__(self, field, value)

    def __setattr__(self, key, value):
        raise AttributeError("%s is immutable; use _replace(%s=...)" % (type(self).__name__, key))

    def _replace(self, **changes):
        unknown = set(changes) - set(self._fields)
        if unknown:
            raise ValueError("Got unexpected field names: %r" % sorted(unknown))
        clone = object.__new__(type(self))
        clone._init([changes.get(f, getattr(self, f)) for f in self._fields])
        return clone

    def _asdict(self):
        return OrderedDict((f, getattr(self, f)) for f in self._fields)

    def __iter__(self):
        return iter([getattr(self, f) for f in self._fields])

    def __eq__(self, other):
        return type(other) is type(self) and list(self) == list(other)

    def __ne__(self, other):
        return not self == other

    def __hash__(self):
        return hash(self.name)

    def __repr__(self):
        return "%s(%s)" % (type(self).__name__, ", ".join("%s=%r" % (f, getattr(self, f)) for f in self._fields))

    def __reduce__(self):
        return (_rebuild, (type(self), list(self)))


def _rebuild(cls, values):
    obj = object.__new__(cls)
    obj._init(values)
    return obj


class SparseFeat(_Column):
    """Single-valued categorical feature (deepctr/feature_column.py:34-57).  ``embedding_dim="auto"`` resolves
    to 6 * floor(V ** 0.25); the table is shared by every column with the same ``embedding_name``."""
    _fields = ("name", "vocabulary_size", "embedding_dim", "use_hash", "vocabulary_path", "dtype",
               "embeddings_initializer", "embedding_name", "group_name", "trainable")
    __slots__ = _fields

    def __init__(self, name, vocabulary_size, embedding_dim=4, use_hash=False, vocabulary_path=None,
                 dtype="int32", embeddings_initializer=None, embedding_name=None,
                 group_name=DEFAULT_GROUP_NAME, trainable=True):
        dim = 6 * int(pow(vocabulary_size, 0.25)) if embedding_dim == "auto" else embedding_dim
        init = embeddings_initializer or RandomNormal(mean=0.0, stddev=0.0001, seed=2020)
        self._init([name, vocabulary_size, dim, use_hash, vocabulary_path, dtype, init,
                    embedding_name or name, group_name, trainable])


class VarLenSparseFeat(_Column):
    """Multi-valued / sequence feature wrapping a SparseFeat (deepctr/feature_column.py:60-109): ids are
    ``[B, maxlen]`` zero-padded; ``combiner`` in {sum, mean, max}; optional length and weight inputs."""
    _fields = ("sparsefeat", "maxlen", "combiner", "length_name", "weight_name", "weight_norm")
    __slots__ = _fields

    def __init__(self, sparsefeat, maxlen, combiner="mean", length_name=None, weight_name=None, weight_norm=True):
        self._init([sparsefeat, maxlen, combiner, length_name, weight_name, weight_norm])

    def __getattr__(self, item):
        # everything a SparseFeat has (name, vocabulary_size, embedding_dim, ...) is answered by the wrapped column
        if item in SparseFeat._fields:
            return getattr(self.sparsefeat, item)
        raise AttributeError(item)


class DenseFeat(_Column):
    """Real-valued feature of width ``dimension`` (deepctr/feature_column.py:112-129); ``transform_fn`` is applied
    as a Lambda layer on the way in."""
    _fields = ("name", "dimension", "dtype", "transform_fn")
    __slots__ = _fields

    def __init__(self, name, dimension=1, dtype="float32", transform_fn=None):
        self._init([name, dimension, dtype, transform_fn])


# --------------------------------------------------------------------------------------------------
_STRING_DTYPES = ("string", "str", "object", "<class 'str'>")


def _require_hash_for_strings(col):
    """String ids cannot index a table: the reference rejects them at input-construction time unless the
    column hashes them first (deepctr/feature_column.py:24-31; same message)."""
    is_string = col.dtype in (str, bytes) or str(col.dtype) in _STRING_DTYPES
    if is_string and not col.use_hash:
        raise ValueError(
            "SparseFeat(name='{}', dtype='string') requires use_hash=True "
            "so string ids can be converted before embedding lookup. "
            "Alternatively, encode the feature values to integer ids before "
            "passing them to DeepCTR.".format(col.name))


def _placeholders(col, prefix):
    """(input key, symbolic Input) pairs one column contributes, in the order the reference emits them."""
    if isinstance(col, DenseFeat):
        return [(col.name, Input(shape=(col.dimension,), name=prefix + col.name, dtype=col.dtype))]
    if isinstance(col, SparseFeat):
        _require_hash_for_strings(col)
        return [(col.name, Input(shape=(1,), name=prefix + col.name, dtype=col.dtype))]
    if isinstance(col, VarLenSparseFeat):
        _require_hash_for_strings(col)
        out = [(col.name, Input(shape=(col.maxlen,), name=prefix + col.name, dtype=col.dtype))]
        if col.weight_name is not None:
            out.append((col.weight_name, Input(shape=(col.maxlen, 1), name=prefix + col.weight_name, dtype="float32")))
        if col.length_name is not None:
            out.append((col.length_name, Input((1,), name=prefix + col.length_name, dtype="int32")))
        return out
    raise TypeError("Invalid feature column type,got", type(col))


def build_input_features(feature_columns, prefix=''):
    """Ordered name -> Input dict: THE input ordering contract of every builder
    (deepctr/feature_column.py:145-168).  Columns repeated across the linear and DNN lists collapse by name."""
    inputs = OrderedDict()
    for col in feature_columns:
        inputs.update(_placeholders(col, prefix))
    return inputs


def get_feature_names(feature_columns):
    return list(build_input_features(feature_columns))


# --------------------------------------------------------------------------------------------------
def _linear_twin(col):
    """The dim-1, zero-initialised copy of a categorical column that carries its first-order weight."""
    if isinstance(col, SparseFeat):
        return col._replace(embedding_dim=1, embeddings_initializer=Zeros())
    if isinstance(col, VarLenSparseFeat):
        return col._replace(sparsefeat=_linear_twin(col.sparsefeat))
    return col


def get_linear_logit(features, feature_columns, units=1, use_bias=False, seed=1024, prefix='linear',
                     l2_reg=0, sparse_feat_refine_weight=None):
    """First-order term (deepctr/feature_column.py:171-210): every categorical column gets a dim-1 table
    (``units`` independent sets, prefixes linear0, linear1, ...), their values are summed and the dense
    features go through a [d, 1] kernel (Linear modes 0 / 1 / 2).  The reference's additional lookup pass whose
    embeddings are discarded (:185) is not issued - only its dense list is used and that does not depend on it."""
    from .layers.utils import Linear, RefineWeight, ZeroLogit, concat_func
    twins = [_linear_twin(col) for col in feature_columns]
    dense_inputs = get_dense_input(features, twins)
    logits = []
    for unit in range(units):
        sparse_embs = input_from_feature_columns(features, twins, l2_reg, seed, prefix=prefix + str(unit))[0]
        if not sparse_embs and not dense_inputs:
            return ZeroLogit()(list(features.values())[0])           # empty feature_columns -> constant 0
        sparse = None
        if sparse_embs:
            sparse = concat_func(sparse_embs)
            if sparse_feat_refine_weight is not None:
                sparse = RefineWeight()([sparse, sparse_feat_refine_weight])
        dense = concat_func(dense_inputs) if dense_inputs else None
        if sparse is not None and dense is not None:
            mode, args = 2, [sparse, dense]
        elif sparse is not None:
            mode, args = 0, sparse
        else:
            mode, args = 1, dense
        logits.append(Linear(l2_reg, mode=mode, use_bias=use_bias, seed=seed)(args))
    return concat_func(logits)


def input_from_feature_columns(features, feature_columns, l2_reg, seed, prefix='', seq_mask_zero=True,
                               support_dense=True, support_group=False):
    """Embeddings + dense values of a column list (deepctr/feature_column.py:213-233).  Returns
    (group-name -> [B,1,E] tensors dict, or their flat list when ``support_group`` is False; dense tensor list).
    At run time all of these are column windows of one buffer written by the fused gather kernel."""
    columns = list(feature_columns) if feature_columns else []
    single = [c for c in columns if isinstance(c, SparseFeat)]
    multi = [c for c in columns if isinstance(c, VarLenSparseFeat)]
    tables = create_embedding_matrix(feature_columns, l2_reg, seed, prefix=prefix, seq_mask_zero=seq_mask_zero)
    by_group = embedding_lookup(tables, features, single)
    dense_values = get_dense_input(features, feature_columns)
    if dense_values and not support_dense:
        raise ValueError("DenseFeat is not supported in dnn_feature_columns")
    sequences = varlen_embedding_lookup(tables, features, multi)
    pooled_by_group = get_varlen_pooling_list(sequences, features, multi)
    merged = mergeDict(by_group, pooled_by_group)
    if support_group:
        return merged, dense_values
    return [t for group in merged.values() for t in group], dense_values
