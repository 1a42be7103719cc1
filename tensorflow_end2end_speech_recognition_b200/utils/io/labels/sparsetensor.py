"""Dense-padded labels <-> (indices, values, dense_shape) triples.

Same wire format and function names as the reference's
``utils/io/labels/sparsetensor.py:12-78`` (the feed format of the
``tf.SparseTensor`` label placeholders, ``models/ctc/ctc.py:246-249``), without
the TensorFlow import.  Decode results are returned as ``SparseTensorValue``.
"""
import collections

import numpy as np

SparseTensorValue = collections.namedtuple("SparseTensorValue", ["indices", "values", "dense_shape"])


def list2sparsetensor(labels, padded_value):
    """labels [B, max_label_len] (padded with ``padded_value``) -> [indices, values, dense_shape]
    (reference: sparsetensor.py:12-39)."""
    dtype_values = np.uint8 if padded_value is None else np.int32
    indices, values = [], []
    for i_utt, each_label in enumerate(labels):
        for i_l, l in enumerate(each_label):
            if l == padded_value:
                break
            indices.append([i_utt, i_l])
            values.append(l)
    dense_shape = [len(labels), np.asarray(indices).max(0)[1] + 1]
    return [np.array(indices, dtype=np.int64), np.array(values, dtype=dtype_values),
            np.array(dense_shape, dtype=np.int64)]


def sparsetensor2list(labels_st, batch_size):
    """(reference: sparsetensor.py:42-78; same boundary rule: a new utterance starts
    where the column index returns to 0)."""
    if isinstance(labels_st, SparseTensorValue):
        indices, values = labels_st.indices, labels_st.values
    else:
        indices, values = labels_st[0], labels_st[1]
    if batch_size == 1:
        return values.reshape((1, -1))
    labels = []
    batch_boundary = np.where(indices[:, 1] == 0)[0]
    for i in range(batch_size - 1):
        labels.append(values[batch_boundary[i]:batch_boundary[i + 1]])
    labels.append(values[batch_boundary[-1]:])
    return labels


def sparse_to_label_lists(labels_st, batch_size):
    """Robust variant used internally: rows may be empty (no boundary heuristic)."""
    if isinstance(labels_st, SparseTensorValue):
        indices, values = labels_st.indices, labels_st.values
    else:
        indices, values = labels_st[0], labels_st[1]
    out = [[] for _ in range(batch_size)]
    indices = np.asarray(indices).reshape(-1, 2)
    for (b, _), v in zip(indices, np.asarray(values)):
        out[int(b)].append(int(v))
    return out
