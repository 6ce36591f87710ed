"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY.

A minimal stand-in for the slice of the TensorFlow 1.2 API that the reference's model code
(/root/reference/models/ops.py, models/tacotron.py) calls, written over torch CPU tensors, so that the
REFERENCE'S OWN SOURCE FILES can be imported and executed unmodified in this container
(tests/golden/make_golden.py).  What this pins and what it does not:

  * pinned: the graph WIRING of the reference -- call order, scopes, reshapes, the slice fed to the pre-net,
    wrapper nesting (OutputProjection(InputProjection(Residual(MultiRNN)))), helper choice, what is
    concatenated with what -- because that comes from executing models/tacotron.py / models/ops.py themselves;
  * not pinned: TensorFlow's kernel numerics and the semantics of each TF primitive, which this shim
    restates from SURVEY.md Appendix A exactly like oracle/tf12.py does (both delegate to the same
    functions).  Parity therefore stays "unpinned" with respect to TensorFlow 1.2 itself.

Variables get TF-style scoped names ('encoder/pre_net/dense/kernel', 'dense_1', ...); a `lookup(name, shape)`
callback supplies their values, so the same weights can be fed to the reference code and to the oracle.
"""
from __future__ import annotations

import collections
import contextlib
import sys
import types

import torch

from . import tf12

float32 = torch.float32
int32 = torch.int32


# ----------------------------------------------------------------------------------------------
# graph state: variable scopes, variable store, injected randomness
# ----------------------------------------------------------------------------------------------
class _State:
    def __init__(self):
        self.reset(None)

    def reset(self, lookup, dropout_masks=None, sample_masks=None):
        self.scope = []                # current variable-scope path
        self.counters = {}             # (scope path, base name) -> next index for default layer names
        self.vars = collections.OrderedDict()
        self.lookup = lookup
        self.dropout_masks = list(dropout_masks or [])    # consumed in call order by layers.dropout(training=True)
        self.sample_masks = list(sample_masks or [])      # consumed per step by ScheduledOutputTrainingHelper
        self.created = []              # variable names in creation order


S = _State()


def _path(name=None):
    p = "/".join(S.scope)
    if name:
        p = f"{p}/{name}" if p else name
    return p


def _unique(base):
    key = (_path(), base)
    i = S.counters.get(key, 0)
    S.counters[key] = i + 1
    return base if i == 0 else f"{base}_{i}"


@contextlib.contextmanager
def variable_scope(name, initializer=None, reuse=None):
    S.scope.append(name)
    try:
        yield
    finally:
        S.scope.pop()


def get_variable(name, shape=None, dtype=None, initializer=None):
    full = _path(name)
    if full not in S.vars:
        v = S.lookup(full, tuple(shape))
        assert tuple(v.shape) == tuple(shape), (full, tuple(v.shape), tuple(shape))
        S.vars[full] = v
        S.created.append(full)
    return S.vars[full]


def Variable(value, name=None, trainable=True):
    return value


def placeholder(dtype, shape=None):
    return None


# ----------------------------------------------------------------------------------------------
# tensor ops used by the reference
# ----------------------------------------------------------------------------------------------
def concat(values, axis):
    return torch.cat(list(values), dim=axis)


def shape(x):
    return list(x.shape)


def reshape(x, shp):
    return x.reshape(tuple(int(s) for s in shp))


def slice(x, begin, size):                      # noqa: A001  (tf.slice)
    idx = tuple(builtins_slice(b, None if s == -1 else b + s) for b, s in zip(begin, size))
    return x[idx]


builtins_slice = __builtins__["slice"] if isinstance(__builtins__, dict) else __builtins__.slice


def transpose(x, perm):
    return x.permute(*perm)


def tile(x, multiples):
    t = torch.as_tensor(x)
    return t.repeat(*[int(m) for m in multiples])


def zeros(shp, dtype=float32):
    if isinstance(shp, int):
        shp = [shp]
    return torch.zeros(*[int(s) for s in shp], dtype=dtype)


def expand_dims(x, axis):
    return x.unsqueeze(axis)


def reduce_sum(x, axis=None):
    return x.sum() if axis is None else x.sum(axis)


def abs(x):                                     # noqa: A001
    return x.abs()


def cast(x, dtype):
    return x.to(dtype)


def global_norm(ts):
    return None


def clip_by_global_norm(ts, c):
    return ts, None


# ----------------------------------------------------------------------------------------------
# tf.nn
# ----------------------------------------------------------------------------------------------
nn = types.SimpleNamespace()
nn.relu = torch.relu
nn.sigmoid = torch.sigmoid
nn.tanh = torch.tanh
nn.embedding_lookup = lambda table, ids: table[ids.to(torch.int64)]


def _bidirectional_dynamic_rnn(cell_fw, cell_bw, inputs, initial_state_fw=None, initial_state_bw=None, dtype=None,
                               sequence_length=None):
    assert sequence_length is None and initial_state_fw is None and initial_state_bw is None
    B, T, _ = inputs.shape
    outs = []
    with variable_scope("bidirectional_rnn"):
        for direction, cell in (("fw", cell_fw), ("bw", cell_bw)):
            with variable_scope(direction):
                h = cell.zero_state(B, inputs.dtype)
                seq = range(T) if direction == "fw" else range(T - 1, -1, -1)
                o = [None] * T
                for t in seq:
                    out, h = cell(inputs[:, t], h)
                    o[t] = out
                outs.append(torch.stack(o, 1))
    return tuple(outs), None


nn.bidirectional_dynamic_rnn = _bidirectional_dynamic_rnn

# ----------------------------------------------------------------------------------------------
# tf.layers  (semantics: oracle/tf12.py)
# ----------------------------------------------------------------------------------------------
layers = types.SimpleNamespace()


def _dense(inputs, units, activation=None, use_bias=True, name=None):
    with variable_scope(name or _unique("dense")):
        W = get_variable("kernel", (inputs.shape[-1], units))
        b = get_variable("bias", (units,)) if use_bias else None
    return tf12.dense(inputs, W, b, activation)


def _conv1d(inputs, filters, kernel_size, padding="valid", activation=None, strides=1):
    assert padding == "same" and strides == 1
    with variable_scope(_unique("conv1d")):
        W = get_variable("kernel", (kernel_size, inputs.shape[-1], filters))
        b = get_variable("bias", (filters,))
    return tf12.conv1d_same(inputs, W, b, activation)


def _batch_normalization(inputs, training=False):
    assert training is False                      # the reference never passes training=
    with variable_scope(_unique("batch_normalization")):
        C = inputs.shape[-1]
        g, b = get_variable("gamma", (C,)), get_variable("beta", (C,))
        m, v = get_variable("moving_mean", (C,)), get_variable("moving_variance", (C,))
    return tf12.batch_norm_inference(inputs, g, b, m, v)


def _max_pooling1d(inputs, pool_size, strides, padding="valid"):
    assert pool_size == 2 and strides == 1 and padding == "same"
    return tf12.max_pool_2_1_same(inputs)


def _dropout(inputs, rate=0.5, training=False):
    if not training:
        return inputs
    mask = S.dropout_masks.pop(0)
    assert tuple(mask.shape) == tuple(inputs.shape), (tuple(mask.shape), tuple(inputs.shape))
    return tf12.dropout(inputs, mask, rate)


layers.dense, layers.conv1d, layers.batch_normalization = _dense, _conv1d, _batch_normalization
layers.max_pooling1d, layers.dropout = _max_pooling1d, _dropout

summary = types.SimpleNamespace(histogram=lambda *a, **k: None, scalar=lambda *a, **k: None, merge_all=lambda: None)


class _Adam:
    def __init__(self, learning_rate=None):
        pass

    def compute_gradients(self, loss):
        return [(None, None)]

    def apply_gradients(self, gv, global_step=None):
        return None


train = types.SimpleNamespace(AdamOptimizer=_Adam)
contrib = types.SimpleNamespace(layers=types.SimpleNamespace(xavier_initializer=lambda: None))


# ----------------------------------------------------------------------------------------------
# tensorflow.contrib.rnn  
# ----------------------------------------------------------------------------------------------
class GRUCell:
    def __init__(self, num_units):
        self._n = num_units

    @property
    def output_size(self):
        return self._n

    def zero_state(self, batch_size, dtype):
        return torch.zeros(batch_size, self._n, dtype=dtype)

    def __call__(self, inputs, state):
        with variable_scope("gru_cell"):
            d = inputs.shape[-1] + self._n
            with variable_scope("gates"):
                Wg, bg = get_variable("kernel", (d, 2 * self._n)), get_variable("bias", (2 * self._n,))
            with variable_scope("candidate"):
                Wc, bc = get_variable("kernel", (d, self._n)), get_variable("bias", (self._n,))
        h = tf12.gru_cell(inputs, state, Wg, bg, Wc, bc)
        return h, h


class MultiRNNCell:
    def __init__(self, cells):
        self._cells = cells

    def zero_state(self, batch_size, dtype):
        return tuple(c.zero_state(batch_size, dtype) for c in self._cells)

    def __call__(self, inputs, state):
        new = []
        cur = inputs
        with variable_scope("multi_rnn_cell"):
            for i, c in enumerate(self._cells):
                with variable_scope(f"cell_{i}"):
                    cur, s = c(cur, state[i])
                    new.append(s)
        return cur, tuple(new)


class ResidualWrapper:
    def __init__(self, cell):
        self._cell = cell

    def zero_state(self, batch_size, dtype):
        return self._cell.zero_state(batch_size, dtype)

    def __call__(self, inputs, state):
        out, s = self._cell(inputs, state)
        return inputs + out, s


class InputProjectionWrapper:
    def __init__(self, cell, num_proj):
        self._cell, self._p = cell, num_proj

    def zero_state(self, batch_size, dtype):
        return self._cell.zero_state(batch_size, dtype)

    def __call__(self, inputs, state):
        with variable_scope("input_projection_wrapper"):
            W, b = get_variable("kernel", (inputs.shape[-1], self._p)), get_variable("bias", (self._p,))
        return self._cell(inputs @ W + b, state)


class OutputProjectionWrapper:
    def __init__(self, cell, output_size):
        self._cell, self._o = cell, output_size

    def zero_state(self, batch_size, dtype):
        return self._cell.zero_state(batch_size, dtype)

    def __call__(self, inputs, state):
        out, s = self._cell(inputs, state)
        with variable_scope("output_projection_wrapper"):
            W, b = get_variable("kernel", (out.shape[-1], self._o)), get_variable("bias", (self._o,))
        return out @ W + b, s


# ----------------------------------------------------------------------------------------------
# tensorflow.contrib.seq2seq  
# ----------------------------------------------------------------------------------------------
class BahdanauAttention:
    def __init__(self, num_units, memory, memory_sequence_length=None, normalize=False):
        assert not normalize
        self._u = num_units
        with variable_scope("memory_layer"):
            W = get_variable("kernel", (memory.shape[-1], num_units))
        self.values, self.keys, self.mask = tf12.attention_prepare(memory, memory_sequence_length, W)
        self.batch_size, self.alignments_size = memory.shape[0], memory.shape[1]

    def __call__(self, query, previous_alignments=None):
        with variable_scope("bahdanau_attention"):
            with variable_scope("query_layer"):
                Wq = get_variable("kernel", (query.shape[-1], self._u))
            v = get_variable("attention_v", (self._u,))
        return tf12.bahdanau_alignments(query, self.keys, self.mask, Wq, v)


AttentionWrapperState = collections.namedtuple("AttentionWrapperState", "cell_state attention time alignments alignment_history")


class _History:
    def __init__(self, items=()):
        self.items = list(items)

    def write(self, t, x):
        return _History(self.items + [x])

    def stack(self):
        return torch.stack(self.items, 0)


class AttentionWrapper:
    def __init__(self, cell, attention_mechanism, attention_layer_size=None, alignment_history=False, cell_input_fn=None,
                 output_attention=True):
        self._cell, self._m, self._als = cell, attention_mechanism, attention_layer_size
        self._hist, self._fn, self._oa = alignment_history, cell_input_fn, output_attention

    def zero_state(self, batch_size, dtype):
        return AttentionWrapperState(self._cell.zero_state(batch_size, dtype), torch.zeros(batch_size, self._als, dtype=dtype), 0,
                                     torch.zeros(batch_size, self._m.alignments_size, dtype=dtype), _History())

    def __call__(self, inputs, state):
        with variable_scope("attention_wrapper"):
            cell_inputs = self._fn(inputs, state.attention)
            cell_output, next_cell_state = self._cell(cell_inputs, state.cell_state)
            alignments = self._m(cell_output, previous_alignments=state.alignments)
            context = torch.bmm(alignments[:, None, :], self._m.values)[:, 0]
            with variable_scope("attention_layer"):
                Wa = get_variable("kernel", (cell_output.shape[-1] + context.shape[-1], self._als))
            attention = torch.cat([cell_output, context], 1) @ Wa
            hist = state.alignment_history.write(state.time, alignments) if self._hist else ()
        ns = AttentionWrapperState(next_cell_state, attention, state.time + 1, alignments, hist)
        return (attention if self._oa else cell_output), ns


class CustomHelper:
    def initialize(self):
        return self._initialize_fn()

    def sample(self, time, outputs, state):
        return self._sample_fn(time=time, outputs=outputs, state=state)

    def next_inputs(self, time, outputs, state, sample_ids):
        return self._next_inputs_fn(time=time, outputs=outputs, state=state, sample_ids=sample_ids)


class TrainingHelper:
    def __init__(self, inputs, sequence_length):
        self._inputs = inputs.transpose(0, 1)          # time major
        self._len = sequence_length.to(torch.int64)
        self._zero = torch.zeros_like(self._inputs[0])

    def initialize(self):
        finished = self._len <= 0
        return finished, (self._zero if bool(finished.all()) else self._inputs[0])

    def sample(self, time, outputs, state):
        return torch.zeros(outputs.shape[0], dtype=torch.int32)

    def next_inputs(self, time, outputs, state, sample_ids):
        nt = time + 1
        finished = nt >= self._len
        nxt = self._zero if bool(finished.all()) else self._inputs[nt]
        return finished, nxt, state


class ScheduledOutputTrainingHelper(TrainingHelper):
    def __init__(self, inputs, sequence_length, sampling_probability):
        super().__init__(inputs, sequence_length)

    def sample(self, time, outputs, state):
        return S.sample_masks.pop(0).to(torch.bool)     # injected Bernoulli(p) draw, one per batch element

    def next_inputs(self, time, outputs, state, sample_ids):
        finished, base, state = super().next_inputs(time, outputs, state, sample_ids)
        if bool(finished.all()):
            return finished, base, state
        return finished, torch.where(sample_ids[:, None], outputs, base), state


BasicDecoderOutput = collections.namedtuple("BasicDecoderOutput", "rnn_output sample_id")


class BasicDecoder:
    def __init__(self, cell, helper, initial_state, output_layer=None):
        assert output_layer is None
        self._cell, self._helper, self._init = cell, helper, initial_state

    def initialize(self):
        return self._helper.initialize() + (self._init,)

    def step(self, time, inputs, state):
        out, cs = self._cell(inputs, state)
        ids = self._helper.sample(time=time, outputs=out, state=cs)
        finished, nxt, ns = self._helper.next_inputs(time=time, outputs=out, state=cs, sample_ids=ids)
        return BasicDecoderOutput(out, ids), ns, nxt, finished


def dynamic_decode(decoder, maximum_iterations=None, impute_finished=False):
    assert not impute_finished
    with variable_scope("decoder"):
        finished, inputs, state = decoder.initialize()
        finished = torch.as_tensor(finished).reshape(-1).to(torch.bool)
        time, outs = 0, []
        # TF traces the while_loop body ONCE: layers created inside it (the decoder pre-net's tf.layers.dense calls)
        # are the same variables at every step.  Eagerly, that means resetting the default-name counters per step.
        counters0 = dict(S.counters)
        while not bool(finished.all()) and (maximum_iterations is None or time < maximum_iterations):
            S.counters = dict(counters0)
            out, state, inputs, dfin = decoder.step(time, inputs, state)
            finished = finished | torch.as_tensor(dfin).reshape(-1).to(torch.bool)
            if maximum_iterations is not None:
                finished = finished | torch.tensor(time + 1 >= maximum_iterations)
            outs.append(out.rnn_output)
            time += 1
    return BasicDecoderOutput(torch.stack(outs, 1), None), state, None


# ----------------------------------------------------------------------------------------------
# module installation: make `import tensorflow as tf`, `from tensorflow.contrib.rnn import *`, ... resolve here
# ----------------------------------------------------------------------------------------------
def install():
    me = sys.modules[__name__]
    tfm = types.ModuleType("tensorflow")
    for k in ("float32", "int32", "variable_scope", "get_variable", "Variable", "placeholder", "concat", "shape", "reshape",
              "slice", "transpose", "tile", "zeros", "expand_dims", "reduce_sum", "abs", "cast", "global_norm",
              "clip_by_global_norm", "nn", "layers", "summary", "train", "contrib"):
        setattr(tfm, k, getattr(me, k))
    rnn = types.ModuleType("tensorflow.contrib.rnn")
    rnn.__all__ = ["GRUCell", "MultiRNNCell", "ResidualWrapper", "InputProjectionWrapper", "OutputProjectionWrapper"]
    for k in rnn.__all__:
        setattr(rnn, k, getattr(me, k))
    aw = types.ModuleType("attention_wrapper"); aw.BahdanauAttention = BahdanauAttention; aw.AttentionWrapper = AttentionWrapper
    hp = types.ModuleType("helper")
    hp.CustomHelper, hp.TrainingHelper, hp.ScheduledOutputTrainingHelper = CustomHelper, TrainingHelper, ScheduledOutputTrainingHelper
    bd = types.ModuleType("basic_decoder"); bd.BasicDecoder = BasicDecoder
    dc = types.ModuleType("decoder"); dc.dynamic_decode = dynamic_decode
    opsm = types.ModuleType("tensorflow.contrib.seq2seq.python.ops")
    opsm.attention_wrapper, opsm.helper, opsm.basic_decoder, opsm.decoder = aw, hp, bd, dc
    mods = {
        "tensorflow": tfm, "tensorflow.contrib": types.ModuleType("tensorflow.contrib"), "tensorflow.contrib.rnn": rnn,
        "tensorflow.contrib.seq2seq": types.ModuleType("tensorflow.contrib.seq2seq"),
        "tensorflow.contrib.seq2seq.python": types.ModuleType("tensorflow.contrib.seq2seq.python"),
        "tensorflow.contrib.seq2seq.python.ops": opsm,
        "tensorflow.contrib.seq2seq.python.ops.helper": hp, "tensorflow.contrib.seq2seq.python.ops.attention_wrapper": aw,
        "tensorflow.contrib.seq2seq.python.ops.basic_decoder": bd, "tensorflow.contrib.seq2seq.python.ops.decoder": dc,
    }
    sys.modules.update(mods)
    return tfm
