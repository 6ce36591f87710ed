"""Host-side mirror of the reference's ``models/ops.py`` over the sm_100a C-ABI.

Same surface as the reference -- ``highway``, ``CBHG``, ``InferenceHelper`` (models/ops.py:5-132)
-- plus the north-star spellings ``pre_net``, ``conv1d_banks``, ``highwaynet``,
``attention_decoder``.  Tensors are fp32 CUDA ``torch.Tensor`` s in TF's NWC layout; every
arithmetic op is a call into ``libtaco_b200.so`` (tacotron_b200/_lib.py).  No CPU fallback.

TF-1 builds a graph inside ``tf.variable_scope`` s; here the equivalent is a ``Scope`` (a
ParamStore + name prefix) taken from the innermost ``variable_scope(...)`` context, so call
sites read like the reference:

    with variable_scope(store, 'enc'):
        encoded = CBHG(pre_out, None, K=16, c=[128,128,128], gru_units=128)
"""
from __future__ import annotations

import contextlib
import ctypes as C

import torch

from .. import _lib as L
from ..params import BN_EPS

# ----------------------------------------------------------------------------------------------
# scopes + derived (kernel-layout) buffers
# ----------------------------------------------------------------------------------------------
_scope_stack = []


class Scope:
    def __init__(self, store, prefix=""):
        self.store = store
        self.prefix = prefix.strip("/")

    def name(self, n):
        return f"{self.prefix}/{n}" if self.prefix else n

    def p(self, n):
        return self.store[self.name(n)]

    def has(self, n):
        return self.name(n) in self.store

    def sub(self, n):
        return Scope(self.store, self.name(n))


@contextlib.contextmanager
def variable_scope(store_or_name, prefix=None):
    """``with variable_scope(store, 'enc'):`` or nested ``with variable_scope('cbhg'):``."""
    if prefix is None and isinstance(store_or_name, str):
        parent = current_scope()
        sc = parent.sub(store_or_name)
    else:
        sc = Scope(store_or_name, prefix or "")
    _scope_stack.append(sc)
    try:
        yield sc
    finally:
        _scope_stack.pop()


def current_scope():
    if not _scope_stack:
        raise RuntimeError("no variable_scope active: wrap the call in `with variable_scope(store, prefix):`")
    return _scope_stack[-1]


# ----------------------------------------------------------------------------------------------
# train-mode activation saving: inside `with saving(S):` every op records the tensors the hand-written
# backward (models/grad.py) needs, under the names grad.py reads, and uses the un-fused variants whose
# intermediates must exist in memory (bank max-pool input, highway pre-activations).
# ----------------------------------------------------------------------------------------------
_save = None


@contextlib.contextmanager
def saving(S):
    global _save
    prev, _save = _save, S
    try:
        yield S
    finally:
        _save = prev


class Runtime:
    """Per-store kernel-side state: precision mode, derived operand buffers, scratch buffers."""

    def __init__(self, store, precision="tf32"):
        assert precision in ("tf32", "fp32", "fp32x3")
        self.store = store
        self.precision = precision
        # 'fp32x3': tcgen05 tensor cores with the error-compensated 3xTF32 split (fp32-grade products, ~1e-6);
        # 'tf32': single-pass TF32 products; 'fp32': exact-product FFMA kernel (cross-check path)
        self.impl = {"tf32": L.IMPL_TC, "fp32": L.IMPL_SIMT, "fp32x3": L.IMPL_TC3}[precision]
        self.derived = {}
        self.derived_version = {}
        self.scratch = {}

    def buf(self, tag, shape, dtype=torch.float32):
        key = (tag, tuple(shape), dtype)
        t = self.scratch.get(key)
        if t is None:
            t = torch.empty(shape, dtype=dtype, device=self.store.device)
            self.scratch[key] = t
        return t

    def get(self, key, builder):
        """Derived buffer cache keyed on the store version (rebuilt in place after an update)."""
        if self.derived_version.get(key) != self.store.version:
            self.derived[key] = builder(self.derived.get(key))
            self.derived_version[key] = self.store.version
        return self.derived[key]


def runtime(store, precision=None):
    rt = getattr(store, "_runtime", None)
    if rt is None or (precision is not None and rt.precision != precision):
        rt = Runtime(store, precision or "tf32")
        store._runtime = rt
    return rt


def _cpad(c):
    return (c + 31) // 32 * 32


def _pack(rt, key, W, taps, Cin, N, ld=None, row0=0, dst=None, n_total=None):
    """TF [taps][C][N] -> K-major TF32 operand rows [row0, row0+N) of dst (tensor-core path).  For the 3xTF32 path
    dst has 2*n_total rows: the hi heads in rows [0, n_total), the lo remainders in rows [n_total, 2 n_total)."""
    ld = ld or taps * _cpad(Cin)
    n_total = n_total or N
    x3 = rt.impl == L.IMPL_TC3
    if dst is None:
        dst = torch.zeros(((2 if x3 else 1) * n_total, ld), dtype=torch.float32, device=W.device)
    if x3:
        L.check(L.lib().taco_pack_weight_x3(L.ptr(W), taps, Cin, N, C.c_void_p(dst.data_ptr() + row0 * ld * 4),
                                            C.c_void_p(dst.data_ptr() + (n_total + row0) * ld * 4), ld, L.current_stream()),
                "taco_pack_weight_x3")
    else:
        L.check(L.lib().taco_pack_weight(L.ptr(W), taps, Cin, N, C.c_void_p(dst.data_ptr() + row0 * ld * 4), ld,
                                         L.current_stream()), "taco_pack_weight")
    return dst


def packed_weight(rt, name, W, taps, Cin, N):
    if rt.impl == L.IMPL_SIMT:
        return None
    return rt.get(("pack", name), lambda old: _pack(rt, name, W.contiguous(), taps, Cin, N, dst=old))


def bn_affine(rt, sc):
    """Inference-mode batch_normalization folded to y = x*scale + shift."""
    def build(old):
        scale = sc.p("bn_gamma") / torch.sqrt(sc.p("bn_var") + BN_EPS)
        shift = sc.p("bn_beta") - sc.p("bn_mean") * scale
        if old is not None:
            old[0].copy_(scale); old[1].copy_(shift)
            return old
        return (scale.contiguous(), shift.contiguous())
    return rt.get(("bn", sc.prefix), build)


# ----------------------------------------------------------------------------------------------
# the one contraction primitive
# ----------------------------------------------------------------------------------------------
def linear(rt, x, W, Wp, N, *, taps=1, bank_K=0, bank_cout=0, bias=None, act=L.ACT_NONE, scale=None, shift=None,
           keep=None, keep_scale=1.0, residual=None, highway_x=None, pool=False, out=None, tag="lin"):
    """x [B,T,C] (contiguous) -> [B,T,N'] through taco_linear_fwd.  See include/taco_b200.h."""
    assert x.dim() == 3 and x.is_cuda and x.dtype == torch.float32 and x.stride(2) == 1
    B, T, Cin = x.shape
    assert x.stride(0) == T * x.stride(1)
    n_out = N // 2 if highway_x is not None else N
    if out is None:
        out = rt.buf(tag, (B, T, n_out))
    d = L.LinearDesc()
    d.X = x.data_ptr(); d.ldx = x.stride(1); d.B = B; d.T = T; d.C = Cin
    d.taps = taps; d.tap0 = -((taps - 1) // 2); d.N = N
    d.bank_K = bank_K; d.bank_cout = bank_cout
    d.W = W.data_ptr() if W is not None else None
    if Wp is not None:
        d.Wp = Wp.data_ptr(); d.ldwp = Wp.stride(0)
    d.Y = out.data_ptr(); d.ldy = out.stride(1)
    d.bias = bias.data_ptr() if bias is not None else None
    d.act = act
    d.scale = scale.data_ptr() if scale is not None else None
    d.shift = shift.data_ptr() if shift is not None else None
    if keep is not None:
        assert keep.dtype == torch.uint8 and keep.is_contiguous() and keep.numel() == B * T * n_out
        d.keep = keep.data_ptr()
    d.keep_scale = float(keep_scale)
    if residual is not None:
        assert residual.stride(-1) == 1
        d.residual = residual.data_ptr(); d.ldr = residual.stride(-2)
    if highway_x is not None:
        d.epilogue = L.EPI_HIGHWAY
        d.hx = highway_x.data_ptr(); d.ldhx = highway_x.stride(-2)
    d.pool = 1 if pool else 0
    d.impl = rt.impl
    L.check(L.lib().taco_linear_fwd(C.byref(d), L.current_stream()), f"taco_linear_fwd[{tag}]")
    return out


def _dropout_mask(shape, rate, device):
    return (torch.rand(shape, device=device) >= rate).to(torch.uint8)


# ----------------------------------------------------------------------------------------------
# Tacotron.pre_net (models/tacotron.py:38-44) -- lives on the model in the reference; the
# north-star surface lists it under ops, so it is defined here and bound as a method there.
# ----------------------------------------------------------------------------------------------
def pre_net(inputs, units=(256, 128), dropout=0.5, train=True, masks=None, scope=None, ids=None):
    """dense(256,relu) -> dropout -> dense(128,relu) -> dropout.  `inputs` [B,T,Cin].
    With `ids` (int32 [B,T]) `inputs` is the embedding TABLE [V,Cin]: layer 1 is then applied to
    the V table rows once and gathered (identical arithmetic per row, V << B*T)."""
    sc = scope or current_scope().sub("prenet")
    rt = runtime(sc.store)
    W1, b1, W2, b2 = sc.p("W1"), sc.p("b1"), sc.p("W2"), sc.p("b2")
    ks = 1.0 / (1.0 - dropout)
    if ids is not None:
        V, Cin = inputs.shape
        B, T = ids.shape
        t1 = linear(rt, inputs.view(1, V, Cin), W1, packed_weight(rt, sc.name("W1"), W1, 1, Cin, units[0]), units[0],
                    bias=b1, act=L.ACT_RELU, tag=sc.name("t1"))
        m1 = m2 = None
        if train:
            m1, m2 = masks if masks is not None else (_dropout_mask((B, T, units[0]), dropout, ids.device),
                                                      _dropout_mask((B, T, units[1]), dropout, ids.device))
        l1 = rt.buf(sc.name("l1"), (B, T, units[0]))
        L.check(L.lib().taco_gather_rows(L.ptr(t1), L.ptr(ids), B * T, units[0], V, L.ptr(m1), ks, L.ptr(l1),
                                         L.current_stream()), "taco_gather_rows")
    else:
        B, T, Cin = inputs.shape
        m1 = m2 = None
        if train:
            m1, m2 = masks if masks is not None else (_dropout_mask((B, T, units[0]), dropout, inputs.device),
                                                      _dropout_mask((B, T, units[1]), dropout, inputs.device))
        l1 = linear(rt, inputs, W1, packed_weight(rt, sc.name("W1"), W1, 1, Cin, units[0]), units[0], bias=b1,
                    act=L.ACT_RELU, keep=m1, keep_scale=ks, tag=sc.name("l1"))
    l2 = linear(rt, l1, W2, packed_weight(rt, sc.name("W2"), W2, 1, units[0], units[1]), units[1], bias=b2,
                act=L.ACT_RELU, keep=m2, keep_scale=ks, tag=sc.name("l2"))
    if _save is not None and ids is not None:
        _save[sc.name("t1")] = t1.view(V, units[0]); _save[sc.name("l1")] = l1; _save[sc.name("l2")] = l2
    return l2


# ----------------------------------------------------------------------------------------------
# conv bank (models/ops.py:54-71): K conv1d 'same' + relu, concat, batch-norm, max-pool(2,1,same)
# ----------------------------------------------------------------------------------------------
def conv1d_banks(inputs, K=16, cout=128, scope=None):
    sc = scope or current_scope().sub("bank")
    rt = runtime(sc.store)
    B, T, Cin = inputs.shape
    Wall = sc.store.span(sc.name("W1"), sc.name(f"W{K}"))
    ball = sc.store.span(sc.name("b1"), sc.name(f"b{K}"))
    scale, shift = bn_affine(rt, sc)
    Wp = None
    if rt.impl != L.IMPL_SIMT:
        def build(old):
            ld = K * _cpad(Cin)
            rows = K * cout * (2 if rt.impl == L.IMPL_TC3 else 1)
            dst = old if old is not None else torch.zeros((rows, ld), dtype=torch.float32, device=inputs.device)
            for k in range(1, K + 1):
                _pack(rt, None, sc.p(f"W{k}"), k, Cin, cout, ld=ld, row0=(k - 1) * cout, dst=dst, n_total=K * cout)
            return dst
        Wp = rt.get(("pack", sc.name("bank")), build)
        if _save is None:
            return linear(rt, inputs, Wall, Wp, K * cout, bank_K=K, bank_cout=cout, bias=ball, act=L.ACT_RELU, scale=scale,
                          shift=shift, pool=True, tag=sc.name("pool"))
    bank = linear(rt, inputs, Wall, Wp, K * cout, bank_K=K, bank_cout=cout, bias=ball, act=L.ACT_RELU, scale=scale,
                  shift=shift, tag=sc.name("bn"))
    pooled = rt.buf(sc.name("pool"), (B, T, K * cout))
    L.check(L.lib().taco_maxpool_fwd(L.ptr(bank), L.ptr(pooled), B, T, K * cout, L.current_stream()), "taco_maxpool_fwd")
    if _save is not None:
        _save[sc.name("bn")] = bank; _save[sc.name("pool")] = pooled; _save[sc.name("bn_affine")] = (scale, shift)
    return pooled


# ----------------------------------------------------------------------------------------------
# highway (models/ops.py:27-46)
# ----------------------------------------------------------------------------------------------
def highway(inputs, units=128, scope=None):
    sc = scope or current_scope()
    rt = runtime(sc.store)
    B, T, Cin = inputs.shape
    if Cin != units:                                            # ops.py:29-30
        Wd, bd = sc.p("Wd"), sc.p("bd")
        inputs = linear(rt, inputs, Wd, packed_weight(rt, sc.name("Wd"), Wd, 1, Cin, units), units, bias=bd,
                        tag=sc.name("d"))

    def build_w(old):
        w = torch.cat([sc.p("WH"), sc.p("WT")], dim=1)
        if old is not None:
            old.copy_(w); return old
        return w.contiguous()

    def build_b(old):
        b = torch.cat([sc.p("bH"), sc.p("bT")])
        if old is not None:
            old.copy_(b); return old
        return b.contiguous()
    Whw = rt.get(("hw_w", sc.prefix), build_w)
    bhw = rt.get(("hw_b", sc.prefix), build_b)
    Wp = packed_weight(rt, sc.name("hw"), Whw, 1, units, 2 * units)
    if _save is not None:                                       # train: keep [h_pre | t_pre] for the backward
        from .. import kernels as K
        Pm = linear(rt, inputs, Whw, Wp, 2 * units, bias=bhw, tag=sc.name("P"))
        out = rt.buf(sc.name("out"), (B, T, units))
        K.highway_fwd(out.view(-1, units), Pm.view(-1, 2 * units), inputs.view(-1, units))
        _save[sc.name("in")] = inputs; _save[sc.name("P")] = Pm
        return out
    return linear(rt, inputs, Whw, Wp, 2 * units, bias=bhw, highway_x=inputs, tag=sc.name("out"))


highwaynet = highway


# ----------------------------------------------------------------------------------------------
# bidirectional GRU (models/ops.py:118-128)
# ----------------------------------------------------------------------------------------------
def bidirectional_gru(inputs, gru_units=128, scope=None):
    sc = scope or current_scope()
    rt = runtime(sc.store)
    assert gru_units == 128 and inputs.shape[-1] == 128, "the persistent GRU kernel is specialised for 128 units"
    B, T, Cin = inputs.shape
    fw, bw = sc.sub("gru_fw"), sc.sub("gru_bw")

    def build_w(old):
        w = torch.cat([fw.p("Wg")[:Cin], fw.p("Wc")[:Cin], bw.p("Wg")[:Cin], bw.p("Wc")[:Cin]], dim=1)
        if old is not None:
            old.copy_(w); return old
        return w.contiguous()

    def build_b(old):
        b = torch.cat([fw.p("bg"), fw.p("bc"), bw.p("bg"), bw.p("bc")])
        if old is not None:
            old.copy_(b); return old
        return b.contiguous()
    Wx = rt.get(("gru_wx", sc.prefix), build_w)                  # [128, 768]
    bx = rt.get(("gru_bx", sc.prefix), build_b)
    xp = linear(rt, inputs, Wx, packed_weight(rt, sc.name("gru_wx"), Wx, 1, Cin, 768), 768, bias=bx, tag=sc.name("xp"))
    out = rt.buf(sc.name("gru_out"), (B, T, 2 * gru_units))
    L.check(L.lib().taco_bigru_fwd(L.ptr(xp), L.ptr(fw.p("Wg")[Cin:]), L.ptr(fw.p("Wc")[Cin:]), L.ptr(bw.p("Wg")[Cin:]),
                                   L.ptr(bw.p("Wc")[Cin:]), L.ptr(out), B, T, L.current_stream()), "taco_bigru_fwd")
    if _save is not None:
        _save[sc.name("hw_out")] = inputs; _save[sc.name("xp")] = xp; _save[sc.name("gru_out")] = out
    return out


# ----------------------------------------------------------------------------------------------
# CBHG (models/ops.py:48-132)
# ----------------------------------------------------------------------------------------------
def CBHG(inputs, speaker_embed=None, K=16, c=[128, 128, 128], gru_units=128, num_highway_layers=4, num_conv_proj=2,
         trace=None):
    if speaker_embed is not None:
        raise NotImplementedError("multi-speaker path (ops.py:101-115) is out of scope: num_speakers == 1")
    assert num_conv_proj == len(c) - 1                           # ops.py:75
    with variable_scope("cbhg") as sc:
        rt = runtime(sc.store)
        B, T, Cin = inputs.shape
        conv_bank = conv1d_banks(inputs, K=K, cout=c[0], scope=sc.sub("bank"))
        conv_proj = conv_bank
        cin = K * c[0]
        for layer in range(num_conv_proj):
            psc = sc.sub(f"proj{layer + 1}")
            last = layer == num_conv_proj - 1
            W, b = psc.p("W"), psc.p("b")
            scale, shift = bn_affine(rt, psc)
            conv_proj = linear(rt, conv_proj, W, packed_weight(rt, psc.name("W"), W, 3, cin, c[layer + 1]), c[layer + 1],
                               taps=3, bias=b, act=L.ACT_NONE if last else L.ACT_RELU, scale=scale, shift=shift,
                               residual=inputs if last else None, tag=psc.name("out"))   # +inputs: ops.py:92
            cin = c[layer + 1]
            if _save is not None:
                _save[psc.name("bn_affine")] = (scale, shift)
                _save[sc.name("res" if last else f"proj{layer + 1}")] = conv_proj
        h = conv_proj
        if trace is not None:
            trace[sc.name("bank_pool")] = conv_bank
            trace[sc.name("res")] = conv_proj
        if _save is not None:
            _save[sc.name("x_in")] = inputs
            _save[sc.name("bank_bn")] = _save[sc.name("bank/bn")]
            _save[sc.name("bank_pool")] = conv_bank
        for layer in range(num_highway_layers):
            h = highway(h, scope=sc.sub(f"highway{layer}"))
            if _save is not None:
                _save[sc.name(f"hw{layer}_in")] = _save[sc.name(f"highway{layer}/in")]
                _save[sc.name(f"hw{layer}_P")] = _save[sc.name(f"highway{layer}/P")]
        if trace is not None:
            trace[sc.name("highway_out")] = h
        out = bidirectional_gru(h, gru_units, scope=sc)
        if trace is not None:
            trace[sc.name("out")] = out
        return out


# ----------------------------------------------------------------------------------------------
# InferenceHelper (models/ops.py:5-25): first input zeros, next input = previous output, never
# finished.  The persistent decoder kernel implements exactly this policy as TACO_DEC_INFER; the
# class carries the policy choice and reproduces the helper protocol for host-side use.
# ----------------------------------------------------------------------------------------------
class InferenceHelper:
    decoder_mode = L.DEC_INFER

    def __init__(self, batch_size, out_size):
        self._batch_size = batch_size
        self._out_size = out_size

    @property
    def batch_size(self):
        return self._batch_size

    def initialize(self, device="cuda"):
        finished = torch.zeros(self._batch_size, dtype=torch.bool, device=device)
        next_inputs = torch.zeros(self._batch_size, self._out_size, dtype=torch.float32, device=device)
        return finished, next_inputs

    def sample(self, time, outputs, state):
        return torch.zeros(32, dtype=torch.int32, device=outputs.device)   # ops.py:15 (hard-coded 32)

    def next_inputs(self, time, outputs, state, sample_ids=None):
        finished = torch.zeros(self._batch_size, dtype=torch.bool, device=outputs.device)
        return finished, outputs, state


# ----------------------------------------------------------------------------------------------
# attention decoder (models/tacotron.py:46-105,136-138): one persistent kernel for all T steps
# ----------------------------------------------------------------------------------------------
def _decoder_weights(sc):
    w = L.DecoderWeights()
    w.pre_W1 = sc.p("prenet/W1").data_ptr(); w.pre_b1 = sc.p("prenet/b1").data_ptr()
    w.pre_W2 = sc.p("prenet/W2").data_ptr(); w.pre_b2 = sc.p("prenet/b2").data_ptr()
    w.in_W = sc.p("in_proj/W").data_ptr(); w.in_b = sc.p("in_proj/b").data_ptr()
    for i in range(3):
        w.gru_Wg[i] = sc.p(f"gru{i+1}/Wg").data_ptr(); w.gru_bg[i] = sc.p(f"gru{i+1}/bg").data_ptr()
        w.gru_Wc[i] = sc.p(f"gru{i+1}/Wc").data_ptr(); w.gru_bc[i] = sc.p(f"gru{i+1}/bc").data_ptr()
    w.out_W = sc.p("out_proj/W").data_ptr(); w.out_b = sc.p("out_proj/b").data_ptr()
    w.att_Wq = sc.p("attn/W_q").data_ptr(); w.att_v = sc.p("attn/v").data_ptr(); w.att_Wa = sc.p("attn/W_a").data_ptr()
    return w


def attention_decoder(encoded, text_length, r, T, mode=L.DEC_INFER, mel=None, sample_mask=None, drop_masks=None,
                      dropout=0.5, scope=None, step_ns=None):
    """encoded [B,Tx,256], text_length int32 [B] -> (seq2seq_output [B,T,80r], alignments [B,T,Tx])."""
    sc = scope or current_scope()
    rt = runtime(sc.store)
    B, Tx, D = encoded.shape
    assert D == 256
    lib = L.lib()
    out_w = 80 * r
    y = rt.buf(sc.name("y"), (B, T, out_w))
    align = rt.buf(sc.name("align"), (B, T, Tx))
    # BahdanauAttention memory: values = length-masked encoder states, keys = values . W_mem (tacotron.py:48-52)
    values = rt.buf(sc.name("values"), (B, Tx, D))
    L.check(lib.taco_mask_rows(L.ptr(encoded), L.ptr(text_length), L.ptr(values), B, Tx, D, L.current_stream()), "taco_mask_rows")
    Wm = sc.p("attn/W_mem")
    keys = linear(rt, values, Wm, packed_weight(rt, sc.name("attn/W_mem"), Wm, 1, D, 256), 256, tag=sc.name("keys"))

    w = _decoder_weights(sc)

    def build_pack(old):
        dst = old if old is not None else torch.empty(lib.taco_decoder_packed_bytes(r) // 4, dtype=torch.float32,
                                                      device=encoded.device)
        L.check(lib.taco_decoder_pack(C.byref(w), r, L.ptr(dst), L.current_stream()), "taco_decoder_pack")
        return dst
    packed = rt.get(("dec_pack", sc.prefix, r), build_pack)
    ws = rt.buf(sc.name("dec_ws"), (lib.taco_decoder_workspace_bytes(32, Tx, T, r) // 4,))
    rt.dec_ws = ws                                               # (per-slot time stamps live in its tail; see decoder.cu)
    ks = 1.0 / (1.0 - dropout)
    hs = None
    if _save is not None:                                        # train: the three GRU state sequences feed the backward
        assert B <= 32, "training decodes at most 32 utterances per rank (one launch)"
        hs = rt.buf(sc.name("H"), (3, T, B, 256))
        _save[sc.name("H")] = hs; _save[sc.name("values")] = values; _save[sc.name("keys")] = keys
        _save[sc.name("y")] = y; _save[sc.name("align")] = align
        if drop_masks is not None:
            _save[sc.name("keep1")], _save[sc.name("keep2")] = drop_masks
        if mode == L.DEC_SCHED:
            _save[sc.name("sample_mask")] = sample_mask
    for b0 in range(0, B, 32):                                   # the kernel handles <= 32 utterances per launch
        nb = min(32, B - b0)
        a = L.DecoderArgs()
        a.weights = C.pointer(w)
        a.packed = packed.data_ptr(); a.keys = keys[b0:].data_ptr(); a.values = values[b0:].data_ptr()
        a.text_length = text_length[b0:].data_ptr()
        a.mode = mode
        if mode != L.DEC_INFER:
            assert mel is not None and mel.is_contiguous() and mel.shape == (B, T, out_w)
            a.mel = mel[b0:].data_ptr()
        if mode == L.DEC_SCHED:
            assert B <= 32, "scheduled sampling with B > 32 needs a [T, B] mask slice per launch"
            assert sample_mask is not None and sample_mask.dtype == torch.uint8 and sample_mask.shape == (T, B)
            a.sample_mask = sample_mask.data_ptr()
        if drop_masks is not None:
            assert B <= 32
            k1, k2 = drop_masks
            assert k1.dtype == torch.uint8 and k1.shape == (T, B, 256) and k2.shape == (T, B, 128)
            a.keep1 = k1.data_ptr(); a.keep2 = k2.data_ptr()
        a.keep_scale = ks
        a.B = nb; a.Tx = Tx; a.T = T; a.r = r
        a.y = y[b0:].data_ptr(); a.align = align[b0:].data_ptr(); a.workspace = ws.data_ptr()
        a.step_ns = step_ns.data_ptr() if step_ns is not None else None
        a.h_save = hs.data_ptr() if hs is not None else None
        L.check(lib.taco_decoder_fwd(C.byref(a), L.current_stream()), "taco_decoder_fwd")
    return y, align
