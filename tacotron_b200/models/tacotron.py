"""Host-side mirror of the reference's ``models/tacotron.py`` (Config, Tacotron) over the sm_100a
C-ABI.  Eager equivalent of the TF-1.2 graph: ``Tacotron(config, inputs, train)`` owns the
variables (one flat device buffer, TF layouts) and ``inference`` / ``add_loss_op`` run the hand
written kernels; the attribute names the reference drivers fetch (``seq2seq_output``, ``output``,
``alignments``, ``loss``) are kept (models/tacotron.py:187-195, train.py:60-72, test.py:52-56).
"""
from __future__ import annotations

import torch

from .. import _lib as L
from ..params import ParamStore, model_shapes
from . import ops


class Config(object):
    """models/tacotron.py:12-33, with the values the reference derives from module globals or
    injects at run time (r, vocab_size, max_decode_iter: train.py:20-22, tacotron.py:13) explicit."""
    max_decode_iter = 108000 // (2 * 300)      # audio.maximum_audio_length // (audio.r * audio.hop_length)
    attention_units = 256
    decoder_units = 256
    mel_features = 80
    embed_dim = 256
    fft_size = 1025

    char_dropout_prob = 0.5
    audio_dropout_prob = 0.5

    num_speakers = 1
    speaker_embed_dim = 16

    scheduled_sample = 0.5

    cap_grads = 5

    init_lr = 0.0005
    annealing_rate = 1

    batch_size = 32

    # injected by the drivers in the reference
    r = 2
    vocab_size = 64

    # B200 path, feed-forward contractions (the recurrent kernels are fp32-grade in every mode):
    #   'fp32x3' (default) = tcgen05 tensor cores with the error-compensated 3xTF32 split: fp32-grade products
    #                        (~1e-6 relative), the reference's own arithmetic class;
    #   'tf32'             = single-pass TF32 products (10-bit mantissa), fp32 accumulate: faster, stated looser tolerance;
    #   'fp32'             = exact-product fp32 FFMA kernel (on-GPU cross-check of the tensor-core paths).
    precision = "fp32x3"
    # replay inference-mode calls from CUDA graphs (captured per input shape)
    cuda_graph = False

    def __init__(self, **kw):
        for k, v in kw.items():
            setattr(self, k, v)


class Tacotron(object):
    def __init__(self, config, inputs=None, train=True, device="cuda", seed=1):
        """Reference signature ``Tacotron(config, inputs, train)`` (tacotron.py:187).  `inputs` may be
        None (build only) or the input dict of tensors; when given the forward (and loss, if train)
        is evaluated immediately so the result attributes exist as they do after a sess.run."""
        if config.num_speakers != 1:
            raise NotImplementedError("multi-speaker (tacotron.py:116-124) is out of scope")
        if not torch.cuda.is_available():
            raise RuntimeError("tacotron_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        L.lib()                                   # fail loudly if the extension is missing
        self.config = config
        self.train = train
        self.device = torch.device(device)
        self.store = ParamStore(model_shapes(config), self.device)
        self.store.init_tf_default(seed)
        self.runtime = ops.runtime(self.store, config.precision)
        self.lr = config.init_lr
        self.global_step = 0
        self.seq2seq_output = self.output = self.alignments = self.loss = None
        self.step_ns = None
        self.section_wait = None
        self.last_graph_kernels = 0                # kernels inside the CUDA graphs replayed by the last call
        self._marks = None                        # bench.py: list of (name, cuda event) section boundaries
        if inputs is not None:
            self(inputs)

    def _mark(self, name):
        if self._marks is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            self._marks.append((name, ev))

    # -- parameters ---------------------------------------------------------------------------
    def load_params(self, params):
        self.store.load(params)

    def state_dict(self):
        return self.store.state_dict()

    # -- Tacotron.pre_net (tacotron.py:38-44) ---------------------------------------------------
    def pre_net(self, inputs, units=[256, 128], dropout=0.5, train=True, masks=None, scope=None, ids=None):
        return ops.pre_net(inputs, units=tuple(units), dropout=dropout, train=train, masks=masks, scope=scope, ids=ids)

    # -- create_decoder + dynamic_decode (tacotron.py:46-105, 136-138) ---------------------------
    def create_decoder(self, encoded, inputs, speaker_embed=None, train=True, T=None, dec_drop_masks=None,
                       sample_mask=None):
        """Returns a zero-argument callable that runs the whole decode (the BasicDecoder +
        dynamic_decode pair of the reference) as one persistent kernel."""
        cfg = self.config
        if train:
            mel = inputs["mel"]
            T = mel.shape[1] if T is None else T
            if cfg.scheduled_sample:                          # ScheduledOutputTrainingHelper (tacotron.py:83-85)
                mode = L.DEC_SCHED
                if sample_mask is None:
                    sample_mask = (torch.rand(T, mel.shape[0], device=mel.device) < cfg.scheduled_sample).to(torch.uint8)
            else:                                             # TrainingHelper (tacotron.py:87)
                mode = L.DEC_TEACHER
            if dec_drop_masks is None:
                B = mel.shape[0]
                dec_drop_masks = ((torch.rand(T, B, 256, device=mel.device) >= cfg.audio_dropout_prob).to(torch.uint8),
                                  (torch.rand(T, B, 128, device=mel.device) >= cfg.audio_dropout_prob).to(torch.uint8))
        else:
            helper = ops.InferenceHelper(inputs["text"].shape[0], cfg.mel_features * cfg.r)   # tacotron.py:89-92
            mode = helper.decoder_mode
            mel = None
            T = cfg.max_decode_iter if T is None else T
            dec_drop_masks = None
            sample_mask = None
        sc = ops.Scope(self.store, "dec")
        step_ns = self.step_ns

        def run():
            return ops.attention_decoder(encoded, inputs["text_length"], cfg.r, T, mode=mode, mel=mel,
                                         sample_mask=sample_mask, drop_masks=dec_drop_masks,
                                         dropout=cfg.audio_dropout_prob, scope=sc, step_ns=step_ns)
        return run

    # -- inference (tacotron.py:107-154) --------------------------------------------------------
    def _encoder(self, inputs, train, enc_drop_masks, trace):
        cfg, store = self.config, self.store
        # embedding lookup fused with the first pre-net layer (tacotron.py:111-114, :128)
        with ops.variable_scope(store, "enc"):
            pre_out = self.pre_net(store["embedding"], dropout=cfg.char_dropout_prob, train=train, masks=enc_drop_masks,
                                   ids=inputs["text"])
            encoded = ops.CBHG(pre_out, None, K=16, c=[128, 128, 128], gru_units=128, trace=trace)      # :131
        if trace is not None:
            trace["enc/prenet_out"] = pre_out
            trace["encoded"] = encoded
        return encoded

    def _postnet(self, seq2seq_output, trace):
        cfg, store = self.config, self.store
        B = seq2seq_output.shape[0]
        # post-processing CBHG + linear-spectrogram dense (:144-151); the reshapes are views
        with ops.variable_scope(store, "post"):
            post_input = seq2seq_output.view(B, -1, cfg.mel_features)
            post = ops.CBHG(post_input, None, K=8, c=[128, 256, 80], gru_units=128, trace=trace)
            W, b = store["post/dense/W"], store["post/dense/b"]
            rt = self.runtime
            dense = ops.linear(rt, post, W, ops.packed_weight(rt, "post/dense/W", W, 1, 256, cfg.fft_size), cfg.fft_size,
                               bias=b, tag="post/dense/out")
        return dense.view(B, -1, cfg.fft_size * cfg.r)

    def inference(self, inputs, train=True, T=None, enc_drop_masks=None, dec_drop_masks=None, sample_mask=None,
                  trace=None):
        """inputs: dict of CUDA tensors: 'text' int32 [B,Tx], 'text_length' int32 [B]; train adds
        'mel' fp32 [B,T,80r] (and 'stft' for the loss).  Returns (seq2seq_output, output).
        With config.cuda_graph the three sections (encoder / decoder / post-net) of an inference-mode call are
        captured once per input shape and replayed: ~26 launches collapse into 3 graph launches."""
        text = inputs["text"]
        assert text.dtype == torch.int32 and text.is_cuda and text.is_contiguous()
        B, Tx = text.shape
        if not 1 <= Tx <= 256:
            raise ValueError(f"text width {Tx} out of range: the persistent decoder keeps keys/values in shared memory, Tx <= 256")
        if getattr(self.config, "cuda_graph", False) and not train and trace is None:
            return self._inference_graphed(inputs, T)
        self._mark("start")
        encoded = self._encoder(inputs, train, enc_drop_masks, trace)
        self._mark("encoder")
        # attention decoder (:135-138)
        dec = self.create_decoder(encoded, inputs, None, train, T=T, dec_drop_masks=dec_drop_masks, sample_mask=sample_mask)
        seq2seq_output, self.alignments = dec()
        self._mark("decoder")
        output = self._postnet(seq2seq_output, trace)
        self._mark("postnet")
        return seq2seq_output, output

    def _inference_graphed(self, inputs, T):
        T = self.config.max_decode_iter if T is None else T
        key = (tuple(inputs["text"].shape), T, self.store.version)
        g = self._graphs.get(key) if hasattr(self, "_graphs") else None
        if g is None:
            if not hasattr(self, "_graphs"):
                self._graphs = {}
            static = {"text": inputs["text"].clone(), "text_length": inputs["text_length"].clone()}
            marks, self._marks = self._marks, None
            cg_flag, self.config.cuda_graph = self.config.cuda_graph, False
            try:
                for _ in range(2):                                  # eager warm-up: allocates every scratch buffer, packs weights
                    self.inference(static, train=False, T=T)
                torch.cuda.synchronize()
                graphs, outs = [], {}
                n_launch0 = L.lib().taco_launch_count()
                pool = torch.cuda.graph_pool_handle()
                ge = torch.cuda.CUDAGraph()
                with torch.cuda.graph(ge, pool=pool):
                    outs["encoded"] = self._encoder(static, False, None, None)
                gd = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gd, pool=pool):
                    dec = self.create_decoder(outs["encoded"], static, None, False, T=T)
                    outs["y"], outs["align"] = dec()
                gp = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gp, pool=pool):
                    outs["out"] = self._postnet(outs["y"], None)
                graphs = [ge, gd, gp]
            finally:
                self._marks = marks
                self.config.cuda_graph = cg_flag
            g = (static, graphs, outs, int(L.lib().taco_launch_count() - n_launch0))
            self._graphs[key] = g
        static, graphs, outs, _ = g
        static["text"].copy_(inputs["text"], non_blocking=True)
        static["text_length"].copy_(inputs["text_length"], non_blocking=True)
        # optional events a pipelined caller sets so that a section does not overwrite a result buffer that an
        # asynchronous device->host copy of the PREVIOUS call is still reading (bench.py e2e)
        waits = getattr(self, "section_wait", None) or {}
        cur = torch.cuda.current_stream()
        self._mark("start")
        graphs[0].replay(); self._mark("encoder")
        if waits.get("decoder") is not None:
            cur.wait_event(waits["decoder"])
        graphs[1].replay(); self._mark("decoder")
        if waits.get("postnet") is not None:
            cur.wait_event(waits["postnet"])
        graphs[2].replay(); self._mark("postnet")
        self.alignments = outs["align"]
        self.last_graph_kernels = g[3]
        return outs["y"], outs["out"]

    # -- add_loss_op (tacotron.py:156-165) -------------------------------------------------------
    def add_loss_op(self, seq2seq_output, output, mel, linear):
        lib = L.lib()
        rt = self.runtime
        part = rt.buf("loss/partial", (lib.taco_l1_partial_count(),))
        res = rt.buf("loss/out", (2,))
        L.check(lib.taco_l1_loss_fwd(L.ptr(seq2seq_output), L.ptr(mel), seq2seq_output.numel(), L.ptr(part),
                                     L.ptr(res[0:1]), L.current_stream()), "taco_l1_loss_fwd")
        part2 = rt.buf("loss/partial2", (lib.taco_l1_partial_count(),))
        L.check(lib.taco_l1_loss_fwd(L.ptr(output), L.ptr(linear), output.numel(), L.ptr(part2), L.ptr(res[1:2]),
                                     L.current_stream()), "taco_l1_loss_fwd")
        self.seq2seq_loss = res[0]
        self.output_loss = res[1]
        return res[0] + res[1]

    # -- add_train_op (tacotron.py:167-185) -------------------------------------------------------
    def add_train_op(self, loss=None):
        """Creates the optimizer state (Adam moments + flat gradient bucket) and returns the callable that plays the
        role of `train_op`: ``train_op(inputs, lr)`` = one ``sess.run([train_op, loss, global_step], {lr})``."""
        from ..optim import FlatAdam
        if getattr(self, "_opt", None) is None:
            self._opt = FlatAdam(self.store.flat)
            self._gviews = self._opt.views(self.store)
        return self.train_step

    def backward(self, S):
        """d(loss)/d(parameters) into the flat gradient bucket (zeroed first): the hand-written reverse of
        inference + add_loss_op (models/grad.py) over the saved activations S of a train-mode forward."""
        from .. import kernels as K
        from . import grad
        self.add_train_op()
        self._opt.zero_grad()
        # GEMM kernel of the backward follows the precision mode: 'tf32' -> 3xTF32 mma.sync tensor cores (fp32-grade,
        # ~1e-6 relative), 'fp32' -> exact-product FFMA.  `self.gemm_impl` (0 / 1) overrides.
        impl = getattr(self, "gemm_impl", None)
        if impl is None:
            impl = 0 if self.config.precision == "fp32" else 1
        prev = K.set_gemm_impl(impl)
        # data gradients of the dense / conv layers run on the tcgen05 forward kernel (a conv data gradient IS a conv):
        # 3xTF32 (fp32-grade) in 'fp32x3' mode, single-pass TF32 in 'tf32' mode; 'fp32' mode keeps the exact-product GEMM.
        # Config.grad_dx_tc overrides (True / False).
        dx_tc = getattr(self.config, "grad_dx_tc", None)
        if dx_tc is None:
            dx_tc = self.config.precision != "fp32"
        prev_dx, K.DX_TC = K.DX_TC, bool(dx_tc)
        prev_impl, K.DX_TC_IMPL = K.DX_TC_IMPL, (L.IMPL_TC if self.config.precision == "tf32" else L.IMPL_TC3)
        # weight gradients of the dense / conv layers: tcgen05 3xTF32 kernel (taco_conv_dw), same switch; Config.grad_dw_tc overrides
        dw_tc = getattr(self.config, "grad_dw_tc", None)
        prev_dw, K.DW_TC = K.DW_TC, bool(dx_tc if dw_tc is None else dw_tc)
        prev_g, K.GEMM_TC = K.GEMM_TC, bool(dx_tc)            # recomputation / data-gradient products on the same kernel
        try:
            grad.model_bwd(K, self.store, self._gviews, S, self.config)
        finally:
            K.set_gemm_impl(prev)
            K.DX_TC = prev_dx
            K.DX_TC_IMPL = prev_impl
            K.DW_TC = prev_dw
            K.GEMM_TC = prev_g
        return self._gviews

    def train_step(self, inputs, lr=None, **kw):
        """forward (train mode, activations saved) -> loss -> backward -> [all-reduce] -> clip -> Adam.
        Returns the loss tensor (device scalar); `self.global_step`, `self.grad_sumsq` are updated."""
        from .. import kernels as K
        self.add_train_op()
        lr = self.lr if lr is None else lr
        S = {}
        with ops.saving(S):
            self.seq2seq_output, self.output = self.inference(inputs, True, **kw)
        self.loss = self.add_loss_op(self.seq2seq_output, self.output, inputs["mel"], inputs["stft"])
        S.update(text=inputs["text"], text_length=inputs["text_length"], mel=inputs["mel"], stft=inputs["stft"])
        S["post/out"] = self.output
        self.saved = S
        self.backward(S)
        # data parallel: one SUM all-reduce of the flat gradient bucket when a process group exists (self.dp = False
        # keeps the step local, e.g. for a per-rank health check before the first collective)
        self.grad_sumsq = self._opt.apply(K, self.store.flat, lr, self.config.cap_grads, allreduce=getattr(self, "dp", True))
        self.store.version += 1                    # derived kernel-layout buffers are refreshed on the next forward
        self.global_step += 1
        return self.loss

    # -- the eager stand-in for sess.run([...]) ---------------------------------------------------
    def __call__(self, inputs, **kw):
        self.seq2seq_output, self.output = self.inference(inputs, self.train, **kw)
        if self.train and "stft" in inputs:
            self.loss = self.add_loss_op(self.seq2seq_output, self.output, inputs["mel"], inputs["stft"])
        return self.seq2seq_output, self.output
