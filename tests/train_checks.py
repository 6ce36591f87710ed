"""GPU checks of the training path, shared by tests/test_gpu_train.py (asserting) and tests/tools/gpu_train_diag.py
(reporting).  Every check runs a CUDA entry point (tacotron_b200/kernels.py) and its torch-CPU mirror
(tests/mirror_kernels.py) on identical inputs and returns {tensor name: (max abs err, max |ref|)}.
"""
from __future__ import annotations

import copy
import math

import torch

from oracle import tacotron_oracle as O
from tests import grad_util, mirror_kernels as MK


def _K():
    from tacotron_b200 import kernels
    return kernels


def _g(seed=0):
    return torch.Generator().manual_seed(seed)


def _rn(g, *shape):
    return torch.randn(*shape, generator=g)


def _cmp(out, name, got, ref):
    got = got.detach().float().cpu()
    ref = ref.detach().float().cpu()
    out[name] = ((got - ref).abs().max().item() if ref.numel() else 0.0, ref.abs().max().item() if ref.numel() else 0.0)


def _to_cuda(x):
    if isinstance(x, torch.Tensor):
        return x.cuda()
    if isinstance(x, (list, tuple)):
        return type(x)(_to_cuda(v) for v in x)
    if isinstance(x, dict):
        return {k: _to_cuda(v) for k, v in x.items()}
    return x


# ----------------------------------------------------------------------------------------------------------------
# element-wise / reduction / gemm kernels
# ----------------------------------------------------------------------------------------------------------------
def check_gemm():
    K = _K()
    g = _g(1)
    res = {}

    def run(name, Cs, As, Bs, cview=None, aview=None, bview=None, **kw):
        """Cs/As/Bs: CPU base tensors; *view: function base -> 2-D view passed to gemm"""
        cview = cview or (lambda t: t); aview = aview or (lambda t: t); bview = bview or (lambda t: t)
        Cc, Ac, Bc = Cs.clone(), As.clone(), Bs.clone()
        Cg, Ag, Bg = Cs.cuda(), As.cuda(), Bs.cuda()
        MK.gemm(cview(Cc), aview(Ac), bview(Bc), **kw)
        K.gemm(cview(Cg), aview(Ag), bview(Bg), **kw)
        torch.cuda.synchronize()
        _cmp(res, name, Cg, Cc)

    M, N, Kd = 70, 45, 37
    run("plain", torch.zeros(M, N), _rn(g, M, Kd), _rn(g, Kd, N))
    run("beta1", _rn(g, M, N), _rn(g, M, Kd), _rn(g, Kd, N), beta=1.0)
    run("tb", torch.zeros(M, N), _rn(g, M, Kd), _rn(g, N, Kd), tb=True)
    run("ta", torch.zeros(M, N), _rn(g, Kd, M), _rn(g, Kd, N), ta=True)
    run("ta_tb_beta1", _rn(g, M, N), _rn(g, Kd, M), _rn(g, N, Kd), ta=True, tb=True, beta=1.0)
    # strided views (column slices of wider buffers)
    run("strided", torch.zeros(M, 2 * N), _rn(g, M, 3 * Kd), _rn(g, Kd, 2 * N), cview=lambda t: t[:, N:], aview=lambda t: t[:, Kd:2 * Kd],
        bview=lambda t: t[:, :N])
    # row shift inside periods (h(t-1) products): 6 blocks of 12 rows
    run("shift-1", torch.zeros(72, N), _rn(g, 72, Kd), _rn(g, Kd, N), shift=-1, period=12)
    run("shift+1", torch.zeros(72, N), _rn(g, 72, Kd), _rn(g, Kd, N), shift=1, period=12)
    run("shift-B", _rn(g, 72, N), _rn(g, 72, Kd), _rn(g, Kd, N), shift=-8, beta=1.0)
    run("ta_shift", _rn(g, 20, N), _rn(g, 72, 20), _rn(g, 72, N), ta=True, shift=-1, period=12, beta=1.0)
    # conv data gradient: taps K-segments, B stored [taps][Cin][Cout]
    taps, Cin, Cout, Bt, T = 3, 24, 16, 4, 9
    W = _rn(g, taps, Cin, Cout)
    run("conv_dx", torch.zeros(Bt * T, Cin), _rn(g, Bt * T, Cout), W.reshape(taps * Cin, Cout), bview=lambda t: t[:Cin], tb=True, shift=1,
        dshift=-1, kper=Cout, taps=taps, b_tap_stride=Cin * Cout, period=T)
    # conv weight gradient: batch over taps, shifted K index, large K (split-K with atomics)
    Bt, T = 16, 100
    run("conv_dw", _rn(g, taps * Cin, Cout), _rn(g, Bt * T, Cin), _rn(g, Bt * T, Cout), cview=lambda t: t[:Cin], ta=True, beta=1.0, shift=-1,
        bshift=1, batch=taps, c_bstride=Cin * Cout, period=T)
    # even filter width (asymmetric 'same' padding): taps=4 -> tap0=-1
    taps = 4
    W = _rn(g, taps, Cin, Cout)
    run("conv_dx_k4", torch.zeros(Bt * T, Cin), _rn(g, Bt * T, Cout), W.reshape(taps * Cin, Cout), bview=lambda t: t[:Cin], tb=True, shift=1,
        dshift=-1, kper=Cout, taps=taps, b_tap_stride=Cin * Cout, period=T)
    # batched per-utterance products (attention context and its transpose)
    Bq, Tq, Tx = 3, 7, 12
    al, va = _rn(g, Bq, Tq, Tx), _rn(g, Bq, Tx, 256)
    run("ctx_batched", torch.zeros(Bq, Tq, 256), al, va, cview=lambda t: t[0], aview=lambda t: t[0], bview=lambda t: t[0], batch=Bq,
        a_bstride=Tq * Tx, b_bstride=Tx * 256, c_bstride=Tq * 256)
    dctx = _rn(g, Tq, Bq, 256)
    run("dvalues_batched", torch.zeros(Bq, Tx, 256), al, dctx, cview=lambda t: t[0], aview=lambda t: t[0], bview=lambda t: t[:, 0], ta=True,
        batch=Bq, a_bstride=Tq * Tx, b_bstride=256, c_bstride=Tx * 256)
    # big weight gradient (rows = 4096) -> many splits
    run("big_dw", _rn(g, 256, 128), _rn(g, 4096, 256), _rn(g, 4096, 128), ta=True, beta=1.0)
    return res


def check_elementwise():
    K = _K()
    g = _g(2)
    res = {}
    M, N = 77, 96
    # colsum (3 forms)
    A, Bm, R = _rn(g, M, N), _rn(g, M, N), _rn(g, M, N)
    for name, args in (("colsum", (A,)), ("colsum_mul", (A, Bm)), ("colsum_mul_sub", (A, Bm, R))):
        oc = _rn(g, N); og = oc.cuda()
        MK.colsum(oc, *args)
        K.colsum(og, *[a.cuda() for a in args])
        _cmp(res, name, og, oc)
    A5 = _rn(g, 5000, 40); oc = torch.zeros(40); og = oc.cuda()
    MK.colsum(oc, A5); K.colsum(og, A5.cuda()); _cmp(res, "colsum_tall", og, oc)
    # bias_act
    for act in (0, 1, 2, 3):
        Cc = _rn(g, M, N); Cg = Cc.cuda(); b = _rn(g, N)
        MK.bias_act_(Cc, b, act); K.bias_act_(Cg, b.cuda(), act); _cmp(res, f"bias_act{act}", Cg, Cc)
    Cc = _rn(g, M, 2 * N); Cg = Cc.cuda()
    MK.bias_act_(Cc[:, N:], None, 2); K.bias_act_(Cg[:, N:], None, 2); _cmp(res, "bias_act_strided", Cg, Cc)
    # mul_shift
    X, Hm = _rn(g, 72, N), _rn(g, 72, N)
    for sh, per in ((-1, 12), (1, 12), (-8, 0)):
        oc = torch.zeros(72, N); og = oc.cuda()
        MK.mul_shift(oc, X, Hm, sh, per); K.mul_shift(og, X.cuda(), Hm.cuda(), sh, per); _cmp(res, f"mul_shift{sh}", og, oc)
    # epi_bwd
    dY, Y = _rn(g, M, N), torch.relu(_rn(g, M, N))
    scale, shift = _rn(g, N), _rn(g, N)
    Ybn = Y * scale + shift
    Rr = _rn(g, M, N)
    cases = {"relu_gain": dict(Y=Y, relu=True, gain=2.0), "relu_bn": dict(Y=Ybn, relu=True, scale=scale, shift=shift),
             "bn_only": dict(Y=Ybn + Rr, relu=False, scale=scale, shift=shift, R=Rr),
             "relu_bn_res": dict(Y=Ybn + Rr, relu=True, scale=scale, shift=shift, R=Rr)}
    for name, kw in cases.items():
        oc = torch.zeros(M, N); og = oc.cuda()
        MK.epi_bwd(oc, dY, kw["Y"], kw["relu"], **{k: v for k, v in kw.items() if k not in ("Y", "relu")})
        K.epi_bwd(og, dY.cuda(), kw["Y"].cuda(), kw["relu"], **{k: (v.cuda() if isinstance(v, torch.Tensor) else v) for k, v in kw.items()
                                                                  if k not in ("Y", "relu")})
        _cmp(res, f"epi_bwd_{name}", og, oc)
    # dropout apply
    Xc = _rn(g, M, N); Xg = Xc.cuda(); keep = (torch.rand(M, N, generator=g) > 0.5).to(torch.uint8)
    MK.epi_fwd_keep_(Xc, keep, 2.0); K.epi_fwd_keep_(Xg, keep.cuda(), 2.0); _cmp(res, "epi_fwd_keep", Xg, Xc)
    # bn_param_grad
    S1, S2, gam, bet = _rn(g, N), _rn(g, N), torch.rand(N, generator=g) + 0.5, _rn(g, N)
    dgc, dbc = _rn(g, N), _rn(g, N); dgg, dbg = dgc.cuda(), dbc.cuda()
    MK.bn_param_grad(dgc, dbc, S1, S2, gam, bet); K.bn_param_grad(dgg, dbg, S1.cuda(), S2.cuda(), gam.cuda(), bet.cuda())
    _cmp(res, "bn_dgamma", dgg, dgc); _cmp(res, "bn_dbeta", dbg, dbc)
    # maxpool backward (with exact ties from relu zeros)
    X3 = torch.relu(_rn(g, 3, 11, 32)); dP = _rn(g, 3, 11, 32)
    dc = torch.zeros_like(X3); dg_ = dc.cuda()
    MK.maxpool_bwd(dc, dP, X3); K.maxpool_bwd(dg_, dP.cuda(), X3.cuda()); _cmp(res, "maxpool_bwd", dg_, dc)
    X1 = _rn(g, 2, 1, 8); dP1 = _rn(g, 2, 1, 8); dc = torch.zeros_like(X1); dg_ = dc.cuda()
    MK.maxpool_bwd(dc, dP1, X1); K.maxpool_bwd(dg_, dP1.cuda(), X1.cuda()); _cmp(res, "maxpool_bwd_T1", dg_, dc)
    # highway
    U = 128
    Pm, Xh, dYh = _rn(g, M, 2 * U), _rn(g, M, U), _rn(g, M, U)
    yc = torch.zeros(M, U); yg = yc.cuda()
    MK.highway_fwd(yc, Pm, Xh); K.highway_fwd(yg, Pm.cuda(), Xh.cuda()); _cmp(res, "highway_fwd", yg, yc)
    dPc, dXc = torch.zeros(M, 2 * U), torch.zeros(M, U); dPg, dXg = dPc.cuda(), dXc.cuda()
    MK.highway_bwd(dPc, dXc, dYh, Pm, Xh); K.highway_bwd(dPg, dXg, dYh.cuda(), Pm.cuda(), Xh.cuda())
    _cmp(res, "highway_bwd_dP", dPg, dPc); _cmp(res, "highway_bwd_dX", dXg, dXc)
    # l1 backward
    A1, B1 = _rn(g, 1000).half().float(), _rn(g, 1000).half().float()
    B1[:10] = A1[:10]
    for beta in (0.0, 1.0):
        dc = _rn(g, 1000); dg_ = dc.cuda()
        MK.l1_bwd(dc, A1, B1, beta); K.l1_bwd(dg_, A1.cuda(), B1.cuda(), beta); _cmp(res, f"l1_bwd_beta{int(beta)}", dg_, dc)
    # embedding scatter-add
    ids = torch.randint(0, 20, (6, 9), generator=g, dtype=torch.int32); rows = _rn(g, 54, 64)
    tc = torch.zeros(20, 64); tg = tc.cuda()
    MK.scatter_add_rows(tc, ids, rows); K.scatter_add_rows(tg, ids.cuda(), rows.cuda()); _cmp(res, "scatter_add_rows", tg, tc)
    # decoder inputs
    for r in (2, 5):
        Bq, Tq = 3, 6
        mel, y = _rn(g, Bq, Tq, 80 * r), _rn(g, Bq, Tq, 80 * r)
        sm = (torch.rand(Tq, Bq, generator=g) < 0.5).to(torch.uint8)
        for sched in (False, True):
            xc, sc = torch.zeros(Tq, Bq, 80), torch.zeros(Tq, Bq, dtype=torch.uint8)
            xg, sg = xc.cuda(), sc.cuda()
            MK.dec_inputs(xc, sc, mel, y, sm if sched else None, r, sched)
            K.dec_inputs(xg, sg, mel.cuda(), y.cuda(), sm.cuda() if sched else None, r, sched)
            _cmp(res, f"dec_inputs_r{r}_s{int(sched)}", xg, xc); _cmp(res, f"dec_sel_r{r}_s{int(sched)}", sg, sc)
    # attention post pass
    Bq, Tq, Tx = 2, 5, 12
    ds = _rn(g, Bq, Tq, Tx); ds[:, :, -3:] = 0
    keys, PQ, v = _rn(g, Bq, Tx, 256), _rn(g, Tq, Bq, 256), _rn(g, 256)
    dkc, dvc = torch.zeros(Bq, Tx, 256), _rn(g, 256); dkg, dvg = dkc.cuda(), dvc.cuda()
    MK.attn_bwd_post(dkc, dvc, ds, keys, PQ, v); K.attn_bwd_post(dkg, dvg, ds.cuda(), keys.cuda(), PQ.cuda(), v.cuda())
    _cmp(res, "attn_post_dkeys", dkg, dkc); _cmp(res, "attn_post_dv", dvg, dvc)
    # mask rows
    src = _rn(g, 3, 8, 256); ln = torch.tensor([8, 3, 5], dtype=torch.int32)
    mc = torch.zeros_like(src); mg = mc.cuda()
    MK.mask_rows(mc, src, ln); K.mask_rows(mg, src.cuda(), ln.cuda()); _cmp(res, "mask_rows", mg, mc)
    # sumsq + adam (clip active and inactive)
    n = 100003
    for clip, tag in ((5.0, "clip"), (1e9, "noclip")):
        p, gr, m, v2 = _rn(g, n), _rn(g, n), _rn(g, n) * 0.1, torch.rand(n, generator=g) * 0.01
        pc, mc, vc = p.clone(), m.clone(), v2.clone(); pg, mg, vg, gg = p.cuda(), m.cuda(), v2.cuda(), gr.cuda()
        ssc = torch.zeros(1); ssg = ssc.cuda()
        MK.sumsq(ssc, gr); K.sumsq(ssg, gg); _cmp(res, f"sumsq_{tag}", ssg, ssc)
        MK.adam_step(pc, gr, mc, vc, 1e-3, 0.9, 0.999, 1e-8, clip, ssc)
        K.adam_step(pg, gg, mg, vg, 1e-3, 0.9, 0.999, 1e-8, clip, ssg)
        _cmp(res, f"adam_p_{tag}", pg, pc); _cmp(res, f"adam_m_{tag}", mg, mc); _cmp(res, f"adam_v_{tag}", vg, vc)
    torch.cuda.synchronize()
    return res


# ----------------------------------------------------------------------------------------------------------------
# recurrent backward kernels
# ----------------------------------------------------------------------------------------------------------------
def check_bigru_bwd(B=3, T=9, seed=3):
    K = _K()
    g = _g(seed)
    out = torch.tanh(_rn(g, B, T, 256))
    ACT = torch.rand(B, T, 768, generator=g)
    ACT[:, :, 256:384] = ACT[:, :, 256:384] * 2 - 1
    ACT[:, :, 640:768] = ACT[:, :, 640:768] * 2 - 1
    dOut = _rn(g, B, T, 256)
    W = [_rn(g, 128, 256) * 0.1, _rn(g, 128, 128) * 0.1, _rn(g, 128, 256) * 0.1, _rn(g, 128, 128) * 0.1]
    dc = torch.zeros(B, T, 768); dg_ = dc.cuda()
    MK.bigru_bwd(dc, dOut, out, ACT, *W)
    K.bigru_bwd(dg_, dOut.cuda(), out.cuda(), ACT.cuda(), *[w.cuda() for w in W])
    torch.cuda.synchronize()
    res = {}
    for d, dn in ((0, "fw"), (1, "bw")):
        for nm, sl in (("dr", slice(0, 128)), ("du", slice(128, 256)), ("dc", slice(256, 384))):
            _cmp(res, f"{dn}_{nm}", dg_[:, :, d * 384:(d + 1) * 384][:, :, sl], dc[:, :, d * 384:(d + 1) * 384][:, :, sl])
    return res


def _small_case(r, sched, B=2, Tx=8, T=5, dtype=torch.float32, ragged=True, seed=0):
    cfg = O.OracleConfig(r=r, vocab_size=20)
    p = O.init_params(cfg, seed=1, trained_like=True, dtype=dtype)
    inp = O.synthetic_inputs(cfg, B, Tx, T, seed=seed, ragged=ragged)
    inp = {k: (v.to(dtype) if v.dtype.is_floating_point else v) for k, v in inp.items()}
    enc_m, dec_m = O.dropout_masks(cfg, B, Tx, T, seed=2)
    sm = O.sched_mask(cfg, B, T, seed=3) if sched else None
    return cfg, p, inp, enc_m, dec_m, sm


def check_decoder_bwd(r=2, sched=True, B=2, Tx=8, T=5):
    """the serial decoder-backward kernel against its mirror, on arguments produced by the CPU host path"""
    from tacotron_b200.models import grad
    K = _K()
    cfg, p, inp, enc_m, dec_m, sm = _small_case(r, sched, B, Tx, T)
    S, y, out = grad_util.saving_forward(p, inp, cfg, enc_m, dec_m, sm)
    S["_capture"] = {}
    G = {k: torch.zeros_like(v) for k, v in p.items()}
    dY = torch.randn(B, T, 80 * r, generator=_g(5))
    grad.decoder_bwd(MK, p, G, S, cfg, dY)
    a_cpu = S["_capture"]["decoder_bwd"]                     # inputs + mirror outputs
    outs = ("DATT", "DY", "DPQ", "DSCORE", "DCTX", "DZ", "DPN2", "DPN1", "DX")
    a_gpu = {}
    for k, v in a_cpu.items():
        if k.startswith("_"):
            continue
        a_gpu[k] = _to_cuda(copy.deepcopy(v))
    for k in outs:                                            # poison the outputs so that unwritten cells show up
        a_gpu[k].fill_(float("nan"))
    a_gpu["DATT"][:T - 1].fill_(float("nan"))
    for i in range(3):
        a_gpu["DG"][i].fill_(float("nan")); a_gpu["DC"][i].fill_(float("nan"))
    K.decoder_bwd(a_gpu)
    torch.cuda.synchronize()
    res = {}
    for k in outs:
        got = torch.nan_to_num(a_gpu[k], nan=1e30)
        _cmp(res, k, got, a_cpu[k])
        if k not in ("DSCORE",) and T > 1:                   # time-major: also report the LAST step alone (first one computed)
            _cmp(res, f"{k}[T-1]", got[T - 1], a_cpu[k][T - 1])
    for i in range(3):
        _cmp(res, f"DG{i}", torch.nan_to_num(a_gpu["DG"][i], nan=1e30), a_cpu["DG"][i])
        _cmp(res, f"DC{i}", torch.nan_to_num(a_gpu["DC"][i], nan=1e30), a_cpu["DC"][i])
    return res


# ----------------------------------------------------------------------------------------------------------------
# whole model
# ----------------------------------------------------------------------------------------------------------------
def _model(cfg_o, p, precision):
    from tacotron_b200.models.tacotron import Config, Tacotron
    cfg = Config(r=cfg_o.r, vocab_size=cfg_o.vocab_size, precision=precision, scheduled_sample=cfg_o.scheduled_sample)
    m = Tacotron(cfg, None, train=True)
    m.load_params(p)
    return m


def check_train_forward(r=2, sched=True, precision="fp32", B=2, Tx=8, T=5):
    """train-mode CUDA forward (un-fused pool, highway pre-activations, decoder state dump) against the oracle's"""
    from tacotron_b200.models import ops
    cfg, p, inp, enc_m, dec_m, sm = _small_case(r, sched, B, Tx, T)
    if not sched:
        cfg.scheduled_sample = 0.0
    Sref, y, out = grad_util.saving_forward(p, inp, cfg, enc_m, dec_m, sm)
    m = _model(cfg, p, precision)
    gi = {k: v.cuda() for k, v in inp.items()}
    S = {}
    with ops.saving(S):
        yg, og = m.inference(gi, True, enc_drop_masks=_to_cuda(enc_m), dec_drop_masks=_to_cuda(dec_m),
                             sample_mask=sm.cuda() if sm is not None else None)
    torch.cuda.synchronize()
    res = {}
    _cmp(res, "y", yg, y); _cmp(res, "out", og, out)
    for k, v in Sref.items():
        if isinstance(v, torch.Tensor) and v.dtype.is_floating_point and k in S:
            _cmp(res, k, S[k].reshape(v.shape), v)
    missing = [k for k, v in Sref.items() if isinstance(v, torch.Tensor) and v.dtype.is_floating_point and k not in S
               and k not in ("mel", "stft", "post/out")]
    res["_missing"] = (float(len(missing)), 0.0)
    if missing:
        res["_missing_names:" + ",".join(missing[:6])] = (float(len(missing)), 0.0)
    return res


def check_model_bwd(r=2, sched=True, precision="fp32", B=2, Tx=8, T=5, gemm_impl=None):
    """all parameter gradients of the CUDA path against torch.autograd over the oracle
    (gemm_impl: None = the model's choice by precision, 0 = FFMA, 1 = 3xTF32 mma.sync)"""
    cfg, p, inp, enc_m, dec_m, sm = _small_case(r, sched, B, Tx, T)
    if not sched:
        cfg.scheduled_sample = 0.0
    _, g_ref = O.loss_and_grads(p, inp, cfg, enc_drop_masks=enc_m, dec_drop_masks=dec_m, sample_mask=sm)
    from tacotron_b200.models import ops
    m = _model(cfg, p, precision)
    m.gemm_impl = gemm_impl
    gi = {k: v.cuda() for k, v in inp.items()}
    S = {}
    with ops.saving(S):
        m.seq2seq_output, m.output = m.inference(gi, True, enc_drop_masks=_to_cuda(enc_m), dec_drop_masks=_to_cuda(dec_m),
                                                 sample_mask=sm.cuda() if sm is not None else None)
    S.update(text=gi["text"], text_length=gi["text_length"], mel=gi["mel"], stft=gi["stft"])
    S["post/out"] = m.output
    G = m.backward(S)
    torch.cuda.synchronize()
    res = {}
    for k, gr in g_ref.items():
        _cmp(res, k, G[k], gr)
    # whole-gradient agreement (what the optimizer sees): relative L2 error and cosine over all parameters
    a = torch.cat([G[k].detach().float().cpu().reshape(-1) for k in g_ref])
    b = torch.cat([g_ref[k].float().reshape(-1) for k in g_ref])
    res["_rel_l2"] = (((a - b).norm() / b.norm()).item(), 1.0)
    res["_cosine"] = (torch.nn.functional.cosine_similarity(a, b, dim=0).item(), 1.0)
    return res


def check_train_step(r=2, sched=True, precision="fp32", steps=2, B=2, Tx=8, T=5):
    """parameters after `steps` optimizer steps against oracle.train_step"""
    cfg, p, inp, enc_m, dec_m, sm = _small_case(r, sched, B, Tx, T)
    if not sched:
        cfg.scheduled_sample = 0.0
    p_ref = {k: v.clone() for k, v in p.items()}
    m_ref = {k: torch.zeros_like(v) for k, v in p.items()}
    v_ref = {k: torch.zeros_like(v) for k, v in p.items()}
    m = _model(cfg, p, precision)
    gi = {k: v.cuda() for k, v in inp.items()}
    res = {}
    for s in range(1, steps + 1):
        loss_ref, gn_ref = O.train_step(p_ref, m_ref, v_ref, inp, cfg, lr=1e-3, step=s, enc_drop_masks=enc_m, dec_drop_masks=dec_m,
                                        sample_mask=sm)
        loss = m.train_step(gi, lr=1e-3, enc_drop_masks=_to_cuda(enc_m), dec_drop_masks=_to_cuda(dec_m),
                            sample_mask=sm.cuda() if sm is not None else None)
        torch.cuda.synchronize()
        _cmp(res, f"loss@{s}", loss.reshape(1), loss_ref.reshape(1))
        _cmp(res, f"gnorm@{s}", m.grad_sumsq.sqrt().reshape(1), torch.as_tensor(gn_ref).reshape(1))
    worst = (0.0, 0.0, "")
    for k in p:
        e = (m.store[k].cpu() - p_ref[k]).abs().max().item()
        if e > worst[0]:
            worst = (e, p_ref[k].abs().max().item(), k)
    res[f"param_worst:{worst[2]}"] = (worst[0], worst[1])
    return res


def fmt(res):
    return "\n".join(f"  {k:34s} err {e:.3e}  ref {r:.3e}" for k, (e, r) in res.items())


def worst_rel(res, floor=1e-6):
    """max over tensors of err / (|ref| + floor)"""
    w = 0.0
    for k, (e, r) in res.items():
        if k.startswith("_"):
            continue
        w = max(w, e / (r + floor))
    return w
