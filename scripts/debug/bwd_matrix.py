import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from tests import train_checks as TC
import tacotron_b200.models.tacotron as TM
for prec in ("fp32", "fp32x3"):
    for impl in (0, 1):
        for dx in (False, True):
            TM.Config.grad_dx_tc = dx
            res = TC.check_model_bwd(2, True, prec, gemm_impl=impl)
            worst = max(((v[0] / (v[1] + 1e-3), k) for k, v in res.items() if not k.startswith("_")))
            print(prec, "gemm_impl", impl, "dx_tc", dx, "rel_l2 %.2e" % res["_rel_l2"][0], "worst %.2e %s" % worst)
