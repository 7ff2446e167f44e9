"""CPU: the drop-in mechanism.  With `dropin/` first on PYTHONPATH the import lines of the unmodified trainers
(train_cardiac_uda.py:27-31, train_camus_echo.py:32-36) resolve to graphecho_b200; with the reference tree behind it,
the modules this build does not replace (utils.lr_scheduler, utils.tools) still resolve from the reference."""
import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent

PROBE = r'''
import sys, json
from models.fpnseg import FPN, Discriminator
from models.graph_matching import GModule
from models.TGCN import TGCN
from utils.sinkhorn_distance import SinkhornDistance
from utils.losses import DiceLoss
import models.vig, models.affinity_layer, models.transformer, models.gradient_reversal
out = {c.__name__: sys.modules[c.__module__].__file__ for c in (FPN, Discriminator, GModule, TGCN, SinkhornDistance, DiceLoss)}
out["vig.Grapher"] = sys.modules[models.vig.Grapher.__module__].__file__
out["Affinity"] = sys.modules[models.affinity_layer.Affinity.__module__].__file__
out["MultiHeadAttention"] = sys.modules[models.transformer.MultiHeadAttention.__module__].__file__
try:
    import utils.lr_scheduler as lrs
    out["lr_scheduler"] = lrs.__file__
except ImportError:
    out["lr_scheduler"] = None
# the trainer's constructor calls (train_cardiac_uda.py:73, 82, 89, 120, 138) work on the replacements
net = FPN([2, 4, 23, 3], num_classes=3, in_channel=1, back_bone="VGG16")
gm = GModule(in_channels=256, num_classes=3, device="cpu")
dis = Discriminator(grad_reverse_lambda=0.02)
tg = TGCN(input_dim=256, hidden_dim=256, clip_shape=(8, 8, 8), soucre_class=10, target_class=10)
sk = SinkhornDistance(eps=0.1, max_iter=5, reduction='mean')
out["n_params"] = sum(p.numel() for p in net.parameters())
print(json.dumps(out))
'''


def _run(pythonpath):
    env = dict(os.environ, PYTHONPATH=os.pathsep.join(pythonpath))
    res = subprocess.run([sys.executable, "-c", PROBE], capture_output=True, text=True, env=env, cwd="/tmp", timeout=300)
    assert res.returncode == 0, res.stderr[-2000:]
    import json
    return json.loads(res.stdout.strip().splitlines()[-1])


def test_trainer_imports_resolve_to_this_build():
    out = _run([str(ROOT / "dropin"), str(ROOT)])
    for name in ("FPN", "Discriminator", "GModule", "TGCN", "SinkhornDistance", "DiceLoss", "vig.Grapher", "Affinity",
                 "MultiHeadAttention"):
        assert "graphecho_b200" in out[name], (name, out[name])
    assert out["n_params"] == 17739971            # FPN(VGG16, nc=3): SURVEY.md Appendix B


@pytest.mark.skipif(not Path("/root/reference/utils/lr_scheduler.py").exists(), reason="reference tree not mounted")
def test_unreplaced_reference_modules_still_resolve_behind_the_dropin():
    out = _run([str(ROOT / "dropin"), str(ROOT), "/root/reference"])
    assert "graphecho_b200" in out["FPN"]
    assert out["lr_scheduler"] and out["lr_scheduler"].startswith("/root/reference")
