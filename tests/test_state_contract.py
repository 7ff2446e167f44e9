"""CPU: the product modules expose the reference's state_dict keys and shapes
(tests/golden/state_contract.json, dumped from the reference by oracle/make_golden.py), so
checkpoints move both ways."""
import contextlib
import io
import json
from pathlib import Path

import pytest

from graphecho_b200.models import fpnseg, graph_matching, TGCN as tgcn_mod, vig, affinity_layer, transformer

CONTRACT = json.loads((Path(__file__).parent / "golden" / "state_contract.json").read_text())


def _build(name):
    if name == "fpn_resnet_nc1":
        return fpnseg.FPN([2, 4, 23, 3], 1, 1, back_bone="resnet")
    if name == "fpn_vgg16_nc3":
        return fpnseg.FPN([2, 4, 23, 3], 3, 1, back_bone="VGG16")
    if name == "discriminator":
        return fpnseg.Discriminator(grad_reverse_lambda=0.02)
    if name == "gmodule_nc3":
        return graph_matching.GModule(256, 3, "cpu")
    if name == "tgcn_nd":
        return tgcn_mod.TGCN(256, 256, (3, 8, 8), 10, 10, None, "node_discriminate")
    if name == "tgcn_sd":
        return tgcn_mod.TGCN(256, 256, (3, 8, 8), 10, 10, None, "sinkhorn_distance")
    if name == "grapher32":
        return vig.Grapher(32, 5, 1, "mr", "gelu", "batch", True, False, 0.0, 1, 64, 0.0, False)
    if name == "grapher256":
        return vig.Grapher(256, 9, 1, "mr", "gelu", "batch", True, False, 0.0, 1, 784, 0.0, False)
    if name == "mrconv32_64":
        return vig.MRConv2d(32, 64, "gelu", None, True)
    if name == "affinity":
        return affinity_layer.Affinity(256)
    if name == "mha":
        return transformer.MultiHeadAttention(256, 1, dropout=0.1, version="v2")
    raise KeyError(name)


@pytest.mark.parametrize("name", [k for k in CONTRACT if not k.startswith("_")])
def test_state_dict_matches_reference(name):
    with contextlib.redirect_stdout(io.StringIO()):
        mod = _build(name)
    ours = {k: list(v.shape) for k, v in mod.state_dict().items()}
    assert ours == CONTRACT[name]
    assert sum(p.numel() for p in mod.parameters()) == CONTRACT["_param_counts"][name]


def test_pvig_factories_build():
    m = vig.pvig_ti_224_gelu()
    assert sum(p.numel() for p in m.parameters()) == 12486314
