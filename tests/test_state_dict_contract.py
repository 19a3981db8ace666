"""CPU: the drop-in modules keep the reference's state_dict contract (key names, shapes, buffers) --
checked against the key list recorded from the reference by oracle/make_goldens.py."""
import json
import os

import numpy as np
import torch

from conftest import GOLDEN
from oracle import index_oracle as ix, swin_oracle as so
from stswincl_b200 import swin


def test_layer_state_dict_matches_reference_keys_and_shapes():
    ref = json.load(open(os.path.join(GOLDEN, "state_dict_contract.json")))
    m = swin.SwinTransformerLayerv5(dim=128, input_resolution=(16, 24), num_heads=2)
    ours = {k: list(v.shape) for k, v in m.state_dict().items()}
    assert ours == ref["layer_128_16x24_h2"]
    m = swin.SwinTransformerLayerv5()
    ours = {k: list(v.shape) for k, v in m.state_dict().items()}
    assert ours == ref["layer_default"]
    assert sum(p.numel() for p in m.parameters()) == ref["layer_default_param_count"]


def test_strict_load_from_oracle_params_and_buffers_exact():
    p = so.make_layer_params(128, (16, 24), 2, seed=51)
    m = swin.SwinTransformerLayerv5(dim=128, input_resolution=(16, 24), num_heads=2)
    missing, unexpected = m.load_state_dict(p, strict=True)
    assert not missing and not unexpected
    blk = m.layers[0][1]
    assert np.array_equal(blk.attn_mask.numpy(), ix.shift_attn_mask(16, 24, 8, 4))
    assert np.array_equal(blk.attn.relative_position_index.numpy(), ix.relative_position_index(8))
    assert m.layers[0][0].attn_mask is None and "layers.0.0.attn_mask" not in m.state_dict()
    # checkpoint key remap of seg18/utils/LoadModel.py:19-20 (pixpro.encoder_2.* <-> swin.*) round-trips
    ck = {"pixpro.encoder_2." + k: v for k, v in m.state_dict().items()}
    back = {"swin" + k[16:]: v for k, v in ck.items()}
    assert set(k[5:] for k in back) == set(m.state_dict().keys())


def test_window_clamp_rule():
    b = swin.SwinTransformerBlock(64, (4, 8), 1, window_size=8, shift_size=4)    # swin_512.py:155-158
    assert (b.window_size, b.shift_size) == (4, 0) and b.attn_mask is None
