"""CPU: host-side logic of the scene-token splice and the ABI surface it relies on (no kernels run)."""
import ctypes

import pytest
import torch

from oracle import modules_oracle as mo


def test_gemm_desc_mirror_matches_library():
    """The ctypes mirror of msr3d_gemm_desc (msr3d_b200/fused.py) has the size the library was built with."""
    from msr3d_b200 import _C
    from msr3d_b200.fused import GemmDesc
    lib = _C.lib()
    lib.msr3d_sizeof_gemm_desc.restype = ctypes.c_int
    assert ctypes.sizeof(GemmDesc) == lib.msr3d_sizeof_gemm_desc()
    assert lib.msr3d_abi_version() == 4


def test_splice_rejects_cpu_tensors():
    from msr3d_b200 import splice
    with pytest.raises(RuntimeError, match="CUDA"):
        splice.splice_scene_tokens(torch.zeros(1, 4, 8), torch.ones(1, 4, dtype=torch.long),
                                   torch.zeros(1, 4, dtype=torch.long), torch.zeros(1, 1, 256),
                                   torch.ones(1, 1, dtype=torch.bool), torch.zeros(8, 256), None)


def test_scene_splice_module_state_dict_keys():
    from msr3d_b200.splice import SceneSplice
    m = SceneSplice(256, 512)
    assert set(m.state_dict()) == {"llm_proj.weight", "llm_proj.bias"}   # model/msr3d/msr3d.py:84-87
    assert m.llm_proj.weight.shape == (512, 256)


def test_oracle_splice_edge_cases():
    """No placeholders: nothing changes but the mask dtype; placeholders in several rows keep torch.where order."""
    W, b = torch.randn(6, 256), torch.randn(6)
    emb = torch.randn(2, 5, 6)
    am = torch.tensor([[0, 1, 1, 1, 1], [1, 1, 1, 1, 0]])
    ids = torch.arange(10).view(2, 5)
    out, m = mo.scene_splice(W, b, emb, am, ids, torch.zeros(2, 0, 256), torch.zeros(2, 0, dtype=torch.bool))
    assert torch.equal(out, emb) and m.dtype == torch.bool and torch.equal(m, am.bool())
    ids[0, 4] = 31495
    ids[1, 0] = 31495
    tok = torch.randn(2, 1, 256)
    masks = torch.tensor([[True], [False]])
    out, m = mo.scene_splice(W, b, emb, am, ids, tok, masks)
    want = torch.nn.functional.linear(tok, W, b)
    assert torch.allclose(out[0, 4], want[0, 0]) and torch.allclose(out[1, 0], want[1, 0])
    assert bool(m[0, 4]) is True and bool(m[1, 0]) is False
    assert torch.equal(out[0, :4], emb[0, :4]) and torch.equal(out[1, 1:], emb[1, 1:])


def test_data_preprocess_rejects_cpu_tensors_and_packs_objects():
    from msr3d_b200 import data
    pts, offs = data.pack_objects([torch.zeros(5, 6), torch.ones(3, 6)])
    assert pts.shape == (8, 6) and offs.tolist() == [0, 5, 8] and offs.dtype == torch.int32
    with pytest.raises(RuntimeError, match="CUDA"):
        data.preprocess_objects(pts, offs, torch.zeros(2, 4, dtype=torch.int32), None, 2)
