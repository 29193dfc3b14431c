"""Generates tests/golden/splice.npz from the UNMODIFIED reference `MSR3D.build_embeds`
(model/msr3d/msr3d.py:208-289): the scene-token splice that follows the prompter — `llm_proj`
(256 -> LLM width), cast to the embedding dtype, masked scatter at the scene placeholder id and the
attention-mask splice — plus its gradients.

Authoring-container only (needs /root/reference):   python tests/golden/gen_golden_splice.py
The method is called unbound on a stand-in `self` that carries exactly the attributes the method
reads (`device`, `llm_model.get_input_embeddings()`, `llm_proj`); everything msr3d.py imports at
module level that is not installed here (clip, hydra, peft, IPython, the data stack) is stubbed
because none of it is touched by build_embeds' scene branch.
"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import refshim  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def load_reference_msr3d():
    refshim.install()

    def stub(name, **attrs):
        m = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
        return m
    stub("clip")
    stub("hydra")
    stub("peft", LoraConfig=object, get_peft_model=None)
    stub("IPython", embed=None)
    d = stub("data")
    d.__path__ = []
    stub("data.data_utils", pad_tensors=None, find_subsequence=None, path_verify=None, save_to_json=None)
    pkg = types.ModuleType("model.msr3d")
    pkg.__path__ = [os.path.join(refshim.REFERENCE_ROOT, "model", "msr3d")]
    sys.modules["model.msr3d"] = pkg
    import importlib
    return importlib.import_module("model.msr3d.msr3d")


def case(ref, seed, B, T, L, D, dtype, n_masked):
    g = torch.Generator().manual_seed(seed)
    SP = 31495
    vocab = 32001
    emb = torch.nn.Embedding(vocab, D)
    with torch.no_grad():
        emb.weight.copy_(torch.randn(vocab, D, generator=g) * 0.02)
    emb = emb.to(dtype)
    proj = torch.nn.Linear(256, D)
    with torch.no_grad():
        proj.weight.copy_(torch.randn(D, 256, generator=g) * 0.02)
        proj.bias.copy_(torch.randn(D, generator=g) * 0.02)
    ids = torch.randint(0, 30000, (B, T), generator=g)
    for b in range(B):   # L consecutive placeholders per prompt at a per-row offset (msr3d.py:280)
        s = int(torch.randint(1, T - L, (1,), generator=g))
        ids[b, s:s + L] = SP
    am = torch.ones(B, T, dtype=torch.long)
    am[:, 0] = 0       # left padding
    tokens = torch.randn(B, L, 256, generator=g).requires_grad_(True)
    masks = torch.ones(B, L, dtype=torch.bool)
    if n_masked:
        masks[:, -n_masked:] = False
    fake = types.SimpleNamespace(
        device=torch.device("cpu"),
        llm_model=types.SimpleNamespace(get_input_embeddings=lambda: emb),
        llm_proj=proj)
    scene = {"msr3d_imgs": torch.zeros(B, 1, 2, 2, 3), "msr3d_img_masks": torch.zeros(B, 1),
             "obj_tokens": tokens, "obj_masks": masks}
    out, am_out = ref.MSR3D.build_embeds(fake, scene_dict=scene, input_ids=ids, attention_mask=am)
    G = torch.randn(out.shape, generator=g).to(dtype)
    (out.float() * G.float()).sum().backward()
    return {
        "ids": ids.numpy(), "attention_mask": am.numpy(), "tokens": tokens.detach().numpy(),
        "masks": masks.numpy(), "W": proj.weight.detach().numpy(), "b": proj.bias.detach().numpy(),
        "emb_rows": emb(ids).detach().float().numpy(),       # input embeddings before the splice
        "out": out.detach().float().numpy(), "am_out": am_out.numpy().astype(np.int64),
        "G": G.float().numpy(), "dW": proj.weight.grad.numpy(), "db": proj.bias.grad.numpy(),
        "dtokens": tokens.grad.numpy(), "dtype": np.array(str(dtype).replace("torch.", "")),
    }


def main():
    ref = load_reference_msr3d()
    out = {}
    for name, kw in {
        "f32": dict(seed=5, B=2, T=24, L=5, D=96, dtype=torch.float32, n_masked=1),
        "f16": dict(seed=6, B=3, T=40, L=8, D=160, dtype=torch.float16, n_masked=2),
        "bf16": dict(seed=7, B=1, T=16, L=3, D=64, dtype=torch.bfloat16, n_masked=0),
    }.items():
        for k, v in case(ref, **kw).items():
            out[f"{name}:{k}"] = v
    np.savez_compressed(os.path.join(HERE, "splice.npz"), **out)
    print("wrote splice.npz", {k: getattr(v, "shape", None) for k, v in out.items() if k.startswith("f16")})


if __name__ == "__main__":
    main()
