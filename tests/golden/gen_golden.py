"""Generates tests/golden/*.npz from the UNMODIFIED reference Python modules.

Authoring-container only (needs /root/reference); run as
    python tests/golden/gen_golden.py
The reference modules are imported where they lie through oracle/refshim.py, with
`pointnet2._ext` served by the C oracle (the reference extension is CUDA-only).
Weights come from msr3d_b200.synthetic.make_state_dict (deterministic per key), are
loaded into the reference model with strict=True — which also pins the state-dict
key/shape compatibility surface, written to state_dict_spec.json — and are NOT
stored; fixtures hold inputs' seeds and the reference's outputs only.
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ext_oracle, refshim  # noqa: E402
from msr3d_b200 import synthetic  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
torch.set_num_threads(8)
ext_oracle.set_threads(8)


def record_ext(ref):
    """Wrap the _ext functions seen by the reference's pointnet2_utils to log index outputs."""
    log = []
    ext = ref.pointnet2_utils._ext
    rec = type("RecExt", (), {})()
    for name in ("gather_points", "gather_points_grad", "group_points", "group_points_grad",
                 "three_nn", "three_interpolate", "three_interpolate_grad"):
        setattr(rec, name, getattr(ext, name))

    def fps(points, n):
        out = ext.furthest_point_sampling(points, n)
        log.append(("fps", out.clone()))
        return out

    def bq(new_xyz, xyz, r, ns):
        out = ext.ball_query(new_xyz, xyz, r, ns)
        log.append(("ball", out.clone()))
        return out
    rec.furthest_point_sampling = fps
    rec.ball_query = bq
    # pointnet2_modules.py:16-22 imports the wrappers a second time under the
    # top-level name `pointnet2_utils`; patch both module objects.
    for mod in (ref.pointnet2_utils, sys.modules.get("pointnet2_utils")):
        if mod is not None:
            mod._ext = rec
    return log, ext


def build_model(ref, situation_type, seed, fusion="cond", obj_loc_encoding=None):
    cfg = refshim.msr3d_prompter_cfg(situation_type=situation_type)
    cfg.model.spatial_encoder.spatial_attn_fusion = fusion
    if obj_loc_encoding is not None:
        cfg.model.spatial_encoder.obj_loc_encoding = obj_loc_encoding
    model = ref.ose3d_situation.OSE3DSituation(cfg)
    spec = [(k, list(v.shape), "int64" if v.dtype == torch.int64 else "float32")
            for k, v in model.state_dict().items()]
    sd = synthetic.make_state_dict(spec, seed)
    model.load_state_dict(sd, strict=True)
    model.eval()   # dropout off; BN eval is forced by freeze_bn anyway
    return model, spec


def run_case(ref, name, B, O, P, seed, situation_type="as_transform_for_objects", with_grad=True, fusion="cond",
             obj_loc_encoding=None):
    model, spec = build_model(ref, situation_type, seed, fusion, obj_loc_encoding)
    batch = synthetic.make_batch(B, O, P, seed=seed)
    log, orig_ext = record_ext(ref)
    layer_out = []
    hooks = [l.register_forward_hook(lambda m, i, o: layer_out.append(o[0].detach().clone()))
             for l in model.spatial_encoder]
    enc_out = []
    hooks.append(model.obj_encoder.register_forward_hook(
        lambda m, i, o: enc_out.append((o[0].detach().clone(), o[1].detach().clone()))))
    data = batch.as_dict()
    out = model(dict(data))
    for h in hooks:
        h.remove()
    for mod in (ref.pointnet2_utils, sys.modules.get("pointnet2_utils")):
        if mod is not None:
            mod._ext = orig_ext
    tokens = out["obj_tokens"]
    fix = {
        "B": B, "O": O, "P": P, "seed": seed,
        "obj_tokens": tokens.detach().numpy(),
        "obj_masks_out": out["obj_masks"].numpy(),
        "obj_embeds": enc_out[0][0].numpy(),
        "obj_sem_cls": enc_out[0][1].numpy(),
    }
    for i, t in enumerate(layer_out):
        fix[f"layer{i}_out"] = t.numpy()
    k = 0
    for kind, t in log:
        fix[f"{kind}_idx_{k // 2}"] = t.numpy()
        k += 1
    if with_grad:
        g = torch.Generator().manual_seed(seed + 77)
        G = torch.randn(tokens.shape, generator=g)
        model.zero_grad()
        (tokens * G).sum().backward()
        names, norms, sums = [], [], []
        for n, p in model.named_parameters():
            if p.requires_grad and p.grad is not None:
                names.append(n)
                norms.append(float(p.grad.norm()))
                sums.append(float(p.grad.double().sum()))
                if p.numel() <= 4096 or n in ("spatial_encoder.0.self_attn.lang_cond_fc.weight",) \
                        or n.endswith("pairwise_loc_fc.weight") or n.endswith("1.multihead_attn.out_proj.weight"):
                    fix["grad:" + n] = p.grad.numpy().copy()
        fix["grad_names"] = np.array(names)
        fix["grad_norms"] = np.array(norms, dtype=np.float64)
        fix["grad_sums"] = np.array(sums, dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **fix)
    print(name, "tokens", tuple(tokens.shape), "abs-mean %.4f" % tokens.abs().mean().item(),
          "ext calls", [k for k, _ in log])
    return spec


def run_mha_cases(ref):
    """MultiHeadAttentionSpatial alone, all five fusions (transformers.py:167-252)."""
    fix = {}
    for fusion in ("cond", "mul", "bias", "ctx", "add"):
        torch.manual_seed(5)
        m = ref.transformers.MultiHeadAttentionSpatial(64, 4, dropout=0.0, spatial_multihead=True,
                                                       spatial_dim=5, spatial_attn_fusion=fusion)
        spec = [(k, list(v.shape), "float32") for k, v in m.state_dict().items()]
        m.load_state_dict(synthetic.make_state_dict(spec, 11))
        m.eval()
        g = torch.Generator().manual_seed(3)
        x = torch.randn(2, 7, 64, generator=g)
        centers = torch.rand(2, 7, 3, generator=g) * 4
        pw = ref.utils.calc_pairwise_locs(centers, None)
        mask = torch.zeros(2, 7, dtype=torch.bool)
        mask[1, 5:] = True
        out, att = m(x, x, x, pw, key_padding_mask=mask)
        fix[f"{fusion}_out"] = out.detach().numpy()
        fix[f"{fusion}_attn"] = att.detach().numpy()
        fix[f"{fusion}_spec"] = np.array(json.dumps(spec))
    fix["pairwise"] = pw.numpy()
    fix["centers"] = centers.numpy()
    np.savez_compressed(os.path.join(HERE, "mha_fusions.npz"), **fix)
    print("mha_fusions ok")


def run_fusion_cases(ref):
    """Whole prompter (forward + backward) with the four fusions no shipped config uses: the fixtures the
    fused engine's fusion-templated attention kernels are checked against (tests/test_modes_gpu.py)."""
    for k, fusion in enumerate(("mul", "bias", "ctx", "add")):
        run_case(ref, f"prompter_fusion_{fusion}", 1, 8, 256, seed=5005 + k, fusion=fusion)


def main():
    ref = refshim.load()
    if "--fusions" in sys.argv:          # only the fixtures added in round 2 (the others stay byte-stable)
        run_fusion_cases(ref)
        return
    if "--as-embedding" in sys.argv:
        run_case(ref, "prompter_as_embedding", 2, 7, 256, seed=7007, situation_type="as_embedding")
        return
    if "--diff-all" in sys.argv:       # obj_loc_encoding: diff_all (a loc layer per transformer layer)
        run_case(ref, "prompter_diff_all", 2, 7, 256, seed=9009, obj_loc_encoding="diff_all")
        run_case(ref, "prompter_diff_all_as_object", 1, 6, 256, seed=9010, situation_type="as_object",
                 obj_loc_encoding="diff_all")
        return
    if "--as-cross-attention" in sys.argv:
        run_case(ref, "prompter_as_cross_attention", 2, 6, 256, seed=8008, situation_type="as_cross_attention")
        return
    spec = run_case(ref, "prompter_cfg1", 1, 8, 256, seed=1001)
    with open(os.path.join(HERE, "state_dict_spec.json"), "w") as fh:
        json.dump(spec, fh, indent=0)
    run_case(ref, "prompter_mid", 2, 12, 1024, seed=2002)
    run_case(ref, "prompter_as_object", 1, 8, 256, seed=3003, situation_type="as_object")
    run_case(ref, "prompter_add_loc", 1, 6, 128, seed=4004, situation_type="as_object_add_loc",
             with_grad=False)
    run_mha_cases(ref)


if __name__ == "__main__":
    main()
