"""CPU: the oracle (oracle/*.py + oracle/pointnet2_oracle.c) against fixtures produced by
the UNMODIFIED reference Python modules (tests/golden/gen_golden.py)."""
import json
import os

import numpy as np
import pytest
import torch

from msr3d_b200 import synthetic
from oracle import ext_oracle, modules_oracle as mo

CASES = {
    "prompter_cfg1": "as_transform_for_objects",
    "prompter_mid": "as_transform_for_objects",
    "prompter_as_object": "as_object",
    "prompter_add_loc": "as_object_add_loc",
    "prompter_as_embedding": "as_embedding",
    "prompter_as_cross_attention": "as_cross_attention",
}


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name + ".npz"), allow_pickle=False)


def _state_dict(sd_spec, name, seed):
    """Weights by NAME (synthetic.make_state_dict): the shipped model's spec, or — for modes that own parameters the
    shipped model does not have — the spec of that mode's drop-in model."""
    if name == "prompter_as_cross_attention":
        from msr3d_b200 import configs
        return configs.model_state_dict(configs.build_prompter(0, CASES[name]), seed)
    return synthetic.make_state_dict(sd_spec, seed)


@pytest.mark.parametrize("name", list(CASES))
def test_prompter_forward_matches_reference(golden_dir, sd_spec, name):
    fx = _load(golden_dir, name)
    seed = int(fx["seed"])
    sd = _state_dict(sd_spec, name, seed)
    batch = synthetic.make_batch(int(fx["B"]), int(fx["O"]), int(fx["P"]), seed=seed)
    trace = {}
    ext_oracle.set_threads(4)
    with torch.no_grad():
        tokens, masks = mo.ose3d_situation(sd, batch.as_dict(), situation_type=CASES[name], trace=trace)
    # index outputs of the native ops: exact
    np.testing.assert_array_equal(trace["obj_encoder.pcd_net.encoder.0.fps_idx"].numpy(), fx["fps_idx_0"])
    np.testing.assert_array_equal(trace["obj_encoder.pcd_net.encoder.0.ball_idx"].numpy(), fx["ball_idx_0"])
    np.testing.assert_array_equal(trace["obj_encoder.pcd_net.encoder.1.fps_idx"].numpy(), fx["fps_idx_1"])
    np.testing.assert_array_equal(trace["obj_encoder.pcd_net.encoder.1.ball_idx"].numpy(), fx["ball_idx_1"])
    for i in range(3):
        np.testing.assert_allclose(trace[f"layer{i}.out"].numpy(), fx[f"layer{i}_out"], rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(tokens.numpy(), fx["obj_tokens"], rtol=1e-4, atol=1e-4)
    np.testing.assert_array_equal(masks.numpy(), fx["obj_masks_out"])


@pytest.mark.parametrize("fusion", ["mul", "bias", "ctx", "add"])
def test_prompter_other_fusions_match_reference(golden_dir, fusion):
    """Whole prompter, forward + backward, with the fusions no shipped config uses against the
    unmodified reference (gen_golden.py --fusions)."""
    from msr3d_b200 import configs
    fx = _load(golden_dir, f"prompter_fusion_{fusion}")
    seed = int(fx["seed"])
    sd = configs.model_state_dict(configs.build_prompter(0, fusion=fusion), seed)
    frozen = lambda k: k.startswith("obj_encoder.") or k == "anchor_size"
    sd = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and not frozen(k) else v) for k, v in sd.items()}
    batch = synthetic.make_batch(int(fx["B"]), int(fx["O"]), int(fx["P"]), seed=seed)
    ext_oracle.set_threads(4)
    tokens, _ = mo.ose3d_situation(sd, batch.as_dict(), fusion=fusion)
    np.testing.assert_allclose(tokens.detach().numpy(), fx["obj_tokens"], rtol=1e-4, atol=1e-4)
    G = torch.randn(tokens.shape, generator=torch.Generator().manual_seed(seed + 77))
    (tokens * G).sum().backward()
    for n, norm in zip([str(n) for n in fx["grad_names"]], fx["grad_norms"]):
        assert sd[n].grad is not None, n
        assert abs(float(sd[n].grad.norm()) - norm) <= 1e-3 * max(norm, 1e-3), n
    for key in fx.files:
        if key.startswith("grad:"):
            ref = fx[key]
            assert np.abs(sd[key[5:]].grad.numpy() - ref).max() <= 1e-4 * max(np.abs(ref).max(), 1e-3), key


def test_encoder_outputs_match_reference(golden_dir, sd_spec):
    fx = _load(golden_dir, "prompter_cfg1")
    seed = int(fx["seed"])
    sd = synthetic.make_state_dict(sd_spec, seed)
    batch = synthetic.make_batch(int(fx["B"]), int(fx["O"]), int(fx["P"]), seed=seed)
    with torch.no_grad():
        emb, cls = mo.pcd_obj_encoder(sd, "obj_encoder", batch.obj_fts)
    scale = np.abs(fx["obj_embeds"]).max()
    assert np.abs(emb.numpy() - fx["obj_embeds"]).max() <= 1e-5 * scale
    np.testing.assert_allclose(cls.numpy(), fx["obj_sem_cls"], rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("name", ["prompter_cfg1", "prompter_as_object", "prompter_as_embedding",
                                  "prompter_as_cross_attention"])
def test_prompter_backward_matches_reference(golden_dir, sd_spec, name):
    fx = _load(golden_dir, name)
    seed = int(fx["seed"])
    sd = _state_dict(sd_spec, name, seed)
    frozen = lambda k: k.startswith("obj_encoder.") or k == "anchor_size"
    sd = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and not frozen(k) else v)
          for k, v in sd.items()}
    batch = synthetic.make_batch(int(fx["B"]), int(fx["O"]), int(fx["P"]), seed=seed)
    tokens, _ = mo.ose3d_situation(sd, batch.as_dict(), situation_type=CASES[name])
    G = torch.randn(tokens.shape, generator=torch.Generator().manual_seed(seed + 77))
    (tokens * G).sum().backward()
    names = [str(n) for n in fx["grad_names"]]
    for n, norm in zip(names, fx["grad_norms"]):
        g = sd[n].grad
        assert g is not None, n
        assert abs(float(g.norm()) - norm) <= 1e-3 * max(norm, 1e-3), n
    for key in fx.files:
        if key.startswith("grad:"):
            n = key[5:]
            ref = fx[key]
            tol = 1e-4 * max(np.abs(ref).max(), 1e-3)
            assert np.abs(sd[n].grad.numpy() - ref).max() <= tol, n
    # parameters the reference leaves without gradient stay that way
    with_grad = {n for n in sd if isinstance(sd[n], torch.Tensor) and sd[n].requires_grad and sd[n].grad is not None}
    assert with_grad == set(names)


@pytest.mark.parametrize("fusion", ["cond", "mul", "bias", "ctx", "add"])
def test_spatial_mha_all_fusions(golden_dir, fusion):
    fx = _load(golden_dir, "mha_fusions")
    spec = [tuple(x) for x in json.loads(str(fx[f"{fusion}_spec"]))]
    sd = {"a." + k: v for k, v in synthetic.make_state_dict(spec, 11).items()}
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 7, 64, generator=g)
    centers = torch.rand(2, 7, 3, generator=g) * 4
    np.testing.assert_allclose(centers.numpy(), fx["centers"])
    pw = mo.pairwise_locs_center(centers)
    np.testing.assert_allclose(pw.numpy(), fx["pairwise"], rtol=1e-5, atol=1e-6)
    mask = torch.zeros(2, 7, dtype=torch.bool)
    mask[1, 5:] = True
    out, att = mo.spatial_mha(sd, "a", x, pw, mask, 4, fusion)
    np.testing.assert_allclose(out.numpy(), fx[f"{fusion}_out"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(att.numpy(), fx[f"{fusion}_attn"], rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize("case", ["f32", "f16", "bf16"])
def test_scene_splice_matches_reference(golden_dir, case):
    """oracle.scene_splice vs the unmodified reference MSR3D.build_embeds (msr3d.py:277-287)."""
    fx = _load(golden_dir, "splice")
    get = lambda k: torch.from_numpy(fx[f"{case}:{k}"])  # noqa: E731
    dtype = getattr(torch, str(fx[f"{case}:dtype"]))
    W, b, tok = get("W").requires_grad_(True), get("b").requires_grad_(True), get("tokens").requires_grad_(True)
    out, am = mo.scene_splice(W, b, get("emb_rows").to(dtype), get("attention_mask"), get("ids"), tok,
                              get("masks"))
    assert out.dtype == dtype and am.dtype == torch.bool
    np.testing.assert_array_equal(out.detach().float().numpy(), fx[f"{case}:out"])   # same ops, same order: exact
    np.testing.assert_array_equal(am.numpy().astype(np.int64), fx[f"{case}:am_out"])
    (out.float() * get("G")).sum().backward()
    for name, t in (("dW", W), ("db", b), ("dtokens", tok)):
        ref = fx[f"{case}:{name}"]
        assert np.abs(t.grad.numpy() - ref).max() <= 1e-5 * max(np.abs(ref).max(), 1e-6), name


@pytest.mark.parametrize("case", ["train", "val"])
def test_preprocess_objects_matches_reference(golden_dir, case):
    """oracle.preprocess_objects vs the unmodified reference preprocess_pcd + padding (msr3d.py:181-215)."""
    from oracle import data_oracle
    fx = _load(golden_dir, "preprocess")
    g = lambda k: fx[f"{case}:{k}"]  # noqa: E731
    rot = g("rot") if bool(g("has_rot")) else None
    fts, locs, masks = data_oracle.preprocess_objects(g("points"), g("offsets"), g("sample_idx"), rot,
                                                      g("obj_fts").shape[0])
    np.testing.assert_array_equal(fts, g("obj_fts"))          # same numpy operations in the same order: exact
    np.testing.assert_array_equal(locs, g("obj_locs"))
    np.testing.assert_array_equal(masks, g("obj_masks"))
    # the degenerate object (all points equal) is left unscaled, the padding rows are ones / zeros / False
    assert np.all(fts[~masks] == 1.0) and np.all(locs[~masks] == 0.0)


def _encoder_train_case(golden_dir, case):
    """Inputs of one case of tests/golden/encoder_train.npz: (fixture, state dict with grads enabled, batch, G)."""
    fx = _load(golden_dir, "encoder_train")
    B, O, P, seed, bn_train = (int(v) for v in fx[f"{case}:meta"])
    names = [k.split(":", 2)[2] for k in fx.files if k.startswith(f"{case}:grad")]
    bufs = [k.split(":", 2)[2] for k in fx.files if k.startswith(f"{case}:buf:")]
    return fx, B, O, P, seed, bool(bn_train), sorted(set(names)), bufs


def check_encoder_grads(fx, case, grads: dict, tol=1e-4):
    """grads: name (relative to pcd_net) -> numpy gradient; compared with the reference's, full or subsampled."""
    worst = ("", 0.0)
    for key in fx.files:
        if key.startswith(f"{case}:grad:"):
            n, ref, got = key.split(":", 2)[2], fx[key], None
            got = grads[n]
        elif key.startswith(f"{case}:gradsub:"):
            n, ref = key.split(":", 2)[2], fx[key]
            got = grads[n].reshape(-1)[::37]
            st = fx[f"{case}:gradstat:{n}"]
            assert abs(np.linalg.norm(grads[n].astype(np.float64)) - st[0]) <= 10 * tol * max(st[0], 1e-6), n
        else:
            continue
        err = np.abs(got - ref).max() / max(np.abs(ref).max(), 1e-6)
        if err > worst[1]:
            worst = (n, err)
        assert err <= tol, (n, err)
    return worst


@pytest.mark.parametrize("case", ["train", "evalbn"])
def test_unfrozen_encoder_matches_reference(golden_dir, case):
    """oracle.pointnetpp with training-mode BatchNorm and autograd vs the unmodified reference PcdObjEncoder(freeze=False)
    (SURVEY 8f-2): embeddings, gradients of all 29 PointNet++ parameters, updated running statistics."""
    from msr3d_b200 import configs
    fx, B, O, P, seed, bn_train, names, bufs = _encoder_train_case(golden_dir, case)
    enc = configs.build_prompter(0).obj_encoder
    spec = [(k, list(v.shape), "int64" if v.dtype == torch.int64 else "float32") for k, v in enc.state_dict().items()]
    sd = synthetic.make_state_dict(spec, seed)
    sd = {"e." + k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v)
          for k, v in sd.items()}
    batch = synthetic.make_batch(B, O, P, seed=seed)
    ext_oracle.set_threads(4)
    stats = {} if bn_train else None
    flat = batch.obj_fts.reshape(B * O, P, 6)
    emb = mo.pointnetpp(sd, "e.pcd_net", flat, (32, 16, None), (32, 32, None), (0.2, 0.4, None), bn_stats=stats)
    np.testing.assert_allclose(emb.detach().numpy().reshape(B, O, -1), fx[f"{case}:emb"], rtol=1e-4, atol=1e-4)
    G = torch.randn(B, O, emb.shape[-1], generator=torch.Generator().manual_seed(seed + 77))
    (emb.view(B, O, -1) * G).sum().backward()
    grads = {n: sd["e.pcd_net." + n].grad.numpy() for n in names}
    worst = check_encoder_grads(fx, case, grads)
    print(case, "worst grad", worst)
    for n in bufs:
        if "num_batches" in n:
            continue
        got = (stats or {}).get("e.pcd_net." + n, sd["e.pcd_net." + n]).detach().numpy()
        np.testing.assert_allclose(got, fx[f"{case}:buf:{n}"], rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize("name,stype", [("prompter_diff_all", "as_transform_for_objects"),
                                        ("prompter_diff_all_as_object", "as_object")])
def test_prompter_diff_all_matches_reference(golden_dir, name, stype):
    """obj_loc_encoding: diff_all (a loc layer per transformer layer, ose3d_situation.py:379-381), forward + backward."""
    from msr3d_b200 import configs
    fx = _load(golden_dir, name)
    seed = int(fx["seed"])
    sd = configs.model_state_dict(configs.build_prompter(0, stype, obj_loc_encoding="diff_all"), seed)
    frozen = lambda k: k.startswith("obj_encoder.") or k == "anchor_size"
    sd = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and not frozen(k) else v) for k, v in sd.items()}
    batch = synthetic.make_batch(int(fx["B"]), int(fx["O"]), int(fx["P"]), seed=seed)
    ext_oracle.set_threads(4)
    tokens, _ = mo.ose3d_situation(sd, batch.as_dict(), situation_type=stype, obj_loc_encoding="diff_all")
    np.testing.assert_allclose(tokens.detach().numpy(), fx["obj_tokens"], rtol=1e-4, atol=1e-4)
    G = torch.randn(tokens.shape, generator=torch.Generator().manual_seed(seed + 77))
    (tokens * G).sum().backward()
    names = [str(n) for n in fx["grad_names"]]
    for n, norm in zip(names, fx["grad_norms"]):
        assert sd[n].grad is not None and abs(float(sd[n].grad.norm()) - norm) <= 1e-3 * max(norm, 1e-3), n
    with_grad = {n for n in sd if isinstance(sd[n], torch.Tensor) and sd[n].requires_grad and sd[n].grad is not None}
    assert with_grad == set(names)
