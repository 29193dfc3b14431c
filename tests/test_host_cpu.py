"""CPU: host-side logic of the drop-in modules — registries, constructor/forward contracts,
state-dict compatibility, and the transformer half of the prompter against the reference
fixtures (the PointNet++ half needs the GPU and is covered by -m gpu tests)."""
import numpy as np
import pytest
import torch

import msr3d_b200
from msr3d_b200 import compat, registry, synthetic
from tests import helpers


def test_state_dict_keys_and_shapes_match_reference(sd_spec):
    model = msr3d_b200.build_model(helpers.prompter_cfg())
    mine = {k: (tuple(v.shape), "int64" if v.dtype == torch.int64 else "float32") for k, v in model.state_dict().items()}
    ref = {k: (tuple(s), d) for k, s, d in sd_spec}
    assert mine == ref
    trainable = sum(p.numel() for p in model.parameters() if p.requires_grad)
    assert trainable == 4224144 and sum(p.numel() for p in model.parameters()) == 6018290


def test_registry_api_and_duplicates():
    assert registry.MODEL_REGISTRY.get("OSE3DSituation") is msr3d_b200.OSE3DSituation
    assert registry.VISION_REGISTRY.get("PcdObjEncoder") is msr3d_b200.PcdObjEncoder
    with pytest.raises(KeyError):
        registry.VISION_REGISTRY.get("nope")
    with pytest.raises(AssertionError):
        registry.VISION_REGISTRY.register(msr3d_b200.PcdObjEncoder)   # fvcore semantics: no duplicates
    with pytest.raises(NotImplementedError):
        registry.build_module("audio", None)


def test_encoder_ctor_keeps_callers_config_and_freezes():
    mlps = [[3, 64, 64, 128], [128, 128, 128, 256], [256, 256, 512, 768]]
    enc = msr3d_b200.PcdObjEncoder(None, sa_mlps=mlps, freeze=True)
    assert mlps[0][0] == 3                       # the reference mutates this to 6; we keep it local
    assert enc.pcd_net.encoder[0].mlps[0][0][0].in_channels == 6
    assert enc.pcd_net.encoder[1].mlps[0][0][0].in_channels == 131
    assert enc.pcd_net.encoder[2].mlps[0][0][0].in_channels == 259
    assert all(not p.requires_grad for p in enc.parameters())


def test_cpu_tensors_are_refused_loudly():
    enc = msr3d_b200.PcdObjEncoder(None, freeze=True)
    with pytest.raises(RuntimeError, match="CPU not supported"):
        enc(torch.randn(1, 2, 64, 6))


@pytest.mark.parametrize("name,stype", [("prompter_cfg1", "as_transform_for_objects"),
                                        ("prompter_as_object", "as_object"),
                                        ("prompter_add_loc", "as_object_add_loc")])
def test_transformer_half_matches_reference(name, stype):
    fx = helpers.golden(name)
    seed = int(fx["seed"])
    model = helpers.build_prompter(seed, stype)
    batch = synthetic.make_batch(int(fx["B"]), int(fx["O"]), int(fx["P"]), seed=seed)
    emb = torch.from_numpy(fx["obj_embeds"])
    model.obj_encoder.encode = lambda pcds: emb       # fixture stands in for the GPU encoder
    with torch.no_grad():
        out = model(batch.as_dict())
    assert out["oatt"] is None
    np.testing.assert_allclose(out["obj_tokens"].numpy(), fx["obj_tokens"], rtol=1e-4, atol=1e-4)
    np.testing.assert_array_equal(out["obj_masks"].numpy(), fx["obj_masks_out"])


def test_transformer_half_gradients_match_reference():
    fx = helpers.golden("prompter_cfg1")
    seed = int(fx["seed"])
    model = helpers.build_prompter(seed)
    batch = synthetic.make_batch(int(fx["B"]), int(fx["O"]), int(fx["P"]), seed=seed)
    emb = torch.from_numpy(fx["obj_embeds"])
    model.obj_encoder.encode = lambda pcds: emb
    out = model(batch.as_dict())["obj_tokens"]
    G = torch.randn(out.shape, generator=torch.Generator().manual_seed(seed + 77))
    (out * G).sum().backward()
    grads = {n: p.grad for n, p in model.named_parameters() if p.grad is not None}
    assert set(grads) == set(str(n) for n in fx["grad_names"])     # same unused-parameter set
    for n, norm in zip(fx["grad_names"], fx["grad_norms"]):
        assert abs(float(grads[str(n)].norm()) - norm) <= 1e-3 * max(norm, 1e-3), n


def test_compat_aliases_resolve():
    compat.install_modules()
    import importlib
    assert importlib.import_module("pointnet2._ext").ball_query is msr3d_b200.ext.ball_query
    assert importlib.import_module("model.ose3d_situation").OSE3DSituation is msr3d_b200.OSE3DSituation
    assert importlib.import_module("modules.build").build_module is registry.build_module
    for name in ("pointnet2", "pointnet2._ext", "modules.build", "model.build", "model.ose3d_situation",
                 "modules.vision.pcd_pointnet_encoder", "modules.layers.pointnet", "modules.layers.transformers"):
        import sys
        sys.modules.pop(name, None)


def test_install_modules_patches_instead_of_shadowing():
    """install_modules() must not replace whole reference modules by 4-name stubs.  A module that is already
    importable keeps everything it defines; only the hot-path classes are swapped."""
    import sys
    import types
    fake = types.ModuleType("modules.layers.transformers")
    fake.GPT2Model = object()                         # what model/msr3d/msr3d.py:27 imports next to the hot-path classes
    fake.TransformerSpatialEncoderLayer = "reference class"
    pkgs = {}
    for pkg in ("modules", "modules.layers"):
        if pkg not in sys.modules:
            pkgs[pkg] = types.ModuleType(pkg)
            pkgs[pkg].__path__ = []
            sys.modules[pkg] = pkgs[pkg]
    sys.modules["modules.layers.transformers"] = fake
    try:
        compat.install_modules()
        m = sys.modules["modules.layers.transformers"]
        assert m is fake and m.GPT2Model is fake.GPT2Model
        assert m.TransformerSpatialEncoderLayer is msr3d_b200.spatial.TransformerSpatialEncoderLayer
        stub = sys.modules["model.ose3d_situation"]
        if getattr(stub, "__msr3d_stub__", False):   # no reference tree here: missing names explain themselves
            with pytest.raises(AttributeError, match="stand-in"):
                stub.SomethingElse
    finally:
        for name in ("pointnet2", "pointnet2._ext", "modules.build", "model.build", "model.ose3d_situation", "model",
                     "modules.vision.pcd_pointnet_encoder", "modules.vision", "modules.layers.pointnet",
                     "modules.layers.transformers", "modules.layers", "modules"):
            sys.modules.pop(name, None)


@pytest.mark.skipif(not __import__("oracle.refshim", fromlist=["x"]).available(), reason="needs /root/reference")
def test_reference_pointnet2_utils_binds_to_ext():
    """SURVEY 8b import side-effects: the reference's OWN pointnet2_utils.py / pointnet2_modules.py, imported where
    they lie, bind to msr3d_b200.ext (`import pointnet2._ext as _ext`, pointnet2_utils.py:22-30) and their autograd
    Functions reach our operators, which refuse CPU tensors exactly like the reference extension.  Own process: the
    shim's sys.modules arrangement is global."""
    import subprocess
    import sys
    code = """
import torch, msr3d_b200
from msr3d_b200 import compat
from oracle import refshim
compat.install_ext()
refshim.install(ext_module=msr3d_b200.ext)
import importlib
u = importlib.import_module("modules.third_party.pointnet2.pointnet2_utils")
m = importlib.import_module("modules.third_party.pointnet2.pointnet2_modules")
assert u._ext is msr3d_b200.ext
for name in ("furthest_point_sampling", "gather_points", "gather_points_grad", "ball_query", "group_points",
             "group_points_grad", "three_nn", "three_interpolate", "three_interpolate_grad"):
    assert hasattr(u._ext, name), name
for fn, args in ((u.furthest_point_sample, (torch.zeros(1, 8, 3), 4)),
                 (u.ball_query, (0.2, 4, torch.zeros(1, 8, 3), torch.zeros(1, 2, 3)))):
    try:
        fn(*args)
    except RuntimeError as e:
        assert "CPU not supported" in str(e), e
    else:
        raise SystemExit("CPU tensors were accepted")
sa = m.PointnetSAModule(npoint=4, radius=0.2, nsample=4, mlp=[3, 8], use_xyz=True)   # builds on the reference's own classes
print("ok")
"""
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=helpers.ROOT, timeout=300)
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr[-2000:]


def test_which_configurations_the_fused_engine_covers():
    """fused.supported() is host logic: the shipped configuration, every attention fusion, five situation modes and
    the unfrozen encoder are on the engine; the rest names its reason and stays on the module tree."""
    from msr3d_b200 import configs, fused
    ok = [dict(), dict(fusion="mul"), dict(fusion="bias"), dict(fusion="ctx"), dict(fusion="add"),
          dict(situation_type="as_object"), dict(situation_type="as_object_add_loc"), dict(situation_type="as_embedding"),
          dict(situation_type="as_cross_attention"),
          dict(freeze=False), dict(fusion="add", spatial_multihead=False), dict(obj_loc_encoding="same_0"),
          dict(obj_loc_encoding="diff_all"), dict(situation_type="as_object", obj_loc_encoding="diff_all")]
    for kw in ok:
        assert fused.supported(configs.build_prompter(0, **kw)) is None, kw
    no = [(dict(activation="relu"), "gelu"), (dict(spatial_multihead=False), "multihead"), (dict(num_attention_heads=4), "heads")]
    for kw, why in no:
        reason = fused.supported(configs.build_prompter(0, **kw))
        assert reason is not None and why in reason, (kw, reason)
    assert fused.encoder_frozen(configs.build_prompter(0)) and not fused.encoder_frozen(configs.build_prompter(0, freeze=False))
