"""CPU, world_size 2, gloo: the N>1 path of the hot-path trainer — flat gradient layout, one
all-reduce, averaged update — checked against a single process that sees both ranks' batches."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from msr3d_b200 import dist as mdist, synthetic
from tests import helpers


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _flat_grads(sd_spec, seed, batch, up):
    """Oracle gradients of one rank, laid out in the engine's flat order."""
    from oracle import modules_oracle as mo
    sd = synthetic.make_state_dict(sd_spec, seed)
    emb = torch.randn(batch.obj_fts.shape[0], batch.obj_fts.shape[1], 768,
                      generator=torch.Generator().manual_seed(99))       # stands in for the frozen encoder
    frozen = lambda k: k.startswith("obj_encoder.") or k == "anchor_size"
    sdg = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and not frozen(k) else v)
           for k, v in sd.items() if not k.startswith("obj_encoder.")}
    tokens, _ = mo.ose3d_situation(sdg, batch.as_dict(), obj_embeds=emb)
    (tokens * up).sum().backward()
    names = [k for k, v in sdg.items() if v.requires_grad]
    w, b, u = mdist.flat_order(names, 3)
    order = w + b + u
    offs, total = mdist.flat_offsets(order, {k: sdg[k].numel() for k in names})
    flat = torch.zeros(total)
    for k in order:
        if sdg[k].grad is not None:
            flat[offs[k]:offs[k] + sdg[k].numel()] = sdg[k].grad.flatten()
    n_opt = offs[u[0]] if u else total
    return flat, n_opt, (w, b, u), offs


def _worker(rank, world, port, spec, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    try:
        batch = synthetic.make_batch(1, 6, 32, seed=mdist.rank_seed(7, rank))   # different scenes per rank
        up = torch.randn(1, 6, 256, generator=torch.Generator().manual_seed(5))
        flat, n_opt, _, _ = _flat_grads(spec, 3, batch, up)
        params = torch.full((n_opt,), float(rank))            # deliberately different replicas
        mdist.broadcast_(params)                               # -> identical
        mdist.allreduce_sum_(flat[:n_opt])
        flat[:n_opt] *= 1.0 / world                            # grad_scale of the optimiser kernel
        out[rank] = (flat.clone(), params.clone(), mdist.world())
    finally:
        dist.destroy_process_group()


def test_two_rank_gradient_exchange_matches_single_process(sd_spec):
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, sd_spec, out), nprocs=world, join=True)
    f0, p0, w0 = out[0]
    f1, p1, w1 = out[1]
    assert w0 == (0, 2) and w1 == (1, 2)
    assert torch.equal(p0, p1) and float(p0[0]) == 0.0          # broadcast from rank 0
    up = torch.randn(1, 6, 256, generator=torch.Generator().manual_seed(5))
    singles = [_flat_grads(sd_spec, 3, synthetic.make_batch(1, 6, 32, seed=mdist.rank_seed(7, r)), up) for r in range(world)]
    n_opt = singles[0][1]
    want = (singles[0][0] + singles[1][0]) / world
    assert torch.equal(f0[:n_opt], f1[:n_opt])                  # every rank holds the same reduced gradient
    torch.testing.assert_close(f0[:n_opt], want[:n_opt], rtol=1e-4, atol=1e-4)   # intra-op thread counts differ
    # the never-used tail is not exchanged and stays zero
    assert float(f0[n_opt:].abs().max()) == 0.0 and f0[n_opt:].numel() > 0


def test_flat_layout_groups_attention_projections(sd_spec):
    names = [k for k, _, _ in sd_spec if not k.startswith("obj_encoder.") and k != "anchor_size"
             and "running" not in k and "num_batches" not in k]
    w, b, u = mdist.flat_order(names, 3)
    assert w[:4] == [f"spatial_encoder.0.self_attn.{n}.weight" for n in ("w_qs", "w_ks", "w_vs", "lang_cond_fc")]
    assert b[:4] == [f"spatial_encoder.0.self_attn.{n}.bias" for n in ("w_qs", "w_ks", "w_vs", "lang_cond_fc")]
    assert set(u) == {"anchor_feat", "orientation_encoder.weight", "orientation_encoder.bias",
                      "loc_layers.0.0.weight", "loc_layers.0.0.bias", "loc_layers.0.1.weight", "loc_layers.0.1.bias"}
    assert all("bias" in n for n in b) and not any("bias" in n for n in w)
    numel = {k: int(torch.tensor(s).prod()) for k, s, _ in sd_spec}
    offs, total = mdist.flat_offsets(w + b + u, numel)
    assert all(o % 4 == 0 for o in offs.values())
    # QKV+cond weights are one contiguous (816 x 256) block
    assert offs[w[3]] + numel[w[3]] - offs[w[0]] == 816 * 256
    assert sorted(w + b + u) == sorted(names)


def _bucket_worker(rank, world, port, e0, n_opt, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(100 + rank)
        flat = torch.randn(n_opt + 16, generator=g)
        whole = flat.clone()
        mdist.allreduce_sum_(whole[:n_opt])                     # one collective after the backward
        w1 = dist.all_reduce(flat[e0:n_opt], async_op=True)      # early bucket, issued under layer 0's backward
        w2 = dist.all_reduce(flat[0:e0], async_op=True)          # late bucket
        w1.wait(); w2.wait()
        out[rank] = (torch.equal(whole, flat), flat.clone())
    finally:
        dist.destroy_process_group()


def test_two_bucket_exchange_equals_single_allreduce(sd_spec):
    """The bucketed gradient exchange of the prefetch schedule (fused.py::train_step_prefetch) covers exactly
    [0, n_opt) and leaves the never-used tail alone."""
    names = [k for k, _, _ in sd_spec if not k.startswith("obj_encoder.") and k != "anchor_size"
             and "running" not in k and "num_batches" not in k]
    numel = {k: int(torch.tensor(s).prod()) for k, s, _ in sd_spec}
    lay = mdist.flat_layout(names, numel, 3)
    assert lay["late"] == (0, lay["early"][0]) and lay["early"][1] == lay["n_opt"]
    # scaled-down copy of the two ranges (the real buffer is 4.2 M floats)
    e0, n_opt = 1000, 3000
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_bucket_worker, args=(2, _free_port(), e0, n_opt, out), nprocs=2, join=True)
    assert out[0][0] and out[1][0]
    assert torch.equal(out[0][1][:n_opt], out[1][1][:n_opt]) and not torch.equal(out[0][1][n_opt:], out[1][1][n_opt:])


def test_flat_layout_buckets_and_decay_range(sd_spec):
    """[late biases | late weights | early weights | early biases | never-used]: one decayed range, the
    gradients of layers >= 1 in one contiguous bucket (reduced under the backward of layer 0)."""
    names = [k for k, _, _ in sd_spec if not k.startswith("obj_encoder.") and k != "anchor_size"
             and "running" not in k and "num_batches" not in k]
    numel = {k: int(torch.tensor(s).prod()) for k, s, _ in sd_spec}
    lay = mdist.flat_layout(names, numel, 3)
    # the reference decays every optimised parameter (one uniform AdamW group, model/msr3d/msr3d.py:165-172)
    assert lay["decay"] == (0, lay["n_opt"])
    lay = mdist.flat_layout(names, numel, 3, no_decay_bias=True)      # opt-in rule of optim/utils.py:1-18
    assert lay["decay"] == lay["weights"]
    offs, (d0, d1), (l0, l1), (e0, e1) = lay["offs"], lay["decay"], lay["late"], lay["early"]
    assert sorted(lay["order"]) == sorted(names)
    assert l0 == 0 and l1 == e0 and e1 == lay["n_opt"] and 0 < d0 < e0 < d1 < e1 < lay["total"]
    for n in names:
        if n.startswith(mdist.UNUSED_IN_SHIPPED_MODE):
            assert offs[n] >= lay["n_opt"], n
            continue
        decayed = d0 <= offs[n] < d1
        assert decayed == ("bias" not in n), n                        # optim/utils.py:1-18
        early = e0 <= offs[n] < e1
        assert early == (n.startswith("spatial_encoder.") and not n.startswith("spatial_encoder.0.")), n
    # QKV+cond weights of every layer are still one contiguous (816 x 256) block
    for i in range(3):
        pre = f"spatial_encoder.{i}.self_attn."
        assert offs[pre + "lang_cond_fc.weight"] + numel[pre + "lang_cond_fc.weight"] - offs[pre + "w_qs.weight"] == 816 * 256


def test_flat_layout_with_attached_caller_parameters(sd_spec):
    """llm_proj / llm_proj_img / clip_proj join the prompter in ONE flat buffer (13.67 M elements): they land in
    the early bucket (their gradients exist before the prompter's backward starts) and inside the decayed range."""
    names = [n for n, shape, _ in sd_spec if not n.startswith("obj_encoder.") and n != "anchor_size"]
    numel = {n: int(__import__("math").prod(shape)) for n, shape, _ in sd_spec}
    extra = {"ext.llm_proj.weight": 4096 * 256, "ext.llm_proj.bias": 4096, "ext.llm_proj_img.weight": 4096 * 1024,
             "ext.llm_proj_img.bias": 4096, "ext.clip_proj.weight": 4096 * 1024, "ext.clip_proj.bias": 4096}
    numel.update(extra)
    lay = mdist.flat_layout(names + list(extra), numel, 3)
    (e0, e1), (l0, l1) = lay["early"], lay["late"]
    for n in extra:
        assert e0 <= lay["offs"][n] < e1, n
        assert lay["decay"][0] <= lay["offs"][n] < lay["decay"][1]
    # 4 224 144 trainable prompter parameters, 24 320 of them never used in the shipped mode (outside n_opt), + 9 449 472
    assert lay["total"] == 13_673_616 and lay["n_opt"] == 13_673_616 - 24_320
    assert l1 == e0 and l0 == 0


@pytest.mark.parametrize("stype,unused", [("as_transform_for_objects", mdist.UNUSED_IN_SHIPPED_MODE),
                                          ("as_object", ()), ("as_embedding", ("anchor_feat", "loc_layers."))])
def test_flat_layout_per_mode_and_with_the_unfrozen_encoder(stype, unused):
    """Layout properties for the round-2 cases: per-mode never-touched segment, and the PointNet++ parameters of an
    unfrozen encoder inside the optimised range (late bucket) with the unevaluated 607-way head outside it."""
    from msr3d_b200 import configs
    model = configs.build_prompter(0, stype, freeze=False)
    named = [(n, p) for n, p in model.named_parameters() if p.requires_grad]
    numel = {n: p.numel() for n, p in named}
    lay = mdist.flat_layout([n for n, _ in named], numel, 3, unused=tuple(unused) + ("obj_encoder.obj3d_clf_pre_head.",))
    offs, n_opt = lay["offs"], lay["n_opt"]
    (l0, l1), (e0, e1) = lay["late"], lay["early"]
    assert l0 == 0 and l1 == e0 and e1 == n_opt and lay["decay"] == (0, n_opt)
    for n, _ in named:
        assert offs[n] % 4 == 0
        touched = not n.startswith(tuple(unused) + ("obj_encoder.obj3d_clf_pre_head.",))
        assert (offs[n] < n_opt) == touched, n
        if n.startswith("obj_encoder.pcd_net."):
            assert offs[n] < l1, n                    # encoder gradients are the last to be final: late bucket
        if n.startswith(("spatial_encoder.1.", "spatial_encoder.2.")):
            assert e0 <= offs[n] < e1, n
    # the attention projections of a layer stay adjacent (one QKV (+cond) GEMM)
    for i in range(3):
        pre = f"spatial_encoder.{i}.self_attn."
        assert offs[pre + "w_ks.weight"] == offs[pre + "w_qs.weight"] + numel[pre + "w_qs.weight"]
        assert offs[pre + "w_vs.weight"] == offs[pre + "w_ks.weight"] + numel[pre + "w_ks.weight"]
