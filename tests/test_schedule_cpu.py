"""CPU: host logic of the training schedules (msr3d_b200/train.py) with a recording stand-in for the engine —
which batch is copied into which slot, which slot each step encodes, and that `flush()` drains exactly the
updates still in flight.  (The numerical equivalence of the schedules is a GPU test:
tests/test_fused_gpu.py::test_pipelined_schedule_equals_sequential.)"""
import types

import pytest
import torch

from msr3d_b200 import train


class _Rec:
    """Stands in for FusedPrompter: records the calls HotPathTrainer makes."""

    def __init__(self):
        self.log = []
        self.n_trainable = 0

    def __getattr__(self, name):
        def call(*a, **kw):
            self.log.append((name, a, kw))
            return torch.tensor(float(len(self.log)))
        return call


def _trainer(pipeline):
    t = train.HotPathTrainer.__new__(train.HotPathTrainer)
    t.engine, t.use_graph, t.pipeline, t._calls, t.n_trainable = _Rec(), True, pipeline, 0, 0
    return t


def _names(t):
    return [n for n, _, _ in t.engine.log]


def test_sequential_schedule_loads_then_steps():
    t = _trainer(False)
    t._step("b0", "u0")
    assert _names(t) == ["load_inputs", "train_step", "loss"]


def test_pipelined_schedule_stages_points_before_and_scene_after_the_step():
    t = _trainer(True)
    t._step("b1", "u1")
    assert _names(t) == ["stage_points", "train_step_pipelined", "loss", "stage_scene"]
    t.flush()
    assert t.engine.log[-2][0] == "train_step_pipelined" and t.engine.log[-2][2] == {"flush": True}


def test_prefetch_schedule_alternates_slots_and_flush_drains_two_updates():
    t = _trainer("prefetch")
    t._calls = 2                      # prime2() filled slots 0 and 1
    for i, b in enumerate(["b2", "b3", "b4"]):
        t._step(b, "u")
    pre = [(a[2]) for n, a, _ in t.engine.log if n == "prefetch"]             # slot each batch is copied into
    enc = [a[0] for n, a, _ in t.engine.log if n == "train_step_prefetch"]    # slot each step encodes
    assert pre == [0, 1, 0]           # batch i -> slot i % 2
    assert enc == [1, 0, 1]           # step i encodes batch i-1 from slot (i-1) % 2: never the slot being filled
    assert all(p != e for p, e in zip(pre, enc))
    n_before = len(t.engine.log)
    out = t.flush()
    tail = t.engine.log[n_before:]
    steps = [(a, kw) for n, a, kw in tail if n == "train_step_prefetch"]
    assert steps == [((0,), {"flush": 1}), ((1,), {"flush": 2})]              # encode the last batch, then update only
    assert isinstance(out, list) and len(out) == 2 and t._calls == 7


def test_prefetch_order_inside_a_step():
    t = _trainer("prefetch")
    t._calls = 2
    t._step("b", "u")
    assert _names(t) == ["prefetch", "train_step_prefetch", "loss"]            # copy is enqueued before the replay
